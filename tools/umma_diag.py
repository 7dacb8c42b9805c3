"""Diagnostic for the UMMA shared-memory descriptor semantics (run on the GPU box): for each probe case print
which layout hypothesis the hardware result matches.  H1 = (LBO: K halves, SBO: 8-row groups), H2 = swapped."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from forge_b200 import ops


def expected(img, off, kstride, gstride, rows):
    r = torch.arange(rows).view(-1, 1)
    k = torch.arange(16).view(1, -1)
    idx = (off + (k // 8) * kstride + (r // 8) * gstride + (r % 8) * 16 + (k % 8) * 2) // 2
    return img[idx].float()


def main():
    torch.manual_seed(0)
    nbytes = 96 * 1024
    img = (torch.randint(-8, 9, (nbytes // 2,)).float() / 4).to(torch.bfloat16)
    dimg = img.view(torch.uint8).cuda()
    b_off = 88 * 1024
    for (a_off, a_lbo, a_sbo) in [(0, 2048, 128), (16, 2048, 128), (128, 2048, 128), (16 * 45, 6912, 128), (16 * 83, 640, 128),
                                  (16 * 164, 16, 128), (0, 128, 256), (0, 256, 128)]:
        try:
            out = ops.umma_probe(dimg, a_off, a_lbo, a_sbo, b_off, 256, 128).cpu()
            torch.cuda.synchronize()
        except Exception as e:            # noqa: BLE001
            print("case", (a_off, a_lbo, a_sbo), "FAILED:", e)
            return
        res = {}
        for an, (ak, ag) in {"H1": (a_lbo, a_sbo), "H2": (a_sbo, a_lbo)}.items():
            for bn, (bk, bg) in {"H1": (256, 128), "H2": (128, 256)}.items():
                A = expected(img, a_off, ak, ag, 128)
                B = expected(img, b_off, bk, bg, 16)
                res["A:%s B:%s" % (an, bn)] = (out - A @ B.t()).abs().max().item()
        print("case", (a_off, a_lbo, a_sbo), {k: round(v, 3) for k, v in res.items()})


if __name__ == "__main__":
    main()
