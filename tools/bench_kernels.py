#!/usr/bin/env python
"""Per-kernel timings on one B200 (development aid; bench.py is the contract benchmark).

    python tools/bench_kernels.py [--reps 20] [--ref] [--only k1|k2|...]

Prints one JSON line per measurement: kernel, ms, achieved GB/s on ALGORITHMIC bytes, fraction of
the measured HBM peak.  L2 is flushed before every timed launch.  --ref also times the oracle
(the reference's PyTorch3D / ATen op sequence) on the same GPU: the denominator of the ">= 10x
the reference GPU renderer" target in BASELINE.json.
"""
import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forge_b200 import ops, synthetic as syn                     # noqa: E402
from forge_b200.models.volume_render import VolRender, camera_to_cam12   # noqa: E402
from forge_b200.models.rotate import Rotate_world                # noqa: E402

DEV = torch.device('cuda')


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, ms, best, nbytes=None, **kw):
    d = {"kernel": name, "ms_median": round(ms, 4), "ms_best": round(best, 4)}
    if nbytes:
        d["algorithmic_MB"] = round(nbytes / 1e6, 2)
        d["GBps"] = round(nbytes / ms / 1e6, 1)
        d["hbm_frac"] = round(nbytes / ms / 1e6 / peak(), 4)
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--objects", type=int, default=4)
    ap.add_argument("--img", type=int, default=256)
    ap.add_argument("--vol", type=int, default=64)
    ap.add_argument("--pts", type=int, default=64)
    ap.add_argument("--bwd", action="store_true")
    args = ap.parse_args()
    want = lambda k: (not args.only) or k in args.only.split(",")   # noqa: E731
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=DEV)
    b, t, img, D, P = args.objects, 5, args.img, args.vol, args.pts
    S = img // 2
    N = b * t
    rays = N * S * S
    inp = syn.render_inputs(b, t, img, D, seed=0, device=DEV)
    cfg = syn.make_config(img_size=img, n_pts_per_ray=P)
    m = VolRender(cfg).to(DEV).eval()
    Kh = inp['K'].clone()
    Kh /= 2
    Kh[:, 2, 2] = 1
    cam12 = camera_to_cam12(inp['R'], inp['T'], Kh, (D, D, D), 1.0).contiguous()
    zs = m._depths(DEV)
    from forge_b200 import _lib
    fp, dq = ops.pack_volume(inp['feat'], inp['dens'])
    o_feat = torch.empty(N, S, S, 16, device=DEV)
    o_sil = torch.empty(N, S, S, device=DEV)
    o_dep = torch.empty(N, S, S, device=DEV)

    def k1(entry="forge_raymarch_fwd"):
        _lib.call(entry, fp.data_ptr(), dq.data_ptr(), inp['view2vol'].data_ptr(), cam12.data_ptr(),
                  zs.data_ptr(), o_feat.data_ptr(), o_sil.data_ptr(), o_dep.data_ptr(), N, b, D, D, D, S, S, P,
                  torch.cuda.current_stream().cuda_stream)

    with torch.no_grad():
        if want("relayout"):
            ms, best = timeit(lambda: ops.pack_volume(inp['feat'], inp['dens']), args.reps, flush)
            report("pack_volume (%dx17x%d^3 NCDHW -> padded NDHWC + quads)" % (b, D), ms, best,
                   (inp['feat'].numel() + inp['dens'].numel()) * 4 + (fp.numel() + dq.numel()) * 4)
        if want("k1"):
            k1_bytes = b * 17 * D ** 3 * 4 + rays * 18 * 4 + N * 48
            ms, best = timeit(k1, args.reps, flush)
            report("raymarch_fwd", ms, best, k1_bytes, Mrays_per_s=round(rays / ms / 1e3, 1),
                   fp32_TFLOPs=round(rays * P * 366 / ms / 1e9, 2))
        if want("k1ab"):       # the two forward formulations side by side
            k1_bytes = b * 17 * D ** 3 * 4 + rays * 18 * 4 + N * 48
            for entry in ("forge_raymarch_fwd_gather", "forge_raymarch_fwd_tma"):
                ms, best = timeit(lambda: k1(entry), args.reps, flush)
                report(entry, ms, best, k1_bytes, Mrays_per_s=round(rays / ms / 1e3, 1))
        if want("k1bwd"):      # the backward kernel alone (C-ABI call, gradient buffers preallocated and accumulated into)
            go = torch.randn(N, S, S, 16, device=DEV)
            gs = torch.randn(N, S, S, device=DEV)
            gfp = torch.zeros_like(fp)
            gdp = torch.zeros(b, D + 2, D + 2, D + 2, device=DEV)
            gc = torch.zeros_like(cam12)
            ws = torch.empty(_lib.load().forge_raymarch_bwd_workspace(N, b, D, D, D, S, S, P) // 4, device=DEV)

            def k1b(gf_, gd_, gc_):
                _lib.call("forge_raymarch_bwd", fp.data_ptr(), dq.data_ptr(), inp['view2vol'].data_ptr(), cam12.data_ptr(),
                          zs.data_ptr(), go.data_ptr(), gs.data_ptr(), gs.data_ptr(), gf_, gd_, gc_, ws.data_ptr(),
                          N, b, D, D, D, S, S, P, torch.cuda.current_stream().cuda_stream)
            for name, a in (("all gradients", (gfp.data_ptr(), gdp.data_ptr(), gc.data_ptr())),
                            ("grad_feat + grad_cam", (gfp.data_ptr(), None, gc.data_ptr())),
                            ("grad_dens + grad_cam", (None, gdp.data_ptr(), gc.data_ptr())),
                            ("pose only", (None, None, gc.data_ptr()))):
                ms, best = timeit(lambda: k1b(*a), args.reps, flush)
                report("raymarch_bwd kernel, " + name, ms, best, None, Mrays_per_s=round(rays / ms / 1e3, 1),
                       merge=os.environ.get("FORGE_K1B_MERGE", "default"))
        if want("volrender"):
            def full():
                cam = dict(R=inp['R'], T=inp['T'], K=inp['K'].clone())
                return m(cam, inp['feat'], inp['dens'], render_depth=True, return_origin_proj=True, view2vol=inp['view2vol'])
            ms, best = timeit(full, args.reps, flush)
            report("VolRender.forward (pack+K1+conv_rgb cuDNN+upsample)", ms, best, None, Mrays_per_s=round(rays / ms / 1e3, 1))
            x = torch.randn(N, S, S, 16, device=DEV)
            ms, best = timeit(lambda: m.decode(x), args.reps, flush)
            report("decoder FUSED fp32 kernel (forge_decoder_fwd)", ms, best, None,
                   fp32_TFLOPs=round(N * (2 * S) ** 2 * 6104 * 2 / ms / 1e9, 2))
            m.decoder_dtype = torch.bfloat16
            ms, best = timeit(lambda: m.decode(x), args.reps, flush)
            report("decoder TENSOR-CORE bf16 kernel (forge_decoder_tc_fwd, tcgen05)", ms, best, (x.numel() + N * 3 * img * img) * 4,
                   bf16_TFLOPs_useful=round(N * (2 * S) ** 2 * 6104 * 2 / ms / 1e9, 2))
            m.decoder_dtype = None
            # constant-weights backward (pose refinement): one kernel from the forward's sign masks vs cuDNN through autograd
            masks = torch.empty(N, img, img, dtype=torch.int32, device=DEV)
            rgb_tmp = torch.empty(N, 3, img, img, device=DEV)
            _lib.call("forge_decoder_fwd", x.data_ptr(), m._decoder_pack(DEV).data_ptr(), rgb_tmp.data_ptr(), masks.data_ptr(),
                      N, S, S, torch.cuda.current_stream().cuda_stream)
            wb = m._decoder_pack(DEV, kind='bwd')
            g_rgb = torch.randn(N, 3, img, img, device=DEV)
            gx = torch.empty_like(x)
            ms, best = timeit(lambda: _lib.call("forge_decoder_bwd_data", g_rgb.data_ptr(), masks.data_ptr(), wb.data_ptr(),
                                                gx.data_ptr(), N, S, S, torch.cuda.current_stream().cuda_stream), args.reps, flush)
            report("decoder BACKWARD-DATA kernel, constant weights (forge_decoder_bwd_data)", ms, best, None,
                   fp32_TFLOPs=round(N * (2 * S) ** 2 * 7930 * 2 / ms / 1e9, 2))
            ms, best = timeit(lambda: _lib.call("forge_decoder_fwd", x.data_ptr(), m._decoder_pack(DEV).data_ptr(), rgb_tmp.data_ptr(),
                                                masks.data_ptr(), N, S, S, torch.cuda.current_stream().cuda_stream), args.reps, flush)
            report("decoder FUSED fp32 kernel + sign-mask emission", ms, best, None)
            for p_ in m.conv_rgb.parameters():
                p_.requires_grad_(False)

            def cudnn_bwd():
                with torch.enable_grad():
                    xi = x.detach().requires_grad_(True)
                    torch.relu(m.conv_rgb(xi.permute(0, 3, 1, 2))).backward(g_rgb)
                return xi.grad
            ms, best = timeit(cudnn_bwd, max(3, args.reps // 2), flush)
            report("decoder conv_rgb forward + backward-data through autograd (cuDNN, constant weights)", ms, best, None)
            del masks, rgb_tmp, g_rgb, gx
            m.fused_decoder = False
            for name, setup in (("fp32 (TF32 allowed, torch default)", lambda: None),
                                ("fp32 strict (allow_tf32=False)", lambda: setattr(torch.backends.cudnn, 'allow_tf32', False)),
                                ("bf16 autocast", lambda: setattr(m, 'decoder_dtype', torch.bfloat16))):
                setup()
                ms, best = timeit(lambda: m.decode(x), args.reps, flush)
                report("decoder conv_rgb cuDNN, " + name, ms, best, None)
            m.decoder_dtype = None
            m.fused_decoder = True
            torch.backends.cudnn.allow_tf32 = True
            xc = x.permute(0, 3, 1, 2).contiguous()
            ms, best = timeit(lambda: torch.relu(m.conv_rgb(xc)), args.reps, flush)
            report("decoder conv_rgb cuDNN, NCHW-contiguous input", ms, best, None)
            sil = torch.randn(N, S, S, device=DEV)
            ms, best = timeit(lambda: torch.nn.functional.interpolate(sil.unsqueeze(1), size=[img, img], mode='bilinear'), args.reps, flush)
            report("bilinear x2 upsample of one [N,1,S,S] map (torch)", ms, best, None)
        if want("k2"):
            C, n = 128, D // 2
            vox, poses = syn.rotate_inputs(b, t, C, n, seed=1, device=DEV)
            rot = Rotate_world(cfg).to(DEV)
            vox_cl = vox.permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4)
            k2_bytes = 2 * b * t * C * n ** 3 * 4
            vcl = vox_cl.permute(0, 1, 3, 4, 5, 2).reshape(b * t, n, n, n, C)
            assert vcl.data_ptr() == vox_cl.data_ptr()
            gxd, gyd, gzd, gmax = rot._device_axes(n, n, n, DEV)
            A = torch.zeros(b, t, 3, 4, device=DEV)
            A[:, 1:] = rot.get_transformation(poses).reshape(b, t - 1, 4, 4)[:, :, :3, :]
            A = A.reshape(b * t, 12).contiguous()
            jobs = rot._jobs(b, t, DEV, None)
            out_cl = torch.empty_like(vcl)

            def k2():
                _lib.call("forge_rotate_fwd", vcl.data_ptr(), A.data_ptr(), jobs.data_ptr(), gxd.data_ptr(), gyd.data_ptr(),
                          gzd.data_ptr(), float(gmax), out_cl.data_ptr(), b * t, C, n, n, n,
                          torch.cuda.current_stream().cuda_stream)
            ms, best = timeit(k2, args.reps, flush)
            report("rotate_fwd kernel (C-ABI call, %dx%dx%dx%d^3 channels-last)" % (b, t, C, n), ms, best, k2_bytes)
            ms, best = timeit(lambda: rot(vox_cl, poses, grid_size=n), args.reps, flush)
            report("Rotate_world.forward channels-last in/out (kernel + torch pose glue)", ms, best, k2_bytes)
            ms, best = timeit(lambda: rot(vox, poses, grid_size=n), args.reps, flush)
            report("Rotate_world.forward NCDHW in (+relayout kernel)", ms, best, k2_bytes)
            # transform jobs only (view 0 aliased): SURVEY 8d's numerator, what bench.py's roofline_rotate times
            jobs_t = rot._jobs_aliased(b, t, DEV, None)
            out_t = torch.empty(b * (t - 1), n, n, n, C, device=DEV)
            ms, best = timeit(lambda: _lib.call("forge_rotate_fwd", vcl.data_ptr(), A.data_ptr(), jobs_t.data_ptr(), gxd.data_ptr(),
                                                gyd.data_ptr(), gzd.data_ptr(), float(gmax), out_t.data_ptr(), b * (t - 1), C, n, n, n,
                                                torch.cuda.current_stream().cuda_stream), args.reps, flush)
            report("rotate_fwd kernel, %d transform jobs only" % (b * (t - 1)), ms, best, 2 * b * (t - 1) * C * n ** 3 * 4,
                   knobs={k: os.environ.get(k) for k in ("FORGE_K2_LEAN", "FORGE_K2_SHAPE", "FORGE_K2_STREAM")})
            if args.ref:
                from oracle import reference_path as rp
                ms, best = timeit(lambda: rp.rotate_world_forward(vox, poses, n, 1.0), max(3, args.reps // 4), flush)
                report("REFERENCE rotate (ATen grid_sample path) on GPU", ms, best, k2_bytes)
        if args.ref and want("k1"):
            from oracle import reference_path as rp
            ren = rp.make_renderer(img, P, 0.5, 2.0).to(DEV)
            idx = inp['view2vol'].long()
            fa, da = inp['feat'][idx], inp['dens'][idx]      # as-called: one materialised volume per view

            def ref():
                cam = dict(R=inp['R'], T=inp['T'], K=inp['K'].clone())
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    return rp.raymarch_core(ren, cam, fa, da, img, 1.0, True)
            ms, best = timeit(ref, max(3, args.reps // 4), flush)
            report("REFERENCE raymarch core (PyTorch3D op sequence) on GPU", ms, best, None, Mrays_per_s=round(rays / ms / 1e3, 2))

    if args.bwd:
        feat_g = inp['feat'].clone().requires_grad_(True)
        dens_g = inp['dens'].clone().requires_grad_(True)
        cam_g = cam12.clone().requires_grad_(True)
        go = torch.randn(N, S, S, 16, device=DEV)
        gs = torch.randn(N, S, S, device=DEV)

        def fb(pose_only):
            f = inp['feat'] if pose_only else feat_g
            d = inp['dens'] if pose_only else dens_g
            o, s, dep = ops.raymarch(f, d, cam_g, inp['view2vol'], zs, S, S, True)
            torch.autograd.backward([o, s, dep], [go, gs, gs])
        for po in (False, True):
            ms, best = timeit(lambda: fb(po), args.reps, flush)
            report("raymarch fwd+bwd" + (" (pose-only grads)" if po else " (all grads)"), ms, best, None,
                   Mrays_per_s=round(rays / ms / 1e3, 1))
        C, n = 128, D // 2
        vox, poses = syn.rotate_inputs(b, t, C, n, seed=1, device=DEV)
        rot = Rotate_world(cfg).to(DEV)
        vox_cl = vox.permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4).requires_grad_(True)
        poses_g = poses.clone().requires_grad_(True)
        g = torch.randn(b, t, n, n, n, C, device=DEV).permute(0, 1, 5, 2, 3, 4)

        def rfb():
            out = rot(vox_cl, poses_g, grid_size=n)
            out.backward(g)
        ms, best = timeit(rfb, args.reps, flush)
        report("rotate fwd+bwd (all grads)", ms, best, 4 * b * t * C * n ** 3 * 4)


if __name__ == "__main__":
    main()
