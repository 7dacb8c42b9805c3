#!/usr/bin/env python
"""fuse stage (ConvGRU over 5 views + fusion_conv + fusion_norm, 128 ch x 32^3) on one B200: cuDNN fp32, cuDNN bf16 autocast
channels-last, forge_conv3d_tc (tcgen05); plus the gate convolution alone against the measured bf16 peak."""
import argparse
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from forge_b200 import ops, synthetic as syn                     # noqa: E402
from forge_b200.models.encoder import Encoder3D                  # noqa: E402

DEV = torch.device('cuda')


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=4)
    ap.add_argument("--grid", type=int, default=32)
    args = ap.parse_args()
    b, t, C, n = args.objects, 5, 128, args.grid
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc = Encoder3D(syn.make_config()).to(DEV).eval()
    x = torch.randn(b, t, n, n, n, C, device=DEV).permute(0, 1, 5, 2, 3, 4)       # K2's output layout
    out = {"bench": "Encoder3D.fuse, %d objects x %d views x %d ch x %d^3" % (b, t, C, n)}
    with torch.no_grad():
        out["cudnn_fp32_ms"] = round(timed(lambda: enc.fuse(x)), 3)
        ref = enc.fuse(x)
        enc.channels_last_3d_()
        out["cudnn_fp32_channels_last_ms"] = round(timed(lambda: enc.fuse(x)), 3)
        enc.compute_dtype = torch.bfloat16
        enc.tc_fusion = False
        out["cudnn_bf16_autocast_ms"] = round(timed(lambda: enc.fuse(x)), 3)
        y_amp = enc.fuse(x)
        enc.tc_fusion = True
        out["forge_conv3d_tc_ms"] = round(timed(lambda: enc.fuse(x)), 3)
        y_tc = enc.fuse(x)
        sc = ref.abs().max().item()
        out["max_abs_vs_fp32"] = {"tc": round((y_tc - ref).abs().max().item() / sc, 5), "cudnn_bf16": round((y_amp - ref).abs().max().item() / sc, 5)}
        # the gate convolution alone
        pk = enc.fusion_feature._tc_packs()
        xt = x[:, 0].to(torch.bfloat16)
        h = torch.randn(b, n, n, n, C, device=DEV).permute(0, 4, 1, 2, 3)
        h16 = h.to(torch.bfloat16)
        ms = timed(lambda: ops.conv3d_tc(xt, pk['wg'], 'gate', pk['bg'], h2=h16, h_state=h), 20)
        flops = 2.0 * b * n ** 3 * 256 * 27 * 256
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
        out["gate_conv"] = {"ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "frac_of_bf16_peak": round(flops / ms / 1e9 / peaks["bf16_tflops"], 3)}
        u, hr, _ = ops.conv3d_tc(xt, pk['wg'], 'gate', pk['bg'], h2=h16, h_state=h)
        ms = timed(lambda: ops.conv3d_tc(xt, pk['wo'], 'out', pk['bo'], h2=hr, h_state=h, u_in=u, want_bf16=True), 20)
        out["out_conv"] = {"ms": round(ms, 4), "tflops": round(flops / 2 / ms / 1e9, 1), "frac_of_bf16_peak": round(flops / 2 / ms / 1e9 / peaks["bf16_tflops"], 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
