#!/usr/bin/env python
"""cuDNN configuration sweep for the FLOP-dominant host part of the step:
Encoder3D.fuse (ConvGRU over 5 views) + both heads, forward + backward, b = 1."""
import json, os, sys, time, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from forge_b200 import synthetic as syn
from forge_b200.models.encoder import Encoder3D

DEV = torch.device('cuda')
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    enc = Encoder3D(syn.make_config()).to(DEV).eval()
x0 = torch.randn(1, 5, 128, 32, 32, 32, device=DEV)


def run(x, dtype, iters=5):
    def step():
        xi = x.clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=dtype, enabled=dtype is not None):
            f = enc.fuse(xi)
            d = enc.get_density3D(f)
            r = enc.get_render_features(f)
        (d.float().mean() + r.float().mean()).backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


x_cl = x0.permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4)
for name, x, dtype, tf32 in (("NCDHW fp32 (TF32 allowed)", x0, None, True), ("NCDHW fp32 strict", x0, None, False),
                             ("channels-last fp32 (TF32 allowed)", x_cl, None, True),
                             ("NCDHW bf16 autocast", x0, torch.bfloat16, True),
                             ("channels-last bf16 autocast", x_cl, torch.bfloat16, True)):
    torch.backends.cudnn.allow_tf32 = tf32
    print(json.dumps({"fuse+heads fwd+bwd": name, "ms": round(run(x, dtype), 2)}), flush=True)
for mod in (enc.fusion_feature, enc.features_head, enc.density_head):
    mod.to(memory_format=torch.channels_last_3d)
for name, x, dtype in (("channels-last weights+input fp32 (TF32)", x_cl, None), ("channels-last weights+input bf16", x_cl, torch.bfloat16)):
    torch.backends.cudnn.allow_tf32 = True
    print(json.dumps({"fuse+heads fwd+bwd": name, "ms": round(run(x, dtype), 2)}), flush=True)
