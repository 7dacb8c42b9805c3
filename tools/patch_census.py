#!/usr/bin/env python
"""How many distinct base voxels / corner voxels do the rays of a pixel patch touch at one sample depth?  (CPU, numpy-free torch.)

The numbers behind K1 backward's same-base merge: at cfg-2 geometry the 16 rays of a 4x4 patch fall into ~6.7 distinct
base voxels per sample, and their 124 corner writes touch ~26.6 distinct voxels.

    python tools/patch_census.py [--img 256] [--vol 64] [--pts 64]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forge_b200 import synthetic as syn                        # noqa: E402


def camera_rays_local(R, T, K_half, S, vol, volume_size=1.0):
    """pin-hole rays in the volume's local [-1, 1] frame: o = -R^T t, d(i, j) = R^T [(u - cx) / fx, (v - cy) / fy, 1], per-axis scale
    (n - 1) / 2 * volume_size / n  ->  o [N,3], d [N,S,S,3]"""
    u = torch.arange(S, dtype=R.dtype) + 0.5
    fx, fy, cx, cy = K_half[:, 0, 0], K_half[:, 1, 1], K_half[:, 0, 2], K_half[:, 1, 2]
    N = R.shape[0]
    dx = ((u[None, None, :] - cx[:, None, None]) / fx[:, None, None]).expand(N, S, S)
    dy = ((u[None, :, None] - cy[:, None, None]) / fy[:, None, None]).expand(N, S, S)
    d_cam = torch.stack([dx, dy, torch.ones(N, S, S, dtype=R.dtype)], -1)
    Rt = R.transpose(1, 2)
    scale = (vol - 1) / 2 * (volume_size / vol)
    return -torch.einsum('nab,nb->na', Rt, T) / scale, torch.einsum('nab,nijb->nija', Rt, d_cam) / scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--img", type=int, default=256)
    ap.add_argument("--vol", type=int, default=64)
    ap.add_argument("--pts", type=int, default=64)
    a = ap.parse_args()
    S, vol, P = a.img // 2, a.vol, a.pts
    inp = syn.render_inputs(1, 5, a.img, vol)
    K = inp['K'].clone()
    K[:, :2] *= 0.5
    o, d = camera_rays_local(inp['R'], inp['T'], K, S, vol)
    zs = torch.linspace(0.5, 2.0, P)
    pts = o[:, None, None, None, :] + zs[None, None, None, :, None] * d[:, :, :, None, :]
    base = torch.floor((pts + 1) / 2 * (vol - 1)).long()
    inside = ((base >= -1) & (base <= vol - 1)).all(-1)
    pitch = vol + 6
    bid = (base[..., 2] + 1) * pitch * pitch + (base[..., 1] + 1) * pitch + (base[..., 0] + 1)
    N = bid.shape[0]
    for ph, pw in ((2, 2), (4, 4), (4, 8), (8, 8), (8, 16)):
        b = bid.reshape(N, S // ph, ph, S // pw, pw, P).permute(0, 1, 3, 5, 2, 4).reshape(-1, ph * pw)
        m = inside.reshape(N, S // ph, ph, S // pw, pw, P).permute(0, 1, 3, 5, 2, 4).reshape(-1, ph * pw)
        b = torch.where(m, b, torch.full_like(b, -1))
        sb, _ = torch.sort(b, dim=1)
        distinct = ((sb[:, 1:] != sb[:, :-1]) & (sb[:, 1:] >= 0)).sum(1) + (sb[:, 0] >= 0).long()
        nact = m.sum(1)
        sel = nact > 0
        offs = [dz * pitch * pitch + dy * pitch + dx for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)]
        allc = torch.stack([torch.where(m, b + o_, torch.full_like(b, -1)) for o_ in offs], 2).reshape(b.shape[0], -1)
        sc, _ = torch.sort(allc, dim=1)
        dc = ((sc[:, 1:] != sc[:, :-1]) & (sc[:, 1:] >= 0)).sum(1) + (sc[:, 0] >= 0).long()
        print("patch %dx%-2d: %.2f active rays, %.2f distinct base voxels (merge ratio %.2f), %.2f distinct corner voxels of %.2f "
              "corner writes (ratio %.2f)" % (ph, pw, nact[sel].float().mean(), distinct[sel].float().mean(),
                                              nact[sel].sum().item() / distinct[sel].sum().item(), dc[sel].float().mean(),
                                              8 * nact[sel].float().mean(), 8 * nact[sel].sum().item() / dc[sel].sum().item()))


if __name__ == "__main__":
    main()
