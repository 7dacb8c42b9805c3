#!/usr/bin/env python
"""Full FORGE.forward step on one B200 (secondary line of SURVEY 8d): b objects x 5 input views -> t_all rendered
views at the step-3.3 geometry (256^2 images, 32^3 x 128 fusion volumes, 64^3 x 17 render volumes, 128^2 rays x 64
samples), ground-truth poses (the pose networks are out of scope), random-init weights, synthetic Kubric-shaped batch.

    python tools/model_bench.py [--objects 4] [--views 10] [--iters 10] [--no-ref] [--train]

Arms: (a) forge_b200 modules, (b) the reference's op sequence (oracle: ATen grid_sample rotate, per-view repeated
volumes, PyTorch3D-style renderer, cuDNN decoder) with the SAME weights and the same cuDNN lift / fuse / heads,
(c) forge_b200 with bf16 fusion + heads and the bf16 tensor-core decoder.  Prints one JSON line with ms per step and a
stage breakdown (CUDA events) of arm (a).
"""
import argparse
import json
import os
import sys
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forge_b200 import synthetic as syn                                   # noqa: E402
from forge_b200.models.model import FORGE, sequence_from_distance          # noqa: E402

DEV = torch.device('cuda')


def timed(fn, iters):
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
    ap.add_argument("--views", type=int, default=10)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--train", action="store_true", help="also time forward + backward (train mode, MSE losses)")
    args = ap.parse_args()
    torch.manual_seed(0)
    b, t_all = args.objects, args.views
    cfg = syn.make_config(img_size=256, n_pts_per_ray=64, use_gt_pose=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FORGE(cfg).to(DEV).eval()
    model.encoder_3d.density_head[6].bias.data.fill_(0.15)     # non-empty density volume: every sample does full work
    sample = syn.kubric_batch(b, n_views_all=t_all, img_size=256, seed=1)
    sample = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in sample.items()}
    out = {"bench": "FORGE.forward, %d objects x 5 input views -> %d rendered views, 256^2 images, GT poses" % (b, t_all)}

    K0 = sample['K_cv2'].clone()

    def fwd():
        # VolRender.forward halves the caller's K in place (reference models/volume_render.py:50-51) and .to(device) of a tensor
        # that already lives on the device is the same tensor: hand every call a fresh copy
        sample['K_cv2'] = K0.clone()
        with torch.no_grad():
            return model(sample, None, DEV)

    out["forge_b200 fp32 ms"] = round(timed(fwd, args.iters), 3)

    # ---- stage breakdown of the same forward ----
    def stage_breakdown():
        from forge_b200.models.model import _view0_first
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        with torch.no_grad():
            clips = sample['images'][:, :5]
            ev[0].record()
            feats = model.lift(clips)
            ev[1].record()
            P = sample['cam_poses_cv2_canonicalized'][:, :5]
            idxs = _view0_first(sequence_from_distance(P[:, :, :3, 3]))
            rot = model.rotate.forward_views(voxels=feats, camPoses_cv2=P, grid_size=feats.shape[3], order=idxs)
            ev[2].record()
            fmv = model.encoder_3d.fuse(rot)
            ev[3].record()
            feat, dens = model.encoder_3d.get_render_volumes(fmv)
            ev[4].record()
            E = sample['cam_extrinsics_cv2_canonicalized'].reshape(b * t_all, 4, 4)
            cam = {'R': E[:, :3, :3], 'T': E[:, :3, 3], 'K': K0.reshape(b * t_all, 3, 3).clone()}
            v2v = torch.arange(b, device=DEV).repeat_interleave(t_all).int()
            model.render(cam, feat, dens, return_origin_proj=True, view2vol=v2v)
            ev[5].record()
            torch.cuda.synchronize()
        names = ["lift (ResNet50 + 3-D conv, cuDNN)", "sort + rotate (K2)", "fuse (ConvGRU)", "heads",
                 "render (camera prep + pack + K1 + decoder + upsample)"]
        return {n: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, n in enumerate(names)}
    stage_breakdown()
    out["stages_ms"] = stage_breakdown()

    if not args.no_ref:
        from oracle import reference_path as rp
        ren = rp.make_renderer(256, 64, cfg.render.min_depth, cfg.render.max_depth).to(DEV)

        def ref():
            with torch.no_grad(), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                feats = model.lift(sample['images'][:, :5])
                P = sample['cam_poses_cv2_canonicalized'][:, :5]
                idxs = rp.sequence_from_distance(P[:, :, :3, 3])
                ft = rp.chose_selected(rp.rotate_world_forward(feats.contiguous(), P, feats.shape[3], 1.0), idxs)
                fmv = model.encoder_3d.fuse(ft)
                dens = model.encoder_3d.get_density3D(fmv)
                feat = model.encoder_3d.get_render_features(fmv)
                fa = feat.unsqueeze(1).repeat(1, t_all, 1, 1, 1, 1).reshape(b * t_all, *feat.shape[1:])
                da = dens.unsqueeze(1).repeat(1, t_all, 1, 1, 1, 1).reshape(b * t_all, *dens.shape[1:])
                E = sample['cam_extrinsics_cv2_canonicalized'].reshape(b * t_all, 4, 4)
                cam = {'R': E[:, :3, :3], 'T': E[:, :3, 3], 'K': K0.reshape(b * t_all, 3, 3).clone()}
                return rp.volrender_forward(ren, model.render.conv_rgb, cam, fa, da, 256, 1.0, return_origin_proj=True)
        out["reference op sequence (oracle on GPU, same cuDNN lift/fuse/heads) ms"] = round(timed(ref, max(2, args.iters // 2)), 3)

    if args.train:
        model.train()
        tgt_rgb = torch.rand(b * t_all, 3, 256, 256, device=DEV)
        tgt_mask = (torch.rand(b * t_all, 1, 256, 256, device=DEV) > 0.5).float()

        def step():
            model.zero_grad(set_to_none=True)
            sample['K_cv2'] = K0.clone()
            rgb, mask = model(sample, None, DEV)
            (F.mse_loss(rgb, tgt_rgb) + F.mse_loss(mask, tgt_mask)).backward()
        out["forge_b200 fp32 train fwd+bwd ms"] = round(timed(step, max(2, args.iters // 2)), 3)
        model.eval()

    # ---- channels-last conv weights (2-D trunk + 3-D stack), then bf16 autocast for them + the tensor-core decoder ----
    model.encoder_3d.channels_last_3d_()
    out["forge_b200 fp32, channels-last conv weights ms"] = round(timed(fwd, args.iters), 3)
    model.encoder_3d.compute_dtype = torch.bfloat16
    model.render.decoder_dtype = torch.bfloat16
    out["forge_b200 bf16 lift/fusion/heads + tensor-core decoder ms"] = round(timed(fwd, args.iters), 3)
    stage_breakdown()
    out["stages_ms bf16 (tensor-core fusion + decoder)"] = stage_breakdown()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
