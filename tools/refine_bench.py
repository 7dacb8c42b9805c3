#!/usr/bin/env python
"""The literal per-object inner loop of the reference: test-time pose optimisation
(reference kubric_eval.py:412-530, demo.py:115-188) -- per Adam iteration

    quat/trans -> SE(3) -> rotate(features) -> sort views -> fuse (ConvGRU) -> heads -> render 5 views
    -> MSE(rgb) + MSE(mask) -> backward to the 4 relative poses -> Adam step

timed for (a) forge_b200 modules (K2 + K1 + fused decoder, cuDNN fusion/heads) and (b) the oracle =
the reference's op sequence (ATen grid_sample, PyTorch3D-style renderer, cuDNN decoder) with the SAME
weights on the same GPU.  b = 1 object, t = 5 views, 32^3 x 128ch feature volumes, 64^3 render volumes,
128^2 rays x 64 samples per view, random-init weights, synthetic targets.
"""
import argparse
import json
import os
import sys
import time
import warnings

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from forge_b200 import synthetic as syn                                   # noqa: E402
from forge_b200.models.model import FORGE, sequence_from_distance, chose_selected   # noqa: E402

DEV = torch.device('cuda')


def quat_to_pose(q, t):
    """[n,4] (w,x,y,z), [n,3] -> [n,4,4]"""
    q = F.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
    P = torch.eye(4, device=q.device).repeat(q.shape[0], 1, 1)
    P[:, :3, :3] = R
    P[:, :3, 3] = t
    return P


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--only", default="", help="run only the arms whose name contains this substring")
    args = ap.parse_args()
    torch.manual_seed(0)
    cfg = syn.make_config(img_size=256, n_pts_per_ray=64, use_gt_pose=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FORGE(cfg).to(DEV).eval()
    # random-init heads end in ReLU and would emit an (almost) empty density volume, which lets the raymarcher
    # skip its feature fetches; bias the last density conv so that every sample does full work
    model.encoder_3d.density_head[6].bias.data.fill_(0.15)
    sample = syn.kubric_batch(1, n_views_all=5, img_size=256, seed=1)
    clips = sample['images'].to(DEV)
    with torch.no_grad():
        features = model.lift(clips).detach()                     # [1,5,128,32,32,32], detached like kubric_eval.py:405
    target_rgb = torch.rand(5, 3, 256, 256, device=DEV)
    target_mask = (torch.rand(5, 1, 256, 256, device=DEV) > 0.5).float()
    canon_pose = sample['cam_poses_cv2_canonicalized'][0, 0].to(DEV)
    rel_gt = sample['cam_poses_rel_cv2'][0, 1:].to(DEV)
    K = sample['K_cv2'][0].to(DEV)

    def make_params():
        # start near the ground-truth relative poses, like the reference does from the pose network's prediction
        # (an identity rotation with the true translation points every camera away from the object: empty views,
        # exactly-zero pose gradients and no raymarching work)
        from forge_b200.models.model import _mat2quat
        q = (_mat2quat(rel_gt)[:, :4] + torch.tensor([0.0, 0.02, -0.01, 0.03], device=DEV)).clone().requires_grad_(True)
        t = (rel_gt[:, :3, 3] + 0.01).clone().requires_grad_(True)
        return q, t

    def poses_from(q, t):
        rel = quat_to_pose(q, t)
        P = torch.cat([canon_pose[None], canon_pose[None] @ rel], dim=0)[None]      # [1,5,4,4]
        E = torch.inverse(P[0])
        return P, E

    def step_forge(q, t):
        P, E = poses_from(q, t)
        idxs = sequence_from_distance(P[:, :, :3, 3])
        feat, dens = model.reconstruct(features, P, idxs)
        cam = {'R': E[:, :3, :3], 'T': E[:, :3, 3], 'K': K.clone()}
        rgb, mask, _ = model.render(cam, feat, dens, return_origin_proj=True,
                                    view2vol=torch.zeros(5, dtype=torch.int32, device=DEV))
        return F.mse_loss(rgb, target_rgb) + F.mse_loss(mask, target_mask)

    from oracle import reference_path as rp
    ren = rp.make_renderer(256, 64, cfg.render.min_depth, cfg.render.max_depth).to(DEV)

    def step_ref(q, t):
        P, E = poses_from(q, t)
        idxs = rp.sequence_from_distance(P[:, :, :3, 3])
        ft = rp.chose_selected(rp.rotate_world_forward(features, P, 32, 1.0), idxs)
        fmv = model.encoder_3d.fuse(ft)
        dens = model.encoder_3d.get_density3D(fmv)
        feat = model.encoder_3d.get_render_features(fmv)
        feat_all = feat.unsqueeze(1).repeat(1, 5, 1, 1, 1, 1).reshape(5, *feat.shape[1:])
        dens_all = dens.unsqueeze(1).repeat(1, 5, 1, 1, 1, 1).reshape(5, *dens.shape[1:])
        cam = {'R': E[:, :3, :3], 'T': E[:, :3, 3], 'K': K.clone()}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rgb, mask, _ = rp.volrender_forward(ren, model.render.conv_rgb, cam, feat_all, dens_all, 256, 1.0,
                                                return_origin_proj=True)
        return F.mse_loss(rgb, target_rgb) + F.mse_loss(mask, target_mask)

    results = {}
    from forge_b200.refine import prepare_for_pose_refinement
    arms = [("forge_b200", step_forge)] + ([] if args.no_ref else [("reference op sequence (oracle on GPU)", step_ref)])
    arms.append(("forge_b200, bf16 fusion+heads (channels-last weights)", step_forge))
    arms.append(("forge_b200, prepare_for_pose_refinement (frozen weights, channels-last, fp32)", step_forge))
    arms.append(("forge_b200, prepare_for_pose_refinement (frozen weights, bf16 fusion+heads, tensor-core decoder)", step_forge))
    for name, fn in arms:
        if args.only and args.only not in name:
            continue
        if "prepare_for" in name:
            bf = torch.bfloat16 if "bf16" in name else None
            prepare_for_pose_refinement(model, fusion_dtype=bf, decoder_dtype=bf)
        elif "bf16" in name:
            model.encoder_3d.channels_last_3d_()
            model.encoder_3d.compute_dtype = torch.bfloat16
        q, t = make_params()
        opt = torch.optim.Adam([q, t], lr=1e-3)
        for _ in range(3):
            opt.zero_grad()
            fn(q, t).backward()
            opt.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            opt.zero_grad()
            loss = fn(q, t)
            loss.backward()
            opt.step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / args.iters * 1e3
        results[name] = {"ms_per_iteration": round(ms, 3), "final_loss": round(loss.item(), 6),
                         "grad_q_norm": float("%.4e" % q.grad.norm().item()), "grad_t_norm": float("%.4e" % t.grad.norm().item())}
    # ---- the same iteration replayed as one CUDA graph (forge_b200.refine.GraphedPoseRefiner) ----
    from forge_b200.refine import GraphedPoseRefiner, make_refine_loss
    if not args.only or "graph" in args.only:
        for name, bf in (("graphed iteration, frozen weights, fp32", None), ("graphed iteration, frozen weights, bf16 fusion+heads, tensor-core decoder", torch.bfloat16)):
            prepare_for_pose_refinement(model, fusion_dtype=bf, decoder_dtype=bf)
            loss_fn = make_refine_loss(model, features, sample['K_cv2'].to(DEV), canon_pose, target_rgb, target_mask)
            q, t = make_params()
            r = GraphedPoseRefiner(loss_fn, q.detach(), t.detach(), lr=1e-3)
            r.step(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r.step(args.iters)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / args.iters * 1e3
            results[name] = {"ms_per_iteration": round(ms, 3), "final_loss": round(r.loss.item(), 6)}
    print(json.dumps({"bench": "test-time pose optimisation iteration (b=1, t=5, 64^3, 5x128^2x64)", **results}))


if __name__ == "__main__":
    main()
