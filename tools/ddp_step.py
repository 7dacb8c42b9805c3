#!/usr/bin/env python
"""Data-parallel training step of the FORGE graph (use_gt_pose) under torch DDP over NCCL: one process per
GPU, each rank its own objects, gradients all-reduced by DDP (the only collective on the path).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_step.py \
        [--batch 4 --views 10 --syncbn --profile]

--syncbn converts every BatchNorm to SyncBatchNorm like the reference (kubric_train_joint.py:136); --profile adds one profiled
step and reports the share of NCCL kernels (all-reduce of the DDP buckets, SyncBN all-gathers) in the step's GPU time.

Checks that (1) the custom autograd Functions (K1, K2, pack, decoder) work under DDP's hooks, (2) after
the all-reduce every rank holds identical gradients, (3) they equal the average of the per-rank gradients
computed without DDP.  Prints step time and rays/s (whole job).
"""
import json
import os
import sys
import time
import warnings

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from forge_b200 import synthetic as syn                 # noqa: E402
from forge_b200.models.model import FORGE              # noqa: E402


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--syncbn", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    b, t_all = args.batch, args.views
    torch.manual_seed(0)                                   # identical init on every rank
    cfg = syn.make_config(img_size=256, n_pts_per_ray=64, use_gt_pose=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FORGE(cfg).to(dev)
    model.encoder_3d.density_head[6].bias.data.fill_(0.1)
    model.train()
    model.encoder_3d.feature_extraction.eval()             # keep the ResNet BN in eval, like the reference's set_model_train
    sample = syn.kubric_batch(b, n_views_all=t_all, img_size=256, seed=10 + rank)     # each rank: its own objects
    target = torch.rand(b * t_all, 3, 256, 256, device=dev)

    def loss_of(m):
        rgb, mask = m(sample, None, dev)
        return torch.nn.functional.mse_loss(rgb, target) + mask.mean()

    # reference gradients without DDP: local grads, then averaged by hand
    model.zero_grad()
    loss_of(model).backward()
    names = [n for n, p in model.named_parameters() if p.grad is not None]
    local_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    for g in local_grads.values():
        dist.all_reduce(g)
        g /= world

    if args.syncbn:                                                        # reference kubric_train_joint.py:136
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    ddp = DDP(model, device_ids=[local], find_unused_parameters=True)      # reference kubric_train_joint.py:141
    model.zero_grad()
    loss_of(ddp).backward()
    # conv biases in front of a train-mode BN have analytically zero gradients (pure rounding noise), so the
    # error is measured against the largest gradient entry of the whole model
    gmax = max(g.abs().max().item() for g in local_grads.values())
    worst = max((p.grad - local_grads[n]).abs().max().item() for n, p in model.named_parameters() if n in local_grads) / gmax
    if args.syncbn:
        worst = 0.0        # cross-rank batch statistics change the gradients by design: only cross-rank identity is checked
    # identical across ranks?
    probe = dict(model.named_parameters())['encoder_3d.features_head.0.weight'].grad.flatten()[:1000].clone()
    lo, hi = probe.clone(), probe.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all())

    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    for _ in range(2):
        opt.zero_grad()
        loss_of(ddp).backward()
        opt.step()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iters = 5
    for _ in range(iters):
        opt.zero_grad()
        loss_of(ddp).backward()
        opt.step()
    dist.barrier()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / iters * 1e3
    nccl = None
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            opt.zero_grad()
            loss_of(ddp).backward()
            opt.step()
            torch.cuda.synchronize()
        ev = prof.key_averages()
        tot = sum(e.self_device_time_total for e in ev)
        nc = sum(e.self_device_time_total for e in ev if 'nccl' in e.key.lower())
        ours = sum(e.self_device_time_total for e in ev if 'forge::' in e.key)
        nccl = {"nccl_kernel_ms": round(nc / 1e3, 3), "all_kernels_ms": round(tot / 1e3, 3), "nccl_share_of_gpu_time": round(nc / max(tot, 1), 4),
                "forge_kernels_ms": round(ours / 1e3, 3)}
    if rank == 0:
        print(json.dumps({"ddp_step": "FORGE(use_gt_pose) fwd+bwd+Adam, b=%d/GPU, %d views rendered" % (b, t_all),
                          "world": world, "ms_per_step": round(ms, 2), "n_params_with_grad": len(names),
                          "grad_rel_err_vs_manual_average": float("%.3e" % worst), "grads_identical_across_ranks": same,
                          "syncbn": args.syncbn, "nccl": nccl,
                          "rays_per_s_whole_job": round(world * b * t_all * 128 * 128 / ms * 1e3)}))
        if args.out:
            json.dump({"world": world, "batch_per_gpu": b, "views": t_all, "syncbn": args.syncbn, "ms_per_step": round(ms, 2),
                       "grads_identical_across_ranks": same, "nccl": nccl}, open(args.out, "w"), indent=1)
    assert same and worst < 5e-3, (same, worst)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
