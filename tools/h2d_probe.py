#!/usr/bin/env python
"""Host-link probe for the e2e scaling question: pinned H2D / D2H copies of the e2e step's sizes
(71.3 MB in, 23.6 MB out) on N ranks at once, nothing else running -- separates the box's PCIe / host-DRAM ceiling from
anything in StreamedRenderer.   torchrun --nproc-per-node N tools/h2d_probe.py [--out file]"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h_in = torch.empty(71_304_928 // 4).pin_memory()
    h_out = torch.empty(23_592_960 // 4).pin_memory()
    d_in, d_out = torch.empty_like(h_in, device=dev), torch.empty_like(h_out, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for name, both in (("h2d_only", False), ("h2d_and_d2h_concurrent", True)):
        for _ in range(5):
            d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a)
        s2.wait_event(a)
        for _ in range(args.iters):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / args.iters
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = {"ms_per_step_max_over_ranks": round(t.item(), 4),
                     "h2d_GBps_per_gpu": round(h_in.numel() * 4 / t.item() / 1e6, 2),
                     "aggregate_GBps": round(world * (h_in.numel() * 4 + (h_out.numel() * 4 if both else 0)) / t.item() / 1e6, 2),
                     "rays_per_s_ceiling_whole_job": round(world * 327680 / t.item() * 1e3)}
    try:
        aff = sorted(os.sched_getaffinity(0))
        res["cpu_affinity"] = "%d cpus, %d..%d" % (len(aff), aff[0], aff[-1])
    except Exception:
        pass
    if rank == 0:
        res["world"] = world
        print(json.dumps(res))
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
