#!/usr/bin/env python
"""bench.py -- rays/s of the FORGE volume-render hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl forge|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], per GPU): 4 objects x 5 views, 128x128 rays per view (img 256
rendered at half resolution, reference models/volume_render.py:18-19), 64 samples per ray, 64^3
voxels x (16 feature + 1 density) channels, fp32.  A step renders all 20 views: re-layout of the
4 DISTINCT volumes to channels-last + the fused raymarch kernel -> 16-ch feature image,
silhouette, depth.  Objects are independent, so N GPUs = N x the work (weak scaling), no
collective on the data path.

One JSON line on stdout (rank 0).  `value` = device-resident inputs, CUDA-event timed, L2 flushed
between steps; `e2e` = pinned-host inputs -> H2D -> public module API -> D2H, all inside the timed
region; `roofline` = the raymarch kernel alone against the measured HBM peak; `cpu_baseline` =
the oracle (restated reference PyTorch3D path) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(objects=4, views=5, img_size=256, vol=64, n_pts=64, min_depth=0.5, max_depth=2.0, volume_size=1.0)
METRIC = "rays/sec (5-view 128^2, 64 samples, 64^3 voxels)"
FLOPS_PER_SAMPLE = 366


def workload_name():
    return "cfg-2: %d objects x %d views %dx%d rays, %d^3 voxels x17ch, %d samples, fp32" % (
        CFG['objects'], CFG['views'], CFG['img_size'] // 2, CFG['img_size'] // 2, CFG['vol'], CFG['n_pts'])


def rays_per_step():
    return CFG['objects'] * CFG['views'] * (CFG['img_size'] // 2) ** 2


def algorithmic_bytes_k1(depth=True):
    """SURVEY 8d contract figure: distinct volumes once + outputs once + cameras."""
    V, D, N = CFG['objects'], CFG['vol'], CFG['objects'] * CFG['views']
    return V * 17 * D ** 3 * 4 + rays_per_step() * (17 + (1 if depth else 0)) * 4 + N * 48


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_inputs(seed, device):
    from forge_b200 import synthetic as syn
    return syn.render_inputs(CFG['objects'], CFG['views'], CFG['img_size'], CFG['vol'], seed=seed, device=device)


def make_config():
    from forge_b200 import synthetic as syn
    return syn.make_config(img_size=CFG['img_size'], n_pts_per_ray=CFG['n_pts'], min_depth=CFG['min_depth'],
                           max_depth=CFG['max_depth'], volume_size=CFG['volume_size'])


# -------------------------------------------------------------------------------------------------
def cpu_reference_rate(n_objects, reps, threads=None):
    """Oracle (= the reference's PyTorch3D op sequence) on host cores; returns (rays/s, cores, sample)."""
    import warnings
    from oracle import reference_path as rp
    from forge_b200 import synthetic as syn
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = syn.render_inputs(n_objects, CFG['views'], CFG['img_size'], CFG['vol'], seed=0)
    ren = rp.make_renderer(CFG['img_size'], CFG['n_pts'], CFG['min_depth'], CFG['max_depth'])
    idx = inp['view2vol'].long()
    feat, dens = inp['feat'][idx], inp['dens'][idx]          # as-called: one volume per view
    best = float('inf')
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in range(reps + 1):                            # first pass is the warm-up
            cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
            t0 = time.perf_counter()
            rp.raymarch_core(ren, cam, feat, dens, CFG['img_size'], CFG['volume_size'], render_depth=True)
            dt = time.perf_counter() - t0
            if r > 0:
                best = min(best, dt)
    rays = n_objects * CFG['views'] * (CFG['img_size'] // 2) ** 2
    sample = "%d object(s) x %d views of the workload, best of %d after 1 warm-up" % (n_objects, CFG['views'], reps)
    return rays / best, threads, sample, best


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (oracle port), rank 0 only."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    import warnings
    from oracle import reference_path as rp
    from forge_b200 import synthetic as syn
    n_obj = 1                                               # bounded sample: 1 of the 4 objects per step
    inp = syn.render_inputs(n_obj, CFG['views'], CFG['img_size'], CFG['vol'], seed=0)
    ren = rp.make_renderer(CFG['img_size'], CFG['n_pts'], CFG['min_depth'], CFG['max_depth'])
    idx = inp['view2vol'].long()
    feat, dens = inp['feat'][idx], inp['dens'][idx]
    rays = n_obj * CFG['views'] * (CFG['img_size'] // 2) ** 2

    def step():
        cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
        rp.raymarch_core(ren, cam, feat, dens, CFG['img_size'], CFG['volume_size'], render_depth=True)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
    value = rays / dt
    sample = "each step = %d object x %d views of the workload (1/%d of a full step)" % (n_obj, CFG['views'], CFG['objects'])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "arm": "oracle port of the reference PyTorch3D path on host CPU"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def reduce_times(total_ms, e2e_ms, k1_ms, world, device):
    """Job-level times = MAX over ranks (every rank processes the same amount of work)."""
    t = torch.tensor([total_ms, e2e_ms, k1_ms], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


# -------------------------------------------------------------------------------------------------
def run_forge(args, rank, world, local_rank):
    import torch.distributed as dist
    from forge_b200 import ops, _lib
    from forge_b200.models.volume_render import VolRender
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = make_config()
    model = VolRender(cfg).to(dev).eval()
    inp = make_inputs(seed=rank, device=dev)
    feat, dens, view2vol = inp['feat'], inp['dens'], inp['view2vol']
    S, D = CFG['img_size'] // 2, CFG['vol']
    Kh = inp['K'].clone()
    Kh /= 2.0
    Kh[:, 2, 2] = 1.0
    cam12, _ = ops.camera_prep(inp['R'], inp['T'], Kh, (D, D, D), CFG['volume_size'])
    zs = model._depths(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2

    launches = [0]
    N = CFG['objects'] * CFG['views']
    o_feat = torch.empty(N, S, S, 16, device=dev)
    o_sil = torch.empty(N, S, S, device=dev)
    o_dep = torch.empty(N, S, S, device=dev)

    def raymarch_packed(fp, dq):
        """the C-ABI call itself (forge_raymarch_fwd) on already-packed volumes"""
        _lib.call("forge_raymarch_fwd", fp.data_ptr(), dq.data_ptr(), view2vol.data_ptr(), cam12.data_ptr(),
                  zs.data_ptr(), o_feat.data_ptr(), o_sil.data_ptr(), o_dep.data_ptr(), N, CFG['objects'], D, D, D,
                  S, S, CFG['n_pts'], torch.cuda.current_stream(dev).cuda_stream)
        return o_feat, o_sil, o_dep

    def run_steps(n):
        """n steps, each: [untimed L2 flush] a | pack (launch 1) | b | raymarch (launch 2) | c"""
        evs = []
        keep = None
        for _ in range(n):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            flush.zero_()
            a.record()
            fp, dq = ops.pack_volume(feat, dens)
            b.record()
            raymarch_packed(fp, dq)
            c.record()
            launches[0] += 2
            keep = (fp, dq)
            evs.append((a, b, c))
        return evs, keep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        # the clock sampler is started BEFORE the warm-up so that nvidia-smi's own start-up (NVML
        # init takes the driver lock for tens of ms) is over when the timed region begins
        sampler = ClockSampler(local_rank)
        if rank == 0 and not args.no_clocks:
            sampler.start()
            time.sleep(1.0)
        run_steps(max(args.warmup, 3))          # same code path as the timed steps (allocator warm too)
        barrier()
        # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between steps ----
        launches[0] = 0
        wall0 = time.perf_counter()
        ev, _ = run_steps(args.steps)
        barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if rank == 0 else None
        step_ms = [a.elapsed_time(c) for a, b, c in ev]
        k1_ms = [b.elapsed_time(c) for a, b, c in ev]
        total_ms = sum(step_ms)
        n_launch = launches[0]
        if args.verbose and rank == 0:
            sys.stderr.write("step ms: %s\n" % " ".join("%.3f" % x for x in step_ms))

        # ---- secondary: K2 (the HBM-bound kernel of the path) on cfg-2's fusion grid, same run ------------
        from forge_b200 import synthetic as syn
        from forge_b200.models.rotate import Rotate_world
        Cr, nr, tr = 128, D // 2, CFG['views']
        rot = Rotate_world(cfg).to(dev)
        vox, poses = syn.rotate_inputs(CFG['objects'], tr, Cr, nr, seed=100 + rank, device=dev)
        vcl = vox.permute(0, 1, 3, 4, 5, 2).contiguous().reshape(CFG['objects'] * tr, nr, nr, nr, Cr)
        del vox
        gxd, gyd, gzd, gmax = rot._device_axes(nr, nr, nr, dev)
        A = ops.pose_affine(poses)
        jobs = rot._jobs(CFG['objects'], tr, dev, None)
        out_cl = torch.empty_like(vcl)
        k2_ms = []
        for it in range(3 + 10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.call("forge_rotate_fwd", vcl.data_ptr(), A.data_ptr(), jobs.data_ptr(), gxd.data_ptr(), gyd.data_ptr(),
                      gzd.data_ptr(), float(gmax), out_cl.data_ptr(), CFG['objects'] * tr, Cr, nr, nr, nr,
                      torch.cuda.current_stream(dev).cuda_stream)
            b.record()
            torch.cuda.synchronize(dev)
            if it >= 3:
                k2_ms.append(a.elapsed_time(b))
        k2_avg_ms = sum(k2_ms) / len(k2_ms)
        k2_bytes = 2 * vcl.numel() * 4
        del vcl, out_cl

        # ---- secondary: the tensor-core (bf16, tcgen05) decoder on the raymarcher's output, same run --------------
        model.decoder_dtype = torch.bfloat16
        wtc = model._decoder_pack(dev, kind='tc')
        rgb = torch.empty(N, 3, 2 * S, 2 * S, device=dev)
        dec_ms = []
        for it in range(3 + 10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.call("forge_decoder_tc_fwd", o_feat.data_ptr(), wtc.data_ptr(), rgb.data_ptr(), None, N, S, S, 0,
                      torch.cuda.current_stream(dev).cuda_stream)
            b.record()
            torch.cuda.synchronize(dev)
            if it >= 3:
                dec_ms.append(a.elapsed_time(b))
        dec_avg_ms = sum(dec_ms) / len(dec_ms)
        dec_flops = 2.0 * 6104 * N * (2 * S) ** 2          # MACs per output pixel: 16*16*9 + 8*16*25 + 3*8*25
        model.decoder_dtype = None
        del rgb

        # ---- end to end: pinned host inputs -> H2D -> public API -> D2H, everything timed ----------
        h_feat, h_dens = feat.cpu().pin_memory(), dens.cpu().pin_memory()
        h_R, h_T, h_K = inp['R'].cpu().pin_memory(), inp['T'].cpu().pin_memory(), inp['K'].cpu().pin_memory()
        h_out = torch.empty(CFG['objects'] * CFG['views'], S, S, 16).pin_memory()
        h_sil = torch.empty(CFG['objects'] * CFG['views'], S, S).pin_memory()
        h_dep = torch.empty(CFG['objects'] * CFG['views'], S, S).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in (h_feat, h_dens, h_R, h_T, h_K)) + 4 * N
        d2h = sum(t.numel() * t.element_size() for t in (h_out, h_sil, h_dep))

        def e2e_step():
            d_feat = h_feat.to(dev, non_blocking=True)
            d_dens = h_dens.to(dev, non_blocking=True)
            cam = dict(R=h_R.to(dev, non_blocking=True), T=h_T.to(dev, non_blocking=True),
                       K=h_K.to(dev, non_blocking=True))        # fresh K every step (forward halves it in place)
            o, s, d, _, _, _ = model.render_features(cam, d_feat, d_dens, True, view2vol=view2vol)
            h_out.copy_(o, non_blocking=True)
            h_sil.copy_(s, non_blocking=True)
            h_dep.copy_(d, non_blocking=True)
        for _ in range(3):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_serial_ms = e0.elapsed_time(e1)

        # the public streamed API: same copies, same kernels, three streams, 3 batches in flight
        from forge_b200.pipeline import StreamedRenderer
        h_v2v = view2vol.cpu().pin_memory()
        sr = StreamedRenderer(model, CFG['objects'], N, D, depth=3, device=dev)
        for _ in range(4):
            sr.submit(h_feat, h_dens, h_R, h_T, h_K, h_v2v, h_out, h_sil, h_dep)
        sr.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sr.launches = 0
        e0.record(sr.s_in)
        for _ in range(args.steps):
            sr.submit(h_feat, h_dens, h_R, h_T, h_K, h_v2v, h_out, h_sil, h_dep)
        sr.drain()
        e1.record(sr.s_out)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e_launches = sr.launches

    total_ms, e2e_ms, k1_total = reduce_times(total_ms, e2e_ms, sum(k1_ms), world, dev)
    e2e_serial_ms = reduce_times(e2e_serial_ms, 0.0, 0.0, world, dev)[0]
    if rank != 0:
        return
    rays = rays_per_step()
    value = world * rays * args.steps / (total_ms * 1e-3)
    e2e_value = world * rays * args.steps / (e2e_ms * 1e-3)
    peaks, peak_kind = measured_peaks()
    k1_avg_s = k1_total / args.steps * 1e-3
    achieved = algorithmic_bytes_k1() / k1_avg_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "per_gpu": True, "l2": "flushed between timed steps (256 MB memset)",
                   "step": "pack 4 distinct NCDHW volumes (padded channels-last + density quads) + fused raymarch (feat16 + sil + depth)",
                   "wall_ms_incl_flush": wall * 1e3 / args.steps},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "forge_b200.pipeline.StreamedRenderer: pinned host volumes+cameras -> H2D -> pack + raymarch -> "
                       "D2H into pinned host images, every step; 3 batches in flight on copy-in/compute/copy-out streams",
                "unpipelined_value": world * rays * args.steps / (e2e_serial_ms * 1e-3),
                "unpipelined_api": "VolRender.render_features, one step at a time on one stream",
                "gpu_launches": e2e_launches},
        "gpu_launches": n_launch,
        "roofline": {"bound": "hbm", "kernel": "raymarch_fwd_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                     "algorithmic_bytes": algorithmic_bytes_k1(), "kernel_ms": k1_avg_s * 1e3,
                     "fp32_tflops": rays * CFG['n_pts'] * FLOPS_PER_SAMPLE / k1_avg_s / 1e12,
                     "note": "K1 is bound by the L1 data pipe (register write-back of the gathered corners, ncu 82 %), not by "
                             "HBM (SURVEY 8d: 82 FLOP/B); bytes are the distinct-volume figure; see roofline_rotate for the "
                             "HBM-bound kernel of the path"},
        "roofline_rotate": {"bound": "hbm", "kernel": "rotate_fwd_kernel (K2, %d view-volumes of 128x%d^3, channels-last)"
                            % (CFG['objects'] * CFG['views'], CFG['vol'] // 2),
                            "achieved": k2_bytes / (k2_avg_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": k2_bytes / (k2_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_kind": peak_kind,
                            "algorithmic_bytes": k2_bytes, "kernel_ms": k2_avg_ms,
                            "note": "secondary line: the HBM-bound kernel of the path, rank 0, timed in the same run"},
        "roofline_decoder": {"bound": "tensor", "kernel": "decoder_tc_kernel (conv_rgb as tcgen05 implicit GEMMs, bf16 in / fp32 "
                             "accumulate, %d x 16 x %dx%d -> 3 x %dx%d)" % (N, S, S, 2 * S, 2 * S),
                             "achieved": dec_flops / (dec_avg_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": dec_flops / (dec_avg_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "peak_kind": peak_kind,
                             "algorithmic_flops": dec_flops, "kernel_ms": dec_avg_ms,
                             "note": "secondary line: useful conv FLOPs (N = 16/8/3 output channels); the binding unit is the "
                                     "shared-memory operand path of the MMAs (ncu l1tex__data_pipe_tc_wavefronts 62 %), see DESIGN.md"},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, secs = cpu_reference_rate(1, reps=3)
        out["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="forge", choices=["forge", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the forge arm has no CPU path (use --impl reference for the CPU oracle)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_forge(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
