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
between steps, step = camera prep + pack + raymarch ("cameras + volumes -> images"); `e2e` =
pinned-host inputs -> H2D -> public API -> D2H, all inside the timed region; `roofline` = the raymarch kernel
alone against the measured HBM peak; `cpu_baseline` = the oracle (restated reference PyTorch3D path) on this
box's host cores.  N = 1 adds the rest of SURVEY 8d's table: `gpu_reference` (the oracle's op sequence on the same
GPU with as-called per-view volumes = the "reference PyTorch3D GPU renderer"), `cpu_baseline_cfg1`
(BASELINE.json configs[0] exactly), `cfg4` (configs[3]: K1 / K2 with their roofline fractions) and `e2e_model`
(pinned host images + cameras -> FORGE.forward -> pinned host RGB / masks, the reference's real host boundary,
models/model.py:42-148).
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
        "config": {"workload": workload_name()},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def reduce_max(vals, world, device):
    """Job-level times = MAX over ranks (every rank processes the same amount of work)."""
    t = torch.tensor(list(vals), dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def timed_launches(fn, flush, dev, reps=10, warm=3):
    """average CUDA-event ms of fn() over reps launches, L2 flushed (untimed) before each"""
    ms = []
    for it in range(warm + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize(dev)
        if it >= warm:
            ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms)


def executed_samples(cam12, zs, S, D):
    """Samples K1 actually evaluates: under zeros padding a sample contributes only if every local coordinate lies inside
    (-1 - 2/(D-1), 1 + 2/(D-1)); everything else is skipped exactly.  Counted on the device, outside any timing."""
    lim = 1.0 + 2.0 / (D - 1)
    px = torch.arange(S, device=cam12.device, dtype=torch.float32) + 0.5
    n = 0
    for c in cam12:
        o, M = c[:3], c[3:].view(3, 3)
        uv1 = torch.stack([px.view(1, S).expand(S, S), px.view(S, 1).expand(S, S), torch.ones(S, S, device=c.device)], -1)
        d = uv1 @ M.t()                                               # [S,S,3]
        p = o.view(1, 1, 1, 3) + zs.view(1, 1, -1, 1) * d.unsqueeze(2)   # [S,S,P,3]
        n += int((p.abs() < lim).all(dim=-1).sum().item())
    return n


def k1_launch(lib_call, fp, dq, view2vol, cam12, zs, o_feat, o_sil, o_dep, N, V, D, S, P, dev):
    lib_call("forge_raymarch_fwd", fp.data_ptr(), dq.data_ptr(), view2vol.data_ptr(), cam12.data_ptr(), zs.data_ptr(),
             o_feat.data_ptr(), o_sil.data_ptr(), o_dep.data_ptr(), N, V, D, D, D, S, S, P,
             torch.cuda.current_stream(dev).cuda_stream)


def k2_setup(cfg, objects, views, C, n, seed, dev):
    from forge_b200 import ops, synthetic as syn
    from forge_b200.models.rotate import Rotate_world
    rot = Rotate_world(cfg).to(dev)
    vox, poses = syn.rotate_inputs(objects, views, C, n, seed=seed, device=dev)
    vcl = vox.permute(0, 1, 3, 4, 5, 2).contiguous().reshape(objects * views, n, n, n, C)
    del vox
    gx, gy, gz, gmax = rot._device_axes(n, n, n, dev)
    A = ops.pose_affine(poses).view(objects, views, 12)[:, 1:].reshape(-1, 12).contiguous()      # one affine per job
    # transform jobs only: the pipeline aliases view 0 of the input instead of copying it (Rotate_world.forward_views)
    jobs = rot._jobs_aliased(objects, views, dev, None)
    out = torch.empty(objects * (views - 1), n, n, n, C, device=dev)
    return vcl, A, jobs, gx, gy, gz, gmax, out


def k2_launch(lib_call, vcl, A, jobs, gx, gy, gz, gmax, out_cl, M, C, n, dev):
    lib_call("forge_rotate_fwd", vcl.data_ptr(), A.data_ptr(), jobs.data_ptr(), gx.data_ptr(), gy.data_ptr(), gz.data_ptr(),
             float(gmax), out_cl.data_ptr(), M, C, n, n, n, torch.cuda.current_stream(dev).cuda_stream)


def leg(out, key, fn):
    """secondary legs never take the headline line down with them"""
    try:
        out[key] = fn()
    except Exception as e:        # noqa: BLE001
        out[key] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


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
    S, D, P = CFG['img_size'] // 2, CFG['vol'], CFG['n_pts']
    V, N = CFG['objects'], CFG['objects'] * CFG['views']
    Kh = inp['K'].clone()
    Kh /= 2.0
    Kh[:, 2, 2] = 1.0
    R_d, T_d = inp['R'].contiguous(), inp['T'].contiguous()
    zs = model._depths(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # 256 MB > 126 MB L2
    peaks, peak_kind = measured_peaks()

    launches = [0]
    o_feat = torch.empty(N, S, S, 16, device=dev)
    o_sil = torch.empty(N, S, S, device=dev)
    o_dep = torch.empty(N, S, S, device=dev)
    cam12 = torch.empty(N, 12, device=dev)
    vs = CFG['volume_size'] / D
    sc = (D - 1) * 0.5 * vs

    def run_steps(n):
        """n steps, each: [untimed L2 flush] a | camera prep (launch 1) + pack (launch 2) | b | raymarch (launch 3) | c"""
        evs = []
        keep = None
        st = torch.cuda.current_stream(dev).cuda_stream
        for _ in range(n):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            flush.zero_()
            a.record()
            _lib.call("forge_camera_prep_fwd", R_d.data_ptr(), T_d.data_ptr(), Kh.data_ptr(), N, sc, sc, sc, 1e-6,
                      cam12.data_ptr(), None, st)
            fp, dq = ops.pack_volume(feat, dens)
            b.record()
            k1_launch(_lib.call, fp, dq, view2vol, cam12, zs, o_feat, o_sil, o_dep, N, V, D, S, P, dev)
            c.record()
            launches[0] += 3
            keep = (fp, dq)
            evs.append((a, b, c))
        return evs, keep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    extra = {}
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
        ev, packed = run_steps(args.steps)
        barrier()
        wall = time.perf_counter() - wall0
        step_ms = [a.elapsed_time(c) for a, b, c in ev]
        k1_ms = [b.elapsed_time(c) for a, b, c in ev]
        total_ms = sum(step_ms)
        n_launch = launches[0]
        if args.verbose and rank == 0:
            sys.stderr.write("step ms: %s\n" % " ".join("%.3f" % x for x in step_ms))
        n_exec = executed_samples(cam12, zs, S, D)

        # ---- secondary: K2 (the HBM-bound kernel of the path) on cfg-2's fusion grid, same run ------------
        Cr, nr, tr = 128, D // 2, CFG['views']
        k2 = k2_setup(cfg, V, tr, Cr, nr, 100 + rank, dev)
        k2_avg_ms = timed_launches(lambda: k2_launch(_lib.call, *k2, V * (tr - 1), Cr, nr, dev), flush, dev)
        k2_bytes = 2 * V * (tr - 1) * Cr * nr ** 3 * 4                      # read + write per TRANSFORM
        k2_bytes_moved = k2_bytes                                           # view 0 is aliased, not copied
        del k2

        # ---- secondary: the tensor-core (bf16, tcgen05) decoder on the raymarcher's output, same run --------------
        model.decoder_dtype = torch.bfloat16
        wtc = model._decoder_pack(dev, kind='tc')
        rgb = torch.empty(N, 3, 2 * S, 2 * S, device=dev)
        dec_avg_ms = timed_launches(
            lambda: _lib.call("forge_decoder_tc_fwd", o_feat.data_ptr(), wtc.data_ptr(), rgb.data_ptr(), None, N, S, S, 0,
                              torch.cuda.current_stream(dev).cuda_stream), flush, dev)
        dec_flops = 2.0 * 6104 * N * (2 * S) ** 2          # MACs per output pixel: 16*16*9 + 8*16*25 + 3*8*25
        model.decoder_dtype = None
        del rgb

        # ---- secondary: the ConvGRU gate convolution (256 -> 256 ch, 3^3 taps, 32^3 voxels x 4 objects) on tcgen05, same run ----
        gC, gn = 128, D // 2
        gx_ = torch.randn(V, gn, gn, gn, gC, device=dev).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
        gh_ = torch.randn(V, gn, gn, gn, gC, device=dev).permute(0, 4, 1, 2, 3)
        gh16 = gh_.to(torch.bfloat16)
        gw = ops.pack_conv3d_weights(torch.randn(2 * gC, 2 * gC, 3, 3, 3, device=dev) / (27 * 2 * gC) ** 0.5)
        gb = torch.zeros(2 * gC, device=dev)
        gru_avg_ms = timed_launches(lambda: ops.conv3d_tc(gx_, gw, 'gate', gb, h2=gh16, h_state=gh_), flush, dev)
        gru_flops = 2.0 * V * gn ** 3 * (2 * gC) * 27 * (2 * gC)
        del gx_, gh_, gh16, gw
        clocks = sampler.stop() if rank == 0 else None

        # ---- end to end: pinned host inputs -> H2D -> public API -> D2H, everything timed ----------
        e2e_steps = max(args.steps, 200)                   # a 20-step region is 28 ms: too short to trust to 10 %
        h_feat, h_dens = feat.cpu().pin_memory(), dens.cpu().pin_memory()
        h_R, h_T, h_K = inp['R'].cpu().pin_memory(), inp['T'].cpu().pin_memory(), inp['K'].cpu().pin_memory()
        h_out = torch.empty(N, S, S, 16).pin_memory()
        h_sil = torch.empty(N, S, S).pin_memory()
        h_dep = torch.empty(N, S, S).pin_memory()
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
        for _ in range(e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_serial_ms = e0.elapsed_time(e1)

        # the public streamed API: same copies, same kernels, three streams, 3 batches in flight
        from forge_b200.pipeline import StreamedRenderer
        h_v2v = view2vol.cpu().pin_memory()
        sr = StreamedRenderer(model, V, N, D, depth=3, device=dev)
        # untimed: full-duplex copies (own download stream) or half-duplex (downloads queued behind the next upload) -- which one is
        # faster depends on the box's host link (one GPU: duplex; 8 GPUs behind one NUMA node: half duplex), so it is measured
        e2e_modes = sr.calibrate(h_feat, h_dens, h_R, h_T, h_K, h_v2v, h_out, h_sil, h_dep, steps=24, sync=barrier)
        cal = reduce_max([e2e_modes["duplex"], e2e_modes["half_duplex"]], world, dev)     # every rank takes the same, box-wide choice
        e2e_modes = {"duplex": cal[0], "half_duplex": cal[1]}
        sr.duplex = cal[0] <= cal[1]
        e2e_mode = "duplex" if sr.duplex else "half_duplex"
        for _ in range(4):
            sr.submit(h_feat, h_dens, h_R, h_T, h_K, h_v2v, h_out, h_sil, h_dep)
        sr.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sr.launches = 0
        e0.record(sr.s_in)
        for _ in range(e2e_steps):
            sr.submit(h_feat, h_dens, h_R, h_T, h_K, h_v2v, h_out, h_sil, h_dep)
        sr.drain()
        e1.record(sr.s_out)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e_launches = sr.launches
        del sr, h_feat, h_dens

        # ---- N = 1 only: the rest of SURVEY 8d's table (rank 0 = the only rank) --------------------------------------
        if world == 1 and not args.quick:
            leg(extra, "gpu_reference", lambda: gpu_reference_leg(inp, dev, flush))
            leg(extra, "cfg4", lambda: cfg4_leg(_lib, ops, dev, flush, peaks, peak_kind))
            leg(extra, "e2e_model", lambda: e2e_model_leg(dev, args))

    total_ms, e2e_ms, k1_total, e2e_serial_ms = reduce_max([total_ms, e2e_ms, sum(k1_ms), e2e_serial_ms], world, dev)
    if rank != 0:
        return
    rays = rays_per_step()
    value = world * rays * args.steps / (total_ms * 1e-3)
    e2e_value = world * rays * e2e_steps / (e2e_ms * 1e-3)
    k1_avg_s = k1_total / args.steps * 1e-3
    achieved = algorithmic_bytes_k1() / k1_avg_s / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    if "gpu_reference" in extra and "value" in extra["gpu_reference"]:
        extra["gpu_reference"]["k1_speedup"] = rays / k1_avg_s / extra["gpu_reference"]["value"]
        extra["gpu_reference"]["step_speedup"] = value / extra["gpu_reference"]["value"]
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "per_gpu": True, "l2": "flushed between timed steps (256 MB memset)",
                   "step": "camera prep (R,T,K -> 12 floats/view) + pack 4 distinct NCDHW volumes (padded channels-last + density "
                           "quads) + fused raymarch (feat16 + sil + depth)",
                   "wall_ms_incl_flush": wall * 1e3 / args.steps},
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "api": "forge_b200.pipeline.StreamedRenderer: pinned host volumes+cameras -> H2D -> camera prep + pack + raymarch "
                       "-> D2H into pinned host images, every step; 3 batches in flight, kernels overlapped with the copies",
                "copy_mode": e2e_mode, "copy_mode_calibration_ms_per_step": e2e_modes,
                "copy_mode_note": "chosen by StreamedRenderer.calibrate() before the timed region (all ranks measure together, max over ranks): duplex = "
                                  "downloads on their own stream, half_duplex = downloads queued behind the next batch's upload",
                "unpipelined_value": world * rays * e2e_steps / (e2e_serial_ms * 1e-3),
                "unpipelined_api": "VolRender.render_features, one step at a time on one stream",
                "gpu_launches": e2e_launches},
        "gpu_launches": n_launch,
        "roofline": {"bound": "hbm", "kernel": "raymarch_fwd_tma_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_kind": peak_kind,
                     "algorithmic_bytes": algorithmic_bytes_k1(), "kernel_ms": k1_avg_s * 1e3,
                     "executed_samples": n_exec, "executed_fraction": n_exec / float(rays * P),
                     # the unit that actually carries K1's operands: every executed sample reads its 8 corners x 64 B from the
                     # shared-memory bricks; the data pipe of an SM moves 128 B per clock
                     "shared_memory": (lambda sm_bytes, sm_peak: {
                         "algorithmic_bytes": sm_bytes, "achieved": sm_bytes / k1_avg_s / 1e9, "peak": sm_peak, "unit": "GB/s",
                         "frac": sm_bytes / k1_avg_s / 1e9 / sm_peak,
                         "note": "executed samples x 8 corners x 64 B of LDS.128 traffic against SMs x 128 B/clk x the SM clock sampled "
                                 "during the run (nominal 1965 MHz when no sample is available); secondary figure, the contract's "
                                 "roofline above stays the HBM one"})(
                         n_exec * 512, torch.cuda.get_device_properties(dev).multi_processor_count * 128 *
                         ((clocks.get("sm_mhz") or 1965.0) * 1e6) / 1e9),
                     "fp32_tflops": n_exec * FLOPS_PER_SAMPLE / k1_avg_s / 1e12,
                     "fp32_tflops_note": "366 FLOP per EXECUTED sample (samples outside the volume are skipped exactly and not counted)",
                     "note": "K1 (TMA-staged bricks in shared memory, conflict-free LDS.128 corner reads) is bound by shared-memory "
                             "latency / instruction issue, not by HBM (SURVEY 8d: 82 FLOP/B); bytes are the distinct-volume figure; "
                             "see roofline_rotate for the HBM-bound kernel of the path and roofline_fusion for the tensor-core one"},
        "roofline_rotate": {"bound": "hbm", "kernel": "rotate_fwd_kernel (K2, %d transforms of 128x%d^3, channels-last; view 0 aliased, not copied)"
                            % (V * (CFG['views'] - 1), CFG['vol'] // 2),
                            "achieved": k2_bytes / (k2_avg_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": k2_bytes / (k2_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_kind": peak_kind,
                            "algorithmic_bytes": k2_bytes, "bytes_moved": k2_bytes_moved,
                            "frac_on_bytes_moved": k2_bytes_moved / (k2_avg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                            "kernel_ms": k2_avg_ms,
                            "note": "secondary line: the HBM-bound kernel of the path, rank 0, timed in the same run; algorithmic bytes = "
                                    "SURVEY 8d's read + write per TRANSFORM (the view-0 passthrough is not counted)"},
        "roofline_decoder": {"bound": "tensor", "kernel": "decoder_tc_kernel (conv_rgb as tcgen05 implicit GEMMs, bf16 in / fp32 "
                             "accumulate, %d x 16 x %dx%d -> 3 x %dx%d)" % (N, S, S, 2 * S, 2 * S),
                             "achieved": dec_flops / (dec_avg_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": dec_flops / (dec_avg_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "peak_kind": peak_kind,
                             "algorithmic_flops": dec_flops, "kernel_ms": dec_avg_ms,
                             "note": "secondary line: useful conv FLOPs (N = 16/8/3 output channels); the binding unit is the "
                                     "shared-memory operand path of the MMAs, see DESIGN.md"},
        "roofline_fusion": {"bound": "tensor", "kernel": "conv3d_tc_kernel<256> (ConvGRU conv_gate as a tcgen05 implicit GEMM: M = %d voxels, "
                            "N = 256, K = 27 x 256; bf16 in / fp32 accumulate in TMEM; sigmoid / h*r epilogue fused)" % (V * gn ** 3),
                            "achieved": gru_flops / (gru_avg_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": gru_flops / (gru_avg_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "peak_kind": peak_kind,
                            "algorithmic_flops": gru_flops, "kernel_ms": gru_avg_ms,
                            "note": "secondary line: the FLOP-dominant kernel of the fusion stage (reference models/fusion.py:29), rank 0, "
                                    "timed in the same run through the Python op (output allocation + tensor-map encode included)"},
        "clocks": clocks,
    }
    out.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, secs = cpu_reference_rate(1, reps=3)
        out["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample}
        if not args.quick:
            leg(out, "cpu_baseline_cfg1", cpu_cfg1_leg)
    print(json.dumps(out))


# ---- N = 1 legs --------------------------------------------------------------------------------------------------------
def gpu_reference_leg(inp, dev, flush):
    """The reference's op sequence (oracle/reference_path.py = PyTorch3D's ATen ops) on the SAME GPU with the as-called
    per-view volumes (reference models/model.py:138-139): the denominator of north_star's ">= 10x the reference GPU renderer".
    A baseline leg like cpu_baseline: the oracle is the thing measured against, never part of a forge number."""
    import warnings
    from oracle import reference_path as rp
    ren = rp.make_renderer(CFG['img_size'], CFG['n_pts'], CFG['min_depth'], CFG['max_depth']).to(dev)
    idx = inp['view2vol'].long()
    feat_all, dens_all = inp['feat'][idx].contiguous(), inp['dens'][idx].contiguous()

    def step():
        cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
        rp.raymarch_core(ren, cam, feat_all, dens_all, CFG['img_size'], CFG['volume_size'], render_depth=True)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ms = timed_launches(step, flush, dev, reps=10, warm=3)
    return {"value": rays_per_step() / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
            "what": "oracle op sequence (inverse, bmm, 2x grid_sample 3-D, cumprod, sum) on cuda, as-called volumes (one per view), "
                    "same cameras / volumes as the forge arm, L2 flushed, mean of 10 after 3 warm-ups"}


def cpu_cfg1_leg():
    """BASELINE.json configs[0] exactly: single object, 5 views of 64x64 rays, 32^3 voxels, 32 samples, reference path on CPU."""
    import warnings
    from oracle import reference_path as rp
    from forge_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    inp = syn.render_inputs(1, 5, 128, 32, seed=0)
    ren = rp.make_renderer(128, 32, CFG['min_depth'], CFG['max_depth'])
    idx = inp['view2vol'].long()
    feat, dens = inp['feat'][idx], inp['dens'][idx]
    best = float('inf')
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for r in range(6):
            cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
            t0 = time.perf_counter()
            rp.raymarch_core(ren, cam, feat, dens, 128, CFG['volume_size'], render_depth=True)
            dt = time.perf_counter() - t0
            if r > 0:
                best = min(best, dt)
    return {"value": 5 * 64 * 64 / best, "unit": "rays/s", "cores": threads, "kind": "port", "ms": best * 1e3,
            "sample": "cfg-1 exactly: 1 object x 5 views 64x64 rays, 32^3 voxels x17ch, 32 samples; best of 5 after 1 warm-up"}


def cfg4_leg(_lib, ops, dev, flush, peaks, peak_kind):
    """BASELINE.json configs[3]: 8 objects x 5 views of 256^2 rays, 128^3 voxels, 128 samples (render), 64^3 fusion grid (rotate)."""
    from forge_b200 import synthetic as syn
    from forge_b200.models.volume_render import VolRender
    b, t, img, D, P = 8, 5, 512, 128, 128
    S, N = img // 2, b * t
    cfg = syn.make_config(img_size=img, n_pts_per_ray=P)
    m = VolRender(cfg).to(dev).eval()
    inp = syn.render_inputs(b, t, img, D, seed=7, device=dev)
    Kh = inp['K'].clone()
    Kh /= 2.0
    Kh[:, 2, 2] = 1.0
    cam12, _ = ops.camera_prep(inp['R'], inp['T'], Kh, (D, D, D), 1.0)
    zs = m._depths(dev)
    fp, dq = ops.pack_volume(inp['feat'], inp['dens'])
    pack_ms = timed_launches(lambda: ops.pack_volume(inp['feat'], inp['dens']), flush, dev, reps=5, warm=2)
    o_feat = torch.empty(N, S, S, 16, device=dev)
    o_sil = torch.empty(N, S, S, device=dev)
    o_dep = torch.empty(N, S, S, device=dev)
    k1_ms = timed_launches(lambda: k1_launch(_lib.call, fp, dq, inp['view2vol'], cam12, zs, o_feat, o_sil, o_dep, N, b, D, S, P, dev),
                           flush, dev, reps=5, warm=2)
    rays = N * S * S
    k1_bytes = b * 17 * D ** 3 * 4 + rays * 18 * 4 + N * 48
    n_exec = executed_samples(cam12, zs, S, D)
    del fp, dq, o_feat, inp
    k2 = k2_setup(cfg, b, t, 128, D // 2, 17, dev)
    k2_ms = timed_launches(lambda: k2_launch(_lib.call, *k2, b * (t - 1), 128, D // 2, dev), flush, dev, reps=5, warm=2)
    k2_bytes = 2 * b * (t - 1) * 128 * (D // 2) ** 3 * 4
    k2_moved = k2_bytes
    del k2
    torch.cuda.empty_cache()
    hb = peaks["hbm_gbs"]
    return {"workload": "cfg-4: 8 objects x 5 views 256x256 rays, 128^3 voxels x17ch, 128 samples, fp32; rotate on 128ch x 64^3",
            "k1_ms": k1_ms, "k1_rays_per_s": rays / (k1_ms * 1e-3), "k1_algorithmic_bytes": k1_bytes,
            "k1_hbm_frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / hb,
            "k1_as_called_bytes": N * 17 * D ** 3 * 4 + rays * 18 * 4,
            "k1_as_called_hbm_frac": (N * 17 * D ** 3 * 4 + rays * 18 * 4) / (k1_ms * 1e-3) / 1e9 / hb,
            "k1_fp32_tflops": n_exec * FLOPS_PER_SAMPLE / (k1_ms * 1e-3) / 1e12, "k1_executed_fraction": n_exec / float(rays * P),
            "pack_ms": pack_ms, "step_rays_per_s": rays / ((k1_ms + pack_ms) * 1e-3),
            "k2_ms": k2_ms, "k2_algorithmic_bytes": k2_bytes, "k2_hbm_frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / hb,
            "k2_frac_on_bytes_moved": k2_moved / (k2_ms * 1e-3) / 1e9 / hb, "peak_kind": peak_kind,
            "note": "a 142.6 MB volume exceeds L2, so each view re-reads its volume (SURVEY 8d): the as-called figure is the one the "
                    "HBM roofline applies to at this size"}


def e2e_model_leg(dev, args):
    """The reference's real host boundary (models/model.py:42-148): pinned host images [4,10,3,256,256] + cameras ->
    FORGE.forward (lift, rotate, fuse, heads, render of 40 views) -> pinned host RGB + masks; H2D and D2H inside the timed region."""
    import warnings
    from forge_b200 import synthetic as syn
    from forge_b200.models.model import FORGE
    b, t_all = CFG['objects'], 10
    cfg = syn.make_config(img_size=CFG['img_size'], n_pts_per_ray=CFG['n_pts'], use_gt_pose=True)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FORGE(cfg).to(dev).eval()
    model.encoder_3d.density_head[6].bias.data.fill_(0.15)     # random-init heads would give an empty density volume
    sample = syn.kubric_batch(b, n_views_all=t_all, img_size=CFG['img_size'], seed=1)
    sample = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in sample.items()}
    h_rgb = torch.empty(b * t_all, 3, CFG['img_size'], CFG['img_size']).pin_memory()
    h_mask = torch.empty(b * t_all, 1, CFG['img_size'], CFG['img_size']).pin_memory()
    used = ('images', 'K_cv2', 'cam_extrinsics_cv2_canonicalized', 'cam_poses_cv2_canonicalized')
    h2d = sum(sample[k].numel() * sample[k].element_size() for k in used)
    d2h = (h_rgb.numel() + h_mask.numel()) * 4

    def step():
        rgb, mask = model(sample, None, dev)          # forward() itself does the .to(device) copies (reference :50, :91-95)
        h_rgb.copy_(rgb, non_blocking=True)
        h_mask.copy_(mask, non_blocking=True)
    res = {}
    steps = 20

    def streamed(tag):
        """the same forward replayed as one CUDA graph with the next batch's uploads and the previous batch's downloads on their own
        streams (forge_b200.pipeline.StreamedForge), in the model's current configuration"""
        from forge_b200.pipeline import StreamedForge
        outs = [(torch.empty_like(h_rgb).pin_memory(), torch.empty_like(h_mask).pin_memory()) for _ in range(2)]
        with torch.no_grad():
            sf = StreamedForge(model, sample, None, depth=2, device=dev)
            for i in range(4):
                sf.submit(sample, *outs[i % 2])
            sf.drain()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                sf.submit(sample, *outs[i % 2])
            sf.drain()
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        res[tag] = {"ms_per_step": ms, "rays_per_s": b * t_all * (CFG['img_size'] // 2) ** 2 / (ms * 1e-3),
                    "views_per_s": b * t_all / (ms * 1e-3), "h2d_bytes_per_step": sf.h2d_bytes, "d2h_bytes_per_step": sf.d2h_bytes,
                    "api": "forge_b200.pipeline.StreamedForge.submit(sample, out_rgb, out_mask): pinned host batch -> H2D (5 input "
                           "views + cameras) -> FORGE.forward as one CUDA graph -> D2H, 2 batches in flight"}
        del sf

    for name, prep in (("fp32", lambda: None),
                       ("bf16_lift_fusion_heads_tc_decoder", lambda: (model.encoder_3d.channels_last_3d_(),
                                                                      setattr(model.encoder_3d, 'compute_dtype', torch.bfloat16),
                                                                      setattr(model.render, 'decoder_dtype', torch.bfloat16)))):
        prep()
        with torch.no_grad():
            for _ in range(3):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        res[name] = {"ms_per_step": ms, "rays_per_s": b * t_all * (CFG['img_size'] // 2) ** 2 / (ms * 1e-3),
                     "views_per_s": b * t_all / (ms * 1e-3)}
        tag = "fp32_graphed_streamed" if name == "fp32" else "bf16_graphed_streamed"
        try:
            streamed(tag)
        except Exception as e:        # noqa: BLE001 -- a secondary number never takes the leg down
            res[tag] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            torch.cuda.synchronize(dev)
    res.update({"api": "forge_b200.models.model.FORGE.forward(sample, dataset, device), eval mode, ground-truth poses, random-init "
                       "weights, %d objects x 5 input views -> %d rendered views at 256^2" % (b, b * t_all),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps})
    del model
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="forge", choices=["forge", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the N = 1 secondary legs (gpu_reference, cfg4, e2e_model, cfg-1 CPU)")
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
