"""StreamedRenderer -- host-to-host rendering with copies overlapped against the kernels.

The reference's scripts hand the renderer tensors that already live on the GPU, but a serving
deployment (and the `e2e` leg of bench.py) starts and ends in host memory.  One object batch needs
H2D of its volumes + cameras, pack + raymarch, and D2H of the images; run back to back that is
PCIe-bound (cfg-2: 71 MB in, 24 MB out per step against 0.5 ms of kernels).  This class keeps
`depth` batches in flight on three CUDA streams (copy-in, compute, copy-out) with per-slot device
buffers and events, so the steady-state rate is that of the slowest stage (the H2D copy) instead
of the sum of all three.

    r = StreamedRenderer(volrender_module, n_views, vol, depth=3)
    for batch in batches:                      # pinned host tensors
        done = r.submit(feat, dens, R, T, K, view2vol_host, out_feat, out_sil, out_depth)
    r.drain()                                  # results are in the pinned output tensors

Every step's inputs are copied from pinned host memory and every step's result is copied back;
nothing is cached across steps.
"""
import torch

from . import _lib


class _Slot:
    def __init__(self, dev, V, N, D, S):
        self.feat = torch.empty(V, 16, D, D, D, device=dev)
        self.dens = torch.empty(V, 1, D, D, D, device=dev)
        self.R = torch.empty(N, 3, 3, device=dev)
        self.T = torch.empty(N, 3, device=dev)
        self.K = torch.empty(N, 3, 3, device=dev)
        self.view2vol = torch.empty(N, dtype=torch.int32, device=dev)
        self.cam12 = torch.empty(N, 12, device=dev)
        self.feat_pad = torch.empty(V, D + 2, D + 2, D + 2, 16, device=dev)
        self.dens_quad = torch.empty(V, D + 2, D + 1, D + 1, 4, device=dev)
        self.out = torch.empty(N, S, S, 16, device=dev)
        self.sil = torch.empty(N, S, S, device=dev)
        self.depth = torch.empty(N, S, S, device=dev)
        self.copied_in = torch.cuda.Event()
        self.computed = torch.cuda.Event()
        self.copied_out = torch.cuda.Event()
        self.busy = False


class StreamedRenderer:
    def __init__(self, volrender, n_volumes, n_views, vol, depth=3, device=None, duplex=True):
        """duplex=True: downloads run on their own stream, concurrently with the next batch's upload (one GPU on its own PCIe
        link: both directions at full rate).  duplex=False: one copy stream carries both directions, the download of batch i
        queued behind the upload of batch i + 1 -- for boxes whose host link does not sustain both directions at once (8 GPUs
        behind one NUMA node: 23.7 GB/s per GPU one way, 13.7 GB/s each way when both run).  Kernels overlap the
        copies either way.  ``calibrate`` measures both on the spot and keeps the faster."""
        self.duplex = bool(duplex)
        self.pending = None                      # duplex=False: (slot, host outputs) whose download is still to be queued
        self.m = volrender
        self.dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.V, self.N, self.D = n_volumes, n_views, vol
        self.S = volrender.img_size // 2
        self.zs = volrender._depths(self.dev)
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.slots = [_Slot(self.dev, n_volumes, n_views, vol, self.S) for _ in range(depth)]
        self.next = 0
        self.launches = 0

    def submit(self, feat, dens, R, T, K, view2vol, out_feat, out_sil, out_depth):
        """All arguments are pinned host tensors; K is the full-resolution intrinsics (halved on the
        device copy, like VolRender.forward does to its argument).  Returns the event that fires when
        the outputs have landed in the host tensors (duplex=False: the event is recorded by the NEXT submit() or by drain(),
        which is when the download is queued -- wait on it only after one of those)."""
        s = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        if s.busy:
            s.copied_out.synchronize()          # the slot's previous batch must have left the device
        s.busy = True
        with torch.cuda.stream(self.s_in):
            for dst, src in ((s.feat, feat), (s.dens, dens), (s.R, R), (s.T, T), (s.K, K), (s.view2vol, view2vol)):
                dst.copy_(src, non_blocking=True)
            s.copied_in.record(self.s_in)
        with torch.cuda.stream(self.s_cmp), torch.no_grad():
            self.s_cmp.wait_event(s.copied_in)
            s.K.mul_(0.5)                       # the device copy is halved, like VolRender.forward does to its argument
            st = self.s_cmp.cuda_stream
            vs = self.m.volume_physical_size / self.D
            sc = (self.D - 1) * 0.5 * vs
            _lib.call("forge_camera_prep_fwd", s.R.data_ptr(), s.T.data_ptr(), s.K.data_ptr(), self.N, sc, sc, sc, 1e-6,
                      s.cam12.data_ptr(), None, st)
            _lib.call("forge_pack_volume", s.feat.data_ptr(), 0, s.dens.data_ptr(), s.feat_pad.data_ptr(),
                      s.dens_quad.data_ptr(), self.V, self.D, self.D, self.D, st)
            _lib.call("forge_raymarch_fwd", s.feat_pad.data_ptr(), s.dens_quad.data_ptr(), s.view2vol.data_ptr(),
                      s.cam12.data_ptr(), self.zs.data_ptr(), s.out.data_ptr(), s.sil.data_ptr(), s.depth.data_ptr(),
                      self.N, self.V, self.D, self.D, self.D, self.S, self.S, self.zs.numel(), st)
            self.launches += 3
            s.computed.record(self.s_cmp)
        if self.duplex:
            self._download(s, (out_feat, out_sil, out_depth), self.s_out)
        else:
            # half duplex: this batch's download goes on the upload stream AFTER the next batch's upload (or at drain), so the
            # copies never run in both directions at once and the upload of batch i + 1 still overlaps the kernels of batch i
            prev, self.pending = self.pending, (s, (out_feat, out_sil, out_depth))
            if prev is not None:
                self._download(prev[0], prev[1], self.s_in)
        return s.copied_out

    def _download(self, s, outs, stream):
        with torch.cuda.stream(stream):
            stream.wait_event(s.computed)
            outs[0].copy_(s.out, non_blocking=True)
            outs[1].copy_(s.sil, non_blocking=True)
            outs[2].copy_(s.depth, non_blocking=True)
            s.copied_out.record(stream)

    def calibrate(self, feat, dens, R, T, K, view2vol, out_feat, out_sil, out_depth, steps=24, sync=None):
        """Run ``steps`` batches in each copy mode and keep the faster one (host links differ: see __init__).  -> dict of ms.
        ``sync`` (optional callable, e.g. a torch.distributed barrier) is called before each measurement so that all the ranks of
        a box measure under the contention they will run under; a multi-rank caller may then override ``self.duplex`` with a
        choice made on the reduced timings."""
        ms = {}
        for mode in (True, False):
            self.drain()
            self.duplex = mode
            for _ in range(3):
                self.submit(feat, dens, R, T, K, view2vol, out_feat, out_sil, out_depth)
            self.drain()
            if sync is not None:
                sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(self.dev)
            e0.record(self.s_in)
            for _ in range(steps):
                self.submit(feat, dens, R, T, K, view2vol, out_feat, out_sil, out_depth)
            self.drain()
            e1.record(self.s_in)
            e1.synchronize()
            ms["duplex" if mode else "half_duplex"] = e0.elapsed_time(e1) / steps
        self.duplex = ms["duplex"] <= ms["half_duplex"]
        return ms

    def drain(self):
        if self.pending is not None:
            self._download(self.pending[0], self.pending[1], self.s_in)
            self.pending = None
        for s in self.slots:
            if s.busy:
                s.copied_out.synchronize()
                s.busy = False


class GraphedVolRender:
    """``VolRender.forward`` for inference captured once in a CUDA graph and replayed.

    At small sizes (BASELINE.json configs[0]: 5 views of 32^2 rays) the kernels of a render take ~0.1 ms while the
    module call costs ~0.3 ms of Python and launch overhead; every entry point of libforge_b200 is capture-safe (no
    allocation, synchronisation or stream creation inside), so the whole call -- camera prep, pack, raymarch, decoder,
    upsample -- replays as one graph launch.  Shapes are fixed at construction; inputs are copied into the graph's
    static buffers, outputs are the graph's static tensors (clone them to keep a result across calls).

        g = GraphedVolRender(volrender.eval(), n_views, n_volumes, vol, render_depth=True, return_origin_proj=True)
        rgb, sil, depth, origin_proj = g(camera_params, feature_3d, density_3d, view2vol)
    """

    def __init__(self, volrender, n_views, n_volumes, vol, render_depth=False, return_origin_proj=False, device=None):
        if volrender.training:
            raise RuntimeError("GraphedVolRender captures the inference path: call .eval() first")
        self.m = volrender
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.dev, self.N, self.V = dev, n_views, n_volumes
        self.render_depth, self.return_origin_proj = render_depth, return_origin_proj
        self.R = torch.eye(3, device=dev).repeat(n_views, 1, 1)
        self.T = torch.zeros(n_views, 3, device=dev)
        self.T[:, 2] = 1.5
        self.K = torch.eye(3, device=dev).repeat(n_views, 1, 1)
        self.K[:, 0, 0] = self.K[:, 1, 1] = float(volrender.img_size)
        self.K[:, 0, 2] = self.K[:, 1, 2] = volrender.img_size / 2.0
        self.K_in = self.K.clone()
        self.feat = torch.zeros(n_volumes, 16, vol, vol, vol, device=dev)
        self.dens = torch.zeros(n_volumes, 1, vol, vol, vol, device=dev)
        self.view2vol = torch.zeros(n_views, dtype=torch.int32, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                     # warm-up: lazy one-time setup (function attributes, caches) happens here
                self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        self.K.copy_(self.K_in)                    # forward halves K in place; the graph owns its working copy
        cam = {'R': self.R, 'T': self.T, 'K': self.K}
        return self.m(cam, self.feat, self.dens, render_depth=self.render_depth,
                      return_origin_proj=self.return_origin_proj, view2vol=self.view2vol)

    def __call__(self, camera_params, feature_3d, density_3d, view2vol=None):
        """Same arguments as VolRender.forward (K is halved on the caller's tensor too, like the module does)."""
        self.R.copy_(camera_params['R'], non_blocking=True)
        self.T.copy_(camera_params['T'], non_blocking=True)
        self.K_in.copy_(camera_params['K'], non_blocking=True)
        camera_params['K'] /= 2.0
        camera_params['K'][:, -1, -1] = 1.0
        self.feat.copy_(feature_3d, non_blocking=True)
        self.dens.copy_(density_3d.reshape(self.dens.shape), non_blocking=True)
        if view2vol is None:
            if feature_3d.shape[0] != self.N:
                raise ValueError("pass view2vol when the volumes are not one per view")
            view2vol = torch.arange(self.N, dtype=torch.int32, device=self.dev)
        self.view2vol.copy_(view2vol, non_blocking=True)
        self.graph.replay()
        return self.out


class StreamedForge:
    """Host-to-host ``FORGE.forward`` for inference (eval mode, ground-truth poses): the whole forward -- lift, rotate, fuse,
    heads, render of every view -- captured once in a CUDA graph, with the uploads of the next batch and the downloads of the
    previous one on their own streams (reference boundary: models/model.py:42-148, a pinned host batch in, images out).

        f = StreamedForge(model.eval(), sample0, dataset)        # sample0: a pinned host batch of the shapes to serve
        for sample, (rgb_host, mask_host) in work:               # pinned host tensors
            ev = f.submit(sample, rgb_host, mask_host)           # returns at once; ev fires when the images are in host memory
        f.drain()

    The eager call spends ~25 % of a 14 ms step (4 objects, 40 rendered views) in Python / launch gaps and in copies that
    run back to back with the kernels; a replayed graph between double-buffered staging tensors leaves the GPU time of the
    forward itself.  Every step uploads its own inputs and downloads its own outputs; nothing is cached across steps."""

    KEYS = ('K_cv2', 'cam_extrinsics_cv2{c}', 'cam_poses_cv2{c}', 'cam_extrinsics_cv2_canonicalized')

    def __init__(self, model, sample, dataset=None, depth=2, device=None):
        if model.training or not model.config.train.use_gt_pose:
            raise RuntimeError("StreamedForge captures the inference path with ground-truth poses: model.eval(), use_gt_pose=True")
        self.m = model
        self.dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        c = '_canonicalized' if model.config.train.canonicalize else ''
        self.keys = tuple(dict.fromkeys(k.format(c=c) for k in self.KEYS))
        self.n_in = 5                                            # FORGE.forward reads images[:, :5] (reference :50)
        dev = self.dev
        self.static = {k: torch.zeros(sample[k].shape, dtype=sample[k].dtype, device=dev) for k in self.keys}
        self.static['images'] = torch.zeros(sample['images'].shape, dtype=sample['images'].dtype, device=dev)
        self.stage_in = [{k: torch.empty_like(v if k != 'images' else v[:, :self.n_in]) for k, v in self.static.items()}
                         for _ in range(depth)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        for k in self.static:                                    # a valid batch for the warm-up / capture passes
            self._upload(self.static[k] if k != 'images' else self.static[k][:, :self.n_in], sample[k], k)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                                   # lazy one-time setup (function attributes, cuDNN plans, weight packs)
                model(self.static, dataset, dev)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.rgb, self.mask = model(self.static, dataset, dev)
        self.stage_out = [(torch.empty_like(self.rgb), torch.empty_like(self.mask)) for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.busy = [False] * depth
        self.next = 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.stage_in[0].values())
        self.d2h_bytes = (self.rgb.numel() + self.mask.numel()) * 4

    def _upload(self, dst, src, key):
        if key == 'images':                                      # the first n_in views of every object: contiguous pinned runs
            for i in range(dst.shape[0]):
                dst[i].copy_(src[i, :self.n_in], non_blocking=True)
        else:
            dst.copy_(src, non_blocking=True)

    def submit(self, sample, out_rgb, out_mask):
        """sample: pinned host batch (same shapes as at construction); out_rgb / out_mask: pinned host tensors."""
        i = self.next
        self.next = (i + 1) % len(self.busy)
        if self.busy[i]:
            self.ev_out[i].synchronize()                         # the slot's previous batch has left the device
        self.busy[i] = True
        with torch.cuda.stream(self.s_in):
            for k, dst in self.stage_in[i].items():
                self._upload(dst, sample[k], k)
            self.ev_in[i].record(self.s_in)
        with torch.cuda.stream(self.s_cmp), torch.no_grad():
            self.s_cmp.wait_event(self.ev_in[i])
            for k, src in self.stage_in[i].items():              # device-to-device into the graph's input tensors
                (self.static[k] if k != 'images' else self.static[k][:, :self.n_in]).copy_(src, non_blocking=True)
            self.graph.replay()
            self.stage_out[i][0].copy_(self.rgb, non_blocking=True)
            self.stage_out[i][1].copy_(self.mask, non_blocking=True)
            self.ev_cmp[i].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[i])
            out_rgb.copy_(self.stage_out[i][0], non_blocking=True)
            out_mask.copy_(self.stage_out[i][1], non_blocking=True)
            self.ev_out[i].record(self.s_out)
        return self.ev_out[i]

    def drain(self):
        for i, b in enumerate(self.busy):
            if b:
                self.ev_out[i].synchronize()
                self.busy[i] = False
