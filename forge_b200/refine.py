"""Helpers for the test-time pose optimisation loop (reference kubric_eval.py:412-530, demo.py:115-188).

The reference's scripts keep every network weight at ``requires_grad=True`` while they optimise only the four
relative poses, so each of the 2000-5000 iterations per object also computes (and throws away) the weight gradients of
the ConvGRU fusion, the heads and the decoder -- about a quarter of the iteration on a B200
(``tools/refine_bench.py``).  ``prepare_for_pose_refinement`` removes that work without touching the scripts' loop:

    model.eval(); prepare_for_pose_refinement(model)      # once, before the loop

Nothing here changes a forward result; with frozen weights K1 / K2 / the decoder run their pose-only backward paths
(NULL volume-gradient outputs in the C ABI).
"""
import torch


def prepare_for_pose_refinement(model, freeze_weights=True, channels_last=True, fusion_dtype=None, decoder_dtype=None):
    """model: forge_b200 ``FORGE`` / ``FORGE_poseEstimator3D`` (or any module holding ``encoder_3d`` / ``render``).

    freeze_weights  set requires_grad=False on every parameter (poses are the only leaves of the loop)
    channels_last   store the 3-D conv weights channels_last_3d (K2 emits channels-last volumes; saves cuDNN's
                    NCDHW<->NDHWC conversion kernels around every conv)
    fusion_dtype    None = fp32 like the reference, torch.bfloat16 = autocast for fusion + heads
    decoder_dtype   None = fp32 fused decoder, torch.bfloat16 = tensor-core decoder
    Returns the model."""
    if freeze_weights:
        for p in model.parameters():
            p.requires_grad_(False)
    enc = getattr(model, 'encoder_3d', None)
    if enc is not None:
        if channels_last and hasattr(enc, 'channels_last_3d_'):
            enc.channels_last_3d_()
        if hasattr(enc, 'compute_dtype'):
            enc.compute_dtype = fusion_dtype
    ren = getattr(model, 'render', None)
    if ren is not None and hasattr(ren, 'decoder_dtype'):
        ren.decoder_dtype = decoder_dtype
    return model


# ---- the refinement iteration itself (reference kubric_eval.py:450-504, demo.py:135-183) ------------------------------------
def quat_to_pose(q, t):
    """[n,4] (w,x,y,z), [n,3] -> [n,4,4]; the normalise + toSE3 steps of the loop (reference kubric_eval.py:452-456 via
    utils/geo_utils.quat2mat_transform :121-137)"""
    import torch.nn.functional as F
    q = F.normalize(q, dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z,
                     2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x,
                     2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z], dim=-1).reshape(-1, 3, 3)
    top = torch.cat([R, t.unsqueeze(-1)], dim=-1)
    n = q.shape[0]
    bottom = torch.cat([torch.zeros(n, 1, 3, device=q.device, dtype=q.dtype), torch.ones(n, 1, 1, device=q.device, dtype=q.dtype)], dim=-1)
    return torch.cat([top, bottom], dim=1)


def rigid_inverse(P):
    """inverse of [n,4,4] rigid transforms, [R^T | -R^T t] (no LU, no host-side error check: capturable in a CUDA graph)"""
    Rt = P[:, :3, :3].transpose(1, 2)
    top = torch.cat([Rt, -(Rt @ P[:, :3, 3:4])], dim=-1)
    return torch.cat([top, P[:, 3:4, :]], dim=1)


def make_refine_loss(model, features, K, canonical_pose, target_imgs, target_masks, w_rgb=1.0, w_mask=1.0, render_depth=False):
    """The loss of one refinement iteration as a function of the relative poses (quat [b(t-1),4], trans [b(t-1),3]):
    poses -> rotate (K2) -> sorted fuse -> heads -> render t views per object (K1 + decoder) -> MSE(rgb) + MSE(mask)
    (reference kubric_eval.py:450-498).  features [b,t,C,D,H,W] detached; K [b,t,3,3]; canonical_pose [4,4]."""
    import torch.nn.functional as F
    from .models.model import sequence_from_distance
    b, t = features.shape[:2]
    dev = features.device
    view2vol = torch.arange(b, device=dev).repeat_interleave(t).int()
    Kflat = K.reshape(b * t, 3, 3).to(dev)
    canon = canonical_pose.to(dev).reshape(1, 4, 4)

    def loss_fn(quat, trans):
        rel = quat_to_pose(quat, trans)                                         # [b(t-1),4,4]
        P = (canon @ rel).reshape(b, t - 1, 4, 4)
        P = torch.cat([canon.reshape(1, 1, 4, 4).expand(b, 1, 4, 4), P], dim=1)  # [b,t,4,4]
        E = rigid_inverse(P.reshape(b * t, 4, 4))
        feat, dens = model.reconstruct(features, P, sequence_from_distance(P[:, :, :3, 3]))
        cam = {'R': E[:, :3, :3], 'T': E[:, :3, 3], 'K': Kflat.clone()}
        out = model.render(cam, feat, dens, return_origin_proj=True, render_depth=render_depth, view2vol=view2vol)
        return w_rgb * F.mse_loss(out[0], target_imgs) + w_mask * F.mse_loss(out[1], target_masks)
    return loss_fn


class GraphedPoseRefiner:
    """One whole refinement iteration -- forward, loss, backward to the poses, Adam step -- captured in a CUDA graph and
    replayed.  The eager loop spends a large part of each ~8 ms iteration in Python and launch overhead
    (~250 kernel launches, most of them tiny); every entry point of libforge_b200 is capture-safe (no allocation,
    synchronisation or stream creation inside), so the iteration replays as ONE graph launch.

        loss_fn = make_refine_loss(model, features, K, canonical_pose, target_imgs, target_masks)
        r = GraphedPoseRefiner(loss_fn, quat0, trans0, lr=1e-3, lr_trans=5e-4)      # the reference's two parameter groups
        for _ in range(2000): r.step()
        quat, trans, loss = r.quat, r.trans, r.loss                                  # device tensors

    Fixed shapes and a constant learning rate (the reference's schedule has lr_end == lr_start, kubric_eval.py:444-449).
    Atomic reductions in the backward kernels make two runs agree to fp32 summation order, not bit for bit -- the same holds
    for two runs of the eager loop."""

    def __init__(self, loss_fn, quat0, trans0, lr=1e-3, lr_trans=None, warmup=3):
        self.quat = quat0.detach().clone().requires_grad_(True)
        self.trans = trans0.detach().clone().requires_grad_(True)
        self.loss_fn = loss_fn
        self.opt = torch.optim.Adam([{'params': [self.quat], 'lr': lr},
                                     {'params': [self.trans], 'lr': lr if lr_trans is None else lr_trans}], lr=lr, capturable=True)
        dev = self.quat.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                   # lazy one-time setup (function attributes, caches, cuDNN plans)
                self._eager_iteration()
        torch.cuda.current_stream(dev).wait_stream(side)
        # the warm-up moved the parameters: restore the initial state before capturing
        with torch.no_grad():
            self.quat.copy_(quat0)
            self.trans.copy_(trans0)
            for st in self.opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self.loss_fn(self.quat, self.trans)
            self.loss.backward()
            self.opt.step()
        self.iterations = 0

    def _eager_iteration(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.quat, self.trans)
        loss.backward()
        self.opt.step()
        return loss

    def step(self, n=1):
        for _ in range(n):
            self.graph.replay()
        self.iterations += n
        return self.loss
