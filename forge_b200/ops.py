"""torch-facing wrappers of the C ABI: layout helpers and the two autograd Functions.

torch is used for device memory, streams and autograd bookkeeping only; all arithmetic on
volumes and images happens in libforge_b200.so.  CUDA tensors are mandatory.
"""
import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("forge_b200 runs on CUDA (sm_100a) only; got a %s tensor. "
                               "There is no CPU fallback." % t.device)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ---- layout -----------------------------------------------------------------------------------
class _ToChannelsLast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        n, C = x.shape[0], x.shape[1]
        S = x[0, 0].numel()
        out = torch.empty([n] + list(x.shape[2:]) + [C], dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("forge_ncs_to_nsc", _ptr(x), _ptr(out), n, C, S, _stream(x))
        return out

    @staticmethod
    def backward(ctx, g):
        return from_channels_last(g)


def to_channels_last(x):
    """[n, C, *spatial] (any strides) -> contiguous [n, *spatial, C]; zero-copy when x already is
    channels-last in memory, otherwise one pass of the re-layout kernel (differentiable)."""
    _require_cuda(x)
    perm = (0,) + tuple(range(2, x.dim())) + (1,)
    xp = x.permute(perm)
    if xp.is_contiguous() and x.dtype == torch.float32:
        return xp
    if x.shape[1] == 1:
        return _f32c(x).reshape([x.shape[0]] + list(x.shape[2:]) + [1])
    return _ToChannelsLast.apply(_f32c(x))


def from_channels_last(x_cl):
    """contiguous [n, *spatial, C] -> contiguous [n, C, *spatial] with the re-layout kernel."""
    _require_cuda(x_cl)
    x_cl = _f32c(x_cl)
    n, C = x_cl.shape[0], x_cl.shape[-1]
    S = x_cl[0, ..., 0].numel()
    out = torch.empty([n, C] + list(x_cl.shape[1:-1]), dtype=torch.float32, device=x_cl.device)
    with torch.cuda.device(x_cl.device):
        _lib.call("forge_nsc_to_ncs", _ptr(x_cl), _ptr(out), n, C, S, _stream(x_cl))
    return out


def sample_points(pts, D, H, W, align_corners):
    """Test hook: base voxel indices [M,3] (x,y,z) and 8-bit in-bounds masks [M] of the device samplers."""
    _require_cuda(pts)
    pts = _f32c(pts.reshape(-1, 3))
    M = pts.shape[0]
    base = torch.empty(M, 3, dtype=torch.int32, device=pts.device)
    mask = torch.empty(M, dtype=torch.uint8, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.call("forge_sample_points", _ptr(pts), M, D, H, W, int(bool(align_corners)), _ptr(base), _ptr(mask),
                  _stream(pts))
    return base, mask


# ---- K1 ---------------------------------------------------------------------------------------
def _feat_layout(feat):
    """feat [V,16,D,H,W] (any strides) -> (tensor to hand to the library, channels_last flag)."""
    if feat.dtype == torch.float32 and feat.permute(0, 2, 3, 4, 1).is_contiguous():
        return feat, 1                      # channels-last in memory already (e.g. channels_last_3d convs)
    return _f32c(feat), 0


def pack_volume(feat, dens):
    """feat [V,16,D,H,W], dens [V,1,D,H,W] or [V,D,H,W] -> (feat_pad [V,D+2,H+2,W+2,16], dens_quad [V,D+2,H+1,W+1,4]),
    the zero-bordered channels-last / density-quad layouts the raymarcher reads (one pass over the data)."""
    _require_cuda(feat, dens)
    V, C, D, H, W = feat.shape
    if C != 16:
        raise ValueError("the render feature volume must have 16 channels (got %d)" % C)
    if dens.numel() != V * D * H * W:
        raise ValueError("density volume %s does not match feature volume %s" % (tuple(dens.shape), tuple(feat.shape)))
    f, cl = _feat_layout(feat)
    d = _f32c(dens)
    feat_pad = torch.empty(V, D + 2, H + 2, W + 2, 16, dtype=torch.float32, device=feat.device)
    dens_quad = torch.empty(V, D + 2, H + 1, W + 1, 4, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.call("forge_pack_volume", _ptr(f), cl, _ptr(d), _ptr(feat_pad), _ptr(dens_quad), V, D, H, W, _stream(feat))
    return feat_pad, dens_quad


class _Raymarch(torch.autograd.Function):
    """(feat [V,16,D,H,W], dens [V,1,D,H,W], cam12 [N,12]) -> feature image, silhouette, depth."""

    @staticmethod
    def forward(ctx, feat, dens, cam12, view2vol, zs, S_h, S_w, render_depth):
        N = cam12.shape[0]
        V, C, D, H, W = feat.shape
        dev = feat.device
        feat_pad, dens_quad = pack_volume(feat, dens)
        out = torch.empty(N, S_h, S_w, C, dtype=torch.float32, device=dev)
        sil = torch.empty(N, S_h, S_w, dtype=torch.float32, device=dev)
        depth = torch.empty(N, S_h, S_w, dtype=torch.float32, device=dev) if render_depth else None
        with torch.cuda.device(dev):
            _lib.call("forge_raymarch_fwd", _ptr(feat_pad), _ptr(dens_quad), _ptr(view2vol), _ptr(cam12), _ptr(zs),
                      _ptr(out), _ptr(sil), _ptr(depth), N, V, D, H, W, S_h, S_w, zs.numel(), _stream(feat))
        if any(ctx.needs_input_grad[:3]):
            ctx.save_for_backward(feat_pad, dens_quad, cam12, view2vol, zs)
        ctx.dims = (N, V, D, H, W, S_h, S_w)
        ctx.render_depth = render_depth
        ctx.feat_cl = _feat_layout(feat)[1]
        ctx.feat_shape, ctx.dens_shape = feat.shape, dens.shape
        if render_depth:
            return out, sil, depth
        dummy = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(dummy)
        return out, sil, dummy

    @staticmethod
    def backward(ctx, g_out, g_sil, g_depth):
        feat_pad, dens_quad, cam12, view2vol, zs = ctx.saved_tensors
        N, V, D, H, W, S_h, S_w = ctx.dims
        dev = feat_pad.device
        need_f, need_d, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g_out = torch.zeros(N, S_h, S_w, 16, device=dev) if g_out is None else _f32c(g_out)
        g_sil = torch.zeros(N, S_h, S_w, device=dev) if g_sil is None else _f32c(g_sil)
        g_depth = _f32c(g_depth) if (ctx.render_depth and g_depth is not None) else None
        gfp = torch.zeros_like(feat_pad) if need_f else None
        gdp = torch.zeros(V, D + 2, H + 2, W + 2, dtype=torch.float32, device=dev) if need_d else None
        gc = torch.zeros_like(cam12) if need_c else None
        gf = gd = None
        if need_f or need_d or need_c:
            P = zs.numel()
            ws = torch.empty(_lib.load().forge_raymarch_bwd_workspace(N, V, D, H, W, S_h, S_w, P) // 4, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.call("forge_raymarch_bwd", _ptr(feat_pad), _ptr(dens_quad), _ptr(view2vol), _ptr(cam12), _ptr(zs),
                          _ptr(g_out), _ptr(g_sil), _ptr(g_depth), _ptr(gfp), _ptr(gdp), _ptr(gc), _ptr(ws),
                          N, V, D, H, W, S_h, S_w, P, _stream(feat_pad))
                if need_f:
                    if ctx.feat_cl:     # interior of the padded gradient, viewed in the caller's layout (no copy)
                        gf = gfp[:, 1:-1, 1:-1, 1:-1, :].permute(0, 4, 1, 2, 3)
                    else:
                        gf = torch.empty(ctx.feat_shape, dtype=torch.float32, device=dev)
                        _lib.call("forge_unpack_volume_grad", _ptr(gfp), _ptr(gf), 0, V, D, H, W, _stream(feat_pad))
            if need_d:
                gd = gdp[:, 1:-1, 1:-1, 1:-1].reshape(ctx.dens_shape)
        return gf, gd, gc, None, None, None, None, None


def raymarch(feat, dens, cam12, view2vol, zs, S_h, S_w, render_depth=False):
    """feat [V,16,D,H,W] (NCDHW or channels-last strides), dens [V,1,D,H,W], cam12 [N,12], view2vol int32 [N],
    zs [P] -> feat image [N,S_h,S_w,16], silhouette [N,S_h,S_w], depth [N,S_h,S_w] or None."""
    _require_cuda(feat, dens, cam12, view2vol, zs)
    if feat.dim() != 5 or feat.shape[1] != 16:
        raise ValueError("the render feature volume must be [V,16,D,H,W] (got %s)" % (tuple(feat.shape),))
    if dens.numel() != feat.numel() // 16:
        raise ValueError("density volume %s does not match feature volume %s" % (tuple(dens.shape), tuple(feat.shape)))
    if cam12.shape[0] != view2vol.numel():
        raise ValueError("cam12 and view2vol disagree on the number of views")
    if feat.dtype != torch.float32:
        feat = feat.float()
    out, sil, depth = _Raymarch.apply(feat, dens.float(), _f32c(cam12), view2vol.int().contiguous(),
                                      _f32c(zs), int(S_h), int(S_w), bool(render_depth))
    return out, sil, (depth if render_depth else None)


# ---- K2 ---------------------------------------------------------------------------------------
class _Rotate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vox_cl, affine12, jobs, gx, gy, gz, gmax, n_dst):
        n_src, D, H, W, C = vox_cl.shape
        M = jobs.shape[0]
        out = torch.empty(n_dst, D, H, W, C, dtype=torch.float32, device=vox_cl.device)
        with torch.cuda.device(vox_cl.device):
            _lib.call("forge_rotate_fwd", _ptr(vox_cl), _ptr(affine12), _ptr(jobs), _ptr(gx), _ptr(gy), _ptr(gz),
                      float(gmax), _ptr(out), M, C, D, H, W, _stream(vox_cl))
        ctx.save_for_backward(vox_cl, affine12, jobs, gx, gy, gz)
        ctx.gmax = float(gmax)
        return out

    @staticmethod
    def backward(ctx, g_out):
        vox_cl, affine12, jobs, gx, gy, gz = ctx.saved_tensors
        n_src, D, H, W, C = vox_cl.shape
        need_v, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gv = torch.zeros_like(vox_cl) if need_v else None
        ga = torch.zeros_like(affine12) if need_a else None
        if need_v or need_a:
            g_out = _f32c(g_out)
            with torch.cuda.device(vox_cl.device):
                _lib.call("forge_rotate_bwd", _ptr(vox_cl), _ptr(affine12), _ptr(jobs), _ptr(gx), _ptr(gy), _ptr(gz),
                          ctx.gmax, _ptr(g_out), _ptr(gv), _ptr(ga), jobs.shape[0], C, D, H, W, _stream(vox_cl))
        return gv, ga, None, None, None, None, None, None


def rotate_resample(vox_cl, affine12, jobs, gx, gy, gz, gmax, n_dst):
    """vox_cl [n_src,D,H,W,C]; affine12 [M,12]; jobs int32 [M,3] = (src, dst, kind) -> [n_dst,D,H,W,C].
    Every dst slot must be written by exactly one job."""
    _require_cuda(vox_cl, affine12, jobs, gx, gy, gz)
    return _Rotate.apply(_f32c(vox_cl), _f32c(affine12), jobs.int().contiguous(), _f32c(gx), _f32c(gy), _f32c(gz),
                         float(gmax), int(n_dst))


# ---- fused decoder -----------------------------------------------------------------------------
def pack_decoder_weights(conv_rgb):
    """BN-folded weight pack for forge_decoder_fwd from the reference-shaped ``conv_rgb`` Sequential
    (ConvT2d, BN, LReLU, Conv2d, BN, LReLU, Conv2d) in eval mode; layout documented in forge_b200.h."""
    ct, bn1, _, c2, bn2, _, c3 = conv_rgb
    with torch.no_grad():
        s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
        s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)
        wt = ct.weight * s1.view(1, 16, 1, 1)                                   # [ci, co, 6, 6]
        w1 = torch.stack([torch.stack([wt[:, :, py::2, px::2].permute(2, 3, 0, 1) for px in (0, 1)]) for py in (0, 1)])
        w2 = (c2.weight * s2.view(8, 1, 1, 1)).permute(2, 3, 1, 0)               # [ky, kx, ci, co]
        w3 = torch.zeros(5, 5, 8, 4, dtype=torch.float32, device=ct.weight.device)
        w3[..., :3] = c3.weight.permute(2, 3, 1, 0)
        b1 = (ct.bias - bn1.running_mean) * s1 + bn1.bias
        b2 = (c2.bias - bn2.running_mean) * s2 + bn2.bias
        b3 = torch.cat([c3.bias, c3.bias.new_zeros(1)])
        pack = torch.cat([t.reshape(-1).float() for t in (w1, w2, w3, b1, b2, b3)]).contiguous()
    assert pack.numel() == _lib.load().forge_decoder_wpack_floats()
    return pack


def pack_decoder_bwd_weights(conv_rgb):
    """fp32 weight pack of forge_decoder_bwd_data (flipped / transposed filters, BN scales folded; forge_b200.h)."""
    ct, bn1, _, c2, bn2, _, c3 = conv_rgb
    with torch.no_grad():
        s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
        s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)
        w3b = c3.weight.flip(2, 3).permute(2, 3, 0, 1)                               # [ky, kx, co 3, c 8]
        w2b = (c2.weight * s2.view(8, 1, 1, 1)).flip(2, 3).permute(2, 3, 0, 1)       # [ky, kx, co 8, ci 16]
        wd = (ct.weight * s1.view(1, 16, 1, 1)).permute(2, 3, 1, 0)                  # [u, v, co 16, ci 16]
        pack = torch.cat([t.reshape(-1).float() for t in (w3b, w2b, wd)]).contiguous()
    assert pack.numel() == _lib.load().forge_decoder_bwd_wpack_floats()
    return pack


def _decoder_backward(ctx, g):
    """Shared backward of the fused decoders -> (grad_x, [grad of every trainable conv_rgb parameter]).  Constant decoder
    weights (pose refinement): one kernel from the sign masks the forward pass saved.  Trainable weights: re-run the
    module's own convs under autograd (cuDNN); the parameters are inputs of the Function, so their gradients are
    returned to autograd (accumulate-grad / DDP hooks fire, torch.autograd.grad has no side effects)."""
    x_nhwc, masks = ctx.saved_tensors
    params = ctx.params
    if not params and masks.numel():
        N, Sh, Sw, _ = x_nhwc.shape
        g = _f32c(g)
        gx = torch.empty_like(x_nhwc)
        wb = pack_decoder_bwd_weights(ctx.conv_rgb) if ctx.bwd_pack is None else ctx.bwd_pack
        with torch.cuda.device(g.device):
            _lib.call("forge_decoder_bwd_data", _ptr(g), _ptr(masks), _ptr(wb), _ptr(gx), N, Sh, Sw, _stream(g))
        return gx, []
    with torch.enable_grad():
        xi = x_nhwc.detach().requires_grad_(ctx.needs_input_grad[0])
        y = torch.relu(ctx.conv_rgb(xi.permute(0, 3, 1, 2)))
        wanted = ([xi] if ctx.needs_input_grad[0] else []) + list(params)
        grads = list(torch.autograd.grad(y, wanted, g, allow_unused=True)) if wanted else []
    gx = grads.pop(0) if ctx.needs_input_grad[0] else None
    return gx, grads


def _trainable(conv_rgb):
    return [p for p in conv_rgb.parameters() if p.requires_grad] if torch.is_grad_enabled() else []


def _decoder_masks(ctx, x_nhwc, conv_rgb):
    """uint32 sign-mask buffer when the input needs a gradient and the decoder weights are constants, else empty."""
    N, Sh, Sw, _ = x_nhwc.shape
    want = ctx.needs_input_grad[0] and not any(p.requires_grad for p in conv_rgb.parameters())
    return torch.empty((N, 2 * Sh, 2 * Sw) if want else (0,), dtype=torch.int32, device=x_nhwc.device)


class _Decoder(torch.autograd.Function):
    """Fused fp32 inference decoder (see _decoder_backward for the backward pass); the trainable conv_rgb parameters
    ride along as inputs so that autograd owns their gradients."""

    @staticmethod
    def forward(ctx, x_nhwc, wpack, conv_rgb, bwd_pack, *params):
        N, Sh, Sw, _ = x_nhwc.shape
        rgb = torch.empty(N, 3, 2 * Sh, 2 * Sw, dtype=torch.float32, device=x_nhwc.device)
        masks = _decoder_masks(ctx, x_nhwc, conv_rgb)
        with torch.cuda.device(x_nhwc.device):
            _lib.call("forge_decoder_fwd", _ptr(x_nhwc), _ptr(wpack), _ptr(rgb), _ptr(masks) if masks.numel() else None,
                      N, Sh, Sw, _stream(x_nhwc))
        ctx.conv_rgb, ctx.bwd_pack, ctx.params = conv_rgb, bwd_pack, params
        ctx.save_for_backward(x_nhwc, masks)
        return rgb

    @staticmethod
    def backward(ctx, g):
        gx, gp = _decoder_backward(ctx, g)
        return (gx, None, None, None, *gp)


def decoder_fused(x_nhwc, wpack, conv_rgb, bwd_pack=None):
    """x [N,S,S,16] NHWC -> relu(conv_rgb(x)) [N,3,2S,2S] through the fused kernel (eval-mode BN).  bwd_pack:
    optional cached pack_decoder_bwd_weights(conv_rgb) for the constant-weights backward."""
    _require_cuda(x_nhwc, wpack)
    if x_nhwc.shape[-1] != 16:
        raise ValueError("the decoder input must have 16 channels")
    return _Decoder.apply(_f32c(x_nhwc), wpack, conv_rgb, bwd_pack, *_trainable(conv_rgb))


# ---- tensor-core (bf16) decoder ------------------------------------------------------------------
def _core_blocks(mat):
    """[n, k] (n, k multiples of 8 after zero padding) -> UMMA K-major tile [k / 8][n][k % 8] (8-row core matrices)."""
    n, k = -(-mat.shape[0] // 8) * 8, -(-mat.shape[1] // 8) * 8
    t = mat.new_zeros(n, k)
    t[:mat.shape[0], :mat.shape[1]] = mat
    return t.view(n, k // 8, 8).permute(1, 0, 2).contiguous()


def _strip(w_kx):
    """One weight strip of decoder_tc.cu: 13 core matrices [8 x zero, W(kx=4), W(3), W(2), W(1), W(0)], each [8 n][8 k].
    The B operand of source column j is the window of 8 blocks starting at block 12 - j."""
    s = w_kx[0].new_zeros(13, 8, 8)
    for kx in range(5):
        blk = w_kx[kx]
        s[12 - kx, :blk.shape[0], :blk.shape[1]] = blk
    return s


def pack_decoder_tc_weights(conv_rgb):
    """Weight pack for forge_decoder_tc_fwd from the reference-shaped ``conv_rgb`` Sequential in eval mode: bf16
    B operands (9 layer-1 tiles, 10 + 6 strips; the BN scales are folded into the weights) followed by the fp32
    epilogue shifts b1[16] b2[8] b3[4] (BN shift with the conv bias folded in); layout documented in forge_b200.h."""
    ct, bn1, _, c2, bn2, _, c3 = conv_rgb
    with torch.no_grad():
        s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
        s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)
        wt = ct.weight.float() * s1.view(1, 16, 1, 1)            # [ci, co, ky, kx], BN scale folded before bf16 rounding
        w2 = c2.weight.float() * s2.view(8, 1, 1, 1)             # [co, ci, ky, kx]
        w3 = c3.weight.float()
        parts = []
        for a in range(3):                       # layer 1: one [64 = (py, px, co)] x [16 ci] tile per input shift (a, b)
            for b in range(3):
                rows = [wt[:, :, py + 4 - 2 * a, px + 4 - 2 * b].t() for py in (0, 1) for px in (0, 1)]   # each [co, ci]
                parts.append(_core_blocks(torch.cat(rows, dim=0)).reshape(-1))
        zeros8 = wt.new_zeros(8 * 64)
        for ky in range(5):                      # layer 2: strips (ky, ci chunk)
            for c in range(2):
                parts.append(_strip([w2[:, 8 * c:8 * c + 8, ky, kx] for kx in range(5)]).reshape(-1))
        parts.append(zeros8)
        for ky in range(5):                      # layer 3: strips ky, a phantom all-zero strip ky = 5
            parts.append(_strip([w3[:, :, ky, kx] for kx in range(5)]).reshape(-1))
        parts.append(wt.new_zeros(13 * 64))
        parts.append(zeros8)
        wb = torch.cat(parts).to(torch.bfloat16).contiguous().view(torch.uint8)
        prm = torch.zeros(64, dtype=torch.float32, device=wt.device)
        prm[0:16] = (ct.bias - bn1.running_mean) * s1 + bn1.bias
        prm[16:24] = (c2.bias - bn2.running_mean) * s2 + bn2.bias
        prm[24:27] = c3.bias
        pack = torch.cat([wb, prm.view(torch.uint8).reshape(-1)]).contiguous()
    assert pack.numel() == _lib.load().forge_decoder_tc_wpack_bytes()
    return pack


class _DecoderTC(torch.autograd.Function):
    """bf16 tensor-core inference decoder (see _decoder_backward for the backward pass)."""

    @staticmethod
    def forward(ctx, x_nhwc, wpack, conv_rgb, max_ctas, bwd_pack, *params):
        N, Sh, Sw, _ = x_nhwc.shape
        rgb = torch.empty(N, 3, 2 * Sh, 2 * Sw, dtype=torch.float32, device=x_nhwc.device)
        masks = _decoder_masks(ctx, x_nhwc, conv_rgb)
        with torch.cuda.device(x_nhwc.device):
            _lib.call("forge_decoder_tc_fwd", _ptr(x_nhwc), _ptr(wpack), _ptr(rgb), _ptr(masks) if masks.numel() else None,
                      N, Sh, Sw, int(max_ctas), _stream(x_nhwc))
        ctx.conv_rgb, ctx.bwd_pack, ctx.params = conv_rgb, bwd_pack, params
        ctx.save_for_backward(x_nhwc, masks)
        return rgb

    @staticmethod
    def backward(ctx, g):
        gx, gp = _decoder_backward(ctx, g)
        return (gx, None, None, None, None, *gp)


def decoder_tc(x_nhwc, wpack, conv_rgb, max_ctas=0, bwd_pack=None):
    """x [N,S,S,16] NHWC fp32 -> relu(conv_rgb(x)) [N,3,2S,2S] fp32 on the tcgen05 tensor cores (bf16 operands,
    fp32 accumulation, eval-mode BN)."""
    _require_cuda(x_nhwc, wpack)
    if x_nhwc.shape[-1] != 16:
        raise ValueError("the decoder input must have 16 channels")
    return _DecoderTC.apply(_f32c(x_nhwc), wpack, conv_rgb, max_ctas, bwd_pack, *_trainable(conv_rgb))


def umma_probe(image_u8, a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo):
    """Test hook: one tcgen05.mma (M=128, N=16, K=16, bf16, no swizzle) over a shared-memory image -> D [128,16]."""
    _require_cuda(image_u8)
    out = torch.empty(128, 16, dtype=torch.float32, device=image_u8.device)
    with torch.cuda.device(image_u8.device):
        _lib.call("forge_umma_probe", _ptr(image_u8), image_u8.numel(), a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo,
                  _ptr(out), _stream(image_u8))
    return out


# ---- camera / pose algebra -------------------------------------------------------------------------
class _CameraPrep(torch.autograd.Function):
    """(R [N,3,3], T [N,3], K_half [N,3,3]) -> cam12 [N,12], origin_proj [N,2]; one launch forward, one backward."""

    @staticmethod
    def forward(ctx, R, T, K_half, scale, eps):
        N = R.shape[0]
        cam12 = torch.empty(N, 12, dtype=torch.float32, device=R.device)
        oproj = torch.empty(N, 2, dtype=torch.float32, device=R.device)
        with torch.cuda.device(R.device):
            _lib.call("forge_camera_prep_fwd", _ptr(R), _ptr(T), _ptr(K_half), N, scale[0], scale[1], scale[2], eps,
                      _ptr(cam12), _ptr(oproj), _stream(R))
        ctx.save_for_backward(R, T, K_half)
        ctx.scale, ctx.eps = scale, eps
        return cam12, oproj

    @staticmethod
    def backward(ctx, g_cam12, g_oproj):
        R, T, K_half = ctx.saved_tensors
        N = R.shape[0]
        need = ctx.needs_input_grad
        gR = torch.empty_like(R) if need[0] else None
        gT = torch.empty_like(T) if need[1] else None
        gK = torch.empty_like(K_half) if need[2] else None
        if any(need[:3]):
            gc = None if g_cam12 is None else _f32c(g_cam12)
            go = None if g_oproj is None else _f32c(g_oproj)
            with torch.cuda.device(R.device):
                _lib.call("forge_camera_prep_bwd", _ptr(R), _ptr(T), _ptr(K_half), N, ctx.scale[0], ctx.scale[1],
                          ctx.scale[2], ctx.eps, _ptr(gc), _ptr(go), _ptr(gR), _ptr(gT), _ptr(gK), _stream(R))
        return gR, gT, gK, None, None


def camera_prep(R, T, K_half, vol_dhw, volume_size, eps=1e-6):
    """OpenCV extrinsics / halved intrinsics -> (cam12 [N,12] for the raymarcher, origin_proj [N,2]); differentiable
    w.r.t. R, T and fx, fy, cx, cy.  Reference models/volume_render.py:53-61, :77-83."""
    _require_cuda(R, T, K_half)
    D, H, W = vol_dhw
    vs = float(volume_size) / D
    scale = ((W - 1) * 0.5 * vs, (H - 1) * 0.5 * vs, (D - 1) * 0.5 * vs)
    return _CameraPrep.apply(_f32c(R), _f32c(T), _f32c(K_half), scale, float(eps))


class _PoseAffine(torch.autograd.Function):
    """poses [B,t,4,4] -> affine12 [B*t,12] (view 0: identity); backward with three small batched matmuls."""

    @staticmethod
    def forward(ctx, poses):
        B, t = poses.shape[:2]
        A = torch.empty(B * t, 12, dtype=torch.float32, device=poses.device)
        need = ctx.needs_input_grad[0]
        pinv = torch.empty(B, t, 4, 4, dtype=torch.float32, device=poses.device) if need else None
        with torch.cuda.device(poses.device):
            _lib.call("forge_pose_affine_fwd", _ptr(poses), B, t, _ptr(A), _ptr(pinv), None, _stream(poses))
        if need:
            ctx.save_for_backward(poses, pinv)
        return A

    @staticmethod
    def backward(ctx, gA):
        poses, X = ctx.saved_tensors                       # X = inverse(poses)
        B, t = poses.shape[:2]
        G = torch.zeros(B, t, 4, 4, dtype=torch.float32, device=poses.device)
        G[:, 1:, :3, :] = gA.reshape(B, t, 3, 4)[:, 1:]    # T = P0 X_v for v >= 1; the identity row is a constant
        Xt = X.transpose(-1, -2)
        gP = -(Xt @ (poses[:, :1].transpose(-1, -2) @ G) @ Xt)          # dL/dP_v = -X^T (P0^T G) X^T
        gP[:, 0] = (G @ Xt).sum(dim=1)                                  # dL/dP_0 = sum_v G_v X_v^T  (G_0 = 0)
        return gP


def pose_affine(camPoses):
    """camPoses [B,t,4,4] -> [B*t,12]: rows of pose_0 @ inverse(pose_v) (reference models/rotate.py:64-89), identity for
    v = 0, in one launch (differentiable)."""
    _require_cuda(camPoses)
    if camPoses.dim() != 4 or camPoses.shape[-2:] != (4, 4):
        raise ValueError("camPoses_cv2 must be [B,t,4,4] (got %s)" % (tuple(camPoses.shape),))
    return _PoseAffine.apply(_f32c(camPoses))


# ---- x2 bilinear upsample of the silhouette / depth maps ---------------------------------------------
class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        M, Sh, Sw = a.shape
        oa = torch.empty(M, 1, 2 * Sh, 2 * Sw, dtype=torch.float32, device=a.device)
        ob = torch.empty_like(oa) if b is not None else None
        with torch.cuda.device(a.device):
            _lib.call("forge_upsample2x_fwd", _ptr(a), _ptr(b), _ptr(oa), _ptr(ob), M, Sh, Sw, _stream(a))
        ctx.dims = (M, Sh, Sw)
        ctx.two = b is not None
        if ob is None:
            ob = torch.empty(0, device=a.device)
            ctx.mark_non_differentiable(ob)
        return oa, ob

    @staticmethod
    def backward(ctx, ga, gb):
        M, Sh, Sw = ctx.dims
        if ga is None and (gb is None or not ctx.two):
            return None, None
        dev = (ga if ga is not None else gb).device
        ga = torch.zeros(M, 1, 2 * Sh, 2 * Sw, device=dev) if ga is None else _f32c(ga)
        gb = (torch.zeros_like(ga) if gb is None else _f32c(gb)) if ctx.two else None
        sa = torch.empty(M, Sh, Sw, dtype=torch.float32, device=dev)
        sb = torch.empty_like(sa) if ctx.two else None
        with torch.cuda.device(dev):
            _lib.call("forge_upsample2x_bwd", _ptr(ga), _ptr(gb), _ptr(sa), _ptr(sb), M, Sh, Sw, _stream(ga))
        return sa, sb


def upsample2x(a, b=None):
    """[M,S_h,S_w] maps (one or two) -> [M,1,2S_h,2S_w] each, bilinear with align_corners=False (reference
    models/volume_render.py:69,74: F.upsample to twice the size), both maps in one launch; differentiable."""
    _require_cuda(a, b)
    oa, ob = _Upsample2x.apply(_f32c(a), None if b is None else _f32c(b))
    return oa, (ob if b is not None else None)


# ---- ConvGRU cell: fused elementwise stages -----------------------------------------------------------
def _gru_layout(g):
    """conv output [B,CC,D,H,W] -> (dense tensor, channels_last flag)"""
    if g.is_contiguous(memory_format=torch.channels_last_3d) and not g.is_contiguous():
        return g, 1
    return (g if g.is_contiguous() else g.contiguous()), 0


def _gru_fmt(cl):
    return torch.channels_last_3d if cl else torch.contiguous_format


def _gru_operand(t, cl):
    """fp32 [B,C,D,H,W] whose inner dims are dense in the given layout (the batch stride is free) -> (tensor, batch stride)"""
    if t.dtype != torch.float32:
        t = t.float()
    B, C, D, H, W = t.shape
    want = (None, 1, H * W * C, W * C, C) if cl else (None, D * H * W, H * W, W, 1)
    ok = all(t.shape[i] == 1 or t.stride(i) == want[i] for i in range(1, 5))
    if not ok:
        t = t.contiguous(memory_format=_gru_fmt(cl))
    return t, (t.stride(0) if B > 1 else C * D * H * W)


def _gru_dense(t, cl, dtype=torch.float32):
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous(memory_format=_gru_fmt(cl))


class _GruGate(torch.autograd.Function):
    """xhr = cat(x, h * sigmoid(g[:, C:])) in one launch."""

    @staticmethod
    def forward(ctx, g, h, x):
        g, cl = _gru_layout(g)
        B, C2, D, H, W = g.shape
        C, S = C2 // 2, D * H * W
        h, hbs = _gru_operand(h, cl)
        x, xbs = _gru_operand(x, cl)
        xhr = torch.empty(g.shape, dtype=torch.float32, device=g.device, memory_format=_gru_fmt(cl))
        bf = int(g.dtype == torch.bfloat16)
        with torch.cuda.device(g.device):
            _lib.call("forge_gru_gate_fwd", _ptr(g), bf, _ptr(h), hbs, _ptr(x), xbs, _ptr(xhr), cl, B, C, S, _stream(g))
        ctx.save_for_backward(g, h)
        ctx.meta = (cl, B, C, S, hbs, bf)
        return xhr

    @staticmethod
    def backward(ctx, d_xhr):
        g, h = ctx.saved_tensors
        cl, B, C, S, hbs, bf = ctx.meta
        d_xhr = _gru_dense(d_xhr, cl)
        dg = torch.empty_like(g)
        dh = torch.empty((B, C) + tuple(g.shape[2:]), dtype=torch.float32, device=g.device, memory_format=_gru_fmt(cl))
        dx = torch.empty_like(dh)
        with torch.cuda.device(g.device):
            _lib.call("forge_gru_gate_bwd", _ptr(d_xhr), _ptr(g), bf, _ptr(h), hbs, _ptr(dg), _ptr(dh), _ptr(dx), cl, B, C, S,
                      _stream(g))
        return dg, dh, dx


class _GruOut(torch.autograd.Function):
    """h' = h (1 - u) + tanh(o) u with u = sigmoid(g[:, :C]) in one launch."""

    @staticmethod
    def forward(ctx, o, g, h):
        g, cl = _gru_layout(g)
        B, C2, D, H, W = g.shape
        C, S = C2 // 2, D * H * W
        o = _gru_dense(o, cl, g.dtype)
        h, hbs = _gru_operand(h, cl)
        hn = torch.empty(o.shape, dtype=torch.float32, device=g.device, memory_format=_gru_fmt(cl))
        bf = int(g.dtype == torch.bfloat16)
        with torch.cuda.device(g.device):
            _lib.call("forge_gru_out_fwd", _ptr(o), _ptr(g), bf, _ptr(h), hbs, _ptr(hn), cl, B, C, S, _stream(g))
        ctx.save_for_backward(o, g, h)
        ctx.meta = (cl, B, C, S, hbs, bf)
        return hn

    @staticmethod
    def backward(ctx, d_hn):
        o, g, h = ctx.saved_tensors
        cl, B, C, S, hbs, bf = ctx.meta
        d_hn = _gru_dense(d_hn, cl)
        d_o = torch.empty_like(o)
        dg = torch.empty_like(g)
        dh = torch.empty(o.shape, dtype=torch.float32, device=g.device, memory_format=_gru_fmt(cl))
        with torch.cuda.device(g.device):
            _lib.call("forge_gru_out_bwd", _ptr(d_hn), _ptr(o), _ptr(g), bf, _ptr(h), hbs, _ptr(d_o), _ptr(dg), _ptr(dh), cl, B, C, S,
                      _stream(g))
        return d_o, dg, dh


def gru_gate(g, h, x):
    """g = conv_gate(cat(x, h)) [B,2C,D,H,W] (fp32 or bf16), h, x [B,C,D,H,W] -> cat(x, h * sigmoid(g[:, C:])) fp32
    (reference models/fusion.py:31-33: sigmoid, split, mul, cat)."""
    _require_cuda(g, h, x)
    if g.dtype not in (torch.float32, torch.bfloat16):
        g = g.float()
    return _GruGate.apply(g, h.float(), x.float())


def gru_out(o, g, h):
    """o = out_gate(cat(x, h r)) [B,C,D,H,W], g as above, h -> h (1 - u) + tanh(o) u, u = sigmoid(g[:, :C])
    (reference models/fusion.py:33-35)."""
    _require_cuda(o, g, h)
    if g.dtype not in (torch.float32, torch.bfloat16):
        g = g.float()
    return _GruOut.apply(o.to(g.dtype), g, h.float())


# ---- tensor-core 3-D convolution with fused ConvGRU epilogues (inference / pose refinement) --------------------------------
def pack_conv3d_weights(weight):
    """[Cout, Cin, 3, 3, 3] fp32 -> bf16 [27][Cin / 64][Cout][64] for forge_conv3d_tc (tap index (dz*3 + dy)*3 + dx)."""
    Cout, Cin = weight.shape[:2]
    if tuple(weight.shape[2:]) != (3, 3, 3) or Cin % 64:
        raise ValueError("forge_conv3d_tc needs a 3x3x3 kernel and Cin a multiple of 64 (got %s)" % (tuple(weight.shape),))
    w = weight.detach().float().permute(2, 3, 4, 1, 0).reshape(27, Cin // 64, 64, Cout).permute(0, 1, 3, 2)
    return w.contiguous().to(torch.bfloat16)


def _bf16_cl(t):
    """[B,C,D,H,W] (any strides) -> bf16 tensor whose memory is [B][D][H][W][C] with dense inner dims; returns (tensor, batch stride)"""
    B, C, D, H, W = t.shape
    want = (1, H * W * C, W * C, C)
    ok = t.dtype == torch.bfloat16 and all(t.shape[i + 1] == 1 or t.stride(i + 1) == want[i] for i in range(4))
    if not ok:
        t = t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    return t, (t.stride(0) if B > 1 else C * D * H * W)


def conv3d_tc(x, wpack, mode, shift, h2=None, scale=None, lrelu=False, norm_shift=None, h_state=None, u_in=None,
              want_f32=True, want_bf16=False, want_norm=False, want_aux=False, max_ctas=0):
    """forge_conv3d_tc on [B,C,D,H,W] tensors (channels-last memory).  Returns (out_f32, out_bf16, out_norm) as
    [B,Cg,D,H,W] channels-last views (None where not requested).  mode: 'plain' | 'gate' | 'out' (see forge_b200.h)."""
    _require_cuda(x, wpack, shift)
    B, Cx, D, H, W = x.shape
    xb, xbs = _bf16_cl(x)
    hb, hbs, Ch = None, 0, 0
    if h2 is not None:
        hb, hbs = _bf16_cl(h2)
        Ch = h2.shape[1]
    Cout = wpack.shape[2]
    m = {'plain': 0, 'gate': 1, 'out': 2}[mode]
    Cg = Cout // 2 if m == 1 else Cout
    dev = x.device

    def cl(dtype):
        return torch.empty(B, D, H, W, Cg, dtype=dtype, device=dev)

    def dense_f32(t):       # fp32 state tensors must be dense channels-last
        if t is None:
            return None
        tt = t.permute(0, 2, 3, 4, 1)
        return tt if (tt.is_contiguous() and t.dtype == torch.float32) else tt.float().contiguous()
    hs, ui = dense_f32(h_state), dense_f32(u_in)
    o32 = cl(torch.float32) if (want_f32 or m != 0) else None
    o16 = cl(torch.bfloat16) if (want_bf16 or m == 1) else None
    on = cl(torch.float32) if want_norm else None
    ax = cl(torch.float32) if want_aux else None
    with torch.cuda.device(dev):
        _lib.call("forge_conv3d_tc", _ptr(xb), xbs, Cx, _ptr(hb), hbs, Ch, _ptr(wpack), m, int(bool(lrelu)), _ptr(scale),
                  _ptr(shift), _ptr(norm_shift), _ptr(hs), _ptr(ui), _ptr(o32), _ptr(o16), _ptr(on), _ptr(ax), 0, 0, B, D, H, W, Cout,
                  int(max_ctas), _stream(x))
    view = lambda t: None if t is None else t.permute(0, 4, 1, 2, 3)     # noqa: E731
    if want_aux:
        return view(o32), view(o16), view(on), view(ax)
    return view(o32), view(o16), view(on)


class _GruTC(torch.autograd.Function):
    """The ConvGRU recurrence over the views + fusion_norm (eval) on forge_conv3d_tc, differentiable w.r.t. the views and the
    initial state (constant weights: the pose-refinement loop).  Backward = the two transposed convolutions per step through the
    same kernel (flipped / transposed weight packs) + the gate derivatives."""

    @staticmethod
    def forward(ctx, pk, h0, *views):
        t = len(views)
        xb = [v.to(torch.bfloat16) for v in views]
        h = h0.float()
        h16 = h.to(torch.bfloat16)
        saved = []
        out = None
        for i in range(t):
            last = i == t - 1
            u, hr, _, r = conv3d_tc(xb[i], pk['wg'], 'gate', pk['bg'], h2=h16, h_state=h, want_aux=True)
            hn, hn16, out, c = conv3d_tc(xb[i], pk['wo'], 'out', pk['bo'], h2=hr, h_state=h, u_in=u, scale=pk['ns'],
                                         norm_shift=pk['nb'], want_bf16=not last, want_norm=last, want_aux=True)
            saved += [h, u, r, c]
            h, h16 = hn, hn16
        ctx.pk, ctx.t = pk, t
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, d_out):
        pk, t = ctx.pk, ctx.t
        saved = ctx.saved_tensors
        B, C, D, H, W = saved[0].shape
        N, dev = B * D * H * W, d_out.device
        zero = pk['zero256']
        st = _stream(d_out)

        def rows(x):                  # [B,C,D,H,W] -> dense channels-last rows [N, C] fp32 (zero-copy for the saved tensors)
            x = x.permute(0, 2, 3, 4, 1)
            return x if (x.is_contiguous() and x.dtype == torch.float32) else x.float().contiguous()

        def new(c, dtype=torch.float32):
            return torch.empty(B, D, H, W, c, dtype=dtype, device=dev)

        dh = rows(d_out * pk['ns'].view(1, -1, 1, 1, 1))                  # through fusion_norm (eval): a per-channel scale
        dxs = [None] * t
        with torch.cuda.device(dev):
            for i in reversed(range(t)):
                h, u, r, c = (rows(x) for x in saved[4 * i:4 * i + 4])
                d_o, dgu, dh_dir = new(C, torch.bfloat16), new(C), new(C)
                _lib.call("forge_gru_tc_bwd", 0, _ptr(dh), _ptr(u), _ptr(c), _ptr(h), None, _ptr(d_o), _ptr(dgu), _ptr(dh_dir), N, C, st)
                g1, _, _ = conv3d_tc(d_o.permute(0, 4, 1, 2, 3), pk['woT'], 'plain', zero)      # [dx_o | d(h r)]
                g1 = rows(g1)
                dg, dh_acc = new(2 * C, torch.bfloat16), new(C)
                _lib.call("forge_gru_tc_bwd", 1, _ptr(g1), _ptr(r), _ptr(h), _ptr(dgu), _ptr(dh_dir), _ptr(dg), _ptr(dh_acc), None, N, C, st)
                g2, _, _ = conv3d_tc(dg.permute(0, 4, 1, 2, 3), pk['wgT'], 'plain', zero)       # [dx_g | dh_g]
                g2 = rows(g2)
                dx, dh = new(C), new(C)
                _lib.call("forge_gru_tc_bwd", 2, _ptr(g1), _ptr(g2), _ptr(dh_acc), None, None, None, _ptr(dx), _ptr(dh), N, C, st)
                dxs[i] = dx.permute(0, 4, 1, 2, 3)
        return (None, dh.permute(0, 4, 1, 2, 3), *dxs)


def gru_tc(pk, h0, views):
    """fusion_norm(GRU(views; h0)) on the tensor cores with autograd support for views / h0 (weights are constants)."""
    return _GruTC.apply(pk, h0, *views)


# ---- render heads on the tensor cores (inference): reference models/encoder.py:16-34 -----------------------------------------------
def _convT_as_conv_weights(ct):
    """ConvTranspose3d(Cin, 32, k 4, stride 2, pad 1) weight [Cin, 32, 4, 4, 4] -> the equivalent 3x3x3 convolution weight
    [8 * 32, Cin, 3, 3, 3] whose 8 output groups are the 8 output parity classes (qz, qy, qx): out[2j + q] takes, per axis,
    input j with tap 1 + q and input j - 1 (q = 0, tap 3) or j + 1 (q = 1, tap 0)."""
    W = ct.weight.detach().float()
    Cin, Co = W.shape[:2]
    taps = {0: ((0, 1), (-1, 3)), 1: ((0, 2), (1, 0))}          # parity -> ((input offset, kernel tap), ...)
    Weq = W.new_zeros(8 * Co, Cin, 3, 3, 3)
    for q in range(8):
        for dz, kz in taps[q >> 2]:
            for dy, ky in taps[(q >> 1) & 1]:
                for dx, kx in taps[q & 1]:
                    Weq[q * Co:(q + 1) * Co, :, dz + 1, dy + 1, dx + 1] = W[:, :, kz, ky, kx].t()
    return Weq


def pack_heads_tc(features_head, density_head):
    """weight / affine packs for heads_tc from the reference-shaped Sequentials (eval-mode BatchNorm folded)"""
    ctf, bnf1, _, cf2, bnf2 = features_head
    ctd, bnd1, _, cd2, bnd2, _, cd3, _ = density_head
    with torch.no_grad():
        def fold(conv, bn):
            s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
            return s, ((conv.bias - bn.running_mean) * s + bn.bias).float()
        sf1, bf1 = fold(ctf, bnf1)
        sd1, bd1 = fold(ctd, bnd1)
        sf2, bf2 = fold(cf2, bnf2)
        sd2, bd2 = fold(cd2, bnd2)
        dev = ctf.weight.device
        W2 = torch.zeros(32, 64, 3, 3, 3, device=dev)
        W2[0:16, 0:32] = cf2.weight.float()                   # render features read stem channels 0..31
        W2[16:24, 32:64] = cd2.weight.float()                 # the density branch reads stem channels 32..63
        return dict(
            wf=pack_conv3d_weights(_convT_as_conv_weights(ctf)), sf=sf1.repeat(8).contiguous(), bf=bf1.repeat(8).contiguous(),
            wd=pack_conv3d_weights(_convT_as_conv_weights(ctd)), sd=sd1.repeat(8).contiguous(), bd=bd1.repeat(8).contiguous(),
            w2=pack_conv3d_weights(W2), s2=torch.cat([sf2, sd2, torch.ones(8, device=dev)]).contiguous(),
            b2=torch.cat([bf2, bd2, torch.zeros(8, device=dev)]).contiguous(),
            w3=cd3.weight.detach().float()[0].permute(1, 2, 3, 0).reshape(27, 8).contiguous(), b3=float(cd3.bias.item()))


def heads_tc(pk, fused):
    """fused [B,128,D,H,W] -> (render features [B,16,2D,2H,2W] fp32 channels-last, density [B,1,2D,2H,2W] fp32): both
    transposed-convolution stems as 27-tap tcgen05 GEMMs with a pixel-shuffle epilogue into ONE 64-channel bf16 tensor, both second
    convolutions as one 64 -> 32 GEMM, the 8 -> 1 density convolution + ReLU as a small direct kernel (4 launches; reference: 8
    cuDNN convolutions / norms / activations)."""
    _require_cuda(fused)
    B, C, D, H, W = fused.shape
    x, xbs = _bf16_cl(fused)
    dev = fused.device
    stem = torch.empty(B, 2 * D, 2 * H, 2 * W, 64, dtype=torch.bfloat16, device=dev)
    feat = torch.empty(B, 2 * D, 2 * H, 2 * W, 16, dtype=torch.float32, device=dev)
    d8 = torch.empty(B, 2 * D, 2 * H, 2 * W, 8, dtype=torch.float32, device=dev)
    dens = torch.empty(B, 1, 2 * D, 2 * H, 2 * W, dtype=torch.float32, device=dev)
    st = _stream(fused)
    with torch.cuda.device(dev):
        for w, s_, b_, off in ((pk['wf'], pk['sf'], pk['bf'], 0), (pk['wd'], pk['sd'], pk['bd'], 32)):
            _lib.call("forge_conv3d_tc", _ptr(x), xbs, C, None, 0, 0, _ptr(w), 3, 1, _ptr(s_), _ptr(b_), None, None, None, None,
                      _ptr(stem), None, None, 64, off, B, D, H, W, 256, 0, st)
        _lib.call("forge_conv3d_tc", _ptr(stem), stem[0].numel(), 64, None, 0, 0, _ptr(pk['w2']), 4, 0, _ptr(pk['s2']), _ptr(pk['b2']),
                  None, None, None, _ptr(feat), None, None, _ptr(d8), 0, 0, B, 2 * D, 2 * H, 2 * W, 32, 0, st)
        _lib.call("forge_conv3d_c8_to_1_relu", _ptr(d8), _ptr(pk['w3']), pk['b3'], _ptr(dens), B, 2 * D, 2 * H, 2 * W, st)
    return feat.permute(0, 4, 1, 2, 3), dens
