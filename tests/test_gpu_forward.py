"""GPU parity tests (forward): CUDA path through the C ABI vs the oracle, golden fixtures, KATs.

Tolerances: fp32 images max-abs <= 1e-4 (BASELINE.json north_star); index math bit-exact.
"""
import warnings

import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

from forge_b200 import ops, synthetic as syn            # noqa: E402
from forge_b200.models.volume_render import VolRender, camera_to_cam12  # noqa: E402
from forge_b200.models.rotate import Rotate_world      # noqa: E402
from oracle import reference_path as rp, closed_form as cf   # noqa: E402

DEV = 'cuda'
TOL = 1e-4
# the decoder stays cuDNN in this round: compare it in true fp32 (TF32 alone costs ~1e-4)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def test_library_loads_and_abi():
    from forge_b200 import _lib
    lib = _lib.load()
    assert lib.forge_abi_version() == _lib.ABI_VERSION
    with pytest.raises(RuntimeError, match="null pointer"):
        _lib.call("forge_ncs_to_nsc", None, None, 1, 1, 1, None)


@pytest.mark.parametrize("shape", [(2, 16, 5, 7, 9), (3, 128, 8, 8, 8), (1, 3, 4, 4, 33)])
def test_relayout_roundtrip_bit_exact(shape):
    x = torch.randn(*shape, device=DEV)
    cl = ops.to_channels_last(x)
    assert cl.shape == (shape[0],) + shape[2:] + (shape[1],)
    assert torch.equal(cl, x.permute(0, 2, 3, 4, 1).contiguous())
    assert torch.equal(ops.from_channels_last(cl), x)
    # already channels-last in memory -> zero copy
    y = cl.permute(0, 4, 1, 2, 3)
    assert ops.to_channels_last(y).data_ptr() == cl.data_ptr()


@pytest.mark.parametrize("ac", [True, False])
def test_sampler_index_math_bit_exact(ac):
    """floor index + in-bounds flags of the device samplers == ATen's formula evaluated in fp32 on CPU."""
    torch.manual_seed(3)
    D, H, W = 64, 48, 33
    pts = torch.cat([torch.rand(200000, 3) * 2.6 - 1.3,
                     torch.tensor([[-1., -1, -1], [1, 1, 1], [0, 0, 0], [1.0000001, -1.0000001, 0.99999994]]),
                     (torch.randint(0, 64, (5000, 3)).float() / 31.5 - 1)])   # many exact-integer hits
    base, mask = ops.sample_points(pts.to(DEV), D, H, W, ac)
    size = torch.tensor([W, H, D], dtype=torch.float32)
    if ac:
        idx = ((pts + 1.0) / 2) * (size - 1)            # GridSampler.h:27-36, align_corners=True
    else:
        idx = ((pts + 1.0) * size - 1) / 2
    b = torch.floor(idx).to(torch.int32)
    assert torch.equal(base.cpu(), b)
    lim = torch.tensor([W, H, D], dtype=torch.int32)
    in0 = (b >= 0) & (b < lim)
    in1 = (b + 1 >= 0) & (b + 1 < lim)
    m = torch.zeros(len(pts), dtype=torch.int32)
    for c in range(8):
        ok = (in1[:, 0] if c & 1 else in0[:, 0]) & (in1[:, 1] if c & 2 else in0[:, 1]) & (in1[:, 2] if c & 4 else in0[:, 2])
        m |= ok.int() << c
    assert torch.equal(mask.cpu().int(), m)


def _module_from_golden(g):
    cfg = syn.make_config(img_size=g['img_size'], n_pts_per_ray=g['n_pts'], min_depth=g['min_depth'],
                          max_depth=g['max_depth'], volume_size=g['volume_size'])
    m = VolRender(cfg)
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith('sd.')}, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize("name", ["volrender_small", "volrender_dense"])
def test_volrender_matches_reference_golden(name):
    g = load_golden(name)
    m = _module_from_golden(g)
    idx = g['view2vol'].long()
    cam = dict(R=g['R'].clone(), T=g['T'].clone(), K=g['K'].clone())          # CPU cameras, CUDA volumes
    with torch.no_grad():
        feat, sil, depth, _, _, _ = m.render_features(cam, g['feat'][idx].to(DEV), g['dens'][idx].to(DEV), True)
    core = g['core'].to(DEV)
    assert (feat - core[..., :16]).abs().max().item() <= TOL
    assert (sil - core[..., 16]).abs().max().item() <= TOL
    assert (depth - core[..., 17]).abs().max().item() <= TOL
    assert torch.allclose(cam['K'][:, 0, 0], g['K'][:, 0, 0] / 2) and bool((cam['K'][:, 2, 2] == 1).all())
    # full forward through the public module API, as-called (one volume per view) and de-duplicated
    for dedup in (False, True):
        cam = dict(R=g['R'].clone(), T=g['T'].clone(), K=g['K'].clone())
        with torch.no_grad():
            if dedup:
                out = m(cam, g['feat'].to(DEV), g['dens'].to(DEV), render_depth=True, return_origin_proj=True,
                        view2vol=g['view2vol'])
            else:
                out = m(cam, g['feat'][idx].to(DEV), g['dens'][idx].to(DEV), render_depth=True, return_origin_proj=True)
        rgb, sil2, depth2, op = out
        assert rgb.shape == g['rgb'].shape and sil2.shape == g['sil'].shape
        assert (rgb.cpu() - g['rgb']).abs().max().item() <= TOL
        assert (sil2.cpu() - g['sil']).abs().max().item() <= TOL
        assert (depth2.cpu() - g['depth']).abs().max().item() <= TOL
        assert (op.cpu() - g['origin_proj']).abs().max().item() <= 1e-3   # pixels
    # return-tuple ordering of the other flag combinations (models/volume_render.py:77-88)
    cam = dict(R=g['R'].clone(), T=g['T'].clone(), K=g['K'].clone())
    with torch.no_grad():
        assert len(m(cam, g['feat'][idx].to(DEV), g['dens'][idx].to(DEV))) == 2
        cam = dict(R=g['R'].clone(), T=g['T'].clone(), K=g['K'].clone())
        o3 = m(cam, g['feat'][idx].to(DEV), g['dens'][idx].to(DEV), return_origin_proj=True)
        assert len(o3) == 3 and o3[2].shape == (len(idx), 2)


def _raymarch_vs_oracle(n_obj, n_views, img, vol, P, dense, seed=0, oracle_dev='cpu'):
    inp = syn.render_inputs(n_obj, n_views, img, vol, seed=seed, dense=dense)
    cfg = syn.make_config(img_size=img, n_pts_per_ray=P)
    m = VolRender(cfg).to(DEV).eval()
    cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
    with torch.no_grad():
        feat, sil, depth, _, _, Kh = m.render_features(cam, inp['feat'].to(DEV), inp['dens'].to(DEV), True,
                                                       view2vol=inp['view2vol'])
    idx = inp['view2vol'].long()
    od = oracle_dev
    worst = 0.0
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for b in range(n_obj):      # one object at a time keeps the oracle's [N,S,S,P,16] intermediates small
            sel = (idx == b).nonzero().squeeze(1)
            ren = rp.make_renderer(img, P, cfg.render.min_depth, cfg.render.max_depth)
            camb = dict(R=inp['R'][sel].to(od), T=inp['T'][sel].to(od), K=inp['K'][sel].clone().to(od))
            core, _ = rp.raymarch_core(ren, camb, inp['feat'][b:b + 1].expand(len(sel), -1, -1, -1, -1).to(od),
                                       inp['dens'][b:b + 1].expand(len(sel), -1, -1, -1, -1).to(od), img, 1.0, True)
            core = core.to(DEV)
            worst = max(worst, (feat[sel] - core[..., :16]).abs().max().item(),
                        (sil[sel] - core[..., 16]).abs().max().item(), (depth[sel] - core[..., 17]).abs().max().item())
    return worst, sil


def test_raymarch_cfg1_vs_oracle():
    """BASELINE.json configs[0]: 1 object, 5 views 64x64 (img 128), 32^3 voxels, 32 samples."""
    worst, sil = _raymarch_vs_oracle(1, 5, 128, 32, 32, dense=False)
    assert worst <= TOL, worst
    assert sil.max().item() > 0.5       # the rays do hit the volume


def test_raymarch_cfg1_dense_sigma_gt_1():
    """relu(randn)*1.5 densities: sign-alternating transmittance must be reproduced."""
    worst, sil = _raymarch_vs_oracle(1, 5, 128, 32, 32, dense=True, seed=4)
    assert sil.min().item() < -0.05 or sil.max().item() > 1.05      # 1-prod(1-s) left [0,1]
    assert worst <= 5e-4, worst      # values grow like prod|1-s|; tolerance scaled accordingly


def test_raymarch_cfg2_vs_oracle():
    """BASELINE.json configs[1]: 4 objects x 5 views 128x128 (img 256), 64^3 voxels, 64 samples."""
    worst, _ = _raymarch_vs_oracle(4, 5, 256, 64, 64, dense=False, seed=1, oracle_dev=DEV)
    assert worst <= TOL, worst


def test_raymarch_ragged_image_and_views_outside():
    """Image side not a multiple of the 8x8 tile, plus a camera that looks away from the volume."""
    inp = syn.render_inputs(1, 3, 2 * 21, 9, seed=7)
    flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0]))      # same camera centre, optical axis reversed
    inp['R'][2] = flip @ inp['R'][2]
    inp['T'][2] = flip @ inp['T'][2]
    cfg = syn.make_config(img_size=42, n_pts_per_ray=19)
    m = VolRender(cfg).to(DEV).eval()
    cam = dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())
    with torch.no_grad():
        feat, sil, depth, _, _, Kh = m.render_features(cam, inp['feat'].to(DEV), inp['dens'].to(DEV), True,
                                                       view2vol=inp['view2vol'])
        f, o, d = cf.raymarch(inp['R'], inp['T'], Kh.cpu(), inp['feat'].expand(3, -1, -1, -1, -1),
                              inp['dens'].expand(3, -1, -1, -1, -1), 21, 19, 0.5, 2.0, 1.0)
    assert (feat.cpu() - f).abs().max().item() <= TOL
    assert (sil.cpu() - o).abs().max().item() <= TOL
    assert (depth.cpu() - d).abs().max().item() <= TOL
    assert sil[2].abs().max().item() == 0.0 and feat[2].abs().max().item() == 0.0


@pytest.mark.parametrize("sigma", [0.25, 1.0, 2.0])
def test_kat_constant_volume(sigma):
    P, D, S = 8, 8, 8
    feat = torch.full((1, 16, D, D, D), 0.5, device=DEV)
    dens = torch.full((1, 1, D, D, D), sigma, device=DEV)
    R, T, _ = syn.ring_cameras(1)
    cfg = syn.make_config(img_size=2 * S, n_pts_per_ray=P, max_depth=1.9)
    m = VolRender(cfg).to(DEV)
    with torch.no_grad():
        f, o, d, _, _, _ = m.render_features(dict(R=R, T=T, K=syn.intrinsics(1, 2 * S)), feat, dens, True)
    Kh = (syn.intrinsics(1, 2 * S) / 2).double()
    Kh[:, 2, 2] = 1.0
    fo, oo, do = cf.raymarch(R.double(), T.double(), Kh, feat.cpu().double(), dens.cpu().double(), S, P, 0.5, 1.9, 1.0)
    assert (o.cpu().double() - oo).abs().max().item() <= 1e-5 * max(1.0, abs(1 - sigma) ** P)
    assert (f.cpu().double() - fo).abs().max().item() <= 1e-5 * max(1.0, abs(1 - sigma) ** P)
    if sigma == 1.0:
        assert (d.cpu().double() - do).abs().max().item() <= 1e-5


# ---- K2 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rotate_g16", "rotate_g32"])
def test_rotate_matches_reference_golden(name):
    g = load_golden(name)
    m = Rotate_world(syn.make_config()).to(DEV)
    with torch.no_grad():
        out = m(g['voxels'].to(DEV), g['poses'].to(DEV), grid_size=g['grid'])
    assert out.shape == g['out'].shape
    assert (out.cpu() - g['out']).abs().max().item() <= TOL
    assert torch.equal(out[:, 0].cpu(), g['voxels'][:, 0])
    gm = torch.tensor([m.grid_coord_max_16, m.grid_coord_max, m.grid_coord_max_48, m.grid_coord_max_64,
                       m.grid_coord_max_128], dtype=torch.float64)
    assert torch.equal(gm, g['grid_coord_max'])


def test_rotate_kat_identity_shrinks_borders():
    m = Rotate_world(syn.make_config()).to(DEV)
    ones = torch.ones(1, 2, 4, 32, 32, 32, device=DEV)
    eye = torch.eye(4, device=DEV).repeat(1, 2, 1, 1)
    r = m(ones, eye, grid_size=32)[0, 1, 0]
    assert r[5, 5, 5].item() == pytest.approx(1.0, abs=1e-6)
    assert r[0, 5, 5].item() == pytest.approx(0.5, abs=1e-6)
    assert r[0, 0, 5].item() == pytest.approx(0.25, abs=1e-6)
    assert r[0, 0, 0].item() == pytest.approx(0.125, abs=1e-6)


@pytest.mark.parametrize("B,t,C,n", [(2, 5, 128, 32), (1, 5, 128, 16), (1, 3, 16, 48), (1, 2, 6, 20)])
def test_rotate_vs_oracle(B, t, C, n):
    vox, poses = syn.rotate_inputs(B, t, C, n, seed=5)
    m = Rotate_world(syn.make_config()).to(DEV)
    with torch.no_grad():
        out = m(vox.to(DEV), poses.to(DEV), grid_size=n)
        ref = rp.rotate_world_forward(vox.to(DEV), poses.to(DEV), n, 1.0)      # same ATen ops as the reference, on GPU
    assert (out - ref).abs().max().item() <= TOL
    # layouts: channels-last input gives the same answer without the re-layout pass
    vox_cl = vox.to(DEV).permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4)
    with torch.no_grad():
        out2 = m(vox_cl, poses.to(DEV), grid_size=n)
    assert torch.equal(out, out2)


def test_rotate_order_folds_view_permutation():
    """order= reproduces chose_selected(rotate(x), sequence_from_distance(...)) (models/model.py:128-129)."""
    vox, poses = syn.rotate_inputs(2, 5, 16, 16, seed=6)
    m = Rotate_world(syn.make_config()).to(DEV)
    idxs = rp.sequence_from_distance(poses[:, :, :3, 3])
    with torch.no_grad():
        plain = m(vox.to(DEV), poses.to(DEV), grid_size=16)
        folded = m(vox.to(DEV), poses.to(DEV), grid_size=16, order=idxs)
    assert torch.equal(folded, rp.chose_selected(plain, idxs.to(DEV)))


def test_cpu_inputs_fail_loudly():
    m = VolRender(syn.make_config(img_size=16, n_pts_per_ray=4))
    inp = syn.render_inputs(1, 1, 16, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(dict(R=inp['R'], T=inp['T'], K=inp['K']), inp['feat'], inp['dens'])
    with pytest.raises(RuntimeError, match="CUDA"):
        Rotate_world(syn.make_config())(torch.zeros(1, 2, 4, 8, 8, 8), torch.eye(4).repeat(1, 2, 1, 1), grid_size=8)


@pytest.mark.parametrize("duplex", [True, False])
def test_streamed_renderer_matches_direct_call(duplex):
    """Host-to-host pipeline (batches in flight; downloads on their own stream, or queued behind the next upload on the copy
    stream) returns exactly what the direct call returns; calibrate() leaves a working pipeline behind."""
    from forge_b200.pipeline import StreamedRenderer
    cfg = syn.make_config(img_size=64, n_pts_per_ray=24)
    m = VolRender(cfg).to(DEV).eval()
    batches = [syn.render_inputs(2, 3, 64, 16, seed=s) for s in range(5)]
    sr = StreamedRenderer(m, 2, 6, 16, depth=2, duplex=duplex)
    if not duplex:
        p0 = {k: v.pin_memory() for k, v in batches[0].items()}
        o0 = (torch.empty(6, 32, 32, 16).pin_memory(), torch.empty(6, 32, 32).pin_memory(), torch.empty(6, 32, 32).pin_memory())
        ms = sr.calibrate(p0['feat'], p0['dens'], p0['R'], p0['T'], p0['K'], p0['view2vol'], *o0, steps=4)
        assert set(ms) == {"duplex", "half_duplex"} and all(v > 0 for v in ms.values())
        sr.duplex = False
    outs = []
    for b in batches:
        pin = {k: v.pin_memory() for k, v in b.items()}
        o = (torch.empty(6, 32, 32, 16).pin_memory(), torch.empty(6, 32, 32).pin_memory(), torch.empty(6, 32, 32).pin_memory())
        sr.submit(pin['feat'], pin['dens'], pin['R'], pin['T'], pin['K'], pin['view2vol'], *o)
        outs.append((pin, o))
    sr.drain()
    for b, (pin, o) in zip(batches, outs):
        with torch.no_grad():
            f, s, d, _, _, _ = m.render_features(dict(R=b['R'], T=b['T'], K=b['K'].clone()), b['feat'].to(DEV),
                                                 b['dens'].to(DEV), True, view2vol=b['view2vol'])
        assert torch.equal(o[0], f.cpu()) and torch.equal(o[1], s.cpu()) and torch.equal(o[2], d.cpu())


@pytest.mark.parametrize("N,S", [(2, 16), (1, 21), (3, 64)])
def test_fused_decoder_matches_module_convs(N, S):
    """forge_decoder_fwd == relu(conv_rgb(x)) in eval mode (strict fp32 cuDNN), incl. ragged tiles."""
    torch.manual_seed(N * 100 + S)
    m = VolRender(syn.make_config(img_size=2 * S, n_pts_per_ray=8)).to(DEV)
    for mod in m.conv_rgb:
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    m.eval()
    x = torch.randn(N, S, S, 16, device=DEV)
    with torch.no_grad():
        fused = m.decode(x)
        m.fused_decoder = False
        ref = m.decode(x)
    assert fused.shape == ref.shape == (N, 3, 2 * S, 2 * S)
    assert (fused - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    # weight-pack cache follows in-place parameter updates
    m.fused_decoder = True
    with torch.no_grad():
        m.conv_rgb[6].bias.add_(0.5)
        fused2 = m.decode(x)
        m.fused_decoder = False
        ref2 = m.decode(x)
    assert (fused2 - ref2).abs().max().item() <= 2e-5 * max(1.0, ref2.abs().max().item())
    # gradients flow through the fused forward (backward re-runs the module convs)
    m.fused_decoder = True
    xg = x.clone().requires_grad_(True)
    m.decode(xg).sum().backward()
    xr = x.clone().requires_grad_(True)
    m.fused_decoder = False
    m.zero_grad()
    m.decode(xr).sum().backward()
    assert torch.allclose(xg.grad, xr.grad, atol=1e-4, rtol=1e-4)


def test_raymarch_cfg4_stress_vs_oracle():
    """BASELINE.json configs[3] geometry: 256x256 rays per view (img 512), 128^3 voxels, 128 samples;
    one object, 2 views, oracle evaluated view by view on the GPU (same ATen ops as the reference)."""
    img, vol, P = 512, 128, 128
    inp = syn.render_inputs(1, 2, img, vol, seed=11)
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV).eval()
    feat_d, dens_d = inp['feat'].to(DEV), inp['dens'].to(DEV)
    with torch.no_grad():
        feat, sil, depth, _, _, Kh = m.render_features(dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone()),
                                                       feat_d, dens_d, True, view2vol=inp['view2vol'])
        worst = 0.0
        for v in range(2):
            f, o, d = cf.raymarch(inp['R'][v:v + 1].to(DEV), inp['T'][v:v + 1].to(DEV), Kh[v:v + 1], feat_d, dens_d,
                                  img // 2, P, 0.5, 2.0, 1.0)
            worst = max(worst, (feat[v:v + 1] - f).abs().max().item(), (sil[v:v + 1] - o).abs().max().item(),
                        (depth[v:v + 1] - d).abs().max().item())
            del f, o, d
    assert worst <= TOL, worst
    # linearity in the feature volume at full size (density fixed): render(a f1 + b f2) = a render(f1) + b render(f2)
    with torch.no_grad():
        f2 = torch.randn_like(feat_d)
        cam = lambda: dict(R=inp['R'].clone(), T=inp['T'].clone(), K=inp['K'].clone())   # noqa: E731
        r2, s2, _, _, _, _ = m.render_features(cam(), f2, dens_d, True, view2vol=inp['view2vol'])
        r12, s12, _, _, _, _ = m.render_features(cam(), 0.5 * feat_d - 2.0 * f2, dens_d, True, view2vol=inp['view2vol'])
    same_sil = bool(torch.equal(s2, sil)) and bool(torch.equal(s12, sil))   # silhouette ignores the features
    lin_err = (r12 - (0.5 * feat - 2.0 * r2)).abs().max().item()
    assert same_sil
    assert lin_err <= 1e-4


def test_rotate_cfg4_stress_vs_oracle():
    """64^3 x 128-channel volumes (the fusion grid of configs[3])."""
    vox, poses = syn.rotate_inputs(1, 2, 128, 64, seed=12)
    m = Rotate_world(syn.make_config()).to(DEV)
    with torch.no_grad():
        out = m(vox.to(DEV), poses.to(DEV), grid_size=64)
        ref = rp.rotate_world_forward(vox.to(DEV), poses.to(DEV), 64, 1.0)
    err = (out - ref).abs().max().item()
    assert err <= TOL
    # idempotence of the passthrough + exact zero outside: a pure translation by more than the volume size empties it
    far = poses.clone()
    far[:, 1:, :3, 3] += 10.0
    with torch.no_grad():
        out2 = m(vox.to(DEV), far.to(DEV), grid_size=64)
    passthrough = bool(torch.equal(out2[:, 0], vox[:, 0].to(DEV)))
    outside = out2[:, 1:].abs().max().item()
    assert passthrough
    assert outside == 0.0


def test_raymarch_non_cubic_volume_and_single_sample():
    """D != H != W volumes (x<->W, y<->H, z<->D axis order) and the P = 1 edge case."""
    torch.manual_seed(5)
    D, H, W, S = 10, 12, 14, 24
    feat = torch.randn(2, 16, D, H, W)
    dens = torch.rand(2, 1, D, H, W) * 0.5
    R, T, _ = syn.ring_cameras(4, seed=9)
    K = syn.intrinsics(4, 2 * S)
    v2v = torch.tensor([0, 1, 1, 0], dtype=torch.int32)
    for P in (17, 1):
        m = VolRender(syn.make_config(img_size=2 * S, n_pts_per_ray=P, min_depth=0.9 if P == 1 else 0.5,
                                      max_depth=1.4 if P == 1 else 2.0)).to(DEV).eval()
        with torch.no_grad():
            f, s, d, _, _, Kh = m.render_features(dict(R=R.clone(), T=T.clone(), K=K.clone()), feat.to(DEV), dens.to(DEV),
                                                  True, view2vol=v2v)
            fo, so, do = cf.raymarch(R, T, Kh.cpu(), feat[v2v.long()], dens[v2v.long()], S, P, m.min_depth, m.max_depth, 1.0)
        assert (f.cpu() - fo).abs().max().item() <= TOL and (s.cpu() - so).abs().max().item() <= TOL
        assert (d.cpu() - do).abs().max().item() <= TOL
        assert s.abs().max().item() > 0.01


def test_raymarch_omniobject3d_geometry_and_density_clamp():
    """omniobject3d config: volume_size 2.0, depths 3..5, camera at 4.0, densities clamped to [0,1] by the model
    (reference config/omniobject3d, models/model.py:140-141)."""
    img, vol, P = 64, 16, 32
    inp = syn.render_inputs(1, 3, img, vol, seed=21, dense=True)
    Rs, Ts, _ = syn.ring_cameras(3, camera_z=4.0, seed=2)
    cfg = syn.make_config(img_size=img, n_pts_per_ray=P, min_depth=3.0, max_depth=5.0, volume_size=2.0, camera_z=4.0,
                          dataset_name='omniobject3d')
    m = VolRender(cfg).to(DEV).eval()
    dens = inp['dens'].clamp(min=0.0, max=1.0)
    with torch.no_grad():
        f, s, d, _, _, Kh = m.render_features(dict(R=Rs.clone(), T=Ts.clone(), K=inp['K'].clone()), inp['feat'].to(DEV),
                                              dens.to(DEV), True, view2vol=inp['view2vol'])
        fo, so, do = cf.raymarch(Rs, Ts, Kh.cpu(), inp['feat'].expand(3, -1, -1, -1, -1), dens.expand(3, -1, -1, -1, -1),
                                 img // 2, P, 3.0, 5.0, 2.0)
    assert (f.cpu() - fo).abs().max().item() <= TOL and (s.cpu() - so).abs().max().item() <= TOL
    assert (d.cpu() - do).abs().max().item() <= 5 * TOL          # depths are ~4: same relative accuracy
    assert 0.2 < s.max().item() <= 1.0 + 1e-6


def test_rotate_non_cubic_volume():
    vox = torch.randn(1, 3, 8, 6, 10, 12)
    _, poses = syn.rotate_inputs(1, 3, 1, 4, seed=3)
    m = Rotate_world(syn.make_config()).to(DEV)
    with torch.no_grad():
        out = m(vox.to(DEV), poses.to(DEV), grid_size=12)
    gx, gmax = cf.rotate_axis(12, 1.0)
    A = cf.rotate_affine(poses)
    gz, gy = cf.rotate_axis(6, 1.0)[0], cf.rotate_axis(10, 1.0)[0]
    Z, Y, X = torch.meshgrid(gz, gy, gx, indexing='ij')
    Pm = torch.stack([X, Y, Z, torch.ones_like(X)], dim=-1)
    grid = torch.einsum('mab,dhwb->mdhwa', A[:, :3, :], Pm) / gmax
    ref = torch.nn.functional.grid_sample(vox[0, 1:], grid, mode='bilinear', padding_mode='zeros', align_corners=False)
    assert (out[0, 1:].cpu() - ref).abs().max().item() <= TOL


@pytest.mark.parametrize("dtype", [None, torch.bfloat16])
def test_volrender_forward_replays_from_a_cuda_graph(dtype):
    """Every C-ABI entry point is capture-safe: the whole VolRender.forward (camera prep, pack, raymarch, decoder,
    upsample) captured in one CUDA graph returns the eager result bit for bit, also for new inputs."""
    from forge_b200.pipeline import GraphedVolRender
    img, vol, P, b, t = 64, 32, 32, 1, 5
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV).eval()
    m.decoder_dtype = dtype
    g = GraphedVolRender(m, b * t, b, vol, render_depth=True, return_origin_proj=True)
    for seed in (0, 1):
        inp = syn.render_inputs(b, t, img, vol, seed=seed, device=DEV)
        with torch.no_grad():
            eager = m(dict(R=inp['R'], T=inp['T'], K=inp['K'].clone()), inp['feat'], inp['dens'], render_depth=True,
                      return_origin_proj=True, view2vol=inp['view2vol'])
        K = inp['K'].clone()
        out = g(dict(R=inp['R'], T=inp['T'], K=K), inp['feat'], inp['dens'], inp['view2vol'])
        assert torch.equal(K, inp['K'] / 2.0 * torch.tensor([[1., 1, 1], [1, 1, 1], [0, 0, 0]], device=DEV)
                           + torch.tensor([[0., 0, 0], [0, 0, 0], [0, 0, 1]], device=DEV))   # halved in place, K[2,2] = 1
        for a, e in zip(out, eager):
            assert torch.equal(a, e)
