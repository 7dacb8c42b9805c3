"""GPU parity tests (backward): gradients of the CUDA path vs autograd through the fp64 oracle.

Tolerance: max-abs gradient error <= 2e-3 of the max-abs reference gradient (fp32 accumulation of
~1e3-1e4 atomically added terms per element; typical observed error is ~1e-5).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from forge_b200 import ops, synthetic as syn                        # noqa: E402
from forge_b200.models.volume_render import VolRender                # noqa: E402
from forge_b200.models.rotate import Rotate_world                    # noqa: E402
from oracle import closed_form as cf                                 # noqa: E402

DEV = 'cuda'
RTOL = 2e-3


def _rel(a, b):
    return (a.double().cpu() - b.double().cpu()).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _render_case(n_obj, n_views, img, vol, P, dense, seed):
    inp = syn.render_inputs(n_obj, n_views, img, vol, seed=seed, dense=dense)
    S = img // 2
    torch.manual_seed(seed)
    g_out = torch.randn(n_obj * n_views, S, S, 16)
    g_sil = torch.randn(n_obj * n_views, S, S)
    g_dep = torch.randn(n_obj * n_views, S, S)
    return inp, S, g_out, g_sil, g_dep


def _oracle_grads(inp, S, P, g_out, g_sil, g_dep, img):
    idx = inp['view2vol'].long()
    feat = inp['feat'].double().requires_grad_(True)
    dens = inp['dens'].double().requires_grad_(True)
    R = inp['R'].double().requires_grad_(True)
    T = inp['T'].double().requires_grad_(True)
    K = inp['K'].double().requires_grad_(True)
    Kh = K / 2.0
    Kh = torch.cat([Kh[:, :2], torch.tensor([0, 0, 1.0], dtype=torch.float64).expand(len(idx), 1, 3)], dim=1)
    f, o, d = cf.raymarch(R, T, Kh, feat[idx], dens[idx], S, P, 0.5, 2.0, 1.0)
    loss = (f * g_out.double()).sum() + (o * g_sil.double()).sum() + (d * g_dep.double()).sum()
    loss.backward()
    return feat.grad, dens.grad, R.grad, T.grad, K.grad


@pytest.mark.parametrize("dense", [False, True])
def test_raymarch_gradients_vs_oracle(dense):
    img, vol, P = 48, 12, 20
    inp, S, g_out, g_sil, g_dep = _render_case(2, 3, img, vol, P, dense, seed=3)
    gf_ref, gd_ref, gR_ref, gT_ref, gK_ref = _oracle_grads(inp, S, P, g_out, g_sil, g_dep, img)

    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV)
    feat = inp['feat'].to(DEV).requires_grad_(True)
    dens = inp['dens'].to(DEV).requires_grad_(True)
    R = inp['R'].to(DEV).requires_grad_(True)
    T = inp['T'].to(DEV).requires_grad_(True)
    K0 = inp['K'].to(DEV).requires_grad_(True)
    cam = dict(R=R, T=T, K=K0 * 1.0)      # non-leaf copy: forward halves it in place
    f, o, d, _, _, _ = m.render_features(cam, feat, dens, True, view2vol=inp['view2vol'])
    loss = (f * g_out.to(DEV)).sum() + (o * g_sil.to(DEV)).sum() + (d * g_dep.to(DEV)).sum()
    loss.backward()
    assert _rel(feat.grad, gf_ref) <= RTOL
    assert _rel(dens.grad, gd_ref) <= RTOL
    assert _rel(R.grad, gR_ref) <= RTOL
    assert _rel(T.grad, gT_ref) <= RTOL
    assert _rel(K0.grad[:, :2], gK_ref[:, :2]) <= RTOL


@pytest.mark.parametrize("n_obj,n_views,img,vol,P,dense", [
    (1, 5, 128, 32, 32, False),       # cfg-1 sizes: 4x4-ray patches spread over ~6-7 base voxels (the merged scatter's common case)
    (1, 2, 70, 20, 33, True),         # partial tiles in both directions, odd sample count, sigma > 1
    (1, 2, 256, 16, 24, False),       # 128^2 rays over a 16^3 volume: whole warps share one base voxel (group size 16)
])
def test_raymarch_gradients_vs_oracle_merge_patterns(n_obj, n_views, img, vol, P, dense):
    """K1 backward pre-reduces, per warp and sample, the feature-gradient contributions of rays that share a base voxel
    (raymarch.cu, kMerge) and accumulates the density gradient as quads: group sizes from 1 to 16 against the fp64 oracle."""
    inp, S, g_out, g_sil, g_dep = _render_case(n_obj, n_views, img, vol, P, dense, seed=11)
    gf_ref, gd_ref, gR_ref, gT_ref, _ = _oracle_grads(inp, S, P, g_out, g_sil, g_dep, img)
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV)
    feat = inp['feat'].to(DEV).requires_grad_(True)
    dens = inp['dens'].to(DEV).requires_grad_(True)
    R = inp['R'].to(DEV).requires_grad_(True)
    T = inp['T'].to(DEV).requires_grad_(True)
    cam = dict(R=R, T=T, K=inp['K'].to(DEV) * 1.0)
    f, o, d, _, _, _ = m.render_features(cam, feat, dens, True, view2vol=inp['view2vol'])
    ((f * g_out.to(DEV)).sum() + (o * g_sil.to(DEV)).sum() + (d * g_dep.to(DEV)).sum()).backward()
    assert _rel(feat.grad, gf_ref) <= RTOL
    assert _rel(dens.grad, gd_ref) <= RTOL
    assert _rel(R.grad, gR_ref) <= RTOL
    assert _rel(T.grad, gT_ref) <= RTOL


def test_raymarch_backward_non_cubic_volume_and_missing_rays():
    """D != H != W (quad / padded strides differ per axis) and a camera that misses the volume entirely."""
    img, P = 64, 24
    inp, S, g_out, g_sil, g_dep = _render_case(1, 3, img, 12, P, False, seed=13)
    g = torch.Generator().manual_seed(4)
    inp['feat'] = torch.randn(1, 16, 10, 14, 18, generator=g)
    inp['dens'] = 0.3 * torch.rand(1, 1, 10, 14, 18, generator=g)
    inp['T'][2] = torch.tensor([3.0, 0.0, 1.5])
    gf_ref, gd_ref, _, gT_ref, _ = _oracle_grads(inp, S, P, g_out, g_sil, g_dep, img)
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV)
    feat = inp['feat'].to(DEV).requires_grad_(True)
    dens = inp['dens'].to(DEV).requires_grad_(True)
    T = inp['T'].to(DEV).requires_grad_(True)
    cam = dict(R=inp['R'].to(DEV), T=T, K=inp['K'].to(DEV) * 1.0)
    f, o, d, _, _, _ = m.render_features(cam, feat, dens, True, view2vol=inp['view2vol'])
    ((f * g_out.to(DEV)).sum() + (o * g_sil.to(DEV)).sum() + (d * g_dep.to(DEV)).sum()).backward()
    assert _rel(feat.grad, gf_ref) <= RTOL
    assert _rel(dens.grad, gd_ref) <= RTOL
    assert _rel(T.grad, gT_ref) <= RTOL


def test_raymarch_pose_only_backward_matches_full():
    """Detached volumes (test-time pose refinement, reference kubric_eval.py:405,450-504): the
    volume scatters are skipped and the camera gradient is unchanged."""
    img, vol, P = 48, 12, 20
    inp, S, g_out, g_sil, g_dep = _render_case(1, 4, img, vol, P, False, seed=5)
    _, _, gR_ref, gT_ref, _ = _oracle_grads(inp, S, P, g_out, g_sil, g_dep, img)
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV)
    R = inp['R'].to(DEV).requires_grad_(True)
    T = inp['T'].to(DEV).requires_grad_(True)
    cam = dict(R=R, T=T, K=inp['K'].to(DEV))
    f, o, d, _, _, _ = m.render_features(cam, inp['feat'].to(DEV), inp['dens'].to(DEV), True, view2vol=inp['view2vol'])
    ((f * g_out.to(DEV)).sum() + (o * g_sil.to(DEV)).sum() + (d * g_dep.to(DEV)).sum()).backward()
    assert _rel(R.grad, gR_ref) <= RTOL
    assert _rel(T.grad, gT_ref) <= RTOL


def test_volrender_forward_backward_through_decoder():
    """Full public forward (as-called NCDHW volumes, one per view) -> loss -> grads reach the volumes,
    the decoder weights and the cameras; compared with the oracle pipeline in fp64."""
    from oracle import reference_path as rp
    img, vol, P = 32, 10, 12
    inp, S, _, _, _ = _render_case(1, 2, img, vol, P, False, seed=8)
    idx = inp['view2vol'].long()
    torch.backends.cudnn.allow_tf32 = False
    m = VolRender(syn.make_config(img_size=img, n_pts_per_ray=P)).to(DEV).eval()
    feat = inp['feat'][idx].to(DEV).requires_grad_(True)
    dens = inp['dens'][idx].to(DEV).requires_grad_(True)
    T = inp['T'].to(DEV).requires_grad_(True)
    rgb, sil, dep = m(dict(R=inp['R'].to(DEV), T=T, K=inp['K'].clone().to(DEV)), feat, dens, render_depth=True)
    (rgb.sum() + 2 * sil.sum() + 0.5 * dep.sum()).backward()
    # oracle in fp64 with the same decoder weights
    conv = rp.make_conv_rgb().double().eval()
    conv.load_state_dict({k: v.double().cpu() for k, v in m.conv_rgb.state_dict().items()})
    ren = rp.make_renderer(img, P, 0.5, 2.0)
    featd = inp['feat'][idx].double().requires_grad_(True)
    densd = inp['dens'][idx].double().requires_grad_(True)
    Td = inp['T'].double().requires_grad_(True)
    Kh = inp['K'].double() / 2
    Kh[:, 2, 2] = 1
    f, o, d = cf.raymarch(inp['R'].double(), Td, Kh, featd, densd, S, P, 0.5, 2.0, 1.0)
    rgb_o = torch.relu(conv(f.permute(0, 3, 1, 2)))
    sil_o = torch.nn.functional.interpolate(o.unsqueeze(1), size=[img] * 2, mode='bilinear')
    dep_o = torch.nn.functional.interpolate(d.unsqueeze(1), size=[img] * 2, mode='bilinear')
    (rgb_o.sum() + 2 * sil_o.sum() + 0.5 * dep_o.sum()).backward()
    assert (rgb.detach().cpu().double() - rgb_o).abs().max().item() <= 1e-4
    assert _rel(feat.grad, featd.grad) <= RTOL
    assert _rel(dens.grad, densd.grad) <= RTOL
    assert _rel(T.grad, Td.grad) <= RTOL
    assert _rel(m.conv_rgb[0].weight.grad, conv[0].weight.grad) <= RTOL


@pytest.mark.parametrize("B,t,C,n", [(2, 3, 8, 12), (1, 2, 128, 16), (1, 3, 6, 10)])
def test_rotate_gradients_vs_oracle(B, t, C, n):
    vox, poses = syn.rotate_inputs(B, t, C, n, seed=9)
    torch.manual_seed(1)
    g = torch.randn(B, t, C, n, n, n)
    voxd = vox.double().requires_grad_(True)
    posd = poses.double().requires_grad_(True)
    (cf.rotate(voxd, posd) * g.double()).sum().backward()

    m = Rotate_world(syn.make_config()).to(DEV)
    v = vox.to(DEV).requires_grad_(True)
    p = poses.to(DEV).requires_grad_(True)
    out = m(v, p, grid_size=n)
    (out * g.to(DEV)).sum().backward()
    assert _rel(v.grad, voxd.grad) <= RTOL
    assert _rel(p.grad[:, :, :3], posd.grad[:, :, :3]) <= RTOL
    # pose-only
    p2 = poses.to(DEV).requires_grad_(True)
    (m(vox.to(DEV), p2, grid_size=n) * g.to(DEV)).sum().backward()
    assert _rel(p2.grad[:, :, :3], posd.grad[:, :, :3]) <= RTOL


def test_rotate_order_backward():
    from oracle import reference_path as rp
    vox, poses = syn.rotate_inputs(2, 4, 8, 10, seed=10)
    idxs = rp.sequence_from_distance(poses[:, :, :3, 3])
    torch.manual_seed(2)
    g = torch.randn(2, 4, 8, 10, 10, 10)
    voxd = vox.double().requires_grad_(True)
    (rp.chose_selected(cf.rotate(voxd, poses.double()), idxs) * g.double()).sum().backward()
    m = Rotate_world(syn.make_config()).to(DEV)
    v = vox.to(DEV).requires_grad_(True)
    (m(v, poses.to(DEV), grid_size=10, order=idxs) * g.to(DEV)).sum().backward()
    assert _rel(v.grad, voxd.grad) <= RTOL
