"""GPU tests of the camera / pose glue kernels against the differentiable torch statements of the same algebra
(which are themselves checked against the oracle in test_gpu_forward.py / test_oracle.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from forge_b200 import ops, synthetic as syn                                   # noqa: E402
from forge_b200.models.volume_render import camera_to_cam12, origin_projection  # noqa: E402
from forge_b200.models.rotate import Rotate_world                              # noqa: E402

DEV = 'cuda'
torch.backends.cuda.matmul.allow_tf32 = False


def _cams(n, seed):
    inp = syn.render_inputs(1, n, 128, 16, seed=seed)
    Kh = inp['K'].clone() / 2.0
    Kh[:, 2, 2] = 1.0
    g = torch.Generator().manual_seed(seed)
    Kh[:, 0, 2] += torch.randn(n, generator=g)          # off-centre principal points
    Kh[:, 1, 2] += torch.randn(n, generator=g)
    return inp['R'].to(DEV), inp['T'].to(DEV), Kh.to(DEV)


@pytest.mark.parametrize("dhw", [(64, 64, 64), (16, 24, 40)])
def test_camera_prep_matches_torch_algebra_values_and_grads(dhw):
    R, T, Kh = _cams(7, 3)
    Rr, Tr, Kr = (t.clone().requires_grad_(True) for t in (R, T, Kh))
    ref12 = camera_to_cam12(Rr, Tr, Kr, dhw, 1.3)
    refo = origin_projection(Tr, Kr)
    Rk, Tk, Kk = (t.clone().requires_grad_(True) for t in (R, T, Kh))
    cam12, oproj = ops.camera_prep(Rk, Tk, Kk, dhw, 1.3)
    assert torch.allclose(cam12, ref12, rtol=2e-6, atol=1e-6)
    assert torch.allclose(oproj, refo, rtol=2e-6, atol=1e-4)
    w12, wo = torch.randn_like(ref12), torch.randn_like(refo)
    ((ref12 * w12).sum() + (refo * wo).sum()).backward()
    ((cam12 * w12).sum() + (oproj * wo).sum()).backward()
    for a, b in ((Rk.grad, Rr.grad), (Tk.grad, Tr.grad)):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item())
    # only fx, fy, cx, cy carry gradient
    for (i, j) in ((0, 0), (1, 1), (0, 2), (1, 2)):
        assert (Kk.grad[:, i, j] - Kr.grad[:, i, j]).abs().max().item() <= 1e-5 * max(1.0, Kr.grad[:, i, j].abs().max().item())
    assert Kk.grad[:, 2].abs().max().item() == 0 and Kk.grad[:, 0, 1].abs().max().item() == 0


def test_origin_projection_clamps_tz_like_the_reference():
    R, T, Kh = _cams(3, 5)
    T = T.clone()
    T[0, 2] = 0.0
    T[1, 2] = -1e-9
    _, oproj = ops.camera_prep(R, T, Kh, (8, 8, 8), 1.0)
    assert torch.allclose(oproj, origin_projection(T, Kh), rtol=1e-6, atol=1e-3)


@pytest.mark.parametrize("B,t", [(1, 2), (4, 5), (3, 1)])
def test_pose_affine_matches_inverse_matmul_values_and_grads(B, t):
    _, poses = syn.rotate_inputs(B, max(t, 2), 4, 8, seed=B * 10 + t)
    poses = poses[:, :t].contiguous().to(DEV)
    rot = Rotate_world(syn.make_config())
    pr = poses.clone().requires_grad_(True)
    pk = poses.clone().requires_grad_(True)
    A = ops.pose_affine(pk).view(B, t, 3, 4)
    eye = torch.eye(4, device=DEV)[:3]
    assert torch.equal(A[:, 0], eye.expand(B, 3, 4))
    if t == 1:
        return
    ref = rot.get_transformation(pr).view(B, t - 1, 4, 4)[:, :, :3, :]
    assert (A[:, 1:] - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
    w = torch.randn_like(ref)
    (ref * w).sum().backward()
    (A[:, 1:] * w).sum().backward()
    assert (pk.grad - pr.grad).abs().max().item() <= 1e-5 * max(1.0, pr.grad.abs().max().item())


def test_pose_affine_flags_singular_pose():
    poses = torch.eye(4, device=DEV).repeat(1, 2, 1, 1)
    poses[0, 1, 2, 2] = 0.0
    A = ops.pose_affine(poses)
    assert torch.isnan(A[1]).all() and torch.isfinite(A[0]).all()


@pytest.mark.parametrize("M,Sh,Sw", [(3, 16, 16), (2, 5, 9), (1, 1, 1), (1, 1, 7), (20, 128, 128)])
def test_upsample2x_matches_interpolate_values_and_grads(M, Sh, Sw):
    import torch.nn.functional as F
    torch.manual_seed(M + Sh)
    a = torch.randn(M, Sh, Sw, device=DEV)
    b = torch.randn(M, Sh, Sw, device=DEV)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ak, bk = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ra = F.interpolate(ar.unsqueeze(1), size=[2 * Sh, 2 * Sw], mode='bilinear')
    rb = F.interpolate(br.unsqueeze(1), size=[2 * Sh, 2 * Sw], mode='bilinear')
    oa, ob = ops.upsample2x(ak, bk)
    assert oa.shape == ra.shape and (oa - ra).abs().max().item() <= 1e-6 and (ob - rb).abs().max().item() <= 1e-6
    wa, wb = torch.randn_like(ra), torch.randn_like(rb)
    ((ra * wa).sum() + (rb * wb).sum()).backward()
    ((oa * wa).sum() + (ob * wb).sum()).backward()
    assert (ak.grad - ar.grad).abs().max().item() <= 1e-5 and (bk.grad - br.grad).abs().max().item() <= 1e-5
    # single map
    o1, none = ops.upsample2x(a)
    assert none is None and torch.equal(o1, oa.detach())
