"""GPU parity of the two forward formulations of K1 behind the C ABI: forge_raymarch_fwd_tma (TMA-staged bricks in shared
memory, raymarch_tma.cu) against forge_raymarch_fwd_gather (direct L1 gathers, the kernel every oracle / golden test pins).
Index math and sample weights are shared code; the two differ only in fp32 summation order (<= 1e-5 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from forge_b200 import _lib, ops, synthetic as syn                     # noqa: E402
from forge_b200.models.volume_render import VolRender                  # noqa: E402

DEV = 'cuda'


def _run(entry, inp, img, D, P, depth=True, min_depth=0.5, max_depth=2.0):
    cfg = syn.make_config(img_size=img, n_pts_per_ray=P, min_depth=min_depth, max_depth=max_depth)
    m = VolRender(cfg).to(DEV).eval()
    S = img // 2
    N, V = inp['R'].shape[0], inp['feat'].shape[0]
    Kh = inp['K'].clone()
    Kh /= 2.0
    Kh[:, 2, 2] = 1.0
    Dz, Dy, Dx = inp['feat'].shape[2:]
    cam12, _ = ops.camera_prep(inp['R'], inp['T'], Kh, (Dz, Dy, Dx), 1.0)
    fp, dq = ops.pack_volume(inp['feat'], inp['dens'])
    zs = m._depths(torch.device(DEV))
    o = torch.full((N, S, S, 16), float('nan'), device=DEV)
    s = torch.full((N, S, S), float('nan'), device=DEV)
    d = torch.full((N, S, S), float('nan'), device=DEV) if depth else None
    _lib.call(entry, fp.data_ptr(), dq.data_ptr(), inp['view2vol'].data_ptr(), cam12.data_ptr(), zs.data_ptr(), o.data_ptr(),
              s.data_ptr(), None if d is None else d.data_ptr(), N, V, Dz, Dy, Dx, S, S, P,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return o, s, d


def _compare(inp, img, D, P, **kw):
    a = _run("forge_raymarch_fwd_gather", inp, img, D, P, **kw)
    b = _run("forge_raymarch_fwd_tma", inp, img, D, P, **kw)
    for name, x, y in zip(("feat", "sil", "depth"), a, b):
        if x is None:
            continue
        assert torch.isfinite(y).all(), name
        scale = max(1.0, x.abs().max().item())
        assert (x - y).abs().max().item() <= 1e-5 * scale, (name, (x - y).abs().max().item())
    return a


@pytest.mark.parametrize("objects,views,img,D,P,dense", [
    (1, 2, 16, 4, 8, False),          # tiny: the volume is smaller than every TMA box shape, 8 x 8 rays
    (1, 3, 64, 16, 24, False),        # small
    (2, 2, 32, 10, 16, True),         # sigma > 1: sign-alternating transmittance
    (1, 5, 128, 32, 32, False),       # cfg-1
    (1, 2, 70, 20, 33, False),        # 35 x 35 rays: partial tiles in both directions, odd sample count
    (4, 5, 256, 64, 64, False),       # cfg-2
    (2, 5, 256, 64, 64, True),        # cfg-2 geometry, dense volume
    (2, 3, 512, 128, 128, False),     # cfg-4 geometry (BASELINE.json configs[3] sizes per view / volume)
])
def test_tma_matches_gather(objects, views, img, D, P, dense):
    inp = syn.render_inputs(objects, views, img, D, seed=3, dense=dense, device=DEV)
    o, s, _ = _compare(inp, img, D, P)
    assert s.max().item() > 0


def test_tma_without_depth_and_shared_volume():
    inp = syn.render_inputs(1, 4, 64, 24, seed=5, device=DEV)
    _compare(inp, 64, 24, 32, depth=False)


def test_tma_rays_missing_the_volume_and_non_cubic_volume():
    """cameras looking away / far off-centre (whole tiles miss the volume) and a D != H != W volume"""
    inp = syn.render_inputs(1, 4, 64, 16, seed=6, device=DEV)
    inp['T'][1] = torch.tensor([3.0, 0.0, 1.5], device=DEV)          # volume far to the side
    inp['T'][2] = torch.tensor([0.0, 0.0, -1.5], device=DEV)         # volume behind the camera
    g = torch.Generator().manual_seed(9)
    inp['feat'] = torch.randn(1, 16, 12, 20, 28, generator=g).to(DEV)
    inp['dens'] = (0.3 * torch.rand(1, 1, 12, 20, 28, generator=g)).to(DEV)
    _compare(inp, 64, 16, 40)


def test_tma_close_camera_boxes_exceed_a_stage():
    """a camera inside / next to the volume: ray spacing shrinks and slabs get long, the per-sample box still has to fit
    (clipped boxes fall back to direct gathers)"""
    inp = syn.render_inputs(1, 3, 128, 64, seed=7, device=DEV)
    inp['T'][:, 2] = 0.6
    _compare(inp, 128, 64, 96, min_depth=0.05, max_depth=1.2)
    inp['K'][:, 0, 0] *= 0.25          # very wide field of view: 16 x 8 pixels cover tens of voxels
    inp['K'][:, 1, 1] *= 0.25
    _compare(inp, 128, 64, 48, min_depth=0.05, max_depth=1.2)
