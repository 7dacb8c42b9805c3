"""CPU tests of round-2 host logic: weight packing for forge_conv3d_tc, the view-0-first order and the aliased job table of
Rotate_world.forward_views, the seeded-weights recipe, and the compat shims the reference's scripts import."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pack_conv3d_weights_layout():
    from forge_b200.ops import pack_conv3d_weights
    Cout, Cin = 128, 192
    w = torch.arange(Cout * Cin * 27, dtype=torch.float32).reshape(Cout, Cin, 3, 3, 3) % 251      # exactly representable in bf16
    p = pack_conv3d_weights(w)
    assert p.shape == (27, Cin // 64, Cout, 64) and p.dtype == torch.bfloat16
    for (co, ci, dz, dy, dx) in [(0, 0, 0, 0, 0), (5, 70, 1, 2, 0), (127, 191, 2, 2, 2), (64, 128, 0, 1, 2)]:
        assert p[(dz * 3 + dy) * 3 + dx, ci // 64, co, ci % 64].item() == w[co, ci, dz, dy, dx].item()
    with pytest.raises(ValueError):
        pack_conv3d_weights(torch.zeros(8, 48, 3, 3, 3))


def test_transposed_conv_pack_is_the_input_gradient():
    """the weight pack used for the backward pass (transpose(0,1).flip) turns conv3d into its own input-gradient operator"""
    import torch.nn.functional as F
    torch.manual_seed(0)
    w = torch.randn(6, 4, 3, 3, 3)
    x = torch.randn(1, 4, 5, 5, 5, requires_grad=True)
    gy = torch.randn(1, 6, 5, 5, 5)
    (F.conv3d(x, w, padding=1) * gy).sum().backward()
    gx = F.conv3d(gy, w.transpose(0, 1).flip(2, 3, 4), padding=1)
    assert torch.allclose(gx, x.grad, atol=1e-5)


def test_view0_first_and_aliased_jobs():
    from forge_b200.models.model import _view0_first, sequence_from_distance
    from forge_b200.models.rotate import Rotate_world
    from forge_b200 import synthetic as syn
    idxs = torch.tensor([[0, 3, 1, 2], [2, 0, 1, 3], [1, 2, 3, 0]])
    out = _view0_first(idxs)
    assert out.tolist() == [[0, 3, 1, 2], [0, 2, 1, 3], [0, 1, 2, 3]]
    # sequence_from_distance already puts view 0 first for distinct camera centres: the helper is then the identity
    _, poses = syn.rotate_inputs(3, 5, 1, 2, seed=4)
    order = sequence_from_distance(poses[:, :, :3, 3])
    assert torch.equal(_view0_first(order), order)
    rot = Rotate_world(syn.make_config())
    jobs = rot._jobs_aliased(3, 5, torch.device('cpu'), order)
    assert jobs.shape == (12, 3) and (jobs[:, 2] == 0).all()
    for b in range(3):
        for v in range(1, 5):
            row = jobs[b * 4 + (v - 1)]
            assert row[0].item() == b * 5 + v                                  # source: view v of object b
            slot = (order[b] == v).nonzero().item()
            assert row[1].item() == b * 4 + slot - 1                           # destination: its sorted slot, minus the aliased slot 0
    assert sorted(jobs[:, 1].tolist()) == list(range(12))                      # every destination written exactly once


def test_seeded_state_dict_depends_on_key_names_only():
    from oracle import seeded
    a = torch.nn.Sequential(torch.nn.Conv3d(4, 8, 3), torch.nn.BatchNorm3d(8))
    b = torch.nn.Sequential(torch.nn.Conv3d(4, 8, 3), torch.nn.BatchNorm3d(8), torch.nn.ConvTranspose3d(8, 4, 4, stride=2))
    sa, sb = seeded.seeded_state_dict(a, 7), seeded.seeded_state_dict(b, 7)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert (sa['1.running_var'] > 0).all()
    assert not torch.equal(sa['0.weight'], seeded.seeded_state_dict(a, 8)['0.weight'])


@pytest.fixture()
def compat_path():
    p = os.path.join(ROOT, "compat")
    sys.path.insert(0, p)
    yield p
    sys.path.remove(p)
    for m in [m for m in sys.modules if m.split('.')[0] in ("skimage", "imageio", "matplotlib", "mpl_toolkits", "lpips")]:
        sys.modules.pop(m, None)


def test_compat_metrics_and_image_io(compat_path, tmp_path):
    import importlib
    metrics = importlib.import_module("skimage.metrics")
    rng = np.random.default_rng(0)
    a = rng.random((32, 32, 3))
    assert metrics.structural_similarity(a, a, multichannel=True, data_range=1) == pytest.approx(1.0)
    b = np.clip(a + 0.1, 0, 1)
    mse = np.mean((a - b) ** 2)
    assert metrics.peak_signal_noise_ratio(a, b, data_range=1) == pytest.approx(10 * np.log10(1.0 / mse))
    s = metrics.structural_similarity(a, rng.random((32, 32, 3)), multichannel=True, data_range=1)
    assert -0.2 < s < 0.3                                   # unrelated noise images are dissimilar
    imageio = importlib.import_module("imageio")
    frames = [np.uint8(255 * rng.random((8, 8, 3))) for _ in range(3)]
    imageio.mimsave(str(tmp_path / "x.gif"), frames, 'GIF', duration=0.1)
    from PIL import Image
    with Image.open(tmp_path / "x.gif") as im:
        assert im.n_frames == 3
    importlib.import_module("matplotlib.pyplot").figure().add_subplot(111).plot([1, 2])     # absorbed, no error
    importlib.import_module("mpl_toolkits.mplot3d")


def test_synthetic_kubric_dataset_item(compat_path):
    import importlib
    from types import SimpleNamespace as NS
    kubric = importlib.import_module("dataset.kubric")
    cfg = NS(dataset=NS(img_size=64, train_all_frame=True, num_frame=5), test=NS(compute_metric=True), render=NS(camera_z=1.5))
    d = kubric.Kubric(cfg, split='train')
    s = d[0]
    assert s['images'].shape == (10, 3, 64, 64) and s['K_cv2'].shape == (10, 3, 3)
    assert s['cam_poses_rel_cv2'].shape == (10, 4, 4) and s['cam_poses_rel_every2_cv2'].shape == (9, 4, 4)
    assert torch.allclose(s['cam_extrinsics_cv2_canonicalized'][0], d.get_canonical_extrinsics_cv2(), atol=1e-5)
    assert torch.allclose(d.get_canonical_pose_cv2() @ d.get_canonical_extrinsics_cv2(), torch.eye(4), atol=1e-6)
    sys.modules.pop("dataset.kubric", None)
    sys.modules.pop("dataset", None)
