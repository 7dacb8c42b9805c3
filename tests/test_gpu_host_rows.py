"""GPU tests of the host rows against fixtures generated from the UNMODIFIED reference modules
on CPU (oracle/make_golden.py).  Weights are rebuilt from a seed on both sides (oracle/seeded.py).  The mirrors run
with the fused ConvGRU kernels (gru.cu) switched on, in NCDHW and in channels-last, strict fp32 (TF32 off)."""
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import load_golden                                      # noqa: E402
from forge_b200 import synthetic as syn                               # noqa: E402
from forge_b200.models.encoder import Encoder3D                       # noqa: E402
from forge_b200.models.fusion import ConvGRU_3D                       # noqa: E402
from forge_b200.models.model_single_pose_estimator import FORGE_poseEstimator3D   # noqa: E402
from oracle import seeded                                             # noqa: E402

DEV = 'cuda'
TOL = 1e-4            # relative to max(1, max|reference|)


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _err(a, b):
    b = b.to(a.device)
    return (a - b).abs().max().item() / max(1.0, b.abs().max().item())


@pytest.mark.parametrize("channels_last", [False, True])
def test_encoder3d_lift_fuse_heads_match_reference(channels_last):
    g = load_golden("encoder_small")
    s = g['seed']
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = seeded.load_seeded(Encoder3D(syn.make_config()), s).to(DEV).eval()
    if channels_last:
        m.channels_last_3d_()
    assert all(c.fused_gates for c in m.fusion_feature.cells)
    views = seeded.seeded_tensor(s, 'views', (1, 3, 128, 8, 8, 8)).to(DEV)
    if channels_last:       # what K2 hands over: [B,t,D,H,W,C] memory viewed as [B,t,C,D,H,W]
        views = views.permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4)
    with torch.no_grad():
        feat3d = m.get_feat3D(seeded.seeded_tensor(s, 'img', (1, 3, 32, 32), kind='rand').to(DEV))
        fused = m.fuse(views)
        dens = m.get_density3D(fused)
        rfeat = m.get_render_features(fused)
        cell = m.fusion_feature.cells[0](views[:, 0], seeded.seeded_tensor(s, 'h0', (1, 128, 8, 8, 8)).to(DEV))
    for name, got in (("feat3d", feat3d), ("fused", fused), ("dens", dens), ("rfeat", rfeat), ("cell", cell)):
        assert got.shape == g[name].shape, name
        assert _err(got, g[name]) <= TOL, (name, _err(got, g[name]))


def test_convgru_train_mode_matches_reference():
    """batch-statistics BN + backward through the fused gate kernels"""
    g = load_golden("convgru_train")
    m = seeded.load_seeded(ConvGRU_3D(syn.make_config(), n_layers=1, input_size=16, hidden_size=16), g['seed']).to(DEV).train()
    x = seeded.seeded_tensor(g['seed'], 'x', (2, 3, 16, 6, 6, 6)).to(DEV)
    out = m(x, [m.fusion_conv(x.mean(dim=1))])
    out.square().sum().backward()
    assert _err(out.detach(), g['out']) <= TOL
    assert _err(m.fusion_norm.running_mean, g['norm_mean']) <= TOL
    assert _err(m.fusion_norm.running_var, g['norm_var']) <= TOL
    assert _err(m.fusion_conv[1].running_mean, g['conv_mean']) <= TOL
    assert _err(m.cells[0].conv_gate.weight.grad, g['g_gate']) <= 5 * TOL
    assert _err(m.fusion_conv[0].weight.grad, g['g_fconv']) <= 5 * TOL


def _pose3d(golden, train_heads):
    g = load_golden(golden)
    cfg = syn.make_config(img_size=256, n_pts_per_ray=32, use_gt_pose=True, parameter='all')
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = FORGE_poseEstimator3D(cfg)
    # the reference model also owns encoder_traj (unused with ground-truth poses); values depend on the key name only
    seeded.load_seeded(m, g['seed'])
    m = m.to(DEV).eval()
    if train_heads:
        m.encoder_3d.fusion_feature.train()
        m.encoder_3d.density_head.train()
        m.encoder_3d.features_head.train()
        m.render.train()
    sample = syn.kubric_batch(2 if train_heads else 1, n_views_all=5, img_size=256, seed=g['seed'])
    with torch.no_grad():
        rgb, mask = m(sample, None, DEV)
    return g, m, rgb, mask


def test_pose_estimator3d_model_matches_reference_values():
    """FORGE_poseEstimator3D.forward vs the reference module's output (models/model_single_pose_estimator.py:26-138):
    lift -> rotate -> 3 fusions -> heads -> 2t renders with the reference's view -> volume order."""
    g, m, rgb, mask = _pose3d("pose3d_model_eval", False)
    assert rgb.shape == (10, 3, 256, 256) and mask.shape == (10, 1, 256, 256)
    # whole-model tolerance: ResNet-50 + 10 ConvGRU steps + heads amplify cuDNN-vs-CPU fp32 differences
    assert _err(seeded.subsample(rgb), g['rgb_sub']) <= 1e-3
    assert _err(seeded.subsample(mask), g['mask_sub']) <= 1e-3
    assert _err(rgb.mean(dim=(1, 2, 3)), g['rgb_mean']) <= 1e-4
    assert _err(mask.mean(dim=(1, 2, 3)), g['mask_mean']) <= 1e-4


def test_pose_estimator3d_model_train_mode_heads_match_reference():
    """train-mode BatchNorm in the heads: the reference runs them on the 2b partial-view volumes and the b all-view volumes
    separately (:110-111, :123-124); statistics and running-stat updates must agree"""
    g, m, rgb, mask = _pose3d("pose3d_model_train", True)
    assert rgb.shape == (20, 3, 256, 256)
    assert _err(seeded.subsample(rgb), g['rgb_sub']) <= 2e-3
    assert _err(seeded.subsample(mask), g['mask_sub']) <= 2e-3
    assert _err(m.encoder_3d.density_head[1].running_mean, g['dens_bn_mean']) <= 1e-4
    assert _err(m.encoder_3d.features_head[4].running_var, g['feat_bn_var']) <= 1e-4
    assert _err(m.render.conv_rgb[1].running_mean, g['rgb_bn_mean']) <= 1e-4


def test_fuse_on_tensor_cores_matches_reference_within_bf16_tolerance():
    """Encoder3D.fuse through forge_conv3d_tc (bf16 operands, fp32 accumulation / state) against the fp32 reference fixture;
    tolerance = bf16 operand rounding through 2 + 2 x 3 convolutions (measured 4e-3 of the output range)."""
    g = load_golden("encoder_small")
    s = g['seed']
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = seeded.load_seeded(Encoder3D(syn.make_config()), s).to(DEV).eval()
    m.channels_last_3d_()
    m.compute_dtype = torch.bfloat16
    views = seeded.seeded_tensor(s, 'views', (1, 3, 128, 8, 8, 8)).to(DEV)
    views = views.permute(0, 1, 3, 4, 5, 2).contiguous().permute(0, 1, 5, 2, 3, 4)
    with torch.no_grad():
        assert m.fusion_feature.tc_eligible(views)
        fused = m.fuse(views)
    assert fused.shape == g['fused'].shape
    assert _err(fused, g['fused']) <= 2e-2
    # with a graph wanted the module falls back to autograd-capable convolutions
    v2 = views.clone().requires_grad_(True)
    assert not m.fusion_feature.tc_eligible(v2)


def test_lift_convolution_on_tensor_cores_matches_reference_within_bf16_tolerance():
    """Encoder3D.get_feat3D with the 64 -> 128 lifting convolution on forge_conv3d_tc (BN folded, LeakyReLU fused) and the
    ResNet trunk under bf16 autocast, against the fp32 reference fixture"""
    g = load_golden("encoder_small")
    s = g['seed']
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = seeded.load_seeded(Encoder3D(syn.make_config()), s).to(DEV).eval()
    m.channels_last_3d_()
    m.compute_dtype = torch.bfloat16
    img = seeded.seeded_tensor(s, 'img', (1, 3, 64, 64), kind='rand').to(DEV)          # 64^2 image -> 8 x 8 feature map (W % 8 == 0)
    with torch.no_grad():
        z = m.feature_extraction(img.contiguous(memory_format=torch.channels_last)).float().contiguous().view(-1, 64, 32, 8, 8)
        assert m._conv1_tc_eligible(z)
        got = m._conv1_tc(z)
        ref = m.conv1(z)                        # fp32 cuDNN on the same trunk output
    assert got.shape == ref.shape
    assert _err(got, ref) <= 2e-2


def test_heads_on_tensor_cores_match_reference_within_bf16_tolerance():
    """features_head / density_head (ConvTranspose3d k4 s2 as a 27-tap GEMM with pixel-shuffle epilogue, merged second
    convolutions, 8 -> 1 density convolution) against the fp32 reference fixture and against the fp32 module on a larger grid"""
    g = load_golden("encoder_small")
    s = g['seed']
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = seeded.load_seeded(Encoder3D(syn.make_config()), s).to(DEV).eval()
    m.compute_dtype = torch.bfloat16
    fused = g['fused'].to(DEV)
    with torch.no_grad():
        assert m.heads_tc_eligible(fused)
        rfeat, dens = m.get_render_volumes(fused)
    assert rfeat.shape == g['rfeat'].shape and dens.shape == g['dens'].shape
    assert _err(rfeat, g['rfeat']) <= 2e-2
    assert _err(dens, g['dens']) <= 2e-2
    assert (dens >= 0).all()
    x = torch.randn(2, 128, 16, 12, 24, device=DEV)
    with torch.no_grad():
        rf, dn = m.get_render_volumes(x)
        m.compute_dtype = None
        rf0, dn0 = m.get_render_features(x), m.get_density3D(x)
    assert rf.shape == rf0.shape and dn.shape == dn0.shape
    assert _err(rf, rf0) <= 2e-2 and _err(dn, dn0) <= 2e-2
