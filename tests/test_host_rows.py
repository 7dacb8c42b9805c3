"""CPU tests of the host rows against fixtures generated from the UNMODIFIED reference modules
(oracle/make_golden.py: models/encoder.py:46-63, models/fusion.py:71-95, utils/geo_utils.py:140-207).  The weights are
rebuilt on both sides from a seed (oracle/seeded.py); on CPU the mirrors run the same ATen ops as the reference, so the
tolerance is fp32 noise."""
import warnings

import pytest
import torch

from conftest import load_golden
from forge_b200 import synthetic as syn
from forge_b200.models.encoder import Encoder3D
from forge_b200.models.fusion import ConvGRU_3D
from forge_b200.models.model import _mat2quat
from oracle import seeded


def _close(a, b, tol):
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_mat2quat_matches_reference_including_half_turns():
    g = load_golden("mat2quat")
    q = _mat2quat(g['T'])
    assert torch.isfinite(q).all()
    assert torch.equal(q, g['quat7'])          # same branch selection, same arithmetic


@pytest.fixture(scope="module")
def encoder():
    g = load_golden("encoder_small")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Encoder3D(syn.make_config())
    return seeded.load_seeded(m, g['seed']).eval(), g


def test_encoder3d_state_dict_is_reference_compatible(encoder):
    """keys and shapes of the mirror equal the reference's (tests/golden/reference_state_dict_shapes.json was dumped from
    the reference model by oracle/dump_reference_keys.py)"""
    import json
    import os
    m, _ = encoder
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_state_dict_shapes.json")) as fh:
        ref = json.load(fh)["FORGE"]
    want = {k[len("encoder_3d."):]: list(v) for k, v in ref.items() if k.startswith("encoder_3d.")}
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == want


def test_encoder3d_lift_fuse_heads_match_reference_cpu(encoder):
    m, g = encoder
    s = g['seed']
    with torch.no_grad():
        feat3d = m.get_feat3D(seeded.seeded_tensor(s, 'img', (1, 3, 32, 32), kind='rand'))
        views = seeded.seeded_tensor(s, 'views', (1, 3, 128, 8, 8, 8))
        fused = m.fuse(views)
        dens = m.get_density3D(fused)
        rfeat = m.get_render_features(fused)
        cell = m.fusion_feature.cells[0](views[:, 0], seeded.seeded_tensor(s, 'h0', (1, 128, 8, 8, 8)))
    for name, got in (("feat3d", feat3d), ("fused", fused), ("dens", dens), ("rfeat", rfeat), ("cell", cell)):
        assert got.shape == g[name].shape, name
        assert _close(got, g[name], 2e-5), name


def test_convgru_train_mode_matches_reference_cpu():
    g = load_golden("convgru_train")
    m = seeded.load_seeded(ConvGRU_3D(syn.make_config(), n_layers=1, input_size=16, hidden_size=16), g['seed']).train()
    x = seeded.seeded_tensor(g['seed'], 'x', (2, 3, 16, 6, 6, 6))
    out = m(x, [m.fusion_conv(x.mean(dim=1))])
    out.square().sum().backward()
    assert _close(out.detach(), g['out'], 2e-5)
    assert _close(m.fusion_norm.running_mean, g['norm_mean'], 1e-5)
    assert _close(m.fusion_norm.running_var, g['norm_var'], 1e-5)
    assert _close(m.fusion_conv[1].running_mean, g['conv_mean'], 1e-5)
    assert _close(m.cells[0].conv_gate.weight.grad, g['g_gate'], 1e-4)
    assert _close(m.fusion_conv[0].weight.grad, g['g_fconv'], 1e-4)
