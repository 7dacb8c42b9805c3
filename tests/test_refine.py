"""prepare_for_pose_refinement: freezes weights / converts layouts without touching values (CPU, no compute)."""
import warnings

import torch

from forge_b200 import synthetic as syn
from forge_b200.models.model import FORGE
from forge_b200.refine import prepare_for_pose_refinement


def test_prepare_for_pose_refinement_freezes_and_preserves_state():
    cfg = syn.make_config(img_size=64, n_pts_per_ray=8, use_gt_pose=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = FORGE(cfg).eval()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    out = prepare_for_pose_refinement(model, fusion_dtype=torch.bfloat16, decoder_dtype=torch.bfloat16)
    assert out is model
    assert all(not p.requires_grad for p in model.parameters())
    assert model.encoder_3d.compute_dtype == torch.bfloat16 and model.render.decoder_dtype == torch.bfloat16
    after = model.state_dict()
    assert before.keys() == after.keys() and all(torch.equal(before[k], after[k]) for k in before)
    w = model.encoder_3d.fusion_feature.cells[0].conv_gate.weight
    assert w.is_contiguous(memory_format=torch.channels_last_3d)
    prepare_for_pose_refinement(model, freeze_weights=False, channels_last=False)
    assert model.encoder_3d.compute_dtype is None and model.render.decoder_dtype is None


def test_gru_operand_layout_logic_accepts_strided_views_without_copies():
    """Host logic of the fused ConvGRU stages: x_t slices of K2's channels-last view sequence and dense NCDHW tensors are passed
    through untouched (only their batch stride is reported); anything else is made dense in the conv output's layout."""
    from forge_b200 import ops
    B, t, C, n = 2, 3, 8, 4
    seq_cl = torch.randn(B, t, n, n, n, C).permute(0, 1, 5, 2, 3, 4)          # [B,t,C,D,H,W], channels-last in memory
    x = seq_cl[:, 1]
    got, bs = ops._gru_operand(x, 1)
    assert got.data_ptr() == x.data_ptr() and bs == t * C * n ** 3
    seq = torch.randn(B, t, C, n, n, n)
    x = seq[:, 2]
    got, bs = ops._gru_operand(x, 0)
    assert got.data_ptr() == x.data_ptr() and bs == t * C * n ** 3
    # NCDHW tensor offered to a channels-last cell: converted, dense batch stride
    got, bs = ops._gru_operand(seq[:, 0], 1)
    assert got.data_ptr() != seq[:, 0].data_ptr() and got.is_contiguous(memory_format=torch.channels_last_3d) and bs == C * n ** 3
    # bf16 state (autocast) is promoted to fp32
    got, _ = ops._gru_operand(torch.randn(1, C, n, n, n).bfloat16(), 0)
    assert got.dtype == torch.float32
    g, cl = ops._gru_layout(torch.randn(B, 2 * C, n, n, n).contiguous(memory_format=torch.channels_last_3d))
    assert cl == 1
    g, cl = ops._gru_layout(torch.randn(B, 2 * C, n, n, n)[:, :, ::2])
    assert cl == 0 and g.is_contiguous()
