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
