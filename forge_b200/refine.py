"""Helpers for the test-time pose optimisation loop (reference kubric_eval.py:412-530, demo.py:115-188).

The reference's scripts keep every network weight at ``requires_grad=True`` while they optimise only the four
relative poses, so each of the 2000-5000 iterations per object also computes (and throws away) the weight gradients of
the ConvGRU fusion, the heads and the decoder -- about a quarter of the iteration on a B200
(``tools/refine_bench.py``).  ``prepare_for_pose_refinement`` removes that work without touching the scripts' loop:

    model.eval(); prepare_for_pose_refinement(model)      # once, before the loop

Nothing here changes a forward result; with frozen weights K1 / K2 / the decoder run their pose-only backward paths
(NULL volume-gradient outputs in the C ABI).
"""
import torch


def prepare_for_pose_refinement(model, freeze_weights=True, channels_last=True, fusion_dtype=None, decoder_dtype=None):
    """model: forge_b200 ``FORGE`` / ``FORGE_poseEstimator3D`` (or any module holding ``encoder_3d`` / ``render``).

    freeze_weights  set requires_grad=False on every parameter (poses are the only leaves of the loop)
    channels_last   store the 3-D conv weights channels_last_3d (K2 emits channels-last volumes; saves cuDNN's
                    NCDHW<->NDHWC conversion kernels around every conv)
    fusion_dtype    None = fp32 like the reference, torch.bfloat16 = autocast for fusion + heads
    decoder_dtype   None = fp32 fused decoder, torch.bfloat16 = tensor-core decoder
    Returns the model."""
    if freeze_weights:
        for p in model.parameters():
            p.requires_grad_(False)
    enc = getattr(model, 'encoder_3d', None)
    if enc is not None:
        if channels_last and hasattr(enc, 'channels_last_3d_'):
            enc.channels_last_3d_()
        if hasattr(enc, 'compute_dtype'):
            enc.compute_dtype = fusion_dtype
    ren = getattr(model, 'render', None)
    if ren is not None and hasattr(ren, 'decoder_dtype'):
        ren.decoder_dtype = decoder_dtype
    return model
