"""GPU tests of the fused ConvGRU cell stages (forge_gru_gate_* / forge_gru_out_*) against the reference's op chain
(models/fusion.py:21-35) evaluated with torch ops on the same cuDNN convolutions."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from forge_b200 import synthetic as syn                       # noqa: E402
from forge_b200.models.fusion import ConvGRUCell_3D, ConvGRU_3D  # noqa: E402

DEV = 'cuda'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _run(cell, x, h, fused, amp):
    cell.fused_gates = fused
    xg, hg = x.clone().requires_grad_(True), h.clone().requires_grad_(True)
    cell.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
        out = cell(xg, hg)
    w = torch.linspace(-1, 1, out.numel(), device=DEV).reshape(out.shape)
    (out.float() * w).sum().backward()
    return out.float(), xg.grad, hg.grad, cell.conv_gate.weight.grad.clone(), cell.out_gate.weight.grad.clone()


@pytest.mark.parametrize("layout", ["ncdhw", "channels_last", "view_slice"])
@pytest.mark.parametrize("amp", [False, True])
def test_fused_gru_cell_matches_op_chain(layout, amp):
    torch.manual_seed(0)
    B, C, n = 2, 16, 6
    cell = ConvGRUCell_3D(None, C, C).to(DEV)
    x = torch.randn(B, C, n, n, n, device=DEV)
    h = torch.randn(B, C, n, n, n, device=DEV)
    if layout == "channels_last":
        cell.to(memory_format=torch.channels_last_3d)
        x = x.contiguous(memory_format=torch.channels_last_3d)
        h = h.contiguous(memory_format=torch.channels_last_3d)
    elif layout == "view_slice":          # x_t as K2 hands it over: a slice of the channels-last [B, t, D, H, W, C] sequence
        cell.to(memory_format=torch.channels_last_3d)
        seq = torch.randn(B, 3, n, n, n, C, device=DEV).permute(0, 1, 5, 2, 3, 4)
        x = seq[:, 1]
        h = h.contiguous(memory_format=torch.channels_last_3d)
    ref = _run(cell, x, h, False, amp)
    got = _run(cell, x, h, True, amp)
    tol = 3e-2 if amp else 1e-5
    for name, a, b in zip(("out", "dx", "dh", "dW_gate", "dW_out"), got, ref):
        scale = max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= tol * scale, name


def test_fused_convgru_sequence_matches_and_odd_sizes():
    """whole ConvGRU_3D (5 views, odd spatial size: scalar kernel path) forward + backward to the views"""
    torch.manual_seed(1)
    B, t, C, n = 1, 5, 8, 5
    gru = ConvGRU_3D(syn.make_config(), n_layers=1, input_size=C, hidden_size=C).to(DEV).eval()
    x = torch.randn(B, t, C, n, n, n, device=DEV)
    outs = []
    for fused in (False, True):
        for cell in gru.cells:
            cell.fused_gates = fused
        xg = x.clone().requires_grad_(True)
        y = gru(xg, [gru.fusion_conv(xg.mean(dim=1))])
        y.square().sum().backward()
        outs.append((y.detach(), xg.grad))
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 1e-5 * max(1.0, outs[0][0].abs().max().item())
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-5 * max(1.0, outs[0][1].abs().max().item())
