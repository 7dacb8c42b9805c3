"""GPU tests of forge_conv3d_tc (tcgen05 implicit-GEMM 3x3x3 convolution with fused ConvGRU epilogues) against torch's
conv3d on the same bf16-rounded operands (fp32 accumulation on both sides): tolerance = fp32 summation-order noise plus
the bf16 rounding of the outputs that are stored as bf16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from forge_b200 import ops                     # noqa: E402

DEV = 'cuda'


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _r(t):
    return t.to(torch.bfloat16).float()


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last_3d)


@pytest.mark.parametrize("B,D,H,W,Cx,Ch,Cout", [(1, 4, 4, 8, 64, 0, 128), (2, 8, 8, 8, 128, 0, 128), (1, 8, 4, 16, 128, 128, 256),
                                               (1, 12, 8, 8, 128, 128, 128), (3, 32, 32, 32, 128, 128, 256)])
def test_plain_conv_matches_torch(B, D, H, W, Cx, Ch, Cout):
    g = torch.Generator().manual_seed(B * 1000 + D + Cout)
    x = torch.randn(B, Cx, D, H, W, generator=g).to(DEV)
    h = torch.randn(B, Ch, D, H, W, generator=g).to(DEV) if Ch else None
    w = (torch.randn(Cout, Cx + Ch, 3, 3, 3, generator=g) / (27 * (Cx + Ch)) ** 0.5).to(DEV)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
    shift = torch.randn(Cout, generator=g).to(DEV)
    y32, y16, _ = ops.conv3d_tc(_cl(x), ops.pack_conv3d_weights(w), 'plain', shift, h2=None if h is None else _cl(h),
                                scale=scale, lrelu=True, want_f32=True, want_bf16=True)
    xin = _r(x) if h is None else torch.cat([_r(x), _r(h)], dim=1)
    ref = F.leaky_relu(F.conv3d(xin, _r(w), padding=1) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1), 0.01)
    assert y32.shape == ref.shape
    assert (y32 - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())
    assert (y16.float() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())


def test_batch_strided_view_source():
    """x_t as K2 hands it over: one view of a [B, t, D, H, W, C] sequence (batch stride t * D*H*W*C)"""
    g = torch.Generator().manual_seed(5)
    B, t, C, n = 2, 3, 128, 8
    seq = torch.randn(B, t, n, n, n, C, generator=g).to(DEV).to(torch.bfloat16)
    x = seq.permute(0, 1, 5, 2, 3, 4)[:, 1]
    w = (torch.randn(128, C, 3, 3, 3, generator=g) / (27 * C) ** 0.5).to(DEV)
    shift = torch.zeros(128, device=DEV)
    y32, _, _ = ops.conv3d_tc(x, ops.pack_conv3d_weights(w), 'plain', shift)
    ref = F.conv3d(x.float(), _r(w), padding=1)
    assert (y32 - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_gru_gate_and_out_epilogues_match_op_chain():
    """reference models/fusion.py:29-35 with bf16-rounded conv operands"""
    g = torch.Generator().manual_seed(7)
    B, C, n = 2, 128, 8
    x = torch.randn(B, C, n, n, n, generator=g).to(DEV)
    h = torch.randn(B, C, n, n, n, generator=g).to(DEV)
    wg = (torch.randn(2 * C, 2 * C, 3, 3, 3, generator=g) / (27 * 2 * C) ** 0.5).to(DEV)
    wo = (torch.randn(C, 2 * C, 3, 3, 3, generator=g) / (27 * 2 * C) ** 0.5).to(DEV)
    bg, bo = torch.randn(2 * C, generator=g).to(DEV) * 0.1, torch.randn(C, generator=g).to(DEV) * 0.1
    ns, nb = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    hcl = _cl(h)
    u, hr, _ = ops.conv3d_tc(_cl(x), ops.pack_conv3d_weights(wg), 'gate', bg, h2=hcl, h_state=hcl)
    hn, hn16, hnorm = ops.conv3d_tc(_cl(x), ops.pack_conv3d_weights(wo), 'out', bo, h2=hr, h_state=hcl, u_in=u, scale=ns,
                                    norm_shift=nb, want_bf16=True, want_norm=True)
    # the op chain on the same rounded operands
    gates = torch.sigmoid(F.conv3d(torch.cat([_r(x), _r(h)], 1), _r(wg), bg, padding=1))
    u_ref, r_ref = gates.split(C, dim=1)
    hr_ref = h * r_ref
    c_ref = torch.tanh(F.conv3d(torch.cat([_r(x), _r(hr_ref)], 1), _r(wo), bo, padding=1))
    hn_ref = h * (1 - u_ref) + c_ref * u_ref
    assert (u - u_ref).abs().max().item() <= 1e-4
    assert (hr.float() - hr_ref).abs().max().item() <= 2e-2 * max(1.0, hr_ref.abs().max().item())
    # hr is rounded to bf16 inside the kernel: feed the kernel's own hr to the reference for a tight comparison of 'out'
    c2 = torch.tanh(F.conv3d(torch.cat([_r(x), hr.float()], 1), _r(wo), bo, padding=1))
    hn2 = h * (1 - u) + c2 * u
    assert (hn - hn2).abs().max().item() <= 2e-4 * max(1.0, hn2.abs().max().item())
    assert (hn - hn_ref).abs().max().item() <= 2e-2 * max(1.0, hn_ref.abs().max().item())
    assert (hn16.float() - hn).abs().max().item() <= 1e-2 * max(1.0, hn.abs().max().item())
    assert (hnorm - (hn * ns.view(1, -1, 1, 1, 1) + nb.view(1, -1, 1, 1, 1))).abs().max().item() <= 1e-4 * max(1.0, hn.abs().max().item() * 2)


def test_differentiable_tensor_core_gru_matches_fp32_autograd():
    """ConvGRU_3D over 3 views with constant weights: output and the gradients to the views / initial state through
    ops.gru_tc (tensor-core forward, transposed convolutions backward) against the fp32 module under autograd;
    tolerance = bf16 operand rounding."""
    from forge_b200 import synthetic as syn
    from forge_b200.models.fusion import ConvGRU_3D
    torch.manual_seed(11)
    B, t, C, n = 1, 3, 128, 8
    gru = ConvGRU_3D(syn.make_config(), n_layers=1, input_size=C, hidden_size=C).to(DEV).eval()
    for p in gru.parameters():
        p.requires_grad_(False)
    gru.fusion_norm.running_var.uniform_(0.5, 1.5)
    gru.fusion_norm.weight.data.uniform_(0.5, 1.5)
    xs = [torch.randn(B, n, n, n, C, device=DEV).permute(0, 4, 1, 2, 3).requires_grad_(True) for _ in range(t)]
    h0 = torch.randn(B, n, n, n, C, device=DEV).permute(0, 4, 1, 2, 3).requires_grad_(True)
    wgt = torch.randn(B, C, n, n, n, device=DEV)
    # fp32 reference: the module's recurrence (cuDNN convolutions + fused gate kernels)
    ref = gru(xs, [h0])
    (ref * wgt).sum().backward()
    g_ref = [x.grad.clone() for x in xs] + [h0.grad.clone()]
    for x in xs + [h0]:
        x.grad = None
    out = ops.gru_tc(gru._tc_packs(), h0, xs)
    (out * wgt).sum().backward()
    g_tc = [x.grad for x in xs] + [h0.grad]
    assert (out - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    for a, b in zip(g_tc, g_ref):
        assert a is not None and torch.isfinite(a).all()
        assert (a - b).abs().max().item() <= 4e-2 * max(1e-6, b.abs().max().item())
        # direction: the gradients are strongly correlated, not merely small
        assert torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item() > 0.999
