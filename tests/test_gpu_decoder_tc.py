"""GPU parity tests of the tensor-core (bf16, tcgen05) decoder and of the UMMA descriptor contract it relies on.

Tolerances (bf16 path, BASELINE.json configs[2]):
  * vs an fp32 emulation that rounds to bf16 at exactly the kernel's rounding points (input, weights, the two
    hidden activations; BN scales folded into the weights before rounding): max-abs <= 4e-3 * max(1, |ref|max) -- only accumulation order and rare rounding-boundary
    flips of a hidden activation differ;
  * vs the fp32 module (what the reference computes): max-abs <= 6e-2 * max(1, |ref|max), mean-abs <= 6e-3.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from forge_b200 import ops, synthetic as syn            # noqa: E402
from forge_b200.models.volume_render import VolRender  # noqa: E402

DEV = 'cuda'
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _probe_expected(img_bf16, off, lbo, sbo, rows):
    """M[r][k] = image[(off + (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2) / 2] (the no-swizzle K-major
    canonical layout of cute::UMMA::make_umma_desc<Major::K>, LayoutType::INTERLEAVE)."""
    r = torch.arange(rows).view(-1, 1)
    k = torch.arange(16).view(1, -1)
    idx = (off + (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2) // 2
    return img_bf16[idx].float()


PROBE_CASES = [
    # (a_off, a_lbo, a_sbo)                    what the decoder needs
    (0, 2048, 128),                           # aligned start, K halves 2 KiB apart
    (16, 2048, 128),                          # start shifted by one record (16 B): a filter tap
    (16 * 45, 6912, 128),                     # layer 1: tap (2, 1) on pitch 22, channel-chunk planes 432 records apart
    (16 * (6 * 152 + 21), 2432, 128),         # layer 2: phase plane 3, ky = 4, group shift 1, chunk planes 152 records apart
    (16 * (5 * 160 + 11), 80, 128),           # layer 3: second K half = one row down (LBO = 5 records = 80 B)
    (16 * 164, 16, 128),                      # K halves overlapping by all but one record (LBO = 16 B)
    (128 * 5, 1664, 128),                     # a weight-strip window: start at block 5, next strip 13 blocks on
]


@pytest.mark.parametrize("a_off,a_lbo,a_sbo", PROBE_CASES)
def test_umma_descriptor_contract(a_off, a_lbo, a_sbo):
    torch.manual_seed(a_off + a_lbo)
    nbytes = 48 * 1024
    img = (torch.randint(-8, 9, (nbytes // 2,)).float() / 4).to(torch.bfloat16)       # exactly representable, exact sums
    b_off = 40 * 1024
    out = ops.umma_probe(img.view(torch.uint8).to(DEV), a_off, a_lbo, a_sbo, b_off, 256, 128).cpu()
    A = _probe_expected(img, a_off, a_lbo, a_sbo, 128)
    B = _probe_expected(img, b_off, 256, 128, 16)
    assert torch.equal(out, A @ B.t())


def _randomise_bn(m):
    for mod in m.conv_rgb:
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)


def _bf16_round(t):
    return t.to(torch.bfloat16).to(t.dtype)


def emulate_bf16_decoder(conv_rgb, x_nhwc):
    """fp64 evaluation with bf16 rounding at the kernel's rounding points."""
    ct, bn1, _, c2, bn2, _, c3 = conv_rgb
    d = torch.float64
    s1 = (bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)).to(d)
    s2 = (bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)).to(d)
    b1 = (ct.bias.to(d) - bn1.running_mean.to(d)) * s1 + bn1.bias.to(d)
    b2 = (c2.bias.to(d) - bn2.running_mean.to(d)) * s2 + bn2.bias.to(d)
    x = _bf16_round(x_nhwc.permute(0, 3, 1, 2).float()).to(d)
    y = F.conv_transpose2d(x, _bf16_round(ct.weight.float() * s1.float().view(1, -1, 1, 1)).to(d), None, stride=2, padding=2)
    y = F.leaky_relu(y + b1.view(1, -1, 1, 1), 0.01)
    y = _bf16_round(y.float()).to(d)
    y = F.conv2d(y, _bf16_round(c2.weight.float() * s2.float().view(-1, 1, 1, 1)).to(d), None, padding=2)
    y = F.leaky_relu(y + b2.view(1, -1, 1, 1), 0.01)
    y = _bf16_round(y.float()).to(d)
    y = F.conv2d(y, _bf16_round(c3.weight.float()).to(d), c3.bias.to(d), padding=2)
    return F.relu(y)


@pytest.mark.parametrize("N,Sh,Sw", [(2, 16, 16), (1, 21, 21), (1, 8, 40), (3, 64, 64), (1, 5, 3), (20, 128, 128)])
def test_decoder_tc_matches_bf16_emulation_and_fp32_module(N, Sh, Sw):
    torch.manual_seed(N * 1000 + Sh * 10 + Sw)
    m = VolRender(syn.make_config(img_size=2 * Sh, n_pts_per_ray=8)).to(DEV)
    _randomise_bn(m)
    m.eval()
    x = torch.randn(N, Sh, Sw, 16, device=DEV)
    with torch.no_grad():
        out = ops.decoder_tc(x, ops.pack_decoder_tc_weights(m.conv_rgb), m.conv_rgb)
        ref32 = F.relu(m.conv_rgb(x.permute(0, 3, 1, 2)))
        emu = emulate_bf16_decoder(m.conv_rgb, x).float()
    assert out.shape == ref32.shape == (N, 3, 2 * Sh, 2 * Sw)
    assert torch.isfinite(out).all()
    scale = max(1.0, ref32.abs().max().item())
    e_emu = (out - emu).abs().max().item()
    e_32 = (out - ref32).abs()
    print("decoder_tc N=%d %dx%d: vs bf16 emulation %.2e, vs fp32 module max %.2e mean %.2e (scale %.2f)"
          % (N, Sh, Sw, e_emu, e_32.max().item(), e_32.mean().item(), scale))
    assert e_emu <= 4e-3 * scale
    assert e_32.max().item() <= 6e-2 * scale and e_32.mean().item() <= 6e-3 * scale


def test_decoder_tc_small_grid_and_module_switch():
    """Persistent loop with fewer CTAs than tiles gives identical bits; VolRender.decode picks the kernel when
    decoder_dtype is bf16 in eval mode; gradients flow (backward = the module's convs)."""
    torch.manual_seed(5)
    m = VolRender(syn.make_config(img_size=96, n_pts_per_ray=8)).to(DEV)
    _randomise_bn(m)
    m.eval()
    x = torch.randn(3, 48, 48, 16, device=DEV)
    pack = ops.pack_decoder_tc_weights(m.conv_rgb)
    with torch.no_grad():
        full = ops.decoder_tc(x, pack, m.conv_rgb)
        few = ops.decoder_tc(x, pack, m.conv_rgb, max_ctas=5)
        m.decoder_dtype = torch.bfloat16
        via_module = m.decode(x)
    assert torch.equal(full, few) and torch.equal(full, via_module)
    xg = x.clone().requires_grad_(True)
    m.decode(xg).sum().backward()
    assert torch.isfinite(xg.grad).all() and xg.grad.abs().sum() > 0


@pytest.mark.parametrize("N,S,tc", [(2, 16, False), (1, 21, False), (3, 64, False), (5, 128, False), (2, 24, True), (1, 13, True)])
def test_decoder_constant_weights_backward_kernel_matches_autograd(N, S, tc):
    """forge_decoder_bwd_data (sign masks from the forward kernel) == autograd through the module's convs when the
    decoder weights are constants; fp32 forward: <= 1e-5 of the gradient scale (same piece-wise linear map, fp32
    accumulation order differs); bf16 tensor-core forward: the kernel differentiates the map the forward kernel
    evaluated (its own sign pattern), checked against the fp32 module's gradient at bf16-level tolerance."""
    torch.manual_seed(N * 10 + S)
    m = VolRender(syn.make_config(img_size=2 * S, n_pts_per_ray=8)).to(DEV)
    _randomise_bn(m)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    if tc:
        m.decoder_dtype = torch.bfloat16
    x = torch.randn(N, S, S, 16, device=DEV)
    g = torch.randn(N, 3, 2 * S, 2 * S, device=DEV)
    xk = x.clone().requires_grad_(True)
    out = m.decode(xk)
    out.backward(g)
    xr = x.clone().requires_grad_(True)
    ref = F.relu(m.conv_rgb(xr.permute(0, 3, 1, 2)))
    ref.backward(g)
    scale = xr.grad.abs().max().item()
    err = (xk.grad - xr.grad).abs()
    print("decoder bwd N=%d S=%d tc=%s: max err %.2e mean %.2e (grad max %.2e)" % (N, S, tc, err.max().item(), err.mean().item(), scale))
    if tc:
        assert err.mean().item() <= 2e-2 * scale       # a few sign flips near zero pre-activations differ from the fp32 map
    else:
        assert err.max().item() <= 1e-5 * max(1.0, scale) + 1e-4 * scale


def test_decoder_backward_falls_back_to_module_convs_for_trainable_weights():
    torch.manual_seed(1)
    m = VolRender(syn.make_config(img_size=64, n_pts_per_ray=8)).to(DEV).eval()
    x = torch.randn(2, 32, 32, 16, device=DEV)
    xk = x.clone().requires_grad_(True)
    m.decode(xk).sum().backward()
    assert m.conv_rgb[0].weight.grad is not None and xk.grad is not None
    xr = x.clone().requires_grad_(True)
    m.zero_grad()
    F.relu(m.conv_rgb(xr.permute(0, 3, 1, 2))).sum().backward()
    assert torch.allclose(xk.grad, xr.grad, atol=1e-4, rtol=1e-4)
