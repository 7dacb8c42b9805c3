"""CPU check of the tensor-core decoder's HOST side: the bf16 B-tile weight pack and the flattened implicit-GEMM
index plan of decoder_tc.cu are replayed in torch (descriptor reads emulated with the canonical no-swizzle K-major
layout) and compared with the straightforward convolutions.  No GPU, no library compute calls."""
import struct

import torch
import torch.nn.functional as F

from forge_b200 import ops, synthetic as syn
from forge_b200.models.volume_render import VolRender

TOX, TOY, IN_W, GP = 32, 20, 22, 5
IN_PLANE, L1_PLANE, L2_PLANE = 432, 152, 160
BLK, STRIP, W1_TILE = 128, 13 * 128, 2048
W1_OFF = 0
W2_OFF = W1_OFF + 9 * W1_TILE
W3_OFF = W2_OFF + 10 * STRIP + 8 * BLK
PRM_OFF = W3_OFF + 6 * STRIP + 8 * BLK


def _operand(mem_bf16, off, lbo, sbo, rows):
    """Rows of a K-major un-swizzled UMMA operand (K = 16) read through a descriptor (start, LBO, SBO) in bytes."""
    r = torch.arange(rows).view(-1, 1)
    k = torch.arange(16).view(1, -1)
    idx = (off + (k // 8) * lbo + (r // 8) * sbo + (r % 8) * 16 + (k % 8) * 2) // 2
    return mem_bf16[idx].float()


def _mma(act, a_off, a_lbo, wmem, b_off, b_lbo):
    return _operand(act, a_off, a_lbo, 128, 128) @ _operand(wmem, b_off, b_lbo, 128, 64).t()


def _lrelu(v):
    return torch.where(v > 0, v, 0.01 * v)


def replay_tile(x_nhwc, pack_u8, n, Y0, X0, out):
    """decoder_tc.cu's plan for one 32x20 output tile, MMA by MMA (same descriptors, same epilogue index math)."""
    Sh, Sw = x_nhwc.shape[1:3]
    OH, OW = 2 * Sh, 2 * Sw
    wmem = pack_u8[:PRM_OFF].view(torch.bfloat16)
    prm = pack_u8[PRM_OFF:].view(torch.float32)
    b1, b2, b3 = prm[0:16], prm[16:24], prm[24:27]
    sL1 = torch.zeros(16 * L1_PLANE * 8, dtype=torch.bfloat16)
    sL2 = torch.zeros(8 * L2_PLANE * 8, dtype=torch.bfloat16)     # the input planes alias this region in the kernel
    sIn = torch.zeros(2 * IN_PLANE * 8, dtype=torch.bfloat16)
    for px in range((TOY // 2 + 6) * IN_W):
        r, c = divmod(px, IN_W)
        iy, ix = Y0 // 2 - 3 + r, X0 // 2 - 3 + c
        if 0 <= iy < Sh and 0 <= ix < Sw:
            for half in (0, 1):
                sIn[(half * IN_PLANE + px) * 8:(half * IN_PLANE + px) * 8 + 8] = x_nhwc[n, iy, ix, half * 8:half * 8 + 8].to(torch.bfloat16)
    # layer 1
    for j in range(3):
        acc = torch.zeros(128, 64)
        for t in range(9):
            a, b = divmod(t, 3)
            acc += _mma(sIn, (j * 128 + a * IN_W + b) * 16, IN_PLANE * 16, wmem, W1_OFF + t * W1_TILE, 1024)
        for row in range(128):
            yr, xr = divmod(j * 128 + row, IN_W)
            if yr < TOY // 2 + 4 and xr < TOX // 2 + 4:
                for py in (0, 1):
                    for px in (0, 1):
                        lr, lc = 2 * yr + py, 2 * xr + px
                        oy, ox = Y0 - 4 + lr, X0 - 4 + lc
                        col = (py * 2 + px) * 16
                        v = _lrelu(acc[row, col:col + 16] + b1) if (0 <= oy < OH and 0 <= ox < OW) else torch.zeros(16)
                        rec = ((lc & 7) * 2) * L1_PLANE + lr * GP + (lc >> 3)
                        sL1[rec * 8:rec * 8 + 8] = v[:8].to(torch.bfloat16)
                        sL1[(rec + L1_PLANE) * 8:(rec + L1_PLANE) * 8 + 8] = v[8:].to(torch.bfloat16)
    # layer 2
    acc = torch.zeros(128, 64)
    for ky in range(5):
        for j in range(12):
            acc += _mma(sL1, (((j & 7) * 2) * L1_PLANE + ky * GP + (j >> 3)) * 16, L1_PLANE * 16,
                        wmem, W2_OFF + ky * 2 * STRIP + (12 - j) * BLK, STRIP)
    for row in range(128):
        yr, xg = divmod(row, GP)
        for delta in range(8):
            oy, ox = Y0 - 2 + yr, X0 - 2 + 8 * xg + delta
            v = _lrelu(acc[row, delta * 8:delta * 8 + 8] + b2) if (0 <= oy < OH and 0 <= ox < OW) else torch.zeros(8)
            rec = delta * L2_PLANE + row
            sL2[rec * 8:rec * 8 + 8] = v.to(torch.bfloat16)
    # layer 3
    acc = torch.zeros(128, 64)
    for kp in range(3):
        for j in range(12):
            acc += _mma(sL2, ((j & 7) * L2_PLANE + 2 * kp * GP + (j >> 3)) * 16, GP * 16,
                        wmem, W3_OFF + 2 * kp * STRIP + (12 - j) * BLK, STRIP)
    for row in range(128):
        yr, xg = divmod(row, GP)
        for delta in range(8):
            oy, ox = Y0 + yr, X0 + 8 * xg + delta
            if yr < TOY and xg < TOX // 8 and oy < OH and ox < OW:
                out[n, :, oy, ox] = torch.relu(acc[row, delta * 8:delta * 8 + 3] + b3)


def test_weight_pack_and_index_plan_replay():
    torch.manual_seed(0)
    m = VolRender(syn.make_config(img_size=80, n_pts_per_ray=8))
    for mod in m.conv_rgb:
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    m.eval()
    import forge_b200._lib as L
    pack = ops.pack_decoder_tc_weights(m.conv_rgb)
    assert pack.dtype == torch.uint8 and pack.numel() == L.load().forge_decoder_tc_wpack_bytes() == PRM_OFF + 256 == 47360
    Sh, Sw = 13, 20                                     # 26 x 40 output: 2 x 2 tiles, ragged in both directions
    x = torch.randn(1, Sh, Sw, 16)
    out = torch.full((1, 3, 2 * Sh, 2 * Sw), float('nan'))
    for Y0 in range(0, 2 * Sh, TOY):
        for X0 in range(0, 2 * Sw, TOX):
            replay_tile(x, pack, 0, Y0, X0, out)
    assert torch.isfinite(out).all()

    def r16(t):
        return t.to(torch.bfloat16).float()
    ct, bn1, _, c2, bn2, _, c3 = m.conv_rgb
    with torch.no_grad():
        s1 = bn1.weight / torch.sqrt(bn1.running_var + bn1.eps)
        s2 = bn2.weight / torch.sqrt(bn2.running_var + bn2.eps)
        b1 = (ct.bias - bn1.running_mean) * s1 + bn1.bias
        b2 = (c2.bias - bn2.running_mean) * s2 + bn2.bias
        y = F.conv_transpose2d(r16(x.permute(0, 3, 1, 2)), r16(ct.weight * s1.view(1, -1, 1, 1)), None, stride=2, padding=2)
        y = r16(F.leaky_relu(y + b1.view(1, -1, 1, 1), 0.01))
        y = F.conv2d(y, r16(c2.weight * s2.view(-1, 1, 1, 1)), None, padding=2)
        y = r16(F.leaky_relu(y + b2.view(1, -1, 1, 1), 0.01))
        ref = F.relu(F.conv2d(y, r16(c3.weight), c3.bias, padding=2))
        ref32 = F.relu(m.conv_rgb(x.permute(0, 3, 1, 2)))
    scale = max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() <= 4e-3 * scale
    assert (out - ref32).abs().max().item() <= 6e-2 * scale


def test_backward_weight_pack_reproduces_autograd_input_gradient():
    """pack_decoder_bwd_weights (flipped / transposed filters, folded BN scales) + the sign masks are all the backward kernel
    uses: chaining plain correlations over the pack reproduces autograd's d relu(conv_rgb(x)) / d x on the CPU."""
    torch.manual_seed(1)
    m = VolRender(syn.make_config(img_size=32, n_pts_per_ray=8)).eval()
    for mod in m.conv_rgb:
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    pack = ops.pack_decoder_bwd_weights(m.conv_rgb)
    w3b, w2b, wd = pack[:600].view(5, 5, 3, 8), pack[600:3800].view(5, 5, 8, 16), pack[3800:].view(6, 6, 16, 16)
    x = torch.randn(2, 16, 16, 16, requires_grad=True)
    ct, bn1, _, c2, bn2, _, c3 = m.conv_rgb
    pre1 = bn1(ct(x.permute(0, 3, 1, 2)))
    pre2 = bn2(c2(F.leaky_relu(pre1, 0.01)))
    pre3 = c3(F.leaky_relu(pre2, 0.01))
    g = torch.randn_like(pre3)
    F.relu(pre3).backward(g)
    with torch.no_grad():
        g3 = g * (pre3 > 0)
        g2 = F.conv2d(g3, w3b.permute(3, 2, 0, 1), padding=2) * torch.where(pre2 > 0, 1.0, 0.01)
        g1 = F.conv2d(g2, w2b.permute(3, 2, 0, 1), padding=2) * torch.where(pre1 > 0, 1.0, 0.01)
        gx = F.conv2d(g1, wd.permute(3, 2, 0, 1), stride=2, padding=2).permute(0, 2, 3, 1)
    assert (gx - x.grad).abs().max().item() <= 1e-6 * max(1.0, x.grad.abs().max().item())
