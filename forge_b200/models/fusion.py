"""ConvGRU_3D -- multi-view voxel fusion, mirror of reference models/fusion.py (:7-95).

A one-layer 3-D convolutional GRU walks over the (distance-sorted) views; the hidden state starts
from ``fusion_conv(mean over views)`` (reference models/encoder.py:59-63) and the result is
``fusion_norm(h_T)``.  Parameter names match the reference (``cells.{i}.conv_gate/out_gate``,
``fusion_norm``, ``fusion_conv.{0,1,3,4}``) so checkpoints load with ``strict=True``.  The convs stay
cuDNN; inputs arriving channels-last (K2's output layout) are consumed as ``channels_last_3d``.
"""
import torch
import torch.nn as nn

from .. import ops


def _conv_bn_act(cin, cout):
    return [nn.Conv3d(cin, cout, 3, padding=1), nn.BatchNorm3d(cout), nn.LeakyReLU(inplace=True)]


class ConvGRUCell_3D(nn.Module):
    """(u, r) = sigmoid(W_g * [x, h]);  c = tanh(W_o * [x, r h]);  h' = h + u (c - h)   (reference :21-35)"""

    def __init__(self, config, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        width = input_size + hidden_size
        self.conv_gate = nn.Conv3d(width, 2 * hidden_size, 3, padding=1)     # update | reset
        self.out_gate = nn.Conv3d(width, hidden_size, 3, padding=1)          # candidate state
        self.fused_gates = True

    def forward(self, x, prev_state=None):
        h = prev_state
        if h is None:
            h = x.new_zeros((x.shape[0], self.hidden_size) + tuple(x.shape[2:]))
        if self.fused_gates and x.is_cuda:
            # the elementwise chain between the two convs as two launches (forge_gru_gate_* / forge_gru_out_*)
            g = self.conv_gate(torch.cat((x, h), 1))
            return ops.gru_out(self.out_gate(ops.gru_gate(g, h, x)), g, h)
        u, r = torch.sigmoid(self.conv_gate(torch.cat((x, h), 1))).split(self.hidden_size, dim=1)
        c = torch.tanh(self.out_gate(torch.cat((x, h * r), 1)))
        return h * (1 - u) + c * u


class ConvGRU_3D(nn.Module):
    def __init__(self, config, n_layers=1, input_size=16, hidden_size=16):
        super(ConvGRU_3D, self).__init__()
        self.input_size, self.hidden_size, self.n_layers = input_size, hidden_size, n_layers
        widths = [input_size] + [hidden_size] * (n_layers - 1)
        self.cells = nn.ModuleList(ConvGRUCell_3D(config, w, hidden_size) for w in widths)
        self.fusion_norm = nn.BatchNorm3d(hidden_size)
        self.fusion_conv = nn.Sequential(*_conv_bn_act(input_size, input_size), *_conv_bn_act(input_size, input_size))

    # ---- tensor-core path (forge_conv3d_tc): bf16 operands, fp32 accumulation and state; inference only ----
    def tc_eligible(self, x):
        """eval mode, no autograd graph wanted, one layer of 128 -> 128 channels on a grid the kernel tiles (z, y % 4, x % 8)"""
        views = list(x) if isinstance(x, (list, tuple)) else None
        x0 = views[0] if views else x[:, 0]
        b, c, d, h, w = x0.shape
        # weight gradients need the cuDNN path; a graph to the views only (pose refinement) is served by ops.gru_tc
        needs_wgrad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        return (x0.is_cuda and not self.training and not needs_wgrad and self.n_layers == 1 and c % 64 == 0 and self.hidden_size == 128
                and c == self.hidden_size and d % 4 == 0 and h % 4 == 0 and w % 8 == 0)

    def _tc_packs(self):
        tensors = list(self.parameters()) + list(self.buffers())
        key = tuple((p.data_ptr(), p._version) for p in tensors)
        if getattr(self, '_tc_cache', None) is None or self._tc_cache[0] != key:
            def bn_fold(conv, bn):
                s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                return s.float().contiguous(), ((conv.bias - bn.running_mean) * s + bn.bias).float().contiguous()
            with torch.no_grad():
                cell = self.cells[0]
                fc = self.fusion_conv
                s1, b1 = bn_fold(fc[0], fc[1])
                s2, b2 = bn_fold(fc[3], fc[4])
                ns = (self.fusion_norm.weight / torch.sqrt(self.fusion_norm.running_var + self.fusion_norm.eps)).float().contiguous()
                nb = (self.fusion_norm.bias - self.fusion_norm.running_mean * ns).float().contiguous()
                wg, wo = cell.conv_gate.weight, cell.out_gate.weight
                packs = dict(wgT=ops.pack_conv3d_weights(wg.transpose(0, 1).flip(2, 3, 4)),        # transposed convolutions
                             woT=ops.pack_conv3d_weights(wo.transpose(0, 1).flip(2, 3, 4)),        # (backward w.r.t. the input)
                             zero256=torch.zeros(256, dtype=torch.float32, device=wg.device),
                             w1=ops.pack_conv3d_weights(fc[0].weight), s1=s1, b1=b1,
                             w2=ops.pack_conv3d_weights(fc[3].weight), s2=s2, b2=b2,
                             wg=ops.pack_conv3d_weights(cell.conv_gate.weight), bg=cell.conv_gate.bias.float().contiguous(),
                             wo=ops.pack_conv3d_weights(cell.out_gate.weight), bo=cell.out_gate.bias.float().contiguous(),
                             ns=ns, nb=nb)
            self._tc_cache = (key, packs)
        return self._tc_cache[1]

    def forward_tc(self, x):
        """fuse(x) = fusion_norm(GRU over the views, h0 = fusion_conv(mean_t x)) (reference models/encoder.py:59-63,
        models/fusion.py:71-95) with every convolution and the gate arithmetic on the tensor cores: 2 + 2 t launches, no
        cat / sigmoid / tanh / lerp passes, the view sequence is read as bf16 straight from K2's channels-last output."""
        pk = self._tc_packs()
        views = list(x) if isinstance(x, (list, tuple)) else list(x.unbind(dim=1))
        t = len(views)
        if torch.is_grad_enabled() and any(v.requires_grad for v in views):
            # a graph to the views is wanted (weights are constants): h0 through the module's own convolutions (autograd),
            # the recurrence through the differentiable tensor-core op
            mean = views[0]
            for v in views[1:]:
                mean = mean + v
            with torch.autocast('cuda', dtype=torch.bfloat16):
                h0 = self.fusion_conv(mean / t)
            return ops.gru_tc(pk, h0.float(), views)
        xb = [v.to(torch.bfloat16) for v in views]                 # .to keeps K2's channels-last memory order
        xm = views[0].float()
        for v in views[1:]:
            xm = xm + v
        xm = xm / t
        _, a16, _ = ops.conv3d_tc(xm, pk['w1'], 'plain', pk['b1'], scale=pk['s1'], lrelu=True, want_f32=False, want_bf16=True)
        h, h16, _ = ops.conv3d_tc(a16, pk['w2'], 'plain', pk['b2'], scale=pk['s2'], lrelu=True, want_f32=True, want_bf16=True)
        out = None
        for i in range(t):
            x_t = xb[i]
            last = i == t - 1
            u, hr, _ = ops.conv3d_tc(x_t, pk['wg'], 'gate', pk['bg'], h2=h16, h_state=h)
            h, h16, out = ops.conv3d_tc(x_t, pk['wo'], 'out', pk['bo'], h2=hr, h_state=h, u_in=u, scale=pk['ns'],
                                        norm_shift=pk['nb'], want_bf16=not last, want_norm=last)
        return out

    def forward(self, x, hidden=None):
        """x [b,t,c,d,h,w] (view sequence); hidden: optional list of initial states, one per layer -> fusion_norm(h_T)"""
        states = list(hidden) if hidden else [None] * self.n_layers
        seq = list(x) if isinstance(x, (list, tuple)) else x.unbind(dim=1)
        h = None
        for cell, h in zip(self.cells, states):
            outs = []
            for x_t in seq:
                h = cell(x_t, h)
                outs.append(h)
            seq = outs                       # the next layer consumes this layer's state sequence
        return self.fusion_norm(h)
