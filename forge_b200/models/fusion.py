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

    def forward(self, x, hidden=None):
        """x [b,t,c,d,h,w] (view sequence); hidden: optional list of initial states, one per layer -> fusion_norm(h_T)"""
        states = list(hidden) if hidden else [None] * self.n_layers
        seq = x.unbind(dim=1)
        h = None
        for cell, h in zip(self.cells, states):
            outs = []
            for x_t in seq:
                h = cell(x_t, h)
                outs.append(h)
            seq = outs                       # the next layer consumes this layer's state sequence
        return self.fusion_norm(h)
