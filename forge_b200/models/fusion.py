"""ConvGRU_3D -- multi-view voxel fusion, mirror of reference models/fusion.py (:7-95).

A one-layer 3-D convolutional GRU walks over the (distance-sorted) views; the hidden state starts
from ``fusion_conv(mean over views)`` (reference models/encoder.py:59-63) and the result is
``fusion_norm(h_T)``.  Parameter names match the reference (``cells.{i}.conv_gate/out_gate``,
``fusion_norm``, ``fusion_conv.{0,1,3,4}``) so checkpoints load with ``strict=True``.  The convs stay
cuDNN; inputs arriving channels-last (K2's output layout) are consumed as ``channels_last_3d``.
"""
import torch
import torch.nn as nn


class ConvGRUCell_3D(nn.Module):
    """h' = (1 - u) h + u tanh(W_o [x, r h]),  (u, r) = sigmoid(W_g [x, h])   (reference :21-35)"""

    def __init__(self, config, input_size, hidden_size):
        super().__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.conv_gate = nn.Conv3d(input_size + hidden_size, 2 * hidden_size, 3, padding=1)
        self.out_gate = nn.Conv3d(input_size + hidden_size, hidden_size, 3, padding=1)

    def forward(self, x, prev_state=None):
        if prev_state is None:
            b, _, d, h, w = x.shape
            prev_state = x.new_zeros(b, self.hidden_size, d, h, w)
        gates = self.conv_gate(torch.cat([x, prev_state], dim=1))
        update, reset = torch.split(gates, self.hidden_size, dim=1)
        update, reset = torch.sigmoid(update), torch.sigmoid(reset)
        candidate = torch.tanh(self.out_gate(torch.cat([x, prev_state * reset], dim=1)))
        return prev_state * (1 - update) + candidate * update


class ConvGRU_3D(nn.Module):
    def __init__(self, config, n_layers=1, input_size=16, hidden_size=16):
        super(ConvGRU_3D, self).__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.n_layers = n_layers
        self.cells = nn.ModuleList(
            [ConvGRUCell_3D(config, input_size if i == 0 else hidden_size, hidden_size) for i in range(n_layers)])
        self.fusion_norm = nn.BatchNorm3d(hidden_size)
        self.fusion_conv = nn.Sequential(
            nn.Conv3d(input_size, input_size, 3, padding=1),
            nn.BatchNorm3d(input_size),
            nn.LeakyReLU(inplace=True),
            nn.Conv3d(input_size, input_size, 3, padding=1),
            nn.BatchNorm3d(input_size),
            nn.LeakyReLU(inplace=True),
        )

    def forward(self, x, hidden=None):
        '''
        x: [b,t,c,d,h,w]; hidden: optional list with one initial state per layer
        '''
        seq_len = x.shape[1]
        if not hidden:
            hidden = [None] * self.n_layers
        layer_input = x
        h = None
        for layer_idx, cell in enumerate(self.cells):
            h = hidden[layer_idx]
            outputs = []
            for t in range(seq_len):
                h = cell(layer_input[:, t], h)
                outputs.append(h)
            if layer_idx + 1 < self.n_layers:
                layer_input = torch.stack(outputs, dim=1)
        return self.fusion_norm(h)
