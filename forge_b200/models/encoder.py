"""Encoder3D -- 2-D -> 3-D lifting and the render heads, mirror of reference models/encoder.py (:8-78).

ResNet-50 trunk with the strides of layer3/layer4 removed (256^2 image -> 2048 x 32 x 32), reshaped to
a 64 x 32^3 volume, one Conv3d to 128 channels; ``features_head`` (128 -> 16 ch, x2) and
``density_head`` (128 -> 1 ch, x2, ReLU) produce the 64^3 render volumes; ``fuse`` runs the ConvGRU.
Same child / parameter names as the reference, so its checkpoints load with ``strict=True``.
All of this stays PyTorch/cuDNN (SURVEY 8a rows a11-a13 are host rows).
"""
import os
import warnings

import torch
import torch.nn as nn
import torchvision

from .fusion import ConvGRU_3D


def _upsample_stem():
    """128 -> 32 channels, x2 resolution (shared shape of both heads)"""
    return [nn.ConvTranspose3d(128, 32, 4, stride=2, padding=1), nn.BatchNorm3d(32), nn.LeakyReLU(inplace=True)]


class Encoder3D(nn.Module):
    def __init__(self, config):
        super(Encoder3D, self).__init__()
        self.feature_extraction = get_resnet50()
        # 16-channel render features: indices 0,1,3,4 carry parameters (reference :16-22)
        self.features_head = nn.Sequential(*_upsample_stem(), nn.Conv3d(32, 16, 3, padding=1), nn.BatchNorm3d(16))
        # non-negative density: indices 0,1,3,4,6 carry parameters (reference :25-34)
        self.density_head = nn.Sequential(*_upsample_stem(), nn.Conv3d(32, 8, 3, padding=1), nn.BatchNorm3d(8),
                                          nn.LeakyReLU(inplace=True), nn.Conv3d(8, 1, 3, padding=1), nn.ReLU(inplace=True))
        # first 3-D conv after the reshape-lift (reference :36-40)
        self.conv1 = nn.Sequential(nn.Conv3d(64, 128, 3, padding=1), nn.BatchNorm3d(128), nn.LeakyReLU(inplace=True))
        self.fusion_feature = ConvGRU_3D(config, n_layers=1, input_size=128, hidden_size=128)
        # arithmetic of the 3-D conv stack (fusion + heads), the FLOP-dominant part of a step:
        # None = fp32 like the reference (cuDNN may use TF32, torch's default); torch.bfloat16 = autocast on the
        # tensor cores (measured fuse+heads fwd+bwd, b=1: 12.7 ms -> 8.8 ms with channels-last weights)
        self.compute_dtype = None
        # with compute_dtype = bfloat16 and no autograd graph wanted, fuse() runs forge_conv3d_tc (tcgen05) instead of
        # cuDNN under autocast
        self.tc_fusion = True
        self._trunk_channels_last = False

    def _conv1_tc_eligible(self, z_3d):
        """the lifting convolution (64 -> 128, BN, LeakyReLU; reference :36-40, :50) on forge_conv3d_tc: bf16 mode, eval-mode BN,
        no autograd graph wanted, a grid the kernel tiles"""
        _, _, D, H, W = z_3d.shape
        wants_graph = torch.is_grad_enabled() and (z_3d.requires_grad or any(p.requires_grad for p in self.conv1.parameters()))
        return (self.compute_dtype == torch.bfloat16 and self.tc_fusion and z_3d.is_cuda and not self.conv1.training
                and not wants_graph and D % 4 == 0 and H % 4 == 0 and W % 8 == 0)

    def _conv1_tc(self, z_3d):
        from .. import ops
        conv, bn = self.conv1[0], self.conv1[1]
        tensors = list(self.conv1.parameters()) + list(self.conv1.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if getattr(self, '_conv1_pack', None) is None or self._conv1_pack[0] != key:
            with torch.no_grad():
                s = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
                b = ((conv.bias - bn.running_mean) * s + bn.bias).float().contiguous()
                self._conv1_pack = (key, ops.pack_conv3d_weights(conv.weight), s, b)
        _, w, s, b = self._conv1_pack
        x = z_3d.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)      # one re-layout pass of the lifted volume
        y, _, _ = ops.conv3d_tc(x, w, 'plain', b, scale=s, lrelu=True)
        return y                                          # fp32, channels-last memory: what K2 consumes zero-copy

    def get_feat3D(self, img):
        with self._amp():
            if self._trunk_channels_last:
                img = img.contiguous(memory_format=torch.channels_last)
            z_2d = self.feature_extraction(img)
            B, C, H, W = z_2d.shape                   # stride-8 feature map
            # the lift is a reshape of the NCHW tensor: 2048 = 64 ch x 32 depth (needs NCHW memory order)
            z_3d = z_2d.contiguous().view(-1, 64, 32, H, W)
            if self._conv1_tc_eligible(z_3d):
                return self._conv1_tc(z_3d)
            if self._trunk_channels_last:
                z_3d = z_3d.contiguous(memory_format=torch.channels_last_3d)
            return self.conv1(z_3d).float()

    def _amp(self):
        return torch.autocast('cuda', dtype=self.compute_dtype, enabled=self.compute_dtype is not None)

    def channels_last_3d_(self):
        """Convert the conv weights (3-D stack and the 2-D ResNet trunk) to channels-last in place (what cuDNN's
        tensor-core kernels want; K2 already emits channels-last volumes).  state_dict keys / values are unaffected."""
        for mod in (self.fusion_feature, self.features_head, self.density_head, self.conv1):
            mod.to(memory_format=torch.channels_last_3d)
        self.feature_extraction.to(memory_format=torch.channels_last)     # the 2-D ResNet trunk of the lift
        self._trunk_channels_last = True
        return self

    def heads_tc_eligible(self, x):
        """both heads on the tensor cores (ops.heads_tc): bf16 mode, eval-mode BN, no autograd graph wanted, tileable grid"""
        _, c, d, h, w = x.shape
        mods = (self.features_head, self.density_head)
        wants_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in mods for p in m.parameters()))
        return (self.compute_dtype == torch.bfloat16 and self.tc_fusion and x.is_cuda and c == 128 and not wants_graph
                and not any(m.training for m in mods) and d % 4 == 0 and h % 4 == 0 and w % 8 == 0)

    def get_render_volumes(self, x):
        """(get_render_features(x), get_density3D(x)); one fused tensor-core path when eligible"""
        if self.heads_tc_eligible(x):
            from .. import ops
            tensors = [t for m in (self.features_head, self.density_head) for t in list(m.parameters()) + list(m.buffers())]
            key = tuple((t.data_ptr(), t._version) for t in tensors)
            if getattr(self, '_heads_pack', None) is None or self._heads_pack[0] != key:
                self._heads_pack = (key, ops.pack_heads_tc(self.features_head, self.density_head))
            return ops.heads_tc(self._heads_pack[1], x)
        return self.get_render_features(x), self.get_density3D(x)

    def get_density3D(self, z_3d):
        with self._amp():
            return self.density_head(z_3d).float()

    def get_render_features(self, x):
        with self._amp():
            return self.features_head(x).float()

    def fuse(self, x):
        # x in [b,t,c,d,h,w]; hidden state initialised from the view mean (reference :59-63)
        if self.compute_dtype == torch.bfloat16 and self.tc_fusion and self.fusion_feature.tc_eligible(x):
            return self.fusion_feature.forward_tc(x)        # tcgen05 convolutions with fused gate epilogues (inference)
        with self._amp():
            if isinstance(x, (list, tuple)):        # per-view tensors (Rotate_world.forward_views: no view-0 copy)
                mean = x[0]
                for v in x[1:]:
                    mean = mean + v
                mean = mean / len(x)
            else:
                mean = x.mean(dim=1)
            return self.fusion_feature(x, [self.fusion_feature.fusion_conv(mean)]).float()

    def forward(self, x):
        raise NotImplementedError


def get_resnet50():
    """ImageNet ResNet-50 without avgpool/fc, layer3/layer4 strides set to 1 (reference :71-78).
    Offline (no weight download possible) the trunk is randomly initialised, with a warning."""
    weights = torchvision.models.ResNet50_Weights.IMAGENET1K_V1
    cached = os.path.join(torch.hub.get_dir(), 'checkpoints', os.path.basename(weights.url))
    if os.path.exists(cached) or os.environ.get('FORGE_ALLOW_DOWNLOAD') == '1':
        try:
            model = torchvision.models.resnet50(weights=weights)
        except Exception as e:   # no network / corrupt cache
            warnings.warn("ImageNet weights for ResNet-50 unavailable (%s); using random init" % type(e).__name__)
            model = torchvision.models.resnet50(weights=None)
    else:    # never touch the network implicitly (the reference downloads at construction time, models/encoder.py:72)
        warnings.warn("ImageNet weights for ResNet-50 are not in the torch hub cache (%s); using random init "
                      "(set FORGE_ALLOW_DOWNLOAD=1 to fetch them)" % cached)
        model = torchvision.models.resnet50(weights=None)
    feature = nn.Sequential(*list(model.children())[:-2])
    for stage in (6, 7):
        feature[stage][0].conv2.stride = (1, 1)
        feature[stage][0].downsample[0].stride = (1, 1)
    return feature
