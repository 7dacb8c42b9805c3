"""Loaded automatically when compat/ is on PYTHONPATH (the reference's scripts are run unchanged).

* aliases the reference still uses but current numpy / Pillow dropped (models/model_utils.py:45, dataset/kubric.py:426);
* offline fallback for the weight downloads the reference triggers at construction time (models/encoder.py:72,
  models/pose_estimator_2d.py:96 / :197, models/perceptual_loss.py:11-14): when the file is neither cached nor
  downloadable the model keeps its random initialisation and a warning is printed -- checkpoints overwrite these weights
  anyway (utils/exp_utils.py loaders, strict=True).
"""
import sys
import warnings

try:
    import numpy as _np
    if not hasattr(_np, "float"):
        _np.float = float
except Exception:
    pass
try:
    from PIL import Image as _Image
    if not hasattr(_Image, "ANTIALIAS"):
        _Image.ANTIALIAS = _Image.LANCZOS
except Exception:
    pass


class _OfflineWeights(dict):
    """returned instead of a downloaded state_dict; Module.load_state_dict treats it as 'keep the current values'"""


def _patch_torch_hub(hub):
    if getattr(hub, "_forge_offline", False):
        return
    orig = hub.load_state_dict_from_url

    def load_state_dict_from_url(url, *a, **kw):
        try:
            return orig(url, *a, **kw)
        except Exception as e:          # no network, no cached file
            warnings.warn("forge_b200 compat: could not fetch %s (%s); keeping the random initialisation" % (url, type(e).__name__))
            return _OfflineWeights()
    hub.load_state_dict_from_url = load_state_dict_from_url
    hub._forge_offline = True
    import torch
    import torch.utils.model_zoo as model_zoo
    model_zoo.load_url = load_state_dict_from_url
    orig_load = torch.nn.Module.load_state_dict

    def load_state_dict(self, state_dict, *a, **kw):
        if isinstance(state_dict, _OfflineWeights):
            return torch.nn.modules.module._IncompatibleKeys([], [])
        return orig_load(self, state_dict, *a, **kw)
    torch.nn.Module.load_state_dict = load_state_dict


def _patch_torchvision(tv):
    if getattr(tv, "_forge_offline", False):
        return
    tv._forge_offline = True
    import torch.hub
    _patch_torch_hub(torch.hub)
    # torchvision binds load_state_dict_from_url at import time in its own modules
    for name in ("torchvision.models._api", "torchvision._internally_replaced_utils"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "load_state_dict_from_url"):
            mod.load_state_dict_from_url = torch.hub.load_state_dict_from_url


# The reference's scripts import torch / torchvision first thing, so patching eagerly costs nothing extra.
try:
    import torch.hub as _hub
    _patch_torch_hub(_hub)
    import torchvision as _tv
    _patch_torchvision(_tv)
except Exception as _e:       # never break interpreter start-up
    warnings.warn("forge_b200 compat: offline-weights fallback not installed: %r" % (_e,))
