"""TEST INFRASTRUCTURE: write tests/golden/reference_state_dict_shapes.json -- the state_dict key/shape
tables of the UNMODIFIED reference models (models/model.py::FORGE, model_single_pose_estimator.py::
FORGE_poseEstimator3D), built offline (weight downloads stubbed out, PyTorch3D stand-in, easydict shim).

Run in the build container only (needs /root/reference):   python -m oracle.dump_reference_keys
"""
import json
import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FORGE_REFERENCE", "/root/reference")


def main():
    sys.path.insert(0, ROOT)
    from oracle import enable_p3d_standin
    enable_p3d_standin()
    sys.path.insert(0, REF)
    import numpy as np
    import torch  # noqa: F401
    import torchvision
    import torch.utils.model_zoo as model_zoo

    orig = torchvision.models.resnet50
    torchvision.models.resnet50 = lambda pretrained=False, **kw: orig(weights=None)      # no download
    if not hasattr(np, 'float'):
        np.float = float                                                                 # models/model_utils.py:45

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)
        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)
    mod = types.ModuleType('easydict')
    mod.EasyDict = EasyDict
    sys.modules['easydict'] = mod

    from config.config import config, update_config
    update_config(os.path.join(REF, 'config/kubric/joint_pose_2d3d.yaml'))
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import models.pose_estimator_2d as pe2d
        real_load = model_zoo.load_url
        pe2d.model_zoo.load_url = lambda *a, **k: pe2d.ResNet(pe2d.Bottleneck, [3, 4, 6, 3]).state_dict()
        from models.model import FORGE
        out['FORGE'] = {k: list(v.shape) for k, v in FORGE(config).state_dict().items()}
        from models.model_single_pose_estimator import FORGE_poseEstimator3D
        out['FORGE_poseEstimator3D'] = {k: list(v.shape) for k, v in FORGE_poseEstimator3D(config).state_dict().items()}
        pe2d.model_zoo.load_url = real_load
    for name, table in out.items():
        scope = sum(1 for n in table if n.split('.')[0] in ('encoder_3d', 'render', 'rotate'))
        print(name, len(table), "keys,", scope, "under encoder_3d / render / rotate")
    with open(os.path.join(ROOT, 'tests', 'golden', 'reference_state_dict_shapes.json'), 'w') as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
