"""CPU tests of the compat shims: the reference's own scripts-side imports resolve, and (when the reference
checkout is present, i.e. in the build container) `models.model.FORGE` built through the shim has exactly the
reference's state_dict keys and shapes, pose networks included."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FORGE_REFERENCE", "/root/reference")


def test_look_at_view_transform_geometry():
    from forge_b200.geometry import look_at_view_transform
    R, T = look_at_view_transform(dist=1.5, elev=[0.0, 30.0, -20.0], azim=[0.0, 90.0, 200.0])
    assert R.shape == (3, 3, 3) and T.shape == (3, 3)
    eye = torch.eye(3).expand(3, 3, 3)
    assert torch.allclose(R @ R.transpose(1, 2), eye, atol=1e-5)
    C = -torch.bmm(R, T[:, :, None])[:, :, 0]                  # camera centre, row-vector convention
    assert torch.allclose(C.norm(dim=1), torch.full((3,), 1.5), atol=1e-5)
    assert torch.allclose(T, torch.tensor([0.0, 0.0, 1.5]).expand(3, 3), atol=1e-5)   # origin sits on the optical axis
    assert torch.allclose(C[0], torch.tensor([0.0, 0.0, 1.5]), atol=1e-6)


def test_easydict_shim():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    try:
        import importlib
        ed = importlib.import_module("easydict")
        c = ed.EasyDict({'a': {'b': 1}, 'l': [{'x': 2}]})
        c.a.c = 3
        c.new = {'k': 'v'}
        assert c.a.b == 1 and c['a']['c'] == 3 and c.new.k == 'v' and c.l[0].x == 2
        with pytest.raises(AttributeError):
            c.missing
    finally:
        sys.path.remove(os.path.join(ROOT, "compat"))
        sys.modules.pop("easydict", None)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not present")
def test_reference_entry_imports_resolve_through_compat():
    """In a fresh interpreter with compat first on PYTHONPATH: the reference's config + models.model.FORGE (with
    the reference's pose networks; weight downloads stubbed) -> identical state_dict tables."""
    code = r'''
import json, sys, warnings
import torch, torchvision
import torch.utils.model_zoo as model_zoo
orig = torchvision.models.resnet50
torchvision.models.resnet50 = lambda pretrained=False, weights=None, **kw: orig(weights=None)
from config.config import config, update_config
update_config(sys.argv[1] + "/config/kubric/joint_pose_2d3d.yaml")
import models, pytorch3d
from pytorch3d.renderer import look_at_view_transform
import models.pose_estimator_2d as pe2d
pe2d.model_zoo.load_url = lambda *a, **k: pe2d.ResNet(pe2d.Bottleneck, [3, 4, 6, 3]).state_dict()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    from models.model import FORGE, sequence_from_distance, chose_selected
    from models.model_single_pose_estimator import FORGE_poseEstimator3D
    from models.volume_render import VolRender
    import forge_b200.models.volume_render as fv
    assert VolRender is fv.VolRender
    out = {"FORGE": {k: list(v.shape) for k, v in FORGE(config).state_dict().items()},
           "FORGE_poseEstimator3D": {k: list(v.shape) for k, v in FORGE_poseEstimator3D(config).state_dict().items()}}
print("JSON" + json.dumps(out))
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT, REF]), FORGE_REFERENCE=REF)
    res = subprocess.run([sys.executable, "-c", code, REF], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout[res.stdout.index("JSON") + 4:])
    with open(os.path.join(ROOT, "tests", "golden", "reference_state_dict_shapes.json")) as fh:
        ref = json.load(fh)
    assert got == ref
