"""CPU checks of the C-ABI boundary: the library builds/loads without a GPU and exports every
symbol include/forge_b200.h declares; the Python mirrors keep the reference's state_dict keys."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "forge_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(forge_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from forge_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared_symbols()
    assert len(names) >= 11
    for name in names:
        assert hasattr(lib, name), "libforge_b200.so does not export %s" % name
    assert set(names) == set(_lib._SIGNATURES), "ctypes binding and header disagree"
    lib.forge_abi_version.restype = ctypes.c_int
    assert lib.forge_abi_version() == _lib.ABI_VERSION


def test_ctypes_signatures_match_header_prototypes():
    """Every prototype of include/forge_b200.h against forge_b200/_lib.py's ctypes table: same number of parameters and the same
    kind per parameter (pointer / int / long long / float / unsigned) -- an ABI change that forgets one side fails here, on CPU."""
    from forge_b200 import _lib
    with open(os.path.join(ROOT, "include", "forge_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = re.findall(r"\b([A-Za-z_][A-Za-z0-9_ ]*?[ \*])\s*(forge_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(_lib._SIGNATURES)

    def kind(decl):
        decl = decl.strip()
        if "*" in decl:
            return "ptr"
        base = " ".join(decl.split()[:-1]) if len(decl.split()) > 1 else decl      # drop the parameter name
        return {"int": "int", "long long": "ll", "float": "float", "unsigned": "uint", "unsigned int": "uint"}[base.replace("const ", "")]

    ckind = {ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr", ctypes.c_int: "int", ctypes.c_longlong: "ll", ctypes.c_float: "float",
             ctypes.c_uint: "uint"}
    for ret, name, params in protos:
        restype, argtypes = _lib._SIGNATURES[name]
        params = params.strip()
        want = [] if params in ("", "void") else [kind(p) for p in params.split(",")]
        assert [ckind[a] for a in argtypes] == want, name
        assert ckind[restype] == ("ptr" if "*" in ret else kind(ret.strip() + " x")), name


def test_error_reporting_without_gpu():
    from forge_b200 import _lib
    with pytest.raises(RuntimeError, match="forge_raymarch_fwd: null pointer"):
        _lib.call("forge_raymarch_fwd", None, None, None, None, None, None, None, None, 1, 1, 4, 4, 4, 8, 8, 4, None)
    with pytest.raises(RuntimeError, match="non-positive"):
        _lib.call("forge_ncs_to_nsc", 16, 16, 0, 1, 1, None)


def test_product_path_refuses_cpu_tensors():
    from forge_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.to_channels_last(torch.zeros(1, 4, 2, 2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.raymarch(torch.zeros(1, 16, 2, 2, 2), torch.zeros(1, 1, 2, 2, 2), torch.zeros(1, 12),
                     torch.zeros(1).int(), torch.zeros(4), 2, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "forge_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                with open(os.path.join(d, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(d, f)
                assert "/root/reference" not in src


def test_state_dict_keys_match_reference_layout():
    """conv_rgb.{0,1,3,4,6}.* and Rotate_world.conv3d_{1..4}.* (SURVEY 8b)."""
    from forge_b200.synthetic import make_config
    from forge_b200.models.volume_render import VolRender
    from forge_b200.models.rotate import Rotate_world
    sd = VolRender(make_config()).state_dict()
    assert sd['conv_rgb.0.weight'].shape == (16, 16, 6, 6)
    assert sd['conv_rgb.3.weight'].shape == (8, 16, 5, 5)
    assert sd['conv_rgb.6.weight'].shape == (3, 8, 5, 5)
    assert {k.split('.')[1] for k in sd} == {'0', '1', '3', '4', '6'}
    rot = Rotate_world(make_config())
    rsd = rot.state_dict()
    assert sorted(rsd) == sorted(['conv3d_%d.%s' % (i, p) for i in (1, 2, 3, 4) for p in ('weight', 'bias')])
    assert rsd['conv3d_3.weight'].shape == (128, 128, 3, 3, 3)
    assert rot.grid_coord_max == 0.484375 and rot.grid_coord_max_64 == 0.4921875
    assert tuple(rot.grid_coord.shape) == (32, 32, 32, 3)
    # golden state_dict from the reference module loads strictly
    from conftest import load_golden
    g = load_golden("volrender_small")
    VolRender(make_config(img_size=32, n_pts_per_ray=24)).load_state_dict(
        {k[3:]: v for k, v in g.items() if k.startswith('sd.')}, strict=True)


def test_model_state_dict_matches_reference_tables():
    """Every encoder_3d / render / rotate key (and shape) of the reference FORGE and FORGE_poseEstimator3D exists
    here and vice versa; the remaining reference keys belong to the out-of-scope pose modules only."""
    import json
    import warnings
    from forge_b200.synthetic import make_config
    from forge_b200.models.model import FORGE
    from forge_b200.models.model_single_pose_estimator import FORGE_poseEstimator3D
    with open(os.path.join(ROOT, "tests", "golden", "reference_state_dict_shapes.json")) as fh:
        ref = json.load(fh)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mine = {"FORGE": FORGE(make_config()).state_dict(), "FORGE_poseEstimator3D": FORGE_poseEstimator3D(make_config()).state_dict()}
    for name, table in ref.items():
        in_scope = {k: v for k, v in table.items() if k.split('.')[0] in ('encoder_3d', 'render', 'rotate')}
        others = {k.split('.')[0] for k in table} - {'encoder_3d', 'render', 'rotate'}
        assert others <= {'encoder_traj', 'encoder_traj_2d', 'pose_head'}
        got = {k: list(v.shape) for k, v in mine[name].items()}
        assert got == in_scope, (sorted(set(got) ^ set(in_scope))[:10])
