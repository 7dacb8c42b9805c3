"""`models` package for the reference's scripts: render / rotate / fusion / model graph from forge_b200,
everything else (pose networks, model_utils, perceptual loss) from the reference checkout."""
import os
import sys


def _reference_models_dir():
    cands = [os.environ.get("FORGE_REFERENCE")] + [p for p in sys.path if p]
    for root in cands:
        if root and os.path.isfile(os.path.join(root, "models", "pose_estimator_3d.py")) \
                and os.path.abspath(os.path.join(root, "models")) != os.path.dirname(os.path.abspath(__file__)):
            return os.path.join(root, "models")
    return None


_ref = _reference_models_dir()
if _ref is not None:
    __path__.append(_ref)      # submodules not shadowed here are imported from the reference
