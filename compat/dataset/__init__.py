"""`dataset` package for the reference's scripts: a synthetic Kubric-shaped dataset (no dataset files are reachable offline,
reference dataset/kubric.py:69 hard-codes /vision/vision_data/...); every other dataset module falls through to the
reference checkout."""
import os
import sys


def _reference_dir():
    for root in [os.environ.get("FORGE_REFERENCE")] + [p for p in sys.path if p]:
        if root and os.path.isfile(os.path.join(root, "dataset", "kubric.py")) \
                and os.path.abspath(os.path.join(root, "dataset")) != os.path.dirname(os.path.abspath(__file__)):
            return os.path.join(root, "dataset")
    return None


_ref = _reference_dir()
if _ref is not None:
    __path__.append(_ref)
