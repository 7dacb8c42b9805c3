"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU (and, for timing only, CUDA-eager) restatement of the reference algorithm for the FORGE
render / rotate hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; ``forge_b200/``
never does (``tests/test_abi.py`` greps the product tree), and the product path raises if its CUDA
library is missing.  The development benches under ``tools/`` (``--ref`` columns of bench_kernels.py,
refine_bench.py, model_bench.py) time it on the GPU as the REFERENCE arm SURVEY 8d asks for -- the
denominator of the "x the reference GPU renderer" figures -- never as part of a forge_b200 measurement.

Layout
  p3d_standin/pytorch3d/   stand-in for the slice of PyTorch3D 0.7.0 the reference imports
                           (third-party dependency, pinned at reference README.md:20-21 plus the
                           depth patch at README.md:26-33; absent from /root/reference and not
                           installable offline -> restated from its published algorithm)
  reference_path.py        the reference's own call sequence (models/volume_render.py:40-103,
                           models/rotate.py:48-156) over that stand-in: same ATen ops, same order
  closed_form.py           independent closed-form derivation (camera centre / pixel rays /
                           affine grid), fp32 or fp64 -- cross-checks reference_path.py
  make_golden.py           imports the UNMODIFIED reference modules from /root/reference (with the
                           stand-in on sys.path) and writes tests/golden/*.npz

PARITY STATUS: "parity unpinned" for the PyTorch3D-internal arithmetic (no reference tests, no
real PyTorch3D here).  Pinned parts: the reference's own glue code is executed for the golden
vectors; ``F.grid_sample`` in ``Rotate_world`` is the real torch op; the grid-max constant
0.4844 quoted at models/rotate.py:23 is reproduced; ray geometry is checked against the
closed-form camera model.
"""
import os
import sys

_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "p3d_standin")


def enable_p3d_standin():
    """Put the PyTorch3D stand-in on sys.path (no-op if a real pytorch3d is importable)."""
    try:
        import pytorch3d  # noqa: F401
        return
    except ImportError:
        pass
    if _STANDIN not in sys.path:
        sys.path.insert(0, _STANDIN)
