"""TEST INFRASTRUCTURE ONLY -- a stand-in for the slice of PyTorch3D 0.7.0 that FORGE imports.

PyTorch3D (pinned by the reference at README.md:20-21, ``pytorch3d-0.7.0``) is a third-party
dependency that is neither vendored under /root/reference nor installable offline.  This
package restates, from the published algorithm of that release, exactly the classes and
functions the reference's hot path touches, so that the UNMODIFIED reference modules
(``/root/reference/models/volume_render.py``, ``models/rotate.py``) can be imported and run
on CPU to (a) produce golden vectors (``oracle/make_golden.py``) and (b) act as the timing
proxy for "the reference PyTorch3D renderer" (PyTorch3D's implicit/volume path is pure
PyTorch, so the same ATen ops run).

PARITY UNPINNED: the reference ships no tests or golden tensors, and the real package cannot
be run here; the arithmetic below is a restatement, cross-checked against an independent
closed-form derivation (``oracle/closed_form.py``) and against the constants the reference
itself states (e.g. ``models/rotate.py:23`` "should be 0.4844").

Nothing in the product path (``forge_b200/``) may import this package.
"""
__version__ = "0.7.0+standin"
