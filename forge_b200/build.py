"""Build libforge_b200.so in-tree with nvcc for sm_100a (no torch / pybind headers involved).

    python -m forge_b200.build [--force] [--verbose]

The shared object lands in forge_b200/lib/ (git-ignored, but it travels to the GPU box with the
repo snapshot).  ``-lineinfo`` keeps ncu's source page usable.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libforge_b200.so")
STAMP = os.path.join(LIBDIR, "libforge_b200.stamp")
SOURCES = ["layout.cu", "raymarch.cu", "rotate.cu", "decoder.cu", "decoder_tc.cu", "decoder_bwd.cu", "camera.cu", "gru.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libforge_b200 cannot be built")


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha256()
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(PKG), "include", "forge_b200.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed. Returns the path of the shared library."""
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    cmd = [_nvcc(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v" if verbose else "-O3", "-o", LIB, *_sources()]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libforge_b200.so (see stderr)")
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
