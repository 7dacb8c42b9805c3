"""Build libforge_b200.so in-tree with nvcc for sm_100a (no torch / pybind headers involved).

    python -m forge_b200.build [--force] [--verbose]

The shared object lands in forge_b200/lib/ (git-ignored, but it travels to the GPU box with the
repo snapshot).  ``-lineinfo`` keeps ncu's source page usable.
"""
import fcntl
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
SOURCES = ["layout.cu", "raymarch.cu", "raymarch_tma.cu", "raymarch_tma1.cu", "rotate.cu", "decoder.cu", "decoder_tc.cu", "decoder_bwd.cu", "camera.cu", "gru.cu", "conv3d_tc.cu", "gru_tc_bwd.cu"]
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


def _up_to_date(digest):
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == digest


def build(force=False, verbose=False):
    """Compile if sources changed. Returns the path of the shared library.

    Safe under torchrun / DDP: the digest check and the compile run under an inter-process file lock, nvcc writes to a
    temporary file that is renamed into place (a concurrent CDLL never sees a half-written library), and the stamp is
    written after the rename."""
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):        # another rank built it while we waited
                return LIB
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [_nvcc(), "-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
                   "-Xptxas", "-v" if verbose else "-O3", "-o", tmp, *_sources(), "-ldl"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libforge_b200.so (see stderr)")
            if os.path.exists(STAMP):
                os.remove(STAMP)                           # never a stamp that describes another library
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as fh:
                fh.write(digest)
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def stamp_matches_sources():
    """True when the library on disk was built from the sources on disk."""
    return _up_to_date(_digest())


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
