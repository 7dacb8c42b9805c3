"""ctypes binding of libforge_b200.so (the C ABI declared in include/forge_b200.h).

There is no fallback: if the shared library is missing it is built with nvcc; if that is not
possible, or the ABI version differs, importing callers get a RuntimeError.
"""
import ctypes
import os
import threading
import warnings

from . import build as _build

_c = ctypes
_F = _c.c_void_p      # device pointers travel as integers (tensor.data_ptr())
_I = _c.c_int

_SIGNATURES = {
    "forge_abi_version": (_c.c_int, []),
    "forge_last_error": (_c.c_char_p, []),
    "forge_ncs_to_nsc": (_c.c_int, [_F, _F, _I, _I, _c.c_longlong, _F]),
    "forge_nsc_to_ncs": (_c.c_int, [_F, _F, _I, _I, _c.c_longlong, _F]),
    "forge_pack_volume": (_c.c_int, [_F, _I, _F, _F, _F, _I, _I, _I, _I, _F]),
    "forge_unpack_volume_grad": (_c.c_int, [_F, _F, _I, _I, _I, _I, _I, _F]),
    "forge_raymarch_fwd": (_c.c_int, [_F] * 8 + [_I] * 8 + [_F]),
    "forge_raymarch_fwd_gather": (_c.c_int, [_F] * 8 + [_I] * 8 + [_F]),
    "forge_raymarch_fwd_tma": (_c.c_int, [_F] * 8 + [_I] * 8 + [_F]),
    "forge_raymarch_bwd": (_c.c_int, [_F] * 12 + [_I] * 8 + [_F]),
    "forge_raymarch_bwd_workspace": (_c.c_longlong, [_I] * 8),
    "forge_rotate_fwd": (_c.c_int, [_F] * 6 + [_c.c_float, _F] + [_I] * 5 + [_F]),
    "forge_rotate_bwd": (_c.c_int, [_F] * 6 + [_c.c_float, _F, _F, _F] + [_I] * 5 + [_F]),
    "forge_decoder_wpack_floats": (_c.c_int, []),
    "forge_decoder_fwd": (_c.c_int, [_F, _F, _F, _F, _I, _I, _I, _F]),
    "forge_decoder_bwd_wpack_floats": (_c.c_int, []),
    "forge_decoder_bwd_data": (_c.c_int, [_F, _F, _F, _F, _I, _I, _I, _F]),
    "forge_decoder_tc_wpack_bytes": (_c.c_int, []),
    "forge_decoder_tc_fwd": (_c.c_int, [_F, _F, _F, _F, _I, _I, _I, _I, _F]),
    "forge_umma_probe": (_c.c_int, [_F, _I] + [_c.c_uint] * 6 + [_F, _F]),
    "forge_camera_prep_fwd": (_c.c_int, [_F, _F, _F, _I] + [_c.c_float] * 4 + [_F, _F, _F]),
    "forge_camera_prep_bwd": (_c.c_int, [_F, _F, _F, _I] + [_c.c_float] * 4 + [_F] * 5 + [_F]),
    "forge_upsample2x_fwd": (_c.c_int, [_F, _F, _F, _F, _I, _I, _I, _F]),
    "forge_upsample2x_bwd": (_c.c_int, [_F, _F, _F, _F, _I, _I, _I, _F]),
    "forge_pose_affine_fwd": (_c.c_int, [_F, _I, _I, _F, _F, _F, _F]),
    "forge_gru_gate_fwd": (_c.c_int, [_F, _I, _F, _c.c_longlong, _F, _c.c_longlong, _F, _I, _I, _I, _I, _F]),
    "forge_gru_gate_bwd": (_c.c_int, [_F, _F, _I, _F, _c.c_longlong, _F, _F, _F, _I, _I, _I, _I, _F]),
    "forge_gru_out_fwd": (_c.c_int, [_F, _F, _I, _F, _c.c_longlong, _F, _I, _I, _I, _I, _F]),
    "forge_gru_out_bwd": (_c.c_int, [_F, _F, _F, _I, _F, _c.c_longlong, _F, _F, _F, _I, _I, _I, _I, _F]),
    "forge_conv3d_tc": (_c.c_int, [_F, _c.c_longlong, _I, _F, _c.c_longlong, _I, _F, _I, _I] + [_F] * 9 + [_I] * 8 + [_F]),
    "forge_conv3d_c8_to_1_relu": (_c.c_int, [_F, _F, _c.c_float, _F, _I, _I, _I, _I, _F]),
    "forge_gru_tc_bwd": (_c.c_int, [_I] + [_F] * 8 + [_c.c_longlong, _I, _F]),
    "forge_sample_points": (_c.c_int, [_F, _I, _I, _I, _I, _I, _F, _F, _F]),
}
ABI_VERSION = 18

_lock = threading.Lock()
_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building if necessary) and return the ctypes handle. Raises RuntimeError on failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        try:
            path = _build.build()       # no-op when the library matches the sources (digest stamp)
        except Exception as e:
            if not os.path.exists(path):  # no silent fallback
                raise RuntimeError(
                    "forge_b200: libforge_b200.so is missing and could not be built (%s). "
                    "The CUDA extension is mandatory; there is no CPU path." % e) from e
            # a library exists but it is not the one these sources describe: refuse unless told otherwise
            if os.environ.get("FORGE_ALLOW_STALE_LIB") != "1":
                raise RuntimeError(
                    "forge_b200: the sources changed but rebuilding libforge_b200.so failed (%s); refusing to load the "
                    "stale library (set FORGE_ALLOW_STALE_LIB=1 to load it anyway)" % e) from e
            warnings.warn("forge_b200: loading a STALE libforge_b200.so (rebuild failed: %s)" % e)
        try:
            lib = _c.CDLL(path)
        except OSError as e:
            raise RuntimeError("forge_b200: cannot load %s: %s" % (path, e)) from e
        for name, sig in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise RuntimeError("forge_b200: %s does not export %s; rebuild with "
                                   "`python -m forge_b200.build --force`" % (path, name)) from e
            fn.restype, fn.argtypes = sig
        got = lib.forge_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError("forge_b200: ABI version mismatch (library %d, binding %d); rebuild with "
                               "`python -m forge_b200.build --force`" % (got, ABI_VERSION))
        _lib = lib
        return _lib


def call(name, *args):
    """Invoke an entry point; raise RuntimeError(forge_last_error()) on a non-zero return."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(lib.forge_last_error().decode("utf-8", "replace"))
    return rc
