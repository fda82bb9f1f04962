"""ctypes binding of libmahakala_b200.so (the C ABI declared in include/mahakala_b200.h).

The product path has NO CPU fallback: if the shared library is missing or CUDA is unavailable every
compute entry point raises ``RuntimeError``.  Importing the package (and loading the library to inspect
its symbols) works without a GPU so that the build can be checked on a CPU-only host.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MAHAKALA_B200_LIB selects another build of the same library (kernel-variant experiments, scripts/build_variant.sh)
LIB_PATH = os.environ.get("MAHAKALA_B200_LIB") or os.path.join(_HERE, "libmahakala_b200.so")

_c = {"d": ctypes.c_double, "l": ctypes.c_long, "i": ctypes.c_int, "p": ctypes.c_void_p}

# name -> argument kinds, in the order of include/mahakala_b200.h
SIGNATURES = {
    "mk_abi_version": "",
    "mk_device_info": "ppppp",
    "mk_measure_fp64_peak": "ipp",
    "mk_fast_math_probe": "plpppp",
    "mk_transcendental_probe": "plpppp",
    "mk_register_metric": "pppppl",
    "mk_metric_set_params": "ip",
    "mk_initial_condition_metric": "idpplpp",
    "mk_camera_grid": "ddddddlipp",
    "mk_camera_points": "ddddpplipp",
    "mk_initial_condition": "dpplpp",
    "mk_integrate": "idllpddpppppl" "pp",
    "mk_integrate_adaptive": "idllp" "dddd" "pppp" "p",
    "mk_fill_frozen_rows": "ppppllp",
    "mk_integrate_paged": "idllpdd" "ppp" "pppp" "l" "pp" "p",
    "mk_integrate_shared": "idllpdd" "ppp" "pppp" "l" "pp" "pp" "li" "p",
    "mk_paged_gather": "ppppp" "lll" "ppp",
    "mk_shadow_bisection": "dddd" "pp" "ll" "dd" "i" "ddd" "ppp",
    "mk_radius_cal": "dpllpp",
    "mk_rhs": "idplpp",
    "mk_rk4_step": "idpplpp",
    "mk_metric": "idplppp",
    "mk_snapshot_create": "llllpppppppppipp",
    "mk_snapshot_create_from_interiors": "llll" "pipii" "ppp" "ppppppp" "i" "ppp",
    "mk_upload_pageable": "pplip",
    "mk_snapshot_unpack": "pppp",
    "mk_snapshot_create_torus": "pp",
    "mk_snapshot_destroy": "p",
    "mk_snapshot_cells": "ppp",
    "mk_sample_scalars": "pdpldpp",
    "mk_sample_prims": "pplpp",
    "mk_rlow_rhigh": "ppplddddddd" "pp",
    "mk_synchrotron": "ppppppl" "id" "ppp",
    "mk_solve_specific_intensity": "pppll" "d" "ppp",
    "mk_solve_attenuated_emissivity": "pppll" "d" "pp",
    "mk_emission_from_states": "ppdpl" "d" "ppp",
    "mk_emission_probe": "pdppl" "ip" "i" "ppp",
    "mk_render": "dddddd" "l" "p" "ll" "dd" "pp" "i" "p" "ppppp" "lll" "pp",
    "mk_render_long": "dddddd" "l" "p" "ll" "dd" "pp" "i" "p" "ppppp" "lll" "p" "ii" "p",
    "mk_render_metric": "i" "dddddd" "l" "p" "ll" "dd" "pp" "i" "p" "ppppp" "lll" "pp",
    "mk_ipc_alloc": "lpp",
    "mk_ipc_open": "pp",
    "mk_ipc_close": "p",
    "mk_ipc_free": "p",
}

# entry points that do not return a status code: name -> (restype, argument kinds)
OTHER = {
    "mk_last_error_string": (ctypes.c_char_p, ""),
    "mk_snapshot_bytes": (ctypes.c_long, "p"),
    "mk_page_rows": (ctypes.c_int, ""),
    "mk_render_patch_count": (ctypes.c_long, "lpl"),
}

_lib = None


class MahakalaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MahakalaB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C mahakala_b200/csrc`).  mahakala_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, sig) in OTHER.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = [_c[k] for k in sig]
        for name, sig in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_c[k] for k in sig]
        _lib = lib
    return _lib


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, int):
        return ctypes.c_void_p(x)
    if hasattr(x, "data_ptr"):            # torch tensor
        return ctypes.c_void_p(x.data_ptr())
    if isinstance(x, ctypes.c_void_p):
        return x
    if isinstance(x, bytes):
        return ctypes.cast(ctypes.c_char_p(x), ctypes.c_void_p)
    if isinstance(x, (ctypes._SimpleCData, ctypes.Array, ctypes.Structure)):
        return ctypes.cast(ctypes.pointer(x), ctypes.c_void_p)
    if type(x).__name__ == "CArgObject":      # ctypes.byref(...)
        return x
    raise TypeError(f"cannot pass {type(x)} as a pointer")


def call(name, *args):
    """Invoke ``name`` with Python scalars / torch tensors / None; raise on a non-zero status."""
    lib = load()
    sig = SIGNATURES[name]
    if len(args) != len(sig):
        raise TypeError(f"{name} takes {len(sig)} arguments, got {len(args)}")
    conv = []
    for kind, a in zip(sig, args):
        conv.append(_ptr(a) if kind == "p" else a)
    rc = getattr(lib, name)(*conv)
    if name == "mk_abi_version":
        return rc
    if rc != 0:
        raise MahakalaB200Error(f"{name} failed ({rc}): {lib.mk_last_error_string().decode()}")
    return rc
