"""ctypes binding of libeigensolver_b200.so (the C ABI in include/eigensolver_b200.h).

The library is built in-tree by `__graft_entry__.build()` /
`make -C eigensolver_b200/csrc`.  There is no fallback: if the shared object is
missing, or no CUDA device is usable, the import / context creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeigensolver_b200.so")

ESB_OK, ESB_ERR_ARG, ESB_ERR_CUDA, ESB_ERR_CAPACITY, ESB_ERR_ALLOC = 0, -1, -2, -3, -4
SLAB_DENSITY, CYLINDER_DENSITY, SLAB_FLOW, CYLINDER_ROTATION, CYLINDER_FLOW = 0, 1, 2, 3, 4
RK4, RK8, RK8N = 0, 1, 2
OMEGA_SHARED, OMEGA_PHASE_SPEED, OMEGA_PER_K = 0, 1, 2
MESH_CLUSTERED, MESH_UNIFORM, MESH_GRADED = 0, 1, 2
ACCEPT_CONVERGED, ACCEPT_REFERENCE, ACCEPT_REFERENCE_SLAB = 0, 1, 2


class EsbError(RuntimeError):
    pass


class esb_model(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scheme", C.c_int32), ("n_steps", C.c_int32), ("mesh", C.c_int32),
        ("c_i0", C.c_double), ("vA_i0", C.c_double), ("vA_e", C.c_double), ("c_e", C.c_double),
        ("gamma", C.c_double), ("rho_i0", C.c_double), ("rho_A", C.c_double),
        ("ext_ic_value", C.c_double), ("ext_ic_slope", C.c_double), ("ext_wavelengths", C.c_double),
        ("s_start", C.c_double), ("s_end", C.c_double), ("U_e", C.c_double),
        ("r_sign", C.c_int32), ("reserved", C.c_int32),
        ("mesh_axis", C.c_double), ("mesh_edge", C.c_double), ("mesh_edge_width", C.c_double),
    ]


class esb_roots(C.Structure):
    _fields_ = [
        ("k_index", C.POINTER(C.c_int32)), ("w_index", C.POINTER(C.c_int32)),
        ("omega", C.POINTER(C.c_double)), ("ext", C.POINTER(C.c_double)),
        ("intq", C.POINTER(C.c_double)), ("accepted", C.POINTER(C.c_int32)),
        ("iterations", C.POINTER(C.c_int32)),
    ]


class esb_scan_result(C.Structure):
    _fields_ = [
        ("n_entries", C.c_int32),
        ("model", C.POINTER(C.c_int32)), ("slot", C.POINTER(C.c_int32)), ("k_index", C.POINTER(C.c_int32)),
        ("w_index", C.POINTER(C.c_int32)), ("accepted", C.POINTER(C.c_int32)),
        ("iterations", C.POINTER(C.c_int32)),
        ("omega", C.POINTER(C.c_double)), ("ext", C.POINTER(C.c_double)), ("intq", C.POINTER(C.c_double)),
    ]


class esb_guard_report(C.Structure):
    _fields_ = [
        ("worst", C.c_double), ("threshold", C.c_double),
        ("slot", C.c_int32), ("k_index", C.c_int32), ("w_index", C.c_int32), ("stride", C.c_int32),
        ("n_checked", C.c_int64), ("n_above", C.c_int64),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_ctx = C.c_void_p

#: every symbol include/eigensolver_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "esb_version": (C.c_int, []),
    "esb_sizeof_model": (C.c_int, []),
    "esb_model_defaults": (C.c_int, [C.c_int32, C.POINTER(esb_model)]),
    "esb_mesh_size": (C.c_int, [C.POINTER(esb_model), _ip]),
    "esb_mesh_nodes": (C.c_int, [C.POINTER(esb_model), _dp]),
    "esb_model_n_fields": (C.c_int, [C.POINTER(esb_model), _ip]),
    "esb_model_max_steps": (C.c_int, [C.POINTER(esb_model), _ip]),
    "esb_create": (C.c_int, [C.c_int32, C.POINTER(_ctx)]),
    "esb_destroy": (C.c_int, [_ctx]),
    "esb_last_error": (C.c_char_p, [_ctx]),
    "esb_set_model": (C.c_int, [_ctx, C.POINTER(esb_model), _dp, _dp, C.c_int32, C.c_double]),
    "esb_set_model_fields": (C.c_int, [_ctx, C.POINTER(esb_model), C.POINTER(_dp), C.c_int32, C.c_int32, _dp,
                                       C.c_int32]),
    "esb_dispersion_grid": (C.c_int, [_ctx, C.c_int32, _dp, C.c_int32, _dp, C.c_int32, C.c_int32, _dp, _dp]),
    "esb_find_roots": (C.c_int, [_ctx, C.c_int32, _dp, C.c_int32, _dp, C.c_int32, C.c_int32, C.c_double,
                                 C.c_int32, C.POINTER(esb_roots), _ip, _ip]),
    "esb_dispersion_grid_dev": (C.c_int, [_ctx, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "esb_brackets_dev": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int32, _ip, C.c_void_p]),
    "esb_upload_axes": (C.c_int, [_ctx, _dp, C.c_int32, _dp, C.c_int32, C.c_int32]),
    "esb_sweep_resident": (C.c_int, [_ctx, C.c_int32, C.c_double, _ip, _ip]),
    "esb_download_roots": (C.c_int, [_ctx, C.POINTER(esb_roots), C.c_int32]),
    "esb_roots_device": (C.c_int, [_ctx, C.c_int32, C.POINTER(esb_roots), _ip]),
    "esb_dispersion_grid_multi": (C.c_int, [_ctx, C.c_int32, _ip, _dp, C.c_int32, _dp, C.c_int32, C.c_int32,
                                            _dp, _dp]),
    "esb_sweep_resident_multi": (C.c_int, [_ctx, C.c_int32, _ip, C.c_double, _ip, _ip]),
    "esb_download_roots_slot": (C.c_int, [_ctx, C.c_int32, C.POINTER(esb_roots), C.c_int32]),
    "esb_scan_models": (C.c_int, [_ctx, C.c_int32, C.POINTER(esb_model), C.POINTER(_dp), C.c_int32, C.c_int32, _dp,
                                  C.c_int32, _ip, C.c_double, C.c_int32, C.c_int32, _ip,
                                  C.POINTER(esb_scan_result)]),
    "esb_scan_device": (C.c_int, [_ctx, C.POINTER(esb_scan_result)]),
    "esb_roots_pinned": (C.c_int, [_ctx, C.c_int32, C.POINTER(esb_roots), _ip]),
    "esb_set_stream": (C.c_int, [_ctx, C.c_void_p]),
    "esb_tables_wait": (C.c_int, [_ctx, C.c_void_p]),
    "esb_pack_modes_dev": (C.c_int, [_ctx, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.c_int32, C.c_void_p]),
    "esb_set_accept_rule": (C.c_int, [_ctx, C.c_int32]),
    "esb_set_guard_fields": (C.c_int, [_ctx, C.POINTER(esb_model), C.POINTER(_dp), C.c_int32, C.c_int32, _dp,
                                       C.c_int32, C.c_int32, C.c_double]),
    "esb_guard_result": (C.c_int, [_ctx, C.POINTER(esb_guard_report)]),
    "esb_set_schedule": (C.c_int, [_ctx, C.c_int32]),
    "esb_fp64_peak": (C.c_int, [_ctx, _dp]),
    "esb_rk_selftest": (C.c_int, [C.c_int32, C.c_int32, C.c_double, _dp]),
    "esb_bessel_ik_scaled": (C.c_int, [C.c_int32, C.c_double, _dp]),
    "esb_bessel_jy": (C.c_int, [C.c_int32, C.c_double, _dp]),
    "esb_bessel_jy_dev": (C.c_int, [_ctx, C.c_int32, _dp, C.c_int32, _dp]),
    "esb_exterior_leaky": (C.c_int, [C.POINTER(esb_model), C.c_int32, C.c_double, C.c_double, _dp]),
    "esb_exterior_leaky_dev": (C.c_int, [_ctx, C.c_int32, _dp, _dp, C.c_int32, _dp]),
    "esb_dispersion_grid_leaky": (C.c_int, [_ctx, C.c_int32, _ip, _dp, C.c_int32, _dp, C.c_int32, C.c_int32,
                                            _dp, _dp]),
    "esb_last_kernel_ms": (C.c_double, [_ctx]),
    "esb_launch_count": (C.c_int64, [_ctx]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises EsbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EsbError(
            "eigensolver_b200: %s not found - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C eigensolver_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(lib, ctx, rc, what):
    if rc == ESB_OK:
        return
    msg = lib.esb_last_error(ctx).decode() if ctx else ""
    raise EsbError("%s failed (status %d): %s" % (what, rc, msg))
