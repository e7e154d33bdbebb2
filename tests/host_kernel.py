"""TEST INFRASTRUCTURE: the kernels' arithmetic on the CPU.

tests/host_harness/host_eval.cpp compiles csrc/core.cuh (the ESB_HD device functions: exterior closed
forms, node coefficients, the Runge-Kutta / Nystrom step, the matching closures) and csrc/model_host.h
(mesh + staged table) for the host.  `evaluate` runs them at arbitrary (k, omega) points for a
`ModelSpec`, so the CPU suite checks the exact formulas the GPU executes against the oracle; the GPU
suite then only has to show that the device build agrees with the host build.  Never imported by the
product."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "host_harness")
_LIB = os.path.join(_DIR, "_build", "libesb_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _DIR])          # no-op when up to date
        l = C.CDLL(_LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        l.esbh_eval_points.restype = C.c_int
        l.esbh_eval_points.argtypes = [C.c_void_p, C.POINTER(dp), C.c_int32, C.c_int32, dp, C.c_int32, C.c_int32, ip,
                                       C.c_int64, dp, dp, dp, dp, dp]
        l.esbh_eval_points_leaky.restype = C.c_int
        l.esbh_eval_points_leaky.argtypes = l.esbh_eval_points.argtypes
        l.esbh_guard_points.restype = C.c_int
        l.esbh_guard_points.argtypes = [C.c_void_p, C.POINTER(dp), C.c_int32, C.c_int32, dp,
                                        C.c_void_p, C.POINTER(dp), C.c_int32, C.c_int32, dp, C.c_int32,
                                        C.c_int32, ip, C.c_int64, dp, dp, C.c_double, dp]
        _lib = l
    return _lib


def guard_grid(spec, fine, modes, k, W, margin=0.15):
    """The discretisation guard's judgement (csrc/esb.cu guard_kernel: core.cuh guard_deviation behind the
    resonance_free filter) of `spec` against `fine` on a phase-speed grid: [n_modes, nk, nw], NaN where the guard
    does not judge the point."""
    k = np.asarray(k, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    kk = np.ascontiguousarray(np.repeat(k, W.size))
    ww = np.ascontiguousarray((k[:, None] * W[None, :]).ravel())
    md = np.asarray(list(modes), dtype=np.int32)
    dev = np.empty((md.size, kk.size))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    a, b = spec.abi_args(), fine.abi_args()
    rc = lib().esbh_guard_points(C.cast(a[0], C.c_void_p), a[1], a[2], a[3], a[4],
                                 C.cast(b[0], C.c_void_p), b[1], b[2], b[3], b[4], a[5],
                                 md.size, md.ctypes.data_as(C.POINTER(C.c_int32)), kk.size, dp(kk), dp(ww),
                                 float(margin), dp(dev))
    if rc:
        raise RuntimeError("esbh_guard_points: status %d" % rc)
    return dev.reshape(md.size, k.size, W.size)


def evaluate(spec, modes, k, w, leaky=False):
    """(ext, int, den), each [n_modes, n]: D = ext - int at the points (k[j], w[j]) (w = omega).
    3 or 2 modes take the fused evaluation of the scan kernel, any other count one mode at a time."""
    k = np.ascontiguousarray(np.broadcast_to(np.asarray(k, dtype=np.float64), np.broadcast(k, w).shape).ravel())
    w = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.float64), k.shape).ravel()) \
        if np.ndim(w) == 0 or np.shape(w) != k.shape else np.ascontiguousarray(np.asarray(w, dtype=np.float64).ravel())
    md = np.asarray(list(modes), dtype=np.int32)
    out = [np.empty((md.size, k.size)) for _ in range(3)]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    args = spec.abi_args()
    model_ptr = C.cast(args[0], C.c_void_p)
    fn = lib().esbh_eval_points_leaky if leaky else lib().esbh_eval_points
    rc = fn(model_ptr, args[1], args[2], args[3], args[4], args[5], md.size,
                                md.ctypes.data_as(C.POINTER(C.c_int32)), k.size, dp(k), dp(w), dp(out[0]),
                                dp(out[1]), dp(out[2]))
    if rc:
        raise RuntimeError("esbh_eval_points%s: status %d" % ("_leaky" if leaky else "", rc))
    return out


def grid(spec, modes, k, W):
    """phase-speed grid like DispersionSolver.dispersion_grid_multi: arrays [n_modes, nk, nw]."""
    k = np.asarray(k, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    kk = np.repeat(k, W.size)
    ww = (k[:, None] * W[None, :]).ravel()
    e, i, d = evaluate(spec, modes, kk, ww)
    shp = (len(list(modes)), k.size, W.size)
    return e.reshape(shp), i.reshape(shp), d.reshape(shp)
