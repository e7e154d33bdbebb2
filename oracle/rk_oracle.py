"""ORACLE (test infrastructure): ctypes binding of oracle/dispersion_rk.c.

See the header of dispersion_rk.c for what it restates and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "liboracle_rk.so")


class ork_model(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_ext", C.c_int), ("n_int", C.c_int), ("leaky", C.c_int),
                ("c_i0", C.c_double), ("vA_i0", C.c_double), ("vA_e", C.c_double), ("c_e", C.c_double),
                ("gamma", C.c_double), ("rho_i0", C.c_double), ("rho_A", C.c_double),
                ("width", C.c_double), ("x0", C.c_double), ("ic_v", C.c_double), ("ic_s", C.c_double),
                ("ext_wavelengths", C.c_double), ("s_start", C.c_double), ("s_end", C.c_double),
                ("U_i0", C.c_double), ("U_e", C.c_double), ("r_sign", C.c_double),
                ("v_twist", C.c_double), ("power", C.c_double), ("profile_kind", C.c_int), ("pad2", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        dp = C.POINTER(C.c_double)
        l.ork_point.restype = C.c_int
        l.ork_point.argtypes = [C.POINTER(ork_model), C.c_int, C.c_double, C.c_double, dp, dp]
        l.ork_grid.restype = None
        l.ork_grid.argtypes = [C.POINTER(ork_model), C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, dp, dp, C.c_int]
        l.ork_refine.restype = C.c_double
        l.ork_refine.argtypes = [C.POINTER(ork_model), C.c_int, C.c_double, C.c_double, C.c_double, dp]
        _lib = l
    return _lib


def make_model(kind, medium=None, width=None, x0=0.0, n_ext=None, n_int=None, rho_A=1.0,
               coordinate="negative", ext_wavelengths=3.0, v_twist=0.25, power=0.8, s_end=None,
               profile="gaussian"):
    """kind: 'slab_density' | 'cylinder_density' | 'slab_flow'; medium: any object with c_i0, vA_i0,
    vA_e, c_e, gamma, rho_i0 attributes (flow: vA_i, c_i, vA_e, c_e, U_i0, U_e, gamma, rho_i);
    defaults: the reference's coronal sets.  coordinate='positive': cylinder scripts in r > 0."""
    m = ork_model()
    cyl = kind == "cylinder_density"
    m.kind = {"slab_density": 0, "cylinder_density": 1, "slab_flow": 2, "cylinder_rotation": 3,
              "cylinder_flow": 4}[kind]
    m.r_sign = -1.0
    m.profile_kind = {"gaussian": 0, "epstein": 1}[profile]
    if kind == "cylinder_flow":
        # Cylinder_method_flow_testing.py:66-69 (coronal speeds), r < 0, P0 = [1e-8, 1e-8] (:774)
        md = medium
        vals = (md.c_i0, md.vA_i0, md.vA_e, md.c_e, md.gamma, md.rho_i0, md.U_i0, md.U_e) if md is not None else (
            1.0, 2.0, 5.0, 0.5, 5.0 / 3.0, 1.0, 0.35, 0.0)
        m.c_i0, m.vA_i0, m.vA_e, m.c_e, m.gamma, m.rho_i0, m.U_i0, m.U_e = vals
        m.rho_A = 1.0
        m.width = width if width is not None else getattr(md, "width", 1.0)
        m.x0 = x0
        m.ic_v, m.ic_s = 1e-8, 1e-8
        m.ext_wavelengths = ext_wavelengths
        m.s_start, m.s_end = -1.0, (s_end if s_end is not None else -0.001)
        m.n_ext = n_ext or 6000
        m.n_int = n_int or 320
        return m
    if kind == "cylinder_rotation":
        # Twisted_photospheric_nonlinear_flow_kink_fast.py:73-76 (photospheric speeds), r > 0
        md = medium
        vals = (md.c_i0, md.vA_i0, md.vA_e, md.c_e, md.gamma, md.rho_i0) if md is not None else (
            1.0, 2.0, 0.5, 1.5, 5.0 / 3.0, 1.0)
        m.c_i0, m.vA_i0, m.vA_e, m.c_e, m.gamma, m.rho_i0 = vals
        m.rho_A, m.width, m.x0 = 1.0, 1.0, 0.0
        m.ic_v, m.ic_s = 1e-8, 1e-8
        m.ext_wavelengths = ext_wavelengths
        m.r_sign, m.s_start = 1.0, 1.0
        m.s_end = s_end if s_end is not None else 0.001
        m.v_twist, m.power = v_twist, power
        m.n_ext = n_ext or 6000
        m.n_int = n_int or 320
        return m
    if kind == "slab_flow":
        vals = medium if medium is not None else type("M", (), dict(
            vA_i=1.0, c_i=0.3, vA_e=2.5, c_e=0.2, U_i0=0.9, U_e=0.0, gamma=5.0 / 3.0, rho_i=1.0))
        m.c_i0, m.vA_i0, m.vA_e, m.c_e = vals.c_i, vals.vA_i, vals.vA_e, vals.c_e
        m.gamma, m.rho_i0, m.U_i0, m.U_e = vals.gamma, vals.rho_i, vals.U_i0, vals.U_e
        m.rho_A = 1.0
        m.width = width if width is not None else 1e5
        m.x0 = x0
        m.ic_v, m.ic_s = 1e-8, 1e-15
        m.ext_wavelengths = ext_wavelengths
        m.s_start, m.s_end = -1.0, 1.0
        m.n_ext = n_ext or 3000
        m.n_int = n_int or 384
        return m
    if medium is None:
        vals = (1.0, 2.0, 5.0, 0.5) if cyl else (1.0, 1.2, 3.0, 0.4)
        m.c_i0, m.vA_i0, m.vA_e, m.c_e = vals
        m.gamma, m.rho_i0 = 5.0 / 3.0, 1.0
    else:
        m.c_i0, m.vA_i0, m.vA_e, m.c_e = medium.c_i0, medium.vA_i0, medium.vA_e, medium.c_e
        m.gamma, m.rho_i0 = medium.gamma, medium.rho_i0
    m.rho_A = rho_A
    m.width = width if width is not None else (0.95 if cyl else 0.9)
    m.x0 = x0
    m.ic_v = 1e-8
    m.ic_s = 1e-15 if cyl else 1e-8
    m.ext_wavelengths = ext_wavelengths
    m.s_start = -1.0
    m.s_end = -0.001 if cyl else 1.0
    if coordinate == "positive":
        assert cyl
        m.r_sign, m.s_start, m.s_end, m.ic_s = 1.0, 1.0, 0.001, 1e-8
    m.n_ext = n_ext or int((6000 if cyl else 3000) * max(1.0, ext_wavelengths / 3.0))
    m.n_int = n_int or (320 if cyl else 384)
    return m


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def point(model, mode, k, w):
    e, q = C.c_double(), C.c_double()
    lib().ork_point(C.byref(model), int(mode), float(k), float(w), C.byref(e), C.byref(q))
    return e.value, q.value


def grid(model, mode, k, w, layout="phase_speed", threads=None):
    lay = {"shared": 0, "phase_speed": 1, "per_k": 2}[layout]
    k = np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nw = w.shape[-1]
    ext = np.empty((k.size, nw))
    inq = np.empty((k.size, nw))
    lib().ork_grid(C.byref(model), int(mode), _dp(k), k.size, _dp(w), nw, lay, _dp(ext), _dp(inq),
                   int(threads or os.cpu_count() or 1))
    return ext, inq


def brackets(D):
    """Sign-change brackets along omega (axis 1): arrays (k_index, w_index), sorted."""
    d0, d1 = D[:, :-1], D[:, 1:]
    ok = np.isfinite(d0) & np.isfinite(d1) & (((d0 < 0) & (d1 > 0)) | ((d0 > 0) & (d1 < 0)))
    ki, wi = np.nonzero(ok)
    return ki.astype(np.int32), wi.astype(np.int32)


def refine(model, mode, k, wlo, whi):
    out = np.zeros(2)
    r = lib().ork_refine(C.byref(model), int(mode), float(k), float(wlo), float(whi), _dp(out))
    return r, out[0], out[1]
