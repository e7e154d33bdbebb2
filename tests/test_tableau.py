"""The Runge-Kutta tableaux the kernels integrate with (csrc/core.cuh), exercised on the host
through esb_rk_selftest: Cooper-Verner must converge with order 8, the classical scheme with
order 4, to the mpmath solution of a non-autonomous linear test equation."""
import ctypes as C

import mpmath as mp
import numpy as np

import eigensolver_b200 as esb
from eigensolver_b200 import _lib as L


def _reference(T):
    mp.mp.dps = 30
    sol = mp.odefun(lambda t, y: [y[1], mp.sin(t) * y[1] - (1 + t * t) * y[0]], 0, [mp.mpf(1), mp.mpf("0.3")])
    return np.array([float(v) for v in sol(T)])


def _run(scheme, n, T):
    lib = esb.load()
    out = (C.c_double * 2)()
    assert lib.esb_rk_selftest(scheme, n, T, out) == 0
    return np.array(out[:])


def test_rk8_is_eighth_order_and_rk4_fourth():
    T = 2.0
    ref = _reference(T)
    e8 = [np.abs(_run(L.RK8, n, T) - ref).max() for n in (4, 8, 16, 32)]
    slopes8 = [np.log2(e8[i] / e8[i + 1]) for i in range(3)]
    assert all(7.3 < s < 8.7 for s in slopes8), (e8, slopes8)
    assert e8[-1] < 1e-11
    e4 = [np.abs(_run(L.RK4, n, T) - ref).max() for n in (32, 64, 128, 256)]
    slopes4 = [np.log2(e4[i] / e4[i + 1]) for i in range(3)]
    assert all(3.7 < s < 4.3 for s in slopes4), (e4, slopes4)
    # converged values agree with each other and the reference
    assert np.abs(_run(L.RK8, 256, T) - ref).max() < 1e-13
    lib = esb.load()
    assert lib.esb_rk_selftest(7, 4, 1.0, (C.c_double * 2)()) == L.ESB_ERR_ARG
