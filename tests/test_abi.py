"""CPU tests of the C-ABI library: it loads, exports every symbol the header declares,
its host-side helpers are right, and it refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import special as sp

import eigensolver_b200 as esb
from eigensolver_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "eigensolver_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(esb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(esb.LIB_PATH)
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    # and the python binding table covers exactly the header
    assert sorted(L.SYMBOLS) == names


def test_version_and_defaults():
    lib = esb.load()
    assert lib.esb_version() == 130
    assert lib.esb_sizeof_model() == C.sizeof(L.esb_model)        # the ctypes mirror of the struct
    m = L.esb_model()
    assert lib.esb_model_defaults(L.CYLINDER_DENSITY, C.byref(m)) == 0
    # Density_cylinder.py:69-72,120,768
    assert (m.c_i0, m.vA_i0, m.vA_e, m.c_e) == (1.0, 2.0, 5.0, 0.5)
    assert (m.s_start, m.s_end) == (-1.0, -0.001)
    assert (m.ext_ic_value, m.ext_ic_slope) == (1e-8, 1e-15)
    assert lib.esb_model_defaults(L.SLAB_DENSITY, C.byref(m)) == 0
    # ..._coronal.py:69-72,91,247
    assert (m.c_i0, m.vA_i0, m.vA_e, m.c_e) == (1.0, 1.2, 3.0, 0.4)
    assert (m.s_start, m.s_end) == (-1.0, 1.0)
    assert (m.ext_ic_value, m.ext_ic_slope) == (1e-8, 1e-8)
    # every kind has defaults with a usable discretisation
    for kind in (L.SLAB_DENSITY, L.CYLINDER_DENSITY, L.SLAB_FLOW, L.CYLINDER_ROTATION, L.CYLINDER_FLOW):
        assert lib.esb_model_defaults(kind, C.byref(m)) == 0
        n = C.c_int32()
        assert lib.esb_mesh_size(C.byref(m), C.byref(n)) == 0 and n.value == 4 * m.n_steps + 1
        x = np.empty(n.value)
        assert lib.esb_mesh_nodes(C.byref(m), x.ctypes.data_as(C.POINTER(C.c_double))) == 0
        d = np.diff(x)
        assert np.all(d > 0) or np.all(d < 0)                      # monotone along the integration
        ends = sorted([x[0], x[-1]])
        assert ends == sorted([m.s_start, m.s_end])
        if kind in (L.SLAB_DENSITY, L.SLAB_FLOW):                  # exactly mirror-symmetric slab meshes
            assert np.array_equal(x - x[0], (x[-1] - x[::-1]))
    assert lib.esb_model_defaults(7, C.byref(m)) == L.ESB_ERR_ARG


@pytest.mark.parametrize("kind,scheme,nps", [(L.CYLINDER_DENSITY, L.RK8, 4), (L.SLAB_DENSITY, L.RK8, 4),
                                             (L.CYLINDER_DENSITY, L.RK4, 2), (L.SLAB_DENSITY, L.RK4, 2)])
def test_mesh_nodes(kind, scheme, nps):
    lib = esb.load()
    m = L.esb_model()
    lib.esb_model_defaults(kind, C.byref(m))
    m.scheme = scheme
    m.n_steps = 64
    n = C.c_int32()
    assert lib.esb_mesh_size(C.byref(m), C.byref(n)) == 0
    assert n.value == 64 * nps + 1
    x = np.empty(n.value)
    assert lib.esb_mesh_nodes(C.byref(m), x.ctypes.data_as(C.POINTER(C.c_double))) == 0
    d = np.diff(x)
    if kind == L.CYLINDER_DENSITY:       # integrated from the axis out to the boundary
        assert x[0] == -0.001 and x[-1] == -1.0 and (d < 0).all()
    else:                                # boundary -> far boundary, through the centre
        assert x[0] == -1.0 and x[-1] == 1.0 and (d > 0).all()
        assert x[(n.value - 1) // 2] == 0.0
    # stage fractions inside each step
    h = x[nps::nps] - x[:-1:nps]
    frac = (x[1:nps] - x[0]) / h[0]
    want = [0.5] if nps == 2 else [(7 - 21**0.5) / 14, 0.5, (7 + 21**0.5) / 14]
    assert np.allclose(frac, want, rtol=0, atol=1e-12)
    # odd step count is invalid for the slab (two half-layers)
    m.n_steps = 63
    rc = lib.esb_mesh_size(C.byref(m), C.byref(n))
    assert (rc == L.ESB_ERR_ARG) == (kind == L.SLAB_DENSITY)


def test_bessel_helper_matches_scipy():
    worst = 0.0
    zs = list(np.logspace(-6, 2.85, 300)) + [1.9999, 2.0, 2.0001, 11.999, 12.0, 12.001, 700.0]
    for n in range(4):
        for z in zs:
            got = esb.bessel_ik_scaled(n, z)
            want = (sp.ive(n, z), sp.ive(n + 1, z) + n / z * sp.ive(n, z), sp.kve(n, z),
                    -sp.kve(n + 1, z) + n / z * sp.kve(n, z))
            worst = max(worst, max(abs(a - b) / abs(b) for a, b in zip(got, want)))
    assert worst < 5e-14, worst
    with pytest.raises(esb.EsbError):
        esb.bessel_ik_scaled(4, 1.0)
    with pytest.raises(esb.EsbError):
        esb.bessel_ik_scaled(0, 0.0)


def test_bessel_jy_matches_scipy():
    """J_n, Y_n and their derivatives (the leaky side the reference skips): error relative to the local
    amplitude sqrt(J^2 + Y^2) - relative accuracy is meaningless next to a zero of an oscillating function."""
    worst = 0.0
    xs = list(np.geomspace(1e-3, 5.0, 150)) + list(np.linspace(4.9, 8.1, 200)) + list(np.geomspace(8.0, 5e3, 300)) + \
        [4.9999, 5.0, 5.0001, 7.9999, 8.0, 8.0001]
    for n in range(4):
        for x in xs:
            J, dJ, Y, dY = esb.bessel_jy(n, x)
            want = (sp.jv(n, x), sp.jvp(n, x), sp.yv(n, x), sp.yvp(n, x))
            amp, damp = np.hypot(want[0], want[2]), np.hypot(want[1], want[3])
            worst = max(worst, abs(J - want[0]) / amp, abs(Y - want[2]) / amp, abs(dJ - want[1]) / damp,
                        abs(dY - want[3]) / damp)
    assert worst < 5e-15, worst
    with pytest.raises(esb.EsbError):
        esb.bessel_jy(4, 1.0)
    with pytest.raises(esb.EsbError):
        esb.bessel_jy(0, 0.0)


def test_leaky_exterior_closed_form_matches_integration():
    """esb_exterior_leaky = the reference's exterior initial-value problem (Density_cylinder.py:765-770:
    P'' + P'/r - (m_e + n^2/r^2) P = 0 from r = -3*2pi/k with P0 = [1e-8, 1e-15]) where m_e < 0 - the points
    its scan loop skips - against a tight numerical integration of the same problem."""
    from scipy.integrate import solve_ivp
    lib = esb.load()
    m = L.esb_model()
    lib.esb_model_defaults(L.CYLINDER_DENSITY, C.byref(m))
    out = (C.c_double * 2)()
    vAe2, ce2 = m.vA_e**2, m.c_e**2
    cTe2 = ce2 * vAe2 / (ce2 + vAe2)
    n_checked = 0
    for n in range(4):
        for k, W in ((0.5, 5.3), (1.0, 6.0), (2.5, 5.05), (1.5, 0.499), (3.0, 0.4985)):   # above vA_e; in (cT_e, c_e) = (0.4975, 0.5)
            w = k * W
            me = (k * k * vAe2 - w * w) * (k * k * ce2 - w * w) / ((vAe2 + ce2) * (k * k * cTe2 - w * w))
            assert me < 0
            assert lib.esb_exterior_leaky(C.byref(m), n, k, w, out) == 0
            r0 = -3.0 * 2.0 * np.pi / k
            sol = solve_ivp(lambda r, y: [y[1], -y[1] / r + (me + n * n / (r * r)) * y[0]], (r0, -1.0),
                            [m.ext_ic_value, m.ext_ic_slope], method="DOP853", rtol=1e-12, atol=1e-30)
            P, dP = sol.y[0, -1], sol.y[1, -1]
            amp = np.hypot(P, dP / np.sqrt(-me))
            assert abs(out[0] - P) < 1e-9 * amp and abs(out[1] - dP) < 1e-9 * amp * np.sqrt(-me), (n, k, W)
            n_checked += 1
    assert n_checked == 20
    assert lib.esb_exterior_leaky(C.byref(m), 0, 1.0, 3.0, out) == 0 and np.isnan(out[0])     # m_e >= 0 there


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    lib = esb.load()
    ctx = L._ctx()
    assert lib.esb_create(0, C.byref(ctx)) == L.ESB_ERR_CUDA
    assert not ctx.value
    with pytest.raises(esb.EsbError, match="no CPU fallback"):
        esb.DispersionSolver("cylinder_density")


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "eigensolver_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "scipy" not in txt, f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/eigensolver_b200.h compiled as strict C99 (-pedantic, warnings are errors) into a program that
    links libeigensolver_b200.so and calls its host-side entry points; without a GPU esb_create refuses."""
    import subprocess
    exe = str(tmp_path / "host_only")
    libdir = os.path.dirname(esb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_c", "host_only.c"), "-o", exe,
                           "-L", libdir, "-leigensolver_b200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode().split()
    version, n_steps, n_nodes, n_fields, max_steps, rc = (int(v) for v in out)
    assert (version, n_steps, n_nodes, n_fields, max_steps) == (130, 152, 4 * 152 + 1, 3, 710)
    import torch
    assert rc == (0 if torch.cuda.is_available() else L.ESB_ERR_CUDA)
