"""Execute the UNMODIFIED reference solver functions in this container.

Test infrastructure only.  This is the harness that `make_golden.py` uses to
obtain outputs of the reference itself.  It can only run where
`/root/reference` exists (the build container); nothing in `tests/ -m gpu`,
`smoke()` or `bench.py` imports it.

The reference solvers are top-level scripts, not importable modules:
  * they plot with matplotlib (absent here)            -> stubbed with MagicMock
  * they pass float sample counts to `np.linspace`     -> rejected by numpy>=1.18;
    the harness wraps `np.linspace` so `num` is cast with int() (what old numpy did)
  * the tail of each file is the multiprocessing driver + plotting; the harness
    executes the source only up to the `wavenumber = np.linspace(...)` driver line,
    so the physics set-up and the `sausage(...)` / `kink(...)` functions are the
    reference's own code, byte for byte.

What is harvested: the reference appends `left - inside` (its dispersion
function D(omega,k)) to a module-global list on every evaluation
(`xi_diff_check` for the cylinder, `P_diff_check[_kink]` for the slab).  Calling
`kink(k, q, q, np.array([w]))` with ONE frequency evaluates exactly one D and
cannot trigger the bisection recursion (the first sign product is with the
initial 0 entry).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types
from unittest import mock

import numpy as np

REF_ROOT = "/root/reference"

SOLVERS = {
    # name: (path, driver-line marker, {mode: (function, D list global)})
    "cylinder_density_coronal": (
        "Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py",
        "wavenumber = np.linspace(0.01,4.5,90)",
        {"sausage": ("sausage", "xi_diff_check"), "kink": ("kink", "xi_diff_check")},
    ),
    "slab_density_coronal": (
        "Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py",
        "wavenumber = np.linspace(0.001,0.75, 25.)",
        {"sausage": ("sausage", "P_diff_check"), "kink": ("kink", "P_diff_check_kink")},
    ),
    "cylinder_density_photospheric": (
        "Cylinder/Non-uniform density/Photospheric/Solvers/Density_cylinder_photospheric.py",
        "wavenumber = np.linspace(0.01,4.5,130)",
        {"sausage": ("sausage", "xi_diff_check"), "kink": ("kink", "xi_diff_check")},
    ),
    "slab_density_photospheric": (
        "Slab/Non uniform density/Photospheric/Solvers/multiprocessor_Inhomogeneous_method.py",
        "wavenumber = np.linspace(",
        {"sausage": ("sausage", "P_diff_check"), "kink": ("kink", "P_diff_check_kink")},
    ),
    "cylinder_rotation_kink": (
        "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_fast.py",
        "wavenumber = np.linspace(0.25,0.37,20)",
        {"kink": ("kink", "xi_diff_check")},
    ),
    "cylinder_rotation_kink_slow": (
        "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_slow.py",
        "wavenumber = np.linspace(0.01,0.5,60)",
        {"kink": ("kink", "xi_diff_check")},
    ),
    "cylinder_rotation_sausage": (
        "Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_flow_sausage.py",
        "wavenumber = np.linspace(0.75,4.,110.)",
        {"sausage": ("sausage", "xi_diff_check")},
    ),
    "cylinder_flow_coronal": (
        "Cylinder/Non-uniform flow/Coronal/solvers/Cylinder_method_flow_testing.py",
        "wavenumber = np.linspace(0.01,4.,150.)",
        {"sausage": ("sausage", "xi_diff_check"), "kink": ("kink", "xi_diff_check")},
    ),
    "slab_flow_photospheric": (
        # the steady (uniform) flow slab: U_i = 0 inside, U_e = -0.15 outside, vA_e = 0, 7 wavelengths
        "Slab/Non uniform flow/Solver/flow_multiprocessor.py",
        "wavenumber = np.linspace(0.01,3.5,350)",
        {"sausage": ("sausage", "P_diff_check"), "kink": ("kink", "P_diff_check_kink")},
    ),
    "slab_flow_coronal": (
        "Slab/Non uniform flow/Solver/flow_multiprocessor_coronal.py",
        "wavenumber = np.linspace(",
        {"sausage": ("sausage", "P_diff_check"), "kink": ("kink", "P_diff_check_kink")},
    ),
}


class _Sink:
    """Stands in for multiprocessing.Queue: the reference only calls .put()."""

    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(list(x))


def _stub_matplotlib():
    names = [
        "matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.gridspec",
        "matplotlib.patches",
    ]
    for n in names:
        if n not in sys.modules:
            m = mock.MagicMock(name=n)
            m.__spec__ = None
            sys.modules[n] = m
    # `from matplotlib import animation` needs attribute access to resolve
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]


@contextlib.contextmanager
def _legacy_linspace(max_interior=None):
    """np.linspace accepting float `num`, as the numpy the reference was written for."""
    orig = np.linspace

    def linspace(start, stop, num=50, *a, **kw):
        num = int(num)
        if max_interior is not None and num > max_interior:
            num = max_interior
        return orig(start, stop, num, *a, **kw)

    np.linspace = linspace
    try:
        yield
    finally:
        np.linspace = orig


@contextlib.contextmanager
def _legacy_odeint():
    """scipy.integrate.odeint accepting y0 = [float, array([x])].

    fsolve hands the objective a shape-(1,) array; the numpy the reference was
    written for coerced the resulting ragged y0 to two floats, numpy>=1.24 raises.
    The shim does that coercion and nothing else."""
    import scipy.integrate as si
    orig = si.odeint

    def odeint(func, y0, t, *a, **kw):
        y0 = [complex(np.asarray(v).reshape(-1)[0]) if np.iscomplexobj(v)
              else float(np.asarray(v).reshape(-1)[0]) for v in y0]
        return orig(func, y0, t, *a, **kw)

    si.odeint = odeint
    try:
        yield
    finally:
        si.odeint = orig


class ReferenceSolver:
    """The reference's physics set-up + sausage()/kink() functions, exec'd as-is."""

    def __init__(self, name, overrides=None, max_interior=None, patches=None):
        """patches: (old, new) source replacements for edits that are not single assignment lines -
        e.g. switching to one of the alternative profile definitions the script carries as comments."""
        path, marker, modes = SOLVERS[name]
        self.name = name
        self.modes = modes
        self.max_interior = max_interior
        with open(os.path.join(REF_ROOT, path)) as fh:
            src = fh.read()
        if name in ("slab_flow_coronal", "slab_density_photospheric"):
            cut = src.find(marker, src.find("kink_ws.put(sol_omegas_kink1)"))
        else:
            cut = src.rfind(marker)
        assert cut > 0, "driver marker not found"
        src = src[:cut]
        for old, new in (patches or []):
            assert src.count(old) == 1, old
            src = src.replace(old, new)
        # parameter overrides (e.g. profile width) are applied by rewriting the
        # single assignment line, exactly what a user of the scripts edits by hand.
        for key, val in (overrides or {}).items():
            import re
            pat = re.compile(r"^%s\s*=.*$" % re.escape(key), re.M)
            assert pat.search(src), key
            src = pat.sub("%s = %r" % (key, val), src, count=1)
        _stub_matplotlib()
        self.ns = {"__name__": "reference_solver_" + name}
        with _legacy_linspace(max_interior), _legacy_odeint(), open(os.devnull, "w") as dn, \
                contextlib.redirect_stdout(dn):
            exec(compile(src, path, "exec"), self.ns)
        if "odeintz" in self.ns:
            # the rotational-flow scripts wrap odeint for complex values (odeintz, :51-72) and
            # build np.array([P_b, dPi]) themselves: same ragged-y0 coercion as _legacy_odeint
            orig_z = self.ns["odeintz"]

            def odeintz(func, z0, t, **kw):
                z0 = [complex(np.asarray(v).reshape(-1)[0]) for v in z0]
                return orig_z(func, z0, t, **kw)

            self.ns["odeintz"] = odeintz

    def D(self, mode, k, w):
        """One evaluation of the reference dispersion function at (k, w).

        Returns nan when the reference skips the point (m_e < 0)."""
        fn, lst = self.modes[mode]
        store = self.ns[lst]
        n0 = len(store)
        with _legacy_linspace(self.max_interior):
            self.ns[fn](float(k), _Sink(), _Sink(), np.array([float(w)]))
        if len(store) == n0:
            return float("nan")
        val = store[n0]
        val = complex(val).real if abs(complex(val).imag) == 0.0 else complex(val)
        # keep the running lists short so memory stays flat
        del store[1:]
        for g in ("sign_check", "sign_check_kink", "sign_check_sausage", "all_ws", "all_ks",
                  "all_ws_kink", "all_ks_kink", "loop_ws"):
            if g in self.ns and isinstance(self.ns[g], list):
                self.ns[g][:] = [0] if g.startswith("sign") else []
        return float(val)

    def roots(self, mode, k, freq):
        """Run the reference's own scan + bisection over `freq` at one k."""
        fn, _ = self.modes[mode]
        qs_w, qs_k = _Sink(), _Sink()
        with _legacy_linspace(self.max_interior):
            self.ns[fn](float(k), qs_w, qs_k, np.asarray(freq, dtype=float))
        ws = list(qs_w.items[0])
        ks = list(qs_k.items[0])
        # the result lists are module globals that keep growing: reset them
        for g in ("sol_omegas", "sol_ks", "sol_omegas_kink", "sol_ks_kink", "sol_omegas1",
                  "sol_ks1", "sol_omegas_kink1", "sol_ks_kink1"):
            if g in self.ns and isinstance(self.ns[g], list):
                self.ns[g][:] = []
        return np.array(ks), np.array(ws)


if __name__ == "__main__":
    import time
    s = ReferenceSolver("cylinder_density_coronal")
    t = time.time()
    print(s.D("kink", 1.0, 3.0), time.time() - t)
    t = time.time()
    print(s.D("sausage", 1.0, 3.0), time.time() - t)
