"""Host side of the B200 dispersion-function path.

`DispersionSolver` owns one C-ABI context (one GPU).  It takes the reference's
inputs - the equilibrium speeds, a density profile, a k range and an omega range -
and returns D(omega,k) grids and per-k root tables.

Reference set-up mirrored here (file:line in /root/reference):
  equilibrium speeds / rho_e         Density_cylinder.py:69-80
  inverted-Gaussian density profile  Density_cylinder.py:124-154   (r0, dr)
                                     ..._coronal.py:93-102          (x0, dx)
No oracle / CPU code is used on this path; everything numerical happens in
libeigensolver_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import warnings

import numpy as np

from . import _lib as L


@dataclasses.dataclass(frozen=True)
class Medium:
    """Characteristic speeds (Density_cylinder.py:69-72)."""
    c_i0: float = 1.0
    vA_i0: float = 2.0
    vA_e: float = 5.0
    c_e: float = 0.5
    gamma: float = 5.0 / 3.0
    rho_i0: float = 1.0

    @property
    def rho_e(self):
        g = self.gamma
        return self.rho_i0 * (self.c_i0**2 + g * 0.5 * self.vA_i0**2) / (self.c_e**2 + g * 0.5 * self.vA_e**2)

    @property
    def cT_e(self):
        return math.sqrt(self.c_e**2 * self.vA_e**2 / (self.c_e**2 + self.vA_e**2))

    @property
    def cT_i0(self):
        return math.sqrt(self.c_i0**2 * self.vA_i0**2 / (self.c_i0**2 + self.vA_i0**2))

    @property
    def c_kink(self):
        re = self.rho_e
        return math.sqrt((self.rho_i0 * self.vA_i0**2 + re * self.vA_e**2) / (self.rho_i0 + re))


CYLINDER_CORONAL = Medium(1.0, 2.0, 5.0, 0.5)
CYLINDER_PHOTOSPHERIC = Medium(1.0, 2.0, 0.5, 1.5)
SLAB_CORONAL = Medium(1.0, 1.2, 3.0, 0.4)
SLAB_PHOTOSPHERIC = Medium(1.0, 1.9, 0.8, 1.3)


@dataclasses.dataclass(frozen=True)
class GaussianDensity:
    """rho_e + (rho_i0 - rho_e) exp(-(x-x0)^2/width^2)   (Density_cylinder.py:135).
    Returns (rho, rho', rho''): the normal-form scheme ("rk8n") needs the second derivative."""
    width: float = 0.95
    x0: float = 0.0

    def __call__(self, medium, x):
        x = np.asarray(x, dtype=np.float64)
        g = np.exp(-((x - self.x0) ** 2) / self.width**2)
        d = medium.rho_i0 - medium.rho_e
        t = -2.0 * (x - self.x0) / self.width**2
        return medium.rho_e + d * g, d * g * t, d * g * (t * t - 2.0 / self.width**2)


@dataclasses.dataclass(frozen=True)
class EpsteinDensity:
    """(rho_i0 - rho_e)/cosh((x-x0)/a)^8 + rho_e: the reference's alternative profile
    (Density_cylinder.py:139-142, `a` = inhomogeneity width).  Like any other callable
    (medium, x) -> (rho, rho') it is simply sampled at the mesh nodes."""
    a: float = 1.0
    x0: float = 0.0

    def __call__(self, medium, x):
        t = (np.asarray(x, dtype=np.float64) - self.x0) / self.a
        d = medium.rho_i0 - medium.rho_e
        ch, sh = np.cosh(t), np.sinh(t)
        return (d / ch**8 + medium.rho_e, d * (-8.0 / self.a) * sh / ch**9,
                d * (8.0 / self.a**2) * (9.0 * sh * sh - ch * ch) / ch**10)


@dataclasses.dataclass(frozen=True)
class FlowMedium:
    """Speeds of the slab flow script (flow_multiprocessor_coronal.py:47-56): uniform density and
    field inside the slab, exterior at rest or streaming with U_e."""
    vA_i: float = 1.0
    c_i: float = 0.3
    vA_e: float = 2.5
    c_e: float = 0.2
    U_i0: float = 0.9
    U_e: float = 0.0
    gamma: float = 5.0 / 3.0
    rho_i: float = 1.0

    @property
    def rho_e(self):
        g = self.gamma
        return self.rho_i * (self.c_i**2 + g * 0.5 * self.vA_i**2) / (self.c_e**2 + g * 0.5 * self.vA_e**2)

    @property
    def cT_i(self):
        return math.sqrt(self.c_i**2 * self.vA_i**2 / (self.c_i**2 + self.vA_i**2))

    @property
    def cT_e(self):
        return math.sqrt(self.c_e**2 * self.vA_e**2 / (self.c_e**2 + self.vA_e**2))


SLAB_FLOW_CORONAL = FlowMedium()


@dataclasses.dataclass(frozen=True)
class GaussianFlow:
    """U(x) = U_e + (U_i0 - U_e) exp(-(x-x0)^2/width^2)   (flow_multiprocessor_coronal.py:77)."""
    width: float = 1e5
    x0: float = 0.0

    def __call__(self, medium, x):
        x = np.asarray(x, dtype=np.float64)
        g = np.exp(-((x - self.x0) ** 2) / self.width**2)
        t = -2.0 * (x - self.x0) / self.width**2
        dU0 = medium.U_i0 - medium.U_e
        return medium.U_e + dU0 * g, dU0 * g * t, dU0 * g * (t * t - 2.0 / self.width**2)


@dataclasses.dataclass(frozen=True)
class AxialFlowMedium(Medium):
    """Speeds of the cylinder axial-flow script (Cylinder_method_flow_testing.py:66-69,130-131):
    uniform density, field and sound speed inside the tube, flow amplitude U_i0 on the axis, U_e
    far from it (the script's exterior itself is at rest: m_e and xi_e use omega unshifted)."""
    U_i0: float = 0.35
    U_e: float = 0.0


CYLINDER_FLOW_CORONAL = AxialFlowMedium(1.0, 2.0, 5.0, 0.5)


@dataclasses.dataclass(frozen=True)
class GaussianAxialFlow:
    """v_z(r) = U_e + (U_i0 - U_e) exp(-(r-r0)^2/width^2)   (Cylinder_method_flow_testing.py:134)."""
    width: float = 1e5
    r0: float = 0.0

    def __call__(self, medium, r):
        r = np.asarray(r, dtype=np.float64)
        g = np.exp(-((r - self.r0) ** 2) / self.width**2)
        dU0 = medium.U_i0 - medium.U_e
        t = -2.0 * (r - self.r0) / self.width**2
        return medium.U_e + dU0 * g, dU0 * g * t, dU0 * g * (t * t - 2.0 / self.width**2)


@dataclasses.dataclass(frozen=True)
class PowerLawRotation:
    """v_phi = v_twist r^power with the pressure that balances it,
    P_i = rho v_twist^2 r^(2 power)/(2 power) + P_0, c_i^2 = gamma P_i/rho
    (Twisted_photospheric_nonlinear_flow_kink_fast.py:105-111)."""
    v_twist: float = 0.25
    power: float = 0.8

    def __call__(self, medium, r):
        r = np.asarray(r, dtype=np.float64)
        v = self.v_twist * r**self.power
        dv = self.v_twist * self.power * r ** (self.power - 1.0)
        P0 = medium.c_i0**2 * medium.rho_i0 / medium.gamma
        Pi = medium.rho_i0 * self.v_twist**2 * r ** (2.0 * self.power) / (2.0 * self.power) + P0
        return v, dv, medium.gamma * Pi / medium.rho_i0


@dataclasses.dataclass
class RootTable:
    """Result of a root search.  `k`, `omega` of accepted modes are what the
    reference stores in sol_ks / sol_omegas.  `k_axis` is the wavenumber axis the table's
    `k_index` refers to; `k` (= k_axis[k_index]) is evaluated on access."""
    k_index: np.ndarray
    w_index: np.ndarray
    k_axis: np.ndarray
    omega: np.ndarray
    ext: np.ndarray
    intq: np.ndarray
    accepted: np.ndarray
    iterations: np.ndarray
    n_brackets: int = 0

    @property
    def k(self):
        return self.k_axis[self.k_index]

    def modes(self):
        m = self.accepted.astype(bool)
        return self.k_axis[self.k_index[m]], self.omega[m]


_KINDS = {"slab_density": L.SLAB_DENSITY, "cylinder_density": L.CYLINDER_DENSITY, "slab_flow": L.SLAB_FLOW,
          "cylinder_rotation": L.CYLINDER_ROTATION, "cylinder_flow": L.CYLINDER_FLOW}
_SCHEMES = {"rk4": L.RK4, "rk8": L.RK8, "rk8n": L.RK8N}
_SCHEME_NAMES = {v: k for k, v in _SCHEMES.items()}
_LAYOUTS = {"shared": L.OMEGA_SHARED, "phase_speed": L.OMEGA_PHASE_SPEED, "per_k": L.OMEGA_PER_K}
_MODES = {"sausage": 0, "kink": 1, "fluting": 2, "fluting2": 2, "fluting3": 3}


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


_DEFAULT_MEDIA = {"cylinder_density": CYLINDER_CORONAL, "slab_density": SLAB_CORONAL,
                  "slab_flow": SLAB_FLOW_CORONAL, "cylinder_rotation": CYLINDER_PHOTOSPHERIC,
                  "cylinder_flow": CYLINDER_FLOW_CORONAL}
_DEFAULT_PROFILES = {"cylinder_density": GaussianDensity(0.95), "slab_density": GaussianDensity(0.9),
                     "slab_flow": GaussianFlow(1e5), "cylinder_rotation": PowerLawRotation(),
                     "cylinder_flow": GaussianAxialFlow(1.0)}


#: (n_steps, mesh_axis) measured for "rk8" on the kinds whose library default is "rk8n"
_RK8_TUNED = {"cylinder_density": (144, 0.16), "cylinder_flow": (144, 0.16)}


class ModelSpec:
    """Host description of one equilibrium + discretisation: the `esb_model` struct, the mesh nodes and
    the profile sampled there - everything `esb_set_model_fields` takes.  Needs no GPU (the mesh
    functions of the library are host code); `DispersionSolver` uploads it, the CPU tests hand the same
    arguments to the host build of the kernels' arithmetic (tests/host_harness)."""

    def __init__(self, kind, medium=None, profile=None, n_steps=None, scheme=None, mesh=None, rho_A=1.0,
                 ext_ic=None, ext_wavelengths=3.0, coordinate="negative", s_end=None, mesh_params=None):
        self.lib = L.load()
        self.kind = kind
        m = L.esb_model()
        L.check(self.lib, None, self.lib.esb_model_defaults(_KINDS[kind], C.byref(m)), "esb_model_defaults")
        self.medium = _DEFAULT_MEDIA[kind] if medium is None else medium
        self.profile = _DEFAULT_PROFILES[kind] if profile is None else profile
        self.rho_A = rho_A
        self._set_medium(m, self.medium)
        if scheme is not None:            # None: the kind's default ("rk8n" where the kind has a normal form)
            m.scheme = _SCHEMES[scheme]
        elif m.scheme == L.RK8N and len(self.profile(self.medium, np.array([m.s_start]))) < 3:
            m.scheme = L.RK8              # a user profile without the second derivative: the (y, y') form
        if m.scheme == L.RK8 and kind in _RK8_TUNED:
            # the discretisation tuned for the first-derivative form (the library's defaults are the
            # normal-form scheme's)
            m.n_steps, m.mesh_axis = _RK8_TUNED[kind]
        if mesh is not None:              # None: the kind's default (graded)
            m.mesh = {"uniform": L.MESH_UNIFORM, "clustered": L.MESH_CLUSTERED, "graded": L.MESH_GRADED}[mesh]
        if mesh_params is not None:       # (mesh_axis, mesh_edge, mesh_edge_width) of the graded mesh
            m.mesh_axis, m.mesh_edge, m.mesh_edge_width = (float(v) for v in mesh_params)
        m.ext_wavelengths = ext_wavelengths
        if n_steps is not None:
            m.n_steps = int(n_steps)
        elif m.scheme == L.RK4:
            m.n_steps = 2048
        if coordinate == "positive":
            if kind != "cylinder_density":
                raise ValueError("coordinate='positive' applies to the cylinder")
            m.r_sign, m.s_start, m.s_end = 1, 1.0, 0.001
            m.ext_ic_slope = 1e-8            # Density_cylinder_photospheric.py: P0 = [1e-8, 1e-8]
        if ext_ic is not None:
            m.ext_ic_value, m.ext_ic_slope = ext_ic
        if s_end is not None:
            m.s_end = float(s_end)
        self.model = m
        n = C.c_int32()
        L.check(self.lib, None, self.lib.esb_mesh_size(C.byref(m), C.byref(n)), "esb_mesh_size")
        self.nodes = np.empty(n.value, dtype=np.float64)
        L.check(self.lib, None, self.lib.esb_mesh_nodes(C.byref(m), _dptr(self.nodes)), "esb_mesh_nodes")
        nf = C.c_int32()
        L.check(self.lib, None, self.lib.esb_model_n_fields(C.byref(m), C.byref(nf)), "esb_model_n_fields")
        self.n_fields = nf.value

    def _set_medium(self, m, medium):
        if self.kind == "slab_flow":
            m.c_i0, m.vA_i0, m.vA_e, m.c_e = medium.c_i, medium.vA_i, medium.vA_e, medium.c_e
            m.gamma, m.rho_i0, m.rho_A, m.U_e = medium.gamma, medium.rho_i, 1.0, medium.U_e
        else:
            m.c_i0, m.vA_i0, m.vA_e, m.c_e = medium.c_i0, medium.vA_i0, medium.vA_e, medium.c_e
            m.gamma, m.rho_i0, m.rho_A = medium.gamma, medium.rho_i0, self.rho_A
            m.U_e = getattr(medium, "U_e", 0.0)

    def replace(self, medium=None, profile=None):
        """Another equilibrium on the same mesh (what a parameter scan changes between sweeps)."""
        if medium is not None:
            self.medium = medium
            self._set_medium(self.model, medium)
        if profile is not None:
            self.profile = profile

    @property
    def scheme(self):
        return _SCHEME_NAMES[self.model.scheme]

    def max_steps(self, scheme=None):
        """The most steps whose table fits the shared memory it is staged in (esb_model_max_steps), for this
        spec's kind and `scheme` (default: its own)."""
        m = type(self.model).from_buffer_copy(self.model)
        if scheme is not None:
            m.scheme = _SCHEMES[scheme]
        n = C.c_int32()
        L.check(self.lib, None, self.lib.esb_model_max_steps(C.byref(m), C.byref(n)), "esb_model_max_steps")
        return n.value

    def sampled(self):
        """-> (fields, boundary): the profile at the mesh nodes, the first field at s_start."""
        scale = self.rho_A if self.kind in ("cylinder_density", "slab_density") else 1.0
        vals = self.profile(self.medium, self.nodes)
        if len(vals) < self.n_fields:
            raise ValueError("scheme %r needs %d profile fields (value, first and second derivative), the "
                             "profile returned %d" % (self.scheme, self.n_fields, len(vals)))
        fields = [np.ascontiguousarray(np.asarray(f, dtype=np.float64) * scale) for f in vals[:self.n_fields]]
        boundary = np.array([float(self.profile(self.medium, np.array([self.model.s_start]))[0][0]) * scale])
        return fields, boundary

    def abi_args(self):
        """The argument list (after ctx) of esb_set_model_fields; keeps the arrays alive on self."""
        self._fields, self._boundary = self.sampled()
        self._fptr = (C.POINTER(C.c_double) * len(self._fields))(*[_dptr(f) for f in self._fields])
        return (C.byref(self.model), self._fptr, len(self._fields), self.nodes.size, _dptr(self._boundary),
                self._boundary.size)

    def solver_kwargs(self):
        """Keyword arguments that rebuild this spec (convergence_check builds a finer copy)."""
        m = self.model
        return dict(kind=self.kind, medium=self.medium, profile=self.profile, n_steps=int(m.n_steps),
                    scheme=self.scheme,
                    mesh={L.MESH_CLUSTERED: "clustered", L.MESH_UNIFORM: "uniform", L.MESH_GRADED: "graded"}[m.mesh],
                    mesh_params=(m.mesh_axis, m.mesh_edge, m.mesh_edge_width),
                    ext_ic=(m.ext_ic_value, m.ext_ic_slope), ext_wavelengths=m.ext_wavelengths, s_end=m.s_end,
                    rho_A=self.rho_A,
                    coordinate="positive" if (self.kind == "cylinder_density" and m.r_sign > 0) else "negative")


class DiscretisationWarning(UserWarning):
    """The fixed-step integration lost digits on this profile: raise n_steps (DispersionSolver.guard_report)."""


#: default sampling of the discretisation guard (ESB_GUARD_AUTO): every sweep re-evaluates about 32 k of its
#: (grid point, mode) pairs at twice the steps, on a side stream - 0.5 % of a 3e7-point sweep
GUARD_STRIDE = -1
GUARD_THRESHOLD = 1e-9


class DispersionSolver:
    """One GPU context evaluating D(omega,k) for one equilibrium model."""

    def __init__(self, kind, medium=None, profile=None, n_steps=None, scheme=None, mesh=None,
                 device=0, rho_A=1.0, ext_ic=None, ext_wavelengths=3.0, coordinate="negative", s_end=None,
                 mesh_params=None, guard=GUARD_STRIDE, guard_threshold=GUARD_THRESHOLD):
        """kind: "cylinder_density" | "slab_density" | "slab_flow" | "cylinder_rotation" |
        "cylinder_flow".
        profile: callable (medium, x) -> (rho, rho', rho'') for the density kinds, (U, U', U'') for
        "slab_flow", (v_phi, v_phi', c_i^2) for "cylinder_rotation", (v_z, v_z', v_z'') for
        "cylinder_flow"; any function may be given (this replaces the reference's sympy profile; a
        profile that returns no second derivative runs on the "rk8" scheme).
        scheme: "rk8n" (normal form, default where the kind has one) | "rk8" | "rk4".
        s_end: far end of the layer (the rotational sausage script stops at r = 0.01, the kink one at 0.001).
        coordinate="positive": the cylinder scripts written in r > 0 (photospheric set:
        layer 1 -> 0.001, exterior slope given as dP/dr).
        guard: stride of the built-in discretisation guard (one 32-point tile per 32 * guard grid points of
        a sweep is re-evaluated at 2 x n_steps; -1 = chosen per sweep, about 32 k samples; a host-returning
        root search warns - DiscretisationWarning - when the worst deviation exceeds guard_threshold;
        guard_report() reads it); 0 / None switches it off."""
        self.spec = ModelSpec(kind, medium, profile, n_steps, scheme, mesh, rho_A, ext_ic, ext_wavelengths,
                              coordinate, s_end, mesh_params)
        self.guard_stride = int(guard or 0)
        self.guard_threshold = float(guard_threshold)
        self.lib = self.spec.lib
        self.kind = kind
        self.ctx = L._ctx()
        rc = self.lib.esb_create(int(device), C.byref(self.ctx))
        if rc != L.ESB_OK:
            self.ctx = None
            raise L.EsbError("esb_create failed (status %d): no usable CUDA device %d; "
                             "eigensolver_b200 has no CPU fallback" % (rc, device))
        self._upload_model()

    # the spec's fields under their old names
    model = property(lambda self: self.spec.model)
    medium = property(lambda self: self.spec.medium)
    profile = property(lambda self: self.spec.profile)
    nodes = property(lambda self: self.spec.nodes)

    def _upload_model(self):
        L.check(self.lib, self.ctx, self.lib.esb_set_model_fields(self.ctx, *self.spec.abi_args()),
                "esb_set_model_fields")
        self._upload_guard()

    def _fine_spec(self, factor=2):
        """The same equilibrium at `factor` x the steps (None if its table cannot be staged)."""
        kw = self.spec.solver_kwargs()
        kw["n_steps"] = int(self.spec.model.n_steps) * int(factor)
        if kw["scheme"] == "rk8n" and kw["n_steps"] > self.spec.max_steps():
            kw["scheme"] = "rk8"          # the eight-field table of that many steps exceeds shared memory
        if kw["n_steps"] > self.spec.max_steps(kw["scheme"]):
            return None
        return ModelSpec(**kw)

    def _upload_guard(self):
        self._guard_on = False
        if self.guard_stride == 0:
            return
        fine = self._fine_spec()
        if fine is None:
            return
        L.check(self.lib, self.ctx,
                self.lib.esb_set_guard_fields(self.ctx, *fine.abi_args(), self.guard_stride, self.guard_threshold),
                "esb_set_guard_fields")
        self._guard_on = True

    def guard_report(self):
        """Discretisation guard of the LAST sweep (waits for it): dict with `worst` (largest deviation of
        g = D Y / (|ext Y| + |int Y|) - the acceptance test's relative mismatch, regular at the poles of D and
        blind to a common factor of int's numerator and denominator - between n_steps and 2 n_steps over the
        sampled points outside the
        resonant continua - the discretisation error of the sweep), where it occurred (`slot`, `k_index`,
        `w_index`), `n_checked`, `n_above` (samples above `threshold`), `stride` (0: guard off)."""
        rep = L.esb_guard_report()
        L.check(self.lib, self.ctx, self.lib.esb_guard_result(self.ctx, C.byref(rep)), "esb_guard_result")
        return {f: getattr(rep, f) for f, _ in L.esb_guard_report._fields_}

    def _warn_guard(self):
        """called by the host-returning root searches (they wait for the sweep anyway)"""
        if not getattr(self, "_guard_on", False):
            return
        rep = self.guard_report()
        if rep["n_checked"] and rep["worst"] > rep["threshold"]:
            warnings.warn("discretisation error %.1e > %.0e at n_steps = %d (%d of %d sampled points above; worst "
                          "at mode slot %d, k index %d, omega index %d): raise n_steps"
                          % (rep["worst"], rep["threshold"], int(self.spec.model.n_steps), rep["n_above"],
                             rep["n_checked"], rep["slot"], rep["k_index"], rep["w_index"]),
                          DiscretisationWarning, stacklevel=3)

    def reconfigure(self, medium=None, profile=None):
        """Swap the equilibrium (speeds and/or profile) on the same context and mesh: one small
        table upload, no reallocation.  This is what a parameter scan does between sweeps."""
        self.spec.replace(medium, profile)
        self._upload_model()

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.esb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------
    @staticmethod
    def _mode(mode):
        return _MODES[mode] if isinstance(mode, str) else int(mode)

    def _axes(self, k, w, layout):
        k = np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        lay = _LAYOUTS[layout]
        if lay == L.OMEGA_PER_K:
            assert w.ndim == 2 and w.shape[0] == k.size
            nw = w.shape[1]
        else:
            w = w.reshape(-1)
            nw = w.size
        return k, w, lay, nw

    def dispersion_grid(self, mode, k, w, layout="phase_speed"):
        """(ext, int) arrays of shape (nk, nw); D = ext - int; NaN where m_e < 0."""
        k, w, lay, nw = self._axes(k, w, layout)
        ext = np.empty((k.size, nw), dtype=np.float64)
        inq = np.empty((k.size, nw), dtype=np.float64)
        rc = self.lib.esb_dispersion_grid(self.ctx, self._mode(mode), _dptr(k), k.size, _dptr(w), nw, lay,
                                          _dptr(ext), _dptr(inq))
        L.check(self.lib, self.ctx, rc, "esb_dispersion_grid")
        return ext, inq

    def dispersion_grid_multi(self, modes, k, w, layout="phase_speed"):
        """Several modes in one fused scan: (ext, int) arrays of shape (n_modes, nk, nw)."""
        k, w, lay, nw = self._axes(k, w, layout)
        md = np.array([self._mode(m) for m in modes], dtype=np.int32)
        ext = np.empty((md.size, k.size, nw), dtype=np.float64)
        inq = np.empty((md.size, k.size, nw), dtype=np.float64)
        rc = self.lib.esb_dispersion_grid_multi(self.ctx, md.size, _iptr(md), _dptr(k), k.size, _dptr(w), nw,
                                                lay, _dptr(ext), _dptr(inq))
        L.check(self.lib, self.ctx, rc, "esb_dispersion_grid_multi")
        return ext, inq

    def dispersion_grid_leaky(self, modes, k, w, layout="phase_speed"):
        """OPT-IN, beyond the reference: (ext, int) of shape (n_modes, nk, nw) over the WHOLE grid - where
        m_e < 0 (the points the reference and every sweep skip) the exterior is the oscillatory solution of
        the same initial-value problem (J_n, Y_n; slab: cos, sin), interior and matching unchanged."""
        k, w, lay, nw = self._axes(k, w, layout)
        md = np.array([self._mode(m) for m in modes], dtype=np.int32)
        ext = np.empty((md.size, k.size, nw), dtype=np.float64)
        inq = np.empty((md.size, k.size, nw), dtype=np.float64)
        rc = self.lib.esb_dispersion_grid_leaky(self.ctx, md.size, _iptr(md), _dptr(k), k.size, _dptr(w), nw,
                                                lay, _dptr(ext), _dptr(inq))
        L.check(self.lib, self.ctx, rc, "esb_dispersion_grid_leaky")
        return ext, inq

    def D(self, mode, k, w, layout="phase_speed"):
        e, i = self.dispersion_grid(mode, k, w, layout)
        return e - i

    def find_roots(self, mode, k, w, layout="phase_speed", tol_percent=1.0, max_roots=None):
        """Scan + brackets + refinement; host arrays in, RootTable (host) out."""
        if max_roots is not None:
            # single C-ABI call with a caller-sized table (ESB_ERR_CAPACITY if too small)
            k, w, lay, nw = self._axes(k, w, layout)
            cap = int(max_roots)
            ki = np.empty(cap, np.int32); wi = np.empty(cap, np.int32)
            om = np.empty(cap, np.float64); ex = np.empty(cap, np.float64); iq = np.empty(cap, np.float64)
            ac = np.empty(cap, np.int32); it = np.empty(cap, np.int32)
            out = L.esb_roots(_iptr(ki), _iptr(wi), _dptr(om), _dptr(ex), _dptr(iq), _iptr(ac), _iptr(it))
            n, nb = C.c_int32(0), C.c_int32(0)
            rc = self.lib.esb_find_roots(self.ctx, self._mode(mode), _dptr(k), k.size, _dptr(w), nw, lay,
                                         float(tol_percent), cap, C.byref(out), C.byref(n), C.byref(nb))
            L.check(self.lib, self.ctx, rc, "esb_find_roots")
            n = n.value
            return RootTable(ki[:n].copy(), wi[:n].copy(), k, om[:n].copy(), ex[:n].copy(),
                             iq[:n].copy(), ac[:n].copy(), it[:n].copy(), nb.value)
        self.upload_axes(k, w, layout)
        n, nb = self.sweep_resident(mode, tol_percent)
        tab = self.download_roots(n)
        tab.n_brackets = nb
        self._warn_guard()
        return tab

    # -- the same pipeline in three steps (axes stay resident in HBM) -----
    def upload_axes(self, k, w, layout="phase_speed"):
        k, w, lay, nw = self._axes(k, w, layout)
        self._k_host = k
        L.check(self.lib, self.ctx, self.lib.esb_upload_axes(self.ctx, _dptr(k), k.size, _dptr(w), nw, lay),
                "esb_upload_axes")

    def sweep_resident(self, mode, tol_percent=1.0):
        """grid -> brackets -> refinement on the uploaded axes; returns (n_roots, n_brackets);
        the root table stays on the device."""
        n, nb = C.c_int32(0), C.c_int32(0)
        L.check(self.lib, self.ctx,
                self.lib.esb_sweep_resident(self.ctx, self._mode(mode), float(tol_percent), C.byref(n),
                                            C.byref(nb)), "esb_sweep_resident")
        return n.value, nb.value

    def sweep_resident_multi(self, modes, tol_percent=1.0):
        """One fused scan for all `modes` (up to 4), then the brackets and the refinement of all of them
        in single launches.  Returns the per-mode table sizes; table m stays on the device in slot m
        (the refinement may still be running: every accessor orders itself after it)."""
        md = np.array([self._mode(m) for m in modes], dtype=np.int32)
        n = np.zeros(md.size, np.int32)
        nb = np.zeros(md.size, np.int32)
        L.check(self.lib, self.ctx,
                self.lib.esb_sweep_resident_multi(self.ctx, md.size, _iptr(md), float(tol_percent), _iptr(n),
                                                  _iptr(nb)), "esb_sweep_resident_multi")
        self.last_n_brackets = [int(x) for x in nb]     # per mode slot: brackets found (>= table entries stored)
        return [int(x) for x in n]

    def set_accept_rule(self, rule):
        """"converged" (default): adjacent-point brackets refined to machine precision, one entry per
        bracket; "reference": the scripts' own scan / bisection rule point for point (header:
        esb_accept_rule) - what their pickles hold."""
        L.check(self.lib, self.ctx,
                self.lib.esb_set_accept_rule(self.ctx, {"converged": L.ACCEPT_CONVERGED,
                                                        "reference": L.ACCEPT_REFERENCE,
                                                        "reference_slab": L.ACCEPT_REFERENCE_SLAB}[rule]),
                "esb_set_accept_rule")

    def find_roots_multi(self, modes, k, w, layout="phase_speed", tol_percent=1.0, pinned=False):
        """Host arrays in, one RootTable per mode out (one fused scan).  pinned=True: the tables are
        views of page-locked buffers owned by the context (see download_roots_pinned)."""
        self.upload_axes(k, w, layout)
        ns = self.sweep_resident_multi(modes, tol_percent)
        if pinned:
            tabs = [self.download_roots_pinned(slot) for slot in range(len(ns))]
        else:
            tabs = [self.download_roots(n, slot) for slot, n in enumerate(ns)]
        for tab, nb in zip(tabs, self.last_n_brackets):
            tab.n_brackets = nb
        self._warn_guard()
        return tabs

    def download_roots_pinned(self, slot=0):
        """Root table of mode slot `slot` copied into page-locked host buffers owned by the context
        (esb_roots_pinned): numpy VIEWS, valid until the next call for the same slot or close()."""
        out = L.esb_roots()
        n = C.c_int32(0)
        L.check(self.lib, self.ctx, self.lib.esb_roots_pinned(self.ctx, int(slot), C.byref(out), C.byref(n)),
                "esb_roots_pinned")
        n = n.value
        if n == 0:
            z4, z8 = np.zeros(0, np.int32), np.zeros(0, np.float64)
            return RootTable(z4, z4, z8, z8, z8, z8, z4, z4, 0)
        view = lambda p: np.ctypeslib.as_array(p, shape=(n,))
        ki = view(out.k_index)
        return RootTable(ki, view(out.w_index), self._k_host, view(out.omega), view(out.ext),
                         view(out.intq), view(out.accepted), view(out.iterations), n)

    def download_roots(self, n, slot=0):
        ki = np.empty(n, np.int32); wi = np.empty(n, np.int32)
        om = np.empty(n, np.float64); ex = np.empty(n, np.float64); iq = np.empty(n, np.float64)
        ac = np.empty(n, np.int32); it = np.empty(n, np.int32)
        out = L.esb_roots(_iptr(ki), _iptr(wi), _dptr(om), _dptr(ex), _dptr(iq), _iptr(ac), _iptr(it))
        L.check(self.lib, self.ctx, self.lib.esb_download_roots_slot(self.ctx, int(slot), C.byref(out), n),
                "esb_download_roots_slot")
        return RootTable(ki, wi, self._k_host, om, ex, iq, ac, it, n)

    def scan_models(self, points, modes, tol_percent=1.0, capacity_per_table=0, download=True):
        """A parameter scan as ONE batched job (esb_scan_models): every equilibrium in `points` (dicts
        with optional 'medium' and 'profile', on this solver's mesh) swept over the uploaded axes with no
        host synchronisation in between.  Returns (table, n_brackets): `table` = dict of numpy VIEWS of
        page-locked buffers owned by the context (model, slot, k_index, w_index, omega, ext, intq,
        accepted, iterations; valid until the next scan), n_brackets[model][mode].
        download=False: the table stays on the device (table = None; see scan_table_device)."""
        import copy
        md = np.array([self._mode(m) for m in modes], dtype=np.int32)
        specs = []
        for p in points:
            sp = copy.copy(self.spec)
            sp.model = type(self.spec.model).from_buffer_copy(self.spec.model)
            sp.replace(p.get("medium"), p.get("profile"))
            specs.append(sp)
        n = len(specs)
        models = (L.esb_model * n)(*[sp.model for sp in specs])
        sampled = [sp.sampled() for sp in specs]
        nf = self.spec.n_fields
        fptr = (C.POINTER(C.c_double) * (n * nf))(*[_dptr(f) for fields, _ in sampled for f in fields])
        boundary = np.ascontiguousarray([b[0] for _, b in sampled], dtype=np.float64)
        nb = np.zeros(n * md.size, np.int32)
        out = L.esb_scan_result()
        rc = self.lib.esb_scan_models(self.ctx, n, models, fptr, nf, self.spec.nodes.size, _dptr(boundary),
                                      md.size, _iptr(md), float(tol_percent), int(capacity_per_table),
                                      1 if download else 0, _iptr(nb), C.byref(out))
        L.check(self.lib, self.ctx, rc, "esb_scan_models")
        if not download:
            return None, nb.reshape(n, md.size)
        ne = out.n_entries
        names = ("model", "slot", "k_index", "w_index", "accepted", "iterations", "omega", "ext", "intq")
        if ne == 0:
            tab = {nm: np.zeros(0, np.float64 if nm in ("omega", "ext", "intq") else np.int32) for nm in names}
        else:
            tab = {nm: np.ctypeslib.as_array(getattr(out, nm), shape=(ne,)) for nm in names}
        return tab, nb.reshape(n, md.size)

    def scan_table_device(self, stream=None):
        """Device pointers of the compact table of the last scan_models: dict name -> (pointer, numpy
        dtype string), plus 'n'.  `stream` as in roots_device."""
        L.check(self.lib, self.ctx,
                self.lib.esb_tables_wait(self.ctx, C.c_void_p(int(stream)) if stream else None),
                "esb_tables_wait")
        out = L.esb_scan_result()
        L.check(self.lib, self.ctx, self.lib.esb_scan_device(self.ctx, C.byref(out)), "esb_scan_device")
        addr = lambda p: C.cast(p, C.c_void_p).value or 0
        tab = {nm: (addr(getattr(out, nm)), "<f8" if nm in ("omega", "ext", "intq") else "<i4")
               for nm in ("model", "slot", "k_index", "w_index", "accepted", "iterations", "omega", "ext", "intq")}
        tab["n"] = out.n_entries
        return tab

    def roots_device(self, slot=0, stream=None):
        """Device pointers of the root table of mode slot `slot` (valid until the next sweep):
        dict name -> (pointer, numpy dtype string), plus 'n'.  Used for device-side gathers.
        stream: the cudaStream_t (integer) that will read them - it is made to wait for the sweep on the
        device; None: this call blocks until the tables are complete (esb_tables_wait)."""
        L.check(self.lib, self.ctx,
                self.lib.esb_tables_wait(self.ctx, C.c_void_p(int(stream)) if stream else None),
                "esb_tables_wait")
        out = L.esb_roots()
        n = C.c_int32(0)
        L.check(self.lib, self.ctx, self.lib.esb_roots_device(self.ctx, int(slot), C.byref(out), C.byref(n)),
                "esb_roots_device")
        addr = lambda p: C.cast(p, C.c_void_p).value or 0
        return {"n": n.value, "k_index": (addr(out.k_index), "<i4"), "w_index": (addr(out.w_index), "<i4"),
                "omega": (addr(out.omega), "<f8"), "ext": (addr(out.ext), "<f8"), "intq": (addr(out.intq), "<f8"),
                "accepted": (addr(out.accepted), "<i4"), "iterations": (addr(out.iterations), "<i4")}

    def set_stream(self, stream_ptr):
        """Run this context on an external cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""
        L.check(self.lib, self.ctx, self.lib.esb_set_stream(self.ctx, C.c_void_p(int(stream_ptr))),
                "esb_set_stream")

    def convergence_check(self, modes, k, w, layout="phase_speed", factor=2):
        """The integrator is fixed-step (uniform work per thread), where the reference's odeint adapts.
        For a profile sharper than the shipped ones this reports whether the step count suffices:
        the scan is repeated on a second context with `factor` times the steps and the largest
        deviation of D relative to max(|ext|, |int|) over the points both evaluate is returned, with
        the (mode, k, omega) where it occurs.  8th order: doubling the steps divides the error by ~256,
        so the value IS (to 0.4 %) the discretisation error of this solver on that grid."""
        kk, ww, lay, nw = self._axes(k, w, layout)
        e0, i0 = self.dispersion_grid_multi(modes, k, w, layout)
        m = self.model
        kw = self.spec.solver_kwargs()
        kw["n_steps"] = int(m.n_steps) * int(factor)
        if kw["scheme"] == "rk8n" and kw["n_steps"] > self.spec.max_steps():
            kw["scheme"] = "rk8"          # the eight-field table of that many steps exceeds shared memory
        with DispersionSolver(guard=0, **kw) as fine:
            e1, i1 = fine.dispersion_grid_multi(modes, k, w, layout)
        ok = np.isfinite(e0) & np.isfinite(i0) & np.isfinite(e1) & np.isfinite(i1)
        dev = np.where(ok, np.abs((e0 - i0) - (e1 - i1)) / np.maximum(np.abs(e1), np.abs(i1)), 0.0)
        j = np.unravel_index(int(np.argmax(dev)), dev.shape)
        om = ww[j[1], j[2]] if lay == L.OMEGA_PER_K else (kk[j[1]] * ww[j[2]] if lay == L.OMEGA_PHASE_SPEED else ww[j[2]])
        return float(dev[j]), {"mode": modes[j[0]], "k": float(kk[j[1]]), "omega": float(om),
                               "n_steps": int(m.n_steps), "n_steps_fine": int(m.n_steps) * int(factor)}

    def _respec(self, n_steps):
        """This solver's equilibrium at another step count (a normal-form table beyond its capacity does not fit
        shared memory: such a count runs on the first-derivative scheme)."""
        kw = self.spec.solver_kwargs()
        kw["n_steps"] = int(n_steps)
        if kw["scheme"] == "rk8n" and kw["n_steps"] > self.spec.max_steps():
            kw["scheme"] = "rk8"
        self.spec = ModelSpec(**kw)
        self._upload_model()

    def resolve_steps(self, modes, k, w, layout="phase_speed", target=None, max_rounds=6, sample=(48, 256)):
        """Error control for a profile the shipped step counts were not tuned for - the fixed-step counterpart
        of the adaptivity the reference gets from odeint (Density_cylinder.py:783 `odeint(dP_dr, ...)`, default
        rtol = atol = 1.49e-8): a `sample` = (rows, columns) grid spanning the caller's (k, w) window in
        the caller's layout is swept with EVERY point re-evaluated at 2 x n_steps by the built-in guard; while the
        worst deviation exceeds `target` (default: this solver's guard_threshold, 1e-9) n_steps is raised by the
        factor the scheme's order predicts ((worst / target)^(1/8) for the eighth-order schemes, 1/4 for
        "rk4", 15 % margin, at least x 1.25) and the model is uploaded again.  The new step count stays in
        force for every later call.  Returns {"n_steps", "scheme", "worst", "resolved", "history": [(n_steps,
        worst), ...]}; resolved = False when the staged table's capacity (guard at twice the steps included)
        was reached first - a DiscretisationWarning says so."""
        if self.guard_stride == 0:
            raise ValueError("resolve_steps needs the discretisation guard (guard != 0)")
        target = self.guard_threshold if target is None else float(target)
        kk, ww, lay, nw = self._axes(k, w, layout)
        # the sample spans the caller's window in the caller's layout, whatever its size (a worker-sized call
        # of one k and ~90 frequencies is below the 8192 pairs the guard samples at all)
        ks = np.linspace(kk.min(), kk.max(), int(sample[0]))
        if lay == L.OMEGA_PER_K:
            o = np.argsort(kk)
            lo, hi = np.interp(ks, kk[o], ww[o].min(axis=1)), np.interp(ks, kk[o], ww[o].max(axis=1))
            ws = lo[:, None] + (hi - lo)[:, None] * np.linspace(0.0, 1.0, int(sample[1]))[None, :]
        else:
            ws = np.linspace(ww.min(), ww.max(), int(sample[1]))
        stride_keep, self.guard_stride = self.guard_stride, 1
        history = []
        try:
            self._upload_guard()
            # the guard needs the table of 2 x n_steps staged as well
            cap = self.spec.max_steps("rk4" if self.spec.scheme == "rk4" else "rk8") // 2
            cap -= cap % 2
            for rnd in range(max_rounds):
                n_now = int(self.spec.model.n_steps)
                if not self._guard_on:
                    break
                self.upload_axes(ks, ws, layout)
                self.sweep_resident_multi(list(modes))
                rep = self.guard_report()
                if rep["n_checked"] == 0:        # the whole window is leaky / inside a continuum: nothing to judge
                    break
                worst = float(rep["worst"])
                history.append((n_now, worst))
                if worst <= target or n_now >= cap or rnd == max_rounds - 1:
                    break
                order = 4.0 if self.spec.scheme == "rk4" else 8.0
                grow = min(4.0, max(1.25, 1.15 * (worst / target) ** (1.0 / order))) if np.isfinite(worst) else 4.0
                self._respec(min(int(-(-n_now * grow // 8) * 8), cap))
        finally:
            self.guard_stride = stride_keep
            self._upload_guard()
        resolved = bool(history) and history[-1][1] <= target
        if not resolved:
            warnings.warn("resolve_steps: discretisation error %s at n_steps = %d, target %.0e not reached within the "
                          "staged table's capacity" % ("%.1e" % history[-1][1] if history else "unknown",
                                                       int(self.spec.model.n_steps), target),
                          DiscretisationWarning, stacklevel=2)
        return {"n_steps": int(self.spec.model.n_steps), "scheme": self.spec.scheme,
                "worst": history[-1][1] if history else float("nan"), "resolved": resolved, "history": history}

    def set_schedule(self, mode):
        """Scan and refinement kernels: "auto" (by size) | "lane" (one lane per point / bracket) |
        "warp" (one warp per point / bracket).  Each is bit-reproducible; they agree to rounding."""
        L.check(self.lib, self.ctx,
                self.lib.esb_set_schedule(self.ctx, {"auto": 0, "lane": 1, "warp": 2}[mode]),
                "esb_set_schedule")

    def bessel_jy_device(self, n, x):
        """(count, 4) array {J_n, J_n', Y_n, Y_n'}(x) evaluated ON THE DEVICE (esb_bessel_jy_dev)."""
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        out = np.empty((x.size, 4))
        L.check(self.lib, self.ctx, self.lib.esb_bessel_jy_dev(self.ctx, int(n), _dptr(x), x.size, _dptr(out)),
                "esb_bessel_jy_dev")
        return out

    def exterior_leaky_device(self, n, k, w):
        """(count, 2) array (P, dP/dr) at |r| = 1 of the exterior solution where m_e < 0 - the points the
        reference and the sweeps skip - in closed form with J_n, Y_n, on the device; NaN where m_e >= 0."""
        k = np.ascontiguousarray(k, dtype=np.float64).ravel()
        w = np.ascontiguousarray(w, dtype=np.float64).ravel()
        assert k.size == w.size
        out = np.empty((k.size, 2))
        L.check(self.lib, self.ctx,
                self.lib.esb_exterior_leaky_dev(self.ctx, int(n), _dptr(k), _dptr(w), k.size, _dptr(out)),
                "esb_exterior_leaky_dev")
        return out

    def fp64_peak_tflops(self):
        v = C.c_double(0.0)
        L.check(self.lib, self.ctx, self.lib.esb_fp64_peak(self.ctx, C.byref(v)), "esb_fp64_peak")
        return v.value

    # ------------------------------------------------------------------
    def last_kernel_ms(self):
        return float(self.lib.esb_last_kernel_ms(self.ctx))

    def launch_count(self):
        return int(self.lib.esb_launch_count(self.ctx))


def bessel_jy(n, x):
    """Host helper: (J_n, J_n', Y_n, Y_n') from the library's evaluators (the leaky side, m_e < 0)."""
    lib = L.load()
    out = (C.c_double * 4)()
    rc = lib.esb_bessel_jy(int(n), float(x), out)
    if rc != L.ESB_OK:
        raise L.EsbError("esb_bessel_jy: bad argument")
    return tuple(out)


def bessel_ik_scaled(n, z):
    """Host helper: (e^-z I_n, d/dz, e^z K_n, d/dz) from the library's evaluators."""
    lib = L.load()
    out = (C.c_double * 4)()
    rc = lib.esb_bessel_ik_scaled(int(n), float(z), out)
    if rc != L.ESB_OK:
        raise L.EsbError("esb_bessel_ik_scaled: bad argument")
    return tuple(out)
