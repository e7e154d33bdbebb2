"""ORACLE (test infrastructure, not product code): CPU restatement of the
reference's numpy/scipy shooting path for the dispersion function D(omega, k).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product path
(`eigensolver_b200/`) never does.

Parity status: PINNED.  `tests/test_oracle_pinned.py` checks this module against
  (a) D values produced by executing the reference's own `sausage()` / `kink()`
      functions in the build container (`tests/golden/make_golden.py`, fixtures
      `tests/golden/ref_D_*.npz`), and
  (b) the reference's shipped root tables (`Example data/*.pickle`, converted to
      `tests/golden/ref_roots_*.npz`).

What the reference does for every (k, omega)  (all four solver families share it):
  1. m_e(omega,k) < 0  -> the point is skipped (leaky regime, not handled).
  2. exterior: integrate the uniform-medium ODE from x = -3*2*pi/k to x = -1 with
     `scipy.integrate.odeint`, initial values (1e-8, 1e-8|1e-15).
  3. interior: integrate the non-uniform-layer ODE from the boundary (-1) across the
     layer with initial values (exterior value, s); `fsolve` finds the slope s that
     satisfies the symmetry condition at the far end (axis / other boundary).
  4. D = (exterior matched quantity at -1) - (interior matched quantity at -1);
     a sign change of D along omega brackets a mode, |D|*100/max(|ext|,|int|) < tol
     accepts it.

Reference sources restated here (file:line refer to /root/reference):
  cylinder, non-uniform density:
      Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py
        physics set-up 69-221, kink 546-824 (scan loop 694-821), sausage 847-1122
      Cylinder/Non-uniform density/Photospheric/Solvers/Density_cylinder_photospheric.py
  cylinder, non-uniform axial flow:
      Cylinder/Non-uniform flow/Coronal/solvers/Cylinder_method_flow_testing.py
        physics set-up 66-218, kink 554-850 (scan loop 701-846), sausage 855-1131
  slab, non-uniform density:
      Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py
        physics set-up 69-185, sausage 461-600 (scan loop), kink 640-790
      Slab/Non uniform density/Photospheric/Solvers/multiprocessor_Inhomogeneous_method.py

The reference rebuilds every coefficient with sympy (`sym.diff` + `lambdify`) at each
(k, omega).  Here the same expressions are written in closed form with numpy; the
pinning test is what shows they are the same functions.

Solver tolerances: `rtol`/`atol`/`xtol` default to None = scipy's defaults
(1.49e-8), i.e. exactly what the reference runs with.  The GPU parity tests call the
oracle with tight tolerances (rtol 1e-12, atol 1e-30) so that the comparison is
against the converged value of the reference's own formulation.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np
from scipy.integrate import odeint
from scipy.optimize import brentq, fsolve

GAMMA = 5.0 / 3.0


# --------------------------------------------------------------------------
# equilibrium models
# --------------------------------------------------------------------------
@dataclasses.dataclass
class Medium:
    """Characteristic speeds of the reference equilibrium (Density_cylinder.py:69-80)."""
    c_i0: float = 1.0
    vA_i0: float = 2.0
    vA_e: float = 5.0
    c_e: float = 0.5
    gamma: float = GAMMA
    rho_i0: float = 1.0

    @property
    def rho_e(self):
        g = self.gamma
        return self.rho_i0 * (self.c_i0**2 + g * 0.5 * self.vA_i0**2) / (
            self.c_e**2 + g * 0.5 * self.vA_e**2)

    @property
    def cT_e(self):
        return math.sqrt(self.c_e**2 * self.vA_e**2 / (self.c_e**2 + self.vA_e**2))

    @property
    def cT_i0(self):
        return math.sqrt(self.c_i0**2 * self.vA_i0**2 / (self.c_i0**2 + self.vA_i0**2))

    @property
    def B_0(self):
        return self.vA_i0 * math.sqrt(self.rho_i0)

    def m_e(self, k, w, U_e=0.0):
        """External m_e^2 (Density_cylinder.py:562 / slab ...coronal.py:172)."""
        W2 = (w - k * U_e) ** 2
        K = k * k
        return ((K * self.vA_e**2 - W2) * (K * self.c_e**2 - W2)) / (
            (self.vA_e**2 + self.c_e**2) * (K * self.cT_e**2 - W2))


# the four parameter sets the reference ships
CYL_CORONAL = Medium(c_i0=1.0, vA_i0=2.0, vA_e=5.0, c_e=0.5)        # Density_cylinder.py:69-72
CYL_PHOTOSPHERIC = Medium(c_i0=1.0, vA_i0=2.0, vA_e=0.5, c_e=1.5)    # Density_cylinder_photospheric.py
SLAB_CORONAL = Medium(c_i0=1.0, vA_i0=1.2, vA_e=3.0, c_e=0.4)        # ...method_coronal.py:69-72
SLAB_PHOTOSPHERIC = Medium(c_i0=1.0, vA_i0=1.9, vA_e=0.8, c_e=1.3)   # ...Inhomogeneous_method.py


@dataclasses.dataclass
class GaussianDensity:
    """rho(x) = rho_A*(rho_e + (rho_i0-rho_e) exp(-(x-x0)^2/width^2))  (Density_cylinder.py:135-154)."""
    medium: Medium
    width: float = 0.95
    x0: float = 0.0
    rho_A: float = 1.0
    #: True: vA from a constant field B_0 (cylinder script); False: slab script's form
    const_B: bool = False

    def rho(self, x):
        m = self.medium
        return self.rho_A * (m.rho_e + (m.rho_i0 - m.rho_e) * np.exp(-(x - self.x0) ** 2 / self.width**2))

    def drho(self, x):
        m = self.medium
        return self.rho_A * (m.rho_i0 - m.rho_e) * np.exp(-(x - self.x0) ** 2 / self.width**2) * (
            -2.0 * (x - self.x0) / self.width**2)

    # Both density solvers keep total pressure balance with a straight field:
    #   c_i^2 = rho_e (c_e^2 + gamma/2 vA_e^2)/rho - gamma/2 vA_i^2   (Density_cylinder.py:210)
    # cylinder: B_i = B_0 const -> vA_i^2 = B_0^2/rho                  (Density_cylinder.py:188-200)
    # slab:     vA_i = vA_i0 sqrt(rho_i0)/sqrt(profile)                (...coronal.py:117)
    # With rho_A = 1 (every shipped script) the two coincide; `const_B` selects the
    # cylinder form when rho_A != 1.
    def vA2(self, x):
        m = self.medium
        prof = self.rho(x) / (1.0 if self.const_B else self.rho_A)
        return m.vA_i0**2 * m.rho_i0 / prof

    def c2(self, x):
        m = self.medium
        return m.rho_e * (m.c_e**2 + 0.5 * m.gamma * m.vA_e**2) / self.rho(x) - 0.5 * m.gamma * self.vA2(x)

    def speeds(self, x):
        """rho, c^2, vA^2 and their x-derivatives."""
        m = self.medium
        rho = self.rho(x)
        drho = self.drho(x)
        ra = 1.0 if self.const_B else self.rho_A
        prof = rho / ra
        dprof = drho / ra
        vA2 = m.vA_i0**2 * m.rho_i0 / prof
        dvA2 = -vA2 * dprof / prof
        Cc = m.rho_e * (m.c_e**2 + 0.5 * m.gamma * m.vA_e**2)
        c2 = Cc / rho - 0.5 * m.gamma * vA2
        dc2 = -Cc * drho / rho**2 - 0.5 * m.gamma * dvA2
        return rho, drho, c2, dc2, vA2, dvA2


@dataclasses.dataclass
class EpsteinDensity(GaussianDensity):
    """The reference's alternative profile (Density_cylinder.py:139-142, commented out in the shipped
    file): rho = (rho_i0 - rho_e)/(cosh(x/a)^4)^2 + rho_e, `width` = the inhomogeneity width a."""

    def rho(self, x):
        m = self.medium
        return self.rho_A * ((m.rho_i0 - m.rho_e) / np.cosh((x - self.x0) / self.width) ** 8 + m.rho_e)

    def drho(self, x):
        m = self.medium
        t = (x - self.x0) / self.width
        return self.rho_A * (m.rho_i0 - m.rho_e) * (-8.0 / self.width) * np.sinh(t) / np.cosh(t) ** 9


# --------------------------------------------------------------------------
# the four geometry/mode closures
# --------------------------------------------------------------------------
class _Base:
    #: initial values of the exterior integration (value, slope)
    ext_ic = (1e-8, 1e-15)
    #: interior integration interval (start at the boundary)
    s0 = -1.0
    s1 = 1.0
    #: reference's fsolve starting guess
    slope_guess = 1.0
    n_ext_out = 500
    n_int_out = 500

    boundary = -1.0
    #: exterior domain = ext_wavelengths*2*pi/k (3 everywhere except the photospheric slab script: 7)
    ext_wavelengths = 3.0

    def ext_start(self, k):
        # "Number of wavelengths/2*pi accomodated in the domain" (Density_cylinder.py:553)
        return self.boundary * self.ext_wavelengths * 2.0 * np.pi / k


class SlabDensity(_Base):
    """Slab with non-uniform density, reference
    Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py."""
    geometry = "slab"
    ext_ic = (1e-8, 1e-8)          # :247  V0 = [1e-8, 1e-8]
    s0, s1 = -1.0, 1.0             # :91   ix = linspace(-1, 1, ...)
    slope_guess = 1.0              # :262  fsolve(objective_dvxi, 1.)
    #: :91 ix = linspace(-1, 1, 1e6).  LSODA's step sequence (hence the 1e-8-level solver
    #: noise) depends on the output grid, so the faithful default keeps the 1e6 points; the
    #: converged (tight-tolerance) value does not, and tests pass n_int_out=500 for speed.
    n_int_out = 10**6

    def __init__(self, profile: GaussianDensity, mode: str, n_int_out=None):
        assert mode in ("sausage", "kink")
        if n_int_out is not None:
            self.n_int_out = int(n_int_out)
        self.profile = profile
        self.medium = profile.medium
        self.mode = mode

    # -- exterior ----------------------------------------------------------
    def ext_rhs(self, k, w):
        m_e = self.medium.m_e(k, w)
        return lambda y, x: [y[1], m_e * y[0]]      # dVx_dx_e  :245

    def ext_match(self, k, w, y_b):
        """-> (value handed to the interior as y(-1), exterior matched quantity)."""
        m = self.medium
        K, A = k * k, w * w
        p_e_const = m.rho_e * (m.vA_e**2 + m.c_e**2) * (K * m.cT_e**2 - A) / (w * (K * m.c_e**2 - A))  # :221
        return y_b[0], p_e_const * y_b[1]           # left_P_solution = p_e_const*Ls[:,1]  :250

    # -- interior ----------------------------------------------------------
    def coeffs(self, x, k, w):
        """y'' = a y' + b y ;  a = -F'/F, b = m0^2   (dVx_dx_i :254)."""
        K, A = k * k, w * w
        rho, drho, c2, dc2, vA2, dvA2 = self.profile.speeds(x)
        s = c2 + vA2
        ds = dc2 + dvA2
        cT2 = c2 * vA2 / s
        dcT2 = (dc2 * vA2 + c2 * dvA2) / s - cT2 * ds / s
        # F = rho (c^2+vA^2)(k^2 cT^2 - w^2)/(k^2 c^2 - w^2)   :222
        dlnF = drho / rho + ds / s + K * dcT2 / (K * cT2 - A) - K * dc2 / (K * c2 - A)
        m0 = (K * c2 - A) * (K * vA2 - A) / (s * (K * cT2 - A))   # :230
        return -dlnF, m0

    def end_residual(self, y_end, y_start0):
        # sausage: vx(1) + vx(-1) = 0 (:259), kink: vx(1) - vx(-1) = 0 (:696)
        return y_end[0] + y_start0 if self.mode == "sausage" else y_end[0] - y_start0

    def end_functional(self):
        """residual = c0*y(end) + c1*y'(end) + cb*y(start)."""
        return (1.0, 0.0, 1.0 if self.mode == "sausage" else -1.0)

    def int_match(self, k, w, y0, slope):
        """interior matched quantity at the boundary: p_i_const[0]*vx'(-1)  (:267)."""
        K, A = k * k, w * w
        rho, _, c2, _, vA2, _ = self.profile.speeds(self.s0)
        cT2 = c2 * vA2 / (c2 + vA2)
        P_Ti = rho * (vA2 + c2) * (K * cT2 - A) / (w * (K * c2 - A))   # :234
        return P_Ti * slope


class CylinderDensity(_Base):
    """Cylinder with non-uniform density, reference
    Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py.

    mode: azimuthal wavenumber m (0 sausage, 1 kink, >=2 fluting)."""
    geometry = "cylinder"
    ext_ic = (1e-8, 1e-15)         # :768  P0 = [1e-8, 1e-15]
    s0, s1 = -1.0, -0.001          # :120  ix = linspace(-1., -0.001, 500)
    slope_guess = -0.001           # :790  fsolve(objective_dPi, -0.001)

    def __init__(self, profile: GaussianDensity, m: int, coordinate: str = "negative"):
        self.profile = profile
        self.medium = profile.medium
        self.m = int(m)
        self.boundary = -1.0
        if coordinate == "positive":
            # Density_cylinder_photospheric.py: ix = linspace(1., 0.001, 1e3) (:120),
            # lx = linspace(3.*2.*pi/k, 1., 500) (:696), P0 = [1e-8, 1e-8] (:771), fsolve(.., 0.001)
            self.s0, self.s1 = 1.0, 0.001
            self.ext_ic = (1e-8, 1e-8)
            self.slope_guess = 0.001
            self.n_int_out = 1000
            self.boundary = 1.0
        self.mode = {0: "sausage", 1: "kink"}.get(self.m, "fluting%d" % self.m)

    def ext_rhs(self, k, w):
        m_e = self.medium.m_e(k, w)
        mm = float(self.m * self.m)
        # dP_dr_e :765 (kink, 1/r^2) and :1061 (sausage, 0/r^2)
        return lambda y, r: [y[1], -y[1] / r + (m_e + mm / (r * r)) * y[0]]

    def ext_match(self, k, w, y_b):
        md = self.medium
        xi_e_const = -1.0 / (md.rho_e * (k * k * md.vA_e**2 - w * w))   # :702
        return y_b[0], xi_e_const * y_b[1]         # left_xi_solution  :773

    def coeffs(self, r, k, w):
        """P'' = a P' + b P ; a = -dF/F, b = g/F   (dP_dr_i :782).

        With v_phi = B_phi = v_z = 0 (Density_cylinder.py:97-108) the reference's
        general coefficients collapse:  Q = T = C1 = 0,  C3 = D rho (w^2 - wA^2),
        F = r D/C3 = r/(rho w^2 - k^2 B^2),  g = -r C2/D,  and
        g/F = m^2/r^2 + k^2 - w^4/((c^2+vA^2)(w^2 - k^2 cT^2))."""
        K, A = k * k, w * w
        rho, drho, c2, dc2, vA2, dvA2 = self.profile.speeds(r)
        s = c2 + vA2
        cT2 = c2 * vA2 / s
        X = rho * (A - K * vA2)                      # rho (w^2 - wA^2)
        dX = drho * (A - K * vA2) - rho * K * dvA2
        dlnF = 1.0 / r - dX / X
        b = self.m**2 / (r * r) + K - A * A / (s * (A - K * cT2))
        return -dlnF, b

    def end_functional(self):
        # kink/fluting: P(axis) = 0 (:787); sausage: P'(axis) = 0 (:1084)
        return (0.0, 1.0, 0.0) if self.m == 0 else (1.0, 0.0, 0.0)

    def end_residual(self, y_end, y_start0):
        return y_end[1] if self.m == 0 else y_end[0]

    def int_match(self, k, w, y0, slope):
        """inside_xi_solution[0] = (C1 P + D P')/C3 at r=-1 (:798) = P'(-1)/(rho (w^2-wA^2))."""
        rho, _, _, _, vA2, _ = self.profile.speeds(self.s0)
        return slope / (rho * (w * w - k * k * vA2))


class CylinderDensityPerPoint(CylinderDensity):
    """The same model with the reference's own per-point cost structure: Density_cylinder.py:705-757
    re-derives D, Q, T, C1, C2, C3, F = r D/C3, dF = diff(F, r) and g = -diff(r C1/C3, r) - r (C2 - C1^2/C3)/D
    with sympy and lambdifies six functions at EVERY (k, omega) of the scan loop, before the two odeint /
    fsolve stages.  `prepare(k, w)` restates exactly that (with v_phi = B_phi = v_z = 0, B_i = B_0 as the
    script sets them, :97-108); the hoisted class above evaluates the collapsed closed forms instead.
    Used by bench.py's CPU legs to report the reference's true per-evaluation rate next to the hoisted
    (faster, conservative) one, and pinned to the hoisted path in tests/test_oracle_pinned.py."""

    def prepare(self, k, w):
        import sympy as sym
        md, pf = self.medium, self.profile
        rr = sym.symbols("r")
        m, v_phi, B_phi, v_z = self.m, 0.0, 0.0, 0.0                                       # :97-108
        B_0 = md.B_0
        rho_e, c_e, vA_e, gamma = md.rho_e, md.c_e, md.vA_e, md.gamma

        def rho_i(r):                                                                      # :135-146
            if isinstance(pf, EpsteinDensity):
                return pf.rho_A * ((md.rho_i0 - rho_e) / (sym.cosh((r - pf.x0) / pf.width) ** 4) ** 2 + rho_e)
            return pf.rho_A * (rho_e + (md.rho_i0 - rho_e) * sym.exp(-(r - pf.x0) ** 2 / pf.width**2))

        B_i = lambda r: B_0                                                                # :196
        vA_i = lambda r: (B_i(r) + B_phi) / sym.sqrt(rho_i(r))                             # :188
        c_i = lambda r: sym.sqrt(rho_e * (c_e**2 + 0.5 * gamma * vA_e**2) / rho_i(r) - 0.5 * gamma * vA_i(r) ** 2)  # :210
        shift = lambda r: w - (m * v_phi / r) + k * v_z                                    # :705
        alfven = lambda r: (m * B_phi / r) + (k * B_i(r)) / sym.sqrt(rho_i(r))             # :708
        cusp = lambda r: alfven(r) * c_i(r) / sym.sqrt(c_i(r) ** 2 + vA_i(r) ** 2)         # :711
        D = lambda r: rho_i(r) * (c_i(r) ** 2 + vA_i(r) ** 2) * (shift(r) ** 2 - alfven(r) ** 2) * (
            shift(r) ** 2 - cusp(r) ** 2)                                                  # :714
        Q = lambda r: (-(shift(r) ** 2 - alfven(r) ** 2) * rho_i(r) * v_phi**2 / r) + (
            2 * shift(r) ** 2 * B_phi**2 / r) + (
            2 * shift(r) * B_phi * v_phi * ((m * B_phi / r) + (k * B_i(r))) / r)           # :719
        T = lambda r: (((m * B_phi / r) + (k * B_i(r))) * B_phi) + rho_i(r) * v_phi * shift(r)   # :722
        C1 = lambda r: (Q(r) * shift(r)) - (2 * m * (c_i(r) ** 2 + vA_i(r) ** 2) * (
            shift(r) ** 2 - cusp(r) ** 2) * T(r) / r**2)                                   # :725
        C2 = lambda r: shift(r) ** 4 - ((c_i(r) ** 2 + vA_i(r) ** 2) * (m**2 / r**2 + k**2) * (
            shift(r) ** 2 - cusp(r) ** 2))                                                 # :730
        C3_diff = lambda r: (B_phi / r) ** 2 - (rho_i(r) * (v_phi / r) ** 2)               # :733
        C3 = lambda r: (D(r) * (rho_i(r) * (shift(r) ** 2 - alfven(r) ** 2) + (r * sym.diff(C3_diff(r), r)))) + (
            Q(r) ** 2 - (4 * (c_i(r) ** 2 + vA_i(r) ** 2) * (shift(r) ** 2 - cusp(r) ** 2) * T(r) ** 2 / r**2))  # :736
        F = lambda r: (r * D(r)) / C3(r)                                                   # :741
        dF = lambda r: sym.diff(F(r), r)                                                   # :746
        g = lambda r: -(sym.diff((r * C1(r) / C3(r)), r)) - (r * (C2(r) - (C1(r) ** 2 / C3(r))) / D(r))   # :752
        self._D = sym.lambdify(rr, D(rr), "numpy")                                         # :717
        self._C1 = sym.lambdify(rr, C1(rr), "numpy")                                       # :728
        self._C3 = sym.lambdify(rr, C3(rr), "numpy")                                       # :739
        self._F = sym.lambdify(rr, F(rr), "numpy")                                         # :744
        self._dF = sym.lambdify(rr, dF(rr), "numpy")                                       # :749
        self._g = sym.lambdify(rr, g(rr), "numpy")                                         # :755

    def coeffs(self, r, k, w):
        F = self._F(r)
        return -self._dF(r) / F, self._g(r) / F                                            # dP_dr_i :782

    def int_match(self, k, w, y0, slope):
        r = self.s0
        return (self._C1(r) * y0 + self._D(r) * slope) / self._C3(r)                       # :798


@dataclasses.dataclass
class FlowMedium:
    """Speeds of the slab flow script (flow_multiprocessor_coronal.py:47-56): uniform density
    and field inside, Gaussian shear flow U(x) = U_e + (U_i0-U_e) exp(-(x-x0)^2/dx^2)."""
    vA_i: float = 1.0
    c_i: float = 0.3
    vA_e: float = 2.5
    c_e: float = 0.2
    U_i0: float = 0.9
    U_e: float = 0.0
    gamma: float = GAMMA
    rho_i: float = 1.0
    width: float = 1e5
    x0: float = 0.0

    @property
    def rho_e(self):
        g = self.gamma
        return self.rho_i * (self.c_i**2 + g * 0.5 * self.vA_i**2) / (self.c_e**2 + g * 0.5 * self.vA_e**2)

    @property
    def cT_i2(self):
        return self.c_i**2 * self.vA_i**2 / (self.c_i**2 + self.vA_i**2)

    @property
    def cT_e2(self):
        return self.c_e**2 * self.vA_e**2 / (self.c_e**2 + self.vA_e**2)

    def U(self, x):
        g = np.exp(-(x - self.x0) ** 2 / self.width**2)
        t = -2.0 * (x - self.x0) / self.width**2
        U = self.U_e + (self.U_i0 - self.U_e) * g
        dU = (self.U_i0 - self.U_e) * g * t
        ddU = (self.U_i0 - self.U_e) * g * (t * t - 2.0 / self.width**2)
        return U, dU, ddU

    def m_e(self, k, w):
        W2 = (w - k * self.U_e) ** 2
        K = k * k
        return ((K * self.vA_e**2 - W2) * (K * self.c_e**2 - W2)) / (
            (self.vA_e**2 + self.c_e**2) * (K * self.cT_e2 - W2))


class SlabFlow(_Base):
    """Slab with a non-uniform (sheared) flow, reference
    Slab/Non uniform flow/Solver/flow_multiprocessor_coronal.py  (sausage :139-330, kink :347-)."""
    geometry = "slab"
    ext_ic = (1e-8, 1e-15)         # :229  V0 = [1e-8, 1e-15]
    s0, s1 = -1.0, 1.0             # :72   ix = linspace(-1, 1, 500)
    slope_guess = 0.0              # :304  fsolve(objective_dvxi, 0.)
    n_int_out = 500

    def __init__(self, medium: FlowMedium, mode: str):
        assert mode in ("sausage", "kink")
        self.medium = medium
        self.mode = mode

    def ext_rhs(self, k, w):
        m_e = self.medium.m_e(k, w)
        return lambda y, x: [y[1], m_e * y[0]]

    def ext_match(self, k, w, y_b):
        md = self.medium
        K = k * k
        We = w - k * md.U_e
        p_e_const = md.rho_e * (md.vA_e**2 + md.c_e**2) * (K * md.cT_e2 - We**2) / (We * (K * md.c_e**2 - We**2))  # :209
        Ub = md.U(self.s0)[0]
        # left_solution = Ls[:,0]*(w - k U_i(-1))/(w - k U_e)  (:290): continuity of the displacement
        return y_b[0] * (w - k * Ub) / We, p_e_const * y_b[1]

    def coeffs(self, x, k, w):
        """vx'' = -D vx' - coeff vx  (dVx_dx_i :297): a = -D, b = -coeff."""
        md = self.medium
        K = k * k
        U, dU, ddU = md.U(x)
        Om = w - k * U
        c2, vA2, cT2 = md.c_i**2, md.vA_i**2, md.cT_i2
        s = c2 + vA2
        m0 = (K * c2 - Om**2) * (K * vA2 - Om**2) / (s * (K * cT2 - Om**2))                      # :211
        t = Om**2 - K * cT2
        Dx = 2.0 * k * dU * (t + K * K * cT2 * c2 / (s * t)) / (Om * (Om**2 - K * c2))           # :215
        coeff = k * ddU / Om + k * dU * Dx / Om - m0                                               # :219
        return -Dx, -coeff

    def end_residual(self, y_end, y_start0):
        return y_end[0] + y_start0 if self.mode == "sausage" else y_end[0] - y_start0

    def int_match(self, k, w, y0, slope):
        md = self.medium
        K = k * k
        Om = w - k * md.U(self.s0)[0]
        c2, vA2, cT2 = md.c_i**2, md.vA_i**2, md.cT_i2
        P_Ti = md.rho_i * (vA2 + c2) * (K * cT2 - Om**2) / (Om * (K * c2 - Om**2))                 # :223
        return P_Ti * slope


class CylinderRotation(_Base):
    """Cylinder with a rotational (twisted) flow v_phi = v_twist r^power, reference
    Cylinder/Rotational flow/Photospheric/Solvers/Twisted_photospheric_nonlinear_flow_kink_fast.py
    (kink: scan loop :259-336) and Twisted_photospheric_flow_sausage.py (sausage).

    Like the reference, the coefficients F, dF/dr and g of the interior ODE are built with sympy
    from D, Q, T, C1, C2, C3 (:264-297) and differentiated symbolically; unlike the reference they
    are lambdified ONCE with (r, omega, k) as arguments instead of once per (k, omega)."""
    geometry = "cylinder"
    ext_ic = (1e-8, 1e-8)          # :302  P0 = [1e-8, 1e-8]
    s0 = 1.0                       # :96   ix = linspace(1., 0.001, 2e3)   (sausage script: 0.01)
    slope_guess = 0.001            # :311  fsolve(objective_dPi, 0.001)
    boundary = 1.0
    n_int_out = 2000

    def __init__(self, medium: Medium, m: int, v_twist=0.25, power=0.8, s_end=None):
        import sympy as sym
        self.medium = medium
        self.m = int(m)
        self.v_twist, self.power = float(v_twist), float(power)
        self.s1 = s_end if s_end is not None else (0.01 if self.m == 0 else 0.001)
        self.mode = {0: "sausage", 1: "kink"}.get(self.m, "fluting%d" % self.m)
        md = medium
        r, w, k = sym.symbols("r w k", positive=True)
        mm = sym.Integer(self.m)
        rho = sym.Float(md.rho_i0)
        B0 = sym.Float(md.B_0)
        P0 = sym.Float(md.c_i0**2 * md.rho_i0 / md.gamma)
        vphi = sym.Float(self.v_twist) * r ** sym.Float(self.power)                     # :107
        P_i = rho * sym.Float(self.v_twist) ** 2 * (r ** (2 * sym.Float(self.power)) /
                                                    (2 * sym.Float(self.power))) + P0    # :109
        c2 = P_i * sym.Float(md.gamma) / rho                                              # :111
        vA2 = B0**2 / rho
        shift = w - mm * vphi / r                                                         # :264 (v_z = 0)
        alf = k * B0 / sym.sqrt(rho)                                                      # :266 (B_phi = 0)
        cusp2 = alf**2 * c2 / (c2 + vA2)
        D = rho * (c2 + vA2) * (shift**2 - alf**2) * (shift**2 - cusp2)                   # :270
        Q = -(shift**2 - alf**2) * rho * vphi**2 / r                                      # :273
        T = rho * vphi * shift                                                            # :275
        C1 = Q * shift**2 - 2 * mm * (c2 + vA2) * (shift**2 - cusp2) * T / r**2           # :277
        C2 = shift**4 - (c2 + vA2) * (mm**2 / r**2 + k**2) * (shift**2 - cusp2)           # :280
        C3_diff = -rho * (vphi / r) ** 2                                                  # :283
        C3 = D * (rho * (shift**2 - alf**2) + r * sym.diff(C3_diff, r)) + (
            Q**2 - 4 * (c2 + vA2) * (shift**2 - cusp2) * T**2 / r**2)                     # :285
        F = r * D / C3                                                                    # :288
        dF = sym.diff(F, r)                                                               # :291
        g = -sym.diff(r * C1 / C3, r) - r * (C2 - C1**2 / C3) / D                         # :294
        self._a = sym.lambdify((r, w, k), -dF / F, "numpy", cse=True)
        self._b = sym.lambdify((r, w, k), g / F, "numpy", cse=True)
        self._xi = sym.lambdify((r, w, k), (C1 / C3, D / C3), "numpy", cse=True)
        self._rho_v2_b = md.rho_i0 * self.v_twist**2          # rho(1) v_phi(1)^2

    def ext_rhs(self, k, w):
        m_e = self.medium.m_e(k, w)
        mm = float(self.m * self.m)
        return lambda y, r: [y[1], -y[1] / r + (m_e + mm / (r * r)) * y[0]]               # :300

    def ext_match(self, k, w, y_b):
        md = self.medium
        xi_e_const = -1.0 / (md.rho_e * (k * k * md.vA_e**2 - w * w))                     # :263
        self._xi_e = xi_e_const * y_b[1]
        return y_b[0], self._xi_e

    def coeffs(self, r, k, w):
        # LSODA may step past the last output point, i.e. to r <= 0; numpy semantics (nan) there,
        # not Python's complex power
        r = np.float64(r)
        return self._a(r, w, k), self._b(r, w, k)

    def end_residual(self, y_end, y_start0):
        if self.m == 0:
            return y_end[1]                                   # sausage: P'(axis) = 0
        # kink (:308): U[:,0][-1] + (B_phi(1)^2 - rho(1) v_phi(1)^2) xi_e(1);  the xi_e term belongs
        # to the inhomogeneous part (it is passed with the start value, see dispersion())
        return y_end[0] - (self._rho_v2_b * self._xi_e if y_start0 != 0.0 else 0.0)

    def int_match(self, k, w, y0, slope):
        c1, d = self._xi(self.s0, w, k)
        return c1 * y0 + d * slope                            # (C1 P + D P')/C3 at r = 1  (:314)


@dataclasses.dataclass
class AxialFlowMedium(Medium):
    """Cylinder_method_flow_testing.py:66-69 (coronal speeds), :130-131 (U_i0, U_e), :123-124 (r0, dr)."""
    U_i0: float = 0.35
    U_e: float = 0.0
    width: float = 1.0
    r0: float = 0.0


class CylinderFlow(_Base):
    """Cylinder with a non-uniform axial flow v_z(r) = U_e + (U_i0-U_e) exp(-(r-r0)^2/dr^2), reference
    Cylinder/Non-uniform flow/Coronal/solvers/Cylinder_method_flow_testing.py (kink :554-850, scan
    loop :701-846; sausage :855-1131).  Uniform rho_i, B_i = B_0, c_i inside (B_phi = v_phi = 0 as
    shipped, :190-196).

    As in the reference, F, dF/dr and g are built with sympy from D, Q, T, C1, C2, C3 (:711-762) and
    differentiated symbolically; they are lambdified once with (r, omega, k) as arguments."""
    geometry = "cylinder"
    ext_ic = (1e-8, 1e-8)          # :774  P0 = [1e-8, 1e-8]
    s0, s1 = -1.0, -0.001          # :120  ix = linspace(-1., -0.001, 1e3)
    slope_guess = -0.001           # :798  fsolve(objective_dPi, -0.001)
    n_int_out = 1000

    def __init__(self, medium: AxialFlowMedium, m: int):
        import sympy as sym
        self.medium = medium
        self.m = int(m)
        self.mode = {0: "sausage", 1: "kink"}.get(self.m, "fluting%d" % self.m)
        md = medium
        r = sym.symbols("r", negative=True)
        w, k = sym.symbols("w k", positive=True)
        mm = sym.Integer(self.m)
        rho = sym.Float(md.rho_i0)                                                        # :145
        B0 = sym.Float(md.B_0)                                                            # :184 (B_phi = 0)
        vA2 = B0**2 / rho                                                                 # :174
        c2 = sym.Float(md.rho_e * (md.c_e**2 + 0.5 * md.gamma * md.vA_e**2)) / rho - sym.Float(
            0.5 * md.gamma) * vA2                                                         # :208
        vz = sym.Float(md.U_e) + sym.Float(md.U_i0 - md.U_e) * sym.exp(
            -(r - sym.Float(md.r0)) ** 2 / sym.Float(md.width) ** 2)                      # :134
        shift = w - k * vz                                                                # :713 (v_phi = 0)
        alf = k * B0 / sym.sqrt(rho)                                                      # :716
        cusp2 = alf**2 * c2 / (c2 + vA2)                                                  # :719
        D = rho * (c2 + vA2) * (shift**2 - alf**2) * (shift**2 - cusp2)                   # :722
        C2 = shift**4 - (c2 + vA2) * (mm**2 / r**2 + k**2) * (shift**2 - cusp2)           # :738
        C3 = D * rho * (shift**2 - alf**2)                                                # :744 (Q = T = 0)
        F = r * D / C3                                                                    # :749
        dF = sym.diff(F, r)                                                               # :754
        g = -r * C2 / D                                                                   # :760 (C1 = 0)
        self._a = sym.lambdify((r, w, k), sym.simplify(-dF / F), "numpy", cse=True)
        self._b = sym.lambdify((r, w, k), g / F, "numpy", cse=True)
        self._xi = sym.lambdify((r, w, k), D / C3, "numpy", cse=True)

    def ext_rhs(self, k, w):
        m_e = self.medium.m_e(k, w)                                                       # :706 (omega unshifted)
        mm = float(self.m * self.m)
        return lambda y, r: [y[1], -y[1] / r + (m_e + mm / (r * r)) * y[0]]               # :771 / :1070

    def ext_match(self, k, w, y_b):
        md = self.medium
        xi_e_const = -1.0 / (md.rho_e * (k * k * md.vA_e**2 - w * w))                     # :709
        return y_b[0], xi_e_const * y_b[1]

    def coeffs(self, r, k, w):
        r = np.float64(r)
        return self._a(r, w, k), self._b(r, w, k)

    def end_residual(self, y_end, y_start0):
        # kink: P(axis) - B_phi(-1)^2 xi_e = P(axis) (:795); sausage: P'(axis) = 0 (:1092)
        return y_end[1] if self.m == 0 else y_end[0]

    def int_match(self, k, w, y0, slope):
        return self._xi(self.s0, w, k) * slope            # (C1 P + D P')/C3 at r = -1, C1 = 0  (:806)


# --------------------------------------------------------------------------
# the dispersion function
# --------------------------------------------------------------------------
def _ode_kw(rtol, atol, scale=1.0):
    """atol="scaled": absolute tolerance = rtol x the magnitude of the initial data, i.e. a
    purely relative control that does not stall where an oscillating solution crosses zero."""
    kw = {}
    if rtol is not None:
        kw["rtol"] = rtol
    if isinstance(atol, str):
        kw["atol"] = (rtol or 1.49012e-8) * abs(scale)
    elif atol is not None:
        kw["atol"] = atol
    if rtol is not None and rtol < 1e-9:
        kw["mxstep"] = 200000
    return kw


def exterior(model, k, w, rtol=None, atol=None):
    """Step 2: boundary values (y, y') of the exterior solution at x=-1, or None if skipped."""
    if model.medium.m_e(k, w) < 0:           # "if m_e < 0: pass"  Density_cylinder.py:760
        return None
    lx = np.linspace(model.ext_start(k), model.boundary, model.n_ext_out)
    Ls = odeint(model.ext_rhs(k, w), list(model.ext_ic), lx,
                **_ode_kw(rtol, atol, 1e-3 * max(abs(model.ext_ic[0]), abs(model.ext_ic[1]))))
    return Ls[-1]


def dispersion(model, k, w, rtol=None, atol=None, xtol=None, shoot="fsolve"):
    """Steps 1-4 for one (k, w).  Returns (exterior quantity, interior quantity);
    D = ext - int.  (nan, nan) where the reference skips the point."""
    if hasattr(model, "prepare"):
        model.prepare(k, w)                  # the reference rebuilds its coefficients before the m_e test (:705-757)
    yb = exterior(model, k, w, rtol, atol)
    if yb is None:
        return float("nan"), float("nan")
    y0, ext_q = model.ext_match(k, w, yb)
    ix = np.linspace(model.s0, model.s1, model.n_int_out)
    okw = _ode_kw(rtol, atol, y0 if y0 != 0 else 1.0)

    def rhs(y, s):
        a, b = model.coeffs(s, k, w)
        return [y[1], a * y[1] + b * y[0]]

    if shoot == "fsolve":
        def objective(sl):
            U = odeint(rhs, [y0, float(np.asarray(sl).reshape(-1)[0])], ix, **okw)
            return model.end_residual(U[-1], y0)
        fkw = {} if xtol is None else {"xtol": xtol}
        slope, = fsolve(objective, model.slope_guess, **fkw)
    else:
        # the interior ODE is linear, so the residual is affine in the slope:
        # two integrations give the same slope fsolve converges to.
        U0 = odeint(rhs, [y0, 0.0], ix, **okw)[-1]
        U1 = odeint(rhs, [0.0, 1.0], ix, **_ode_kw(rtol, atol, 1.0))[-1]
        r0 = model.end_residual(U0, y0)
        r1 = model.end_residual(U1, 0.0)
        slope = -r0 / r1
    return ext_q, model.int_match(k, w, y0, slope)


def D(model, k, w, **kw):
    e, i = dispersion(model, k, w, **kw)
    return e - i


def mismatch_percent(e, i):
    """The reference's acceptance test (Density_cylinder.py:809)."""
    return abs(e - i) * 100.0 / max(abs(e), abs(i))


def scan(model, k, freq, **kw):
    """D over an omega grid at one k -> (ext[], int[]) arrays (nan = skipped)."""
    out = np.array([dispersion(model, k, w, **kw) for w in freq])
    return out[:, 0], out[:, 1]


def brackets(Dvals):
    """Indices j with a sign change between grid points j and j+1 (both evaluated)."""
    d = np.asarray(Dvals)
    ok = np.isfinite(d[:-1]) & np.isfinite(d[1:])
    return np.nonzero(ok & (d[:-1] * d[1:] < 0))[0]


def refine(model, k, w_lo, w_hi, xtol=1e-14, **kw):
    """Brent refinement of one bracket; returns (omega, ext, int)."""
    f = lambda w: D(model, k, w, **kw)
    w = brentq(f, w_lo, w_hi, xtol=xtol * max(abs(w_lo), abs(w_hi)), rtol=8.9e-16, maxiter=200)
    e, i = dispersion(model, k, w, **kw)
    return w, e, i


def find_roots(model, k, freq, tol_percent=1.0, **kw):
    """Reference semantics, restated with a proper root polish:
    brackets along omega -> refine -> keep those that pass the reference's
    acceptance test (poles of D change sign too but never pass it)."""
    e, i = scan(model, k, freq, **kw)
    roots = []
    for j in brackets(e - i):
        w, er, ir = refine(model, k, freq[j], freq[j + 1], **kw)
        if mismatch_percent(er, ir) < tol_percent:
            roots.append(w)
    return np.array(roots)
