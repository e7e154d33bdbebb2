// Device core of the dispersion-function hot path: exterior closed forms,
// per-node ODE coefficients from the staged profile table, fixed-step RK4 / RK8
// shooting integrator and the matching closures.  One thread evaluates one
// (k, omega) point (eval_point_multi), or one warp does (WARP = true, warp_transfer).
//
// Reference path restated (file:line in /root/reference):
//   cylinder  Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py:694-821
//   slab      Slab/Non uniform density/Coronal/Solvers/
//                 multiprocessor_Inhomogeneous_method_coronal.py:461-600
//
// What is different from the reference (same mathematics, B200-first mechanics):
//   * exterior: closed form (cosh/sinh, I_n/K_n) of the reference's own initial-value
//     problem instead of a 500-point odeint integration;
//   * interior: the ODE is linear, so the slope that fsolve searches for is obtained
//     from one integration (cylinder: from the axis outwards, where the end condition
//     is an initial condition) or two fundamental solutions (slab);
//   * sympy/lambdify coefficients -> rational expressions of the tabulated profile
//     (rho, rho') at fixed Runge-Kutta stage nodes, staged in shared memory.
#pragma once
#include "bessel.cuh"
#include "bessel_jy.cuh"

namespace esb {
#define ESB_RKN_ENUM
#include "rkn_tableau.h"
#undef ESB_RKN_ENUM
#if defined(__CUDACC__)
// NOT const-qualified: nvcc would fold the values back into 64-bit immediates
#define ESB_TABLE_QUAL __device__ __constant__
namespace dev_tables {
#include "rkn_tableau.h"
}
#undef ESB_TABLE_QUAL
#endif
#define ESB_TABLE_QUAL static const
namespace host_tables {
#include "rkn_tableau.h"
}
#undef ESB_TABLE_QUAL
#define RKN(name) ESB_TAB(esb_rkn8)[RKN_##name]
}  // namespace esb

namespace esb {

enum { KIND_SLAB_DENSITY = 0, KIND_CYL_DENSITY = 1, KIND_SLAB_FLOW = 2, KIND_CYL_ROTATION = 3, KIND_CYL_FLOW = 4 };

template <int KIND>
constexpr bool is_cyl_second_order = (KIND == KIND_CYL_DENSITY || KIND == KIND_CYL_FLOW);
// SCHEME_RK8N: the same Cooper-Verner method in Nystrom form on the normal form u'' = q u of the
// second-order kinds (integrate_layer_nform below)
enum { SCHEME_RK4 = 0, SCHEME_RK8 = 1, SCHEME_RK8N = 2 };
enum { OMEGA_SHARED = 0, OMEGA_PHASE_SPEED = 1, OMEGA_PER_K = 2 };

constexpr int TAB_FIELDS = 4;   // doubles per node in the staged table
constexpr int NF_FIELDS = 8;    // ... for the normal-form scheme (SCHEME_RK8N)
constexpr double NF_SQRT34 = 0.86602540378443864676;   // sqrt(3/4): see model_host.h

struct DevModel {
    int kind, scheme, n_steps, n_nodes;
    // exterior (uniform) medium
    double vAe2, ce2, cTe2, se2, rho_e;
    double ic_v, ic_s;        // exterior initial values
    double ext_len;           // exterior start = -ext_len / k
    // interior: c^2 = alpha/rho, vA^2 = beta/rho, cT^2 = tau/rho, S = alpha + beta
    double alpha, beta, tau, S, invS;
    double rho_b;             // density at the boundary s_start
    double s_start;           // boundary position (-1)
    double r_sign;            // cylinder: -1 scripts written in r<0 (coronal), +1 in r>0 (photospheric)
    // slab with a sheared flow U(x) / cylinder with an axial flow v_z(r): uniform interior c_i,
    // vA_i, rho_i; U_b = flow speed at the boundary s_start
    double ci2, vAi2, cTi2, si, rho_i, U_e, U_b;
    // slab kinds: the staged profile is mirror-symmetric about the mid-plane (checked by the host when the
    // model is set): the even and the odd solution are integrated from the mid-plane over HALF the layer
    int symmetric;
    // cylinder with rotational flow v_phi(r): uniform rho_i, vA_i (vAi2, rho_i above);
    // rho v_phi^2 at the boundary enters the kink end condition
    double rho_vb2;
    double r_axis;            // cylinder kinds: position of the first node (the axis end of the layer)
    double f_lo, f_hi;        // range of the first profile field (rho / v_z) over the layer (normal-form scheme)
};

// ---------------------------------------------------------------- tableau ----
constexpr double SQ21 = 4.58257569495584000658804719373;
constexpr double C8_M = (7.0 - SQ21) / 14.0;
constexpr double C8_P = (7.0 + SQ21) / 14.0;
// Cooper-Verner 8th-order, 11 stages (verified to order 8 in tests/test_tableau.py)
constexpr double a21 = 0.5;
constexpr double a31 = 0.25, a32 = 0.25;
constexpr double a41 = 1.0 / 7.0, a42 = (-7.0 - 3.0 * SQ21) / 98.0, a43 = (21.0 + 5.0 * SQ21) / 49.0;
constexpr double a51 = (11.0 + SQ21) / 84.0, a53 = (18.0 + 4.0 * SQ21) / 63.0, a54 = (21.0 - SQ21) / 252.0;
constexpr double a61 = (5.0 + SQ21) / 48.0, a63 = (9.0 + SQ21) / 36.0,
                 a64 = (-231.0 + 14.0 * SQ21) / 360.0, a65 = (63.0 - 7.0 * SQ21) / 80.0;
constexpr double a71 = (10.0 - SQ21) / 42.0, a73 = (-432.0 + 92.0 * SQ21) / 315.0,
                 a74 = (633.0 - 145.0 * SQ21) / 90.0, a75 = (-504.0 + 115.0 * SQ21) / 70.0,
                 a76 = (63.0 - 13.0 * SQ21) / 35.0;
constexpr double a81 = 1.0 / 14.0, a85 = (14.0 - 3.0 * SQ21) / 126.0, a86 = (13.0 - 3.0 * SQ21) / 63.0,
                 a87 = 1.0 / 9.0;
constexpr double a91 = 1.0 / 32.0, a95 = (91.0 - 21.0 * SQ21) / 576.0, a96 = 11.0 / 72.0,
                 a97 = (-385.0 - 75.0 * SQ21) / 1152.0, a98 = (63.0 + 13.0 * SQ21) / 128.0;
constexpr double a101 = 1.0 / 14.0, a105 = 1.0 / 9.0, a106 = (-733.0 - 147.0 * SQ21) / 2205.0,
                 a107 = (515.0 + 111.0 * SQ21) / 504.0, a108 = (-51.0 - 11.0 * SQ21) / 56.0,
                 a109 = (132.0 + 28.0 * SQ21) / 245.0;
constexpr double a115 = (-42.0 + 7.0 * SQ21) / 18.0, a116 = (-18.0 + 28.0 * SQ21) / 45.0,
                 a117 = (-273.0 - 53.0 * SQ21) / 72.0, a118 = (301.0 + 53.0 * SQ21) / 72.0,
                 a119 = (28.0 - 28.0 * SQ21) / 45.0, a1110 = (49.0 - 7.0 * SQ21) / 18.0;
constexpr double b8_1 = 1.0 / 20.0, b8_8 = 49.0 / 180.0, b8_9 = 16.0 / 45.0, b8_10 = 49.0 / 180.0,
                 b8_11 = 1.0 / 20.0;

ESB_HD int nodes_per_step(int scheme) { return scheme == SCHEME_RK4 ? 2 : 4; }

// ------------------------------------------------------------- point data ----
struct Point {
    double K, A;       // k^2, omega^2
    double k, w;       // k, omega
    // products that do not change along the layer
    double Kalpha, Kbeta, Ktau, SKtau, AKc;
};

ESB_HD Point make_point(const DevModel& M, double k, double w) {
    Point p;
    p.K = k * k;
    p.A = w * w;
    p.k = k;
    p.w = w;
    p.Kalpha = p.K * M.alpha;
    p.Kbeta = p.K * M.beta;
    p.Ktau = p.K * M.tau;
    p.SKtau = M.S * p.Ktau;
    p.AKc = -p.A * p.K * (M.tau - M.alpha);
    return p;
}

// Four reciprocals for the price of one: a double-precision division costs ~8 FP64-pipe instructions (and
// a slow-path check), and every step needs one per new stage node.  1/(p1 p2 p3 p4) and nine products give
// all four to ~3 ulp.  The arguments are products of two O(1)..O(1e32) factors (the w -> 0 clamps below keep
// the product of all four inside the double range); one of them vanishing or overflowing (a resonance
// exactly on a node: inside a continuum) spoils the four, which are noise there anyway.
ESB_HD void reciprocal4(const double (&p)[4], double (&inv)[4]) {
    const double p01 = p[0] * p[1], p23 = p[2] * p[3];
    const double r = 1.0 / (p01 * p23);
    const double i01 = r * p23, i23 = r * p01;
    inv[0] = i01 * p[1];
    inv[1] = i01 * p[0];
    inv[2] = i23 * p[3];
    inv[3] = i23 * p[2];
}

// y'' = a y' + b y  at one staged node (4 doubles f[0..3]).
//   cylinder: f = {1/r, 1/r^2, rho, rho'}
//       a = -1/r + rho' w^2/(rho w^2 - k^2 beta)
//       b = m^2/r^2 + k^2 - (rho w^2)^2/(S (rho w^2 - k^2 tau))
//     (Density_cylinder.py:742-756 with v_phi = B_phi = v_z = 0, B_i = B_0:
//      F = r/(rho w^2 - k^2 B_0^2), a = -F'/F, b = g/F.)
//   slab: f = {rho, rho', -, -}
//       a = -F'/F = -w^2 k^2 (tau-alpha) rho' / ((k^2 alpha - rho w^2)(k^2 tau - rho w^2))
//       b = m0^2 = (k^2 alpha - rho w^2)(k^2 beta - rho w^2)/(S (k^2 tau - rho w^2))
//     (..._coronal.py:222-230.)
// `b` excludes the azimuthal term; `bm` is its factor: b_total = b + m^2 * bm (bm = 1/r^2 for the
// cylinder, 0 for the slab), so that several azimuthal orders share one coefficient evaluation.
// node_coeffs in two halves around its one division, so that the four new stage nodes of a step share ONE
// reciprocal (reciprocal4 below): `pre` returns the product to invert and keeps what the second half needs.
struct CoefPre { double u, X, Y, Om, O2, t, sc, va, st; };

template <int KIND>
ESB_HD double node_coeffs_pre(const DevModel& M, const Point& p, const double* f, CoefPre& c) {
    if (KIND == KIND_CYL_DENSITY) {
        c.u = f[2] * p.A;
        c.X = c.u - p.Kbeta;
        c.Y = fma(M.S, c.u, -p.SKtau);
        return c.X * c.Y;
    } else if (KIND == KIND_CYL_FLOW) {
        c.Om = fma(-p.k, f[2], p.w);
        c.O2 = c.Om * c.Om;
        c.X = fma(-p.K, M.vAi2, c.O2);
        c.Y = M.si * fma(-p.K, M.cTi2, c.O2);
        return c.X * c.Y;
    } else if (KIND == KIND_SLAB_FLOW) {
        c.Om = fma(-p.k, f[0], p.w);
        c.O2 = c.Om * c.Om;
        c.t = fma(-p.K, M.cTi2, c.O2);
        c.sc = fma(-p.K, M.ci2, c.O2);          // Om^2 - k^2 c^2
        c.va = fma(-p.K, M.vAi2, c.O2);         // Om^2 - k^2 vA^2
        c.st = M.si * c.t;
        c.X = c.Om * c.sc;
        return c.st * c.X;
    } else {
        c.u = f[0] * p.A;
        c.X = p.Kalpha - c.u;                   // p1
        c.Y = p.Ktau - c.u;                     // p3
        return c.X * c.Y;
    }
}

template <int KIND>
ESB_HD void node_coeffs_fin(const DevModel& M, const Point& p, const double* f, const CoefPre& c, double inv,
                            double& a, double& b, double& bm) {
    if (KIND == KIND_CYL_DENSITY) {
        a = fma(f[3] * p.A * c.Y, inv, -f[0]);
        b = fma(-(c.u * c.u) * c.X, inv, p.K);
        bm = f[1];
    } else if (KIND == KIND_CYL_FLOW) {
        // f = {1/r, 1/r^2, v_z, v_z'}.  Cylinder_method_flow_testing.py:711-762 with v_phi = B_phi = 0
        // and uniform rho, c, vA:  Om = w - k v_z,  Q = T = C1 = 0,  C3 = D rho (Om^2 - wA^2),
        //   F = r D/C3 = r/(rho (Om^2 - k^2 vA^2)),   g = -r C2/D
        //   a = -F'/F = -1/r - 2 k v_z' Om/(Om^2 - k^2 vA^2)
        //   b = g/F   = m^2/r^2 + k^2 - Om^4/(s (Om^2 - k^2 cT^2))
        a = fma(-2.0 * p.k * f[3] * c.Om * c.Y, inv, -f[0]);
        b = fma(-(c.O2 * c.O2) * c.X, inv, p.K);
        bm = f[1];
    } else if (KIND == KIND_SLAB_FLOW) {
        // f = {U, U', U''}.  vx'' = -D vx' - coeff vx  (flow_multiprocessor_coronal.py:211-219,297):
        //   Om = w - k U,  t = Om^2 - k^2 cT^2
        //   m0 = (k^2 c^2 - Om^2)(k^2 vA^2 - Om^2)/(s (k^2 cT^2 - Om^2))
        //   D  = 2 k U' (t + k^4 cT^2 c^2/(s t)) / (Om (Om^2 - k^2 c^2))
        //   a = -D,  b = m0 - k U''/Om - k U' D/Om
        // one reciprocal for 1/(s t), 1/(Om sc): inv = 1/(st * Om * sc)
        const double inv_st = inv * c.X;
        const double inv_osc = inv * c.st;               // 1/(Om sc)
        const double m0 = -(c.sc * c.va) * inv_st;       // (Kc^2-O2)(KvA^2-O2)/(s(KcT^2-O2)) = -(sc va)/(s t)
        const double kdU = p.k * f[1];
        const double Dx = 2.0 * kdU * fma(p.K * p.K * M.cTi2 * M.ci2, inv_st, c.t) * inv_osc;
        const double invOm = inv_osc * c.sc;             // 1/Om
        a = -Dx;
        b = m0 - (p.k * f[2] + kdU * Dx) * invOm;
        bm = 0.0;
    } else {
        const double p2 = p.Kbeta - c.u;
        a = p.AKc * f[1] * inv;
        b = (c.X * c.X) * p2 * inv * M.invS;
        bm = 0.0;
    }
}

template <int KIND>
ESB_HD void node_coeffs(const DevModel& M, const Point& p, const double* f, double& a, double& b,
                        double& bm) {
    CoefPre c;
    const double prod = node_coeffs_pre<KIND>(M, p, f, c);
    node_coeffs_fin<KIND>(M, p, f, c, 1.0 / prod, a, b, bm);
}

// ------------------------------------------------------------- integrator ----
// One Cooper-Verner step of the linear system (u, v)' = f(node; u, v) for NS independent
// solutions.  RHS(s, n, U, V, hFU, hFV) returns the right-hand side of solution s at stage node n
// ALREADY MULTIPLIED BY THE STEP h (the step is folded into the node coefficients once per step),
// so every stage value is a pure FMA chain  y + sum_j a_ij (h f_j).
// stage -> node: 1:0  2:2 3:2  4:3 5:3  6:2  7:1 8:1  9:2  10:3  11:4
template <int NS, class RHS>
ESB_HD void rk8_generic(double (&y)[NS], double (&yp)[NS], const RHS& rhs) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double p = y[s], q = yp[s];
        double F1, G1, F2, G2, F3, G3, F4, G4, F5, G5, F6, G6, F7, G7, F8, G8, F9, G9, F10, G10, F11, G11;
        rhs(s, 0, p, q, F1, G1);
        const double P2 = fma(a21, F1, p), Q2 = fma(a21, G1, q);
        rhs(s, 2, P2, Q2, F2, G2);
        const double P3 = fma(a32, F2, fma(a31, F1, p)), Q3 = fma(a32, G2, fma(a31, G1, q));
        rhs(s, 2, P3, Q3, F3, G3);
        const double P4 = fma(a43, F3, fma(a42, F2, fma(a41, F1, p))),
                     Q4 = fma(a43, G3, fma(a42, G2, fma(a41, G1, q)));
        rhs(s, 3, P4, Q4, F4, G4);
        const double P5 = fma(a54, F4, fma(a53, F3, fma(a51, F1, p))),
                     Q5 = fma(a54, G4, fma(a53, G3, fma(a51, G1, q)));
        rhs(s, 3, P5, Q5, F5, G5);
        const double P6 = fma(a65, F5, fma(a64, F4, fma(a63, F3, fma(a61, F1, p)))),
                     Q6 = fma(a65, G5, fma(a64, G4, fma(a63, G3, fma(a61, G1, q))));
        rhs(s, 2, P6, Q6, F6, G6);
        const double P7 = fma(a76, F6, fma(a75, F5, fma(a74, F4, fma(a73, F3, fma(a71, F1, p))))),
                     Q7 = fma(a76, G6, fma(a75, G5, fma(a74, G4, fma(a73, G3, fma(a71, G1, q)))));
        rhs(s, 1, P7, Q7, F7, G7);
        const double P8 = fma(a87, F7, fma(a86, F6, fma(a85, F5, fma(a81, F1, p)))),
                     Q8 = fma(a87, G7, fma(a86, G6, fma(a85, G5, fma(a81, G1, q))));
        rhs(s, 1, P8, Q8, F8, G8);
        const double P9 = fma(a98, F8, fma(a97, F7, fma(a96, F6, fma(a95, F5, fma(a91, F1, p))))),
                     Q9 = fma(a98, G8, fma(a97, G7, fma(a96, G6, fma(a95, G5, fma(a91, G1, q)))));
        rhs(s, 2, P9, Q9, F9, G9);
        const double P10 = fma(a109, F9, fma(a108, F8, fma(a107, F7, fma(a106, F6, fma(a105, F5,
                                  fma(a101, F1, p)))))),
                     Q10 = fma(a109, G9, fma(a108, G8, fma(a107, G7, fma(a106, G6, fma(a105, G5,
                                  fma(a101, G1, q))))));
        rhs(s, 3, P10, Q10, F10, G10);
        const double P11 = fma(a1110, F10, fma(a119, F9, fma(a118, F8, fma(a117, F7, fma(a116, F6,
                                  fma(a115, F5, p)))))),
                     Q11 = fma(a1110, G10, fma(a119, G9, fma(a118, G8, fma(a117, G7, fma(a116, G6,
                                  fma(a115, G5, q))))));
        rhs(s, 4, P11, Q11, F11, G11);
        y[s] = fma(b8_11, F11, fma(b8_10, F10, fma(b8_9, F9, fma(b8_8, F8, fma(b8_1, F1, p)))));
        yp[s] = fma(b8_11, G11, fma(b8_10, G10, fma(b8_9, G9, fma(b8_8, G8, fma(b8_1, G1, q)))));
    }
}

// y'' = a y' + b_s y in the step-scaled variables (u, z) = (y, h y'):  h u' = z,
// h z' = (h a) z + (h^2 b_s) u.  ha = h a, h2bs = h^2 b_s: the first component's stage slope is the
// stage value of z itself (no multiplication), the second is one multiply + one FMA.
template <int NS>
struct RhsSecondOrder {
    const double (&ha)[5];
    const double (&h2bs)[NS][5];
    ESB_HDM void operator()(int s, int n, double U, double Z, double& FU, double& FZ) const {
        FU = Z;
        FZ = fma(ha[n], Z, h2bs[s][n] * U);
    }
};

// general 2x2 system (rotational-flow cylinder): (u, v)' = [[m11, m12], [m21, m22]] (u, v),
// the four entries pre-multiplied by h
struct RhsSystem {
    const double (&m11)[5];
    const double (&m12)[5];
    const double (&m21)[5];
    const double (&m22)[5];
    ESB_HDM void operator()(int, int n, double U, double V, double& FU, double& FV) const {
        FU = fma(m11[n], U, m12[n] * V);
        FV = fma(m21[n], U, m22[n] * V);
    }
};

// one step for (y, z = h y'); ha = h a, h2bs = h^2 b_s
template <int NS>
ESB_HD void rk8_step(double (&y)[NS], double (&z)[NS], const double (&ha)[5], const double (&h2bs)[NS][5]) {
    const RhsSecondOrder<NS> rhs{ha, h2bs};
    rk8_generic<NS>(y, z, rhs);
}

template <int NS>
ESB_HD void rk4_step(double (&y)[NS], double (&yp)[NS], double h, const double (&ca)[3],
                     const double (&cbs)[NS][3]) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double p = y[s], q = yp[s];
        const double(&cb)[3] = cbs[s];
        const double hh = 0.5 * h;
        const double G1 = fma(ca[0], q, cb[0] * p);
        const double P2 = fma(hh, q, p), Q2 = fma(hh, G1, q), G2 = fma(ca[1], Q2, cb[1] * P2);
        const double P3 = fma(hh, Q2, p), Q3 = fma(hh, G2, q), G3 = fma(ca[1], Q3, cb[1] * P3);
        const double P4 = fma(h, Q3, p), Q4 = fma(h, G3, q), G4 = fma(ca[2], Q4, cb[2] * P4);
        const double h6 = h * (1.0 / 6.0);
        y[s] = fma(h6, q + 2.0 * (Q2 + Q3) + Q4, p);
        yp[s] = fma(h6, G1 + 2.0 * (G2 + G3) + G4, q);
    }
}

// Integrate NS solutions along the staged mesh.  tab: [n_nodes][TAB_FIELDS], then h[n_steps], then
// g[n_steps] = h[i+1]/h[i] (g[n_steps-1] = 1/h[n_steps-1]).
// m2[s] = (azimuthal order)^2 of solution s (cylinder); the node coefficients are evaluated
// once per node and shared by all solutions.  RK8 runs in the step-scaled variables (y, z = h y'):
// g[i] rescales z from one step to the next and back to y' after the last one.
template <int KIND, int SCHEME, int NS, bool RANGE = false>
ESB_HD void integrate_layer(const DevModel& M, const Point& pt, const double* __restrict__ tab,
                            const double (&m2)[NS], double (&y)[NS], double (&yp)[NS], int r0 = 0,
                            int r1 = 0) {
    // RANGE: steps [r0, r1) of the mesh, else all of it; (y, y') in and out are unscaled
    const int i0 = RANGE ? r0 : 0;
    const int iend = RANGE ? r1 : M.n_steps;
    const double* hs = tab + (size_t)M.n_nodes * TAB_FIELDS;
    const double* gs = hs + M.n_steps;
    constexpr int NPS = (SCHEME == SCHEME_RK8) ? 4 : 2;
    constexpr int NN = NPS + 1;
    double a0, b0, bm0;
    node_coeffs<KIND>(M, pt, tab + (size_t)(i0 * NPS) * TAB_FIELDS, a0, b0, bm0);
    if constexpr (SCHEME == SCHEME_RK8) {
        const double h0 = hs[i0];
#pragma unroll
        for (int s = 0; s < NS; ++s) yp[s] *= h0;
    }
    for (int i = i0; i < iend; ++i) {
        const double* f = tab + (size_t)(i * NPS) * TAB_FIELDS;
        const double h = hs[i];
        double ca[NN], cb[NN], bm[NN], cbs[NS][NN];
        ca[0] = a0; cb[0] = b0; bm[0] = bm0;
        if constexpr (NN == 5) {                 // one reciprocal for the four new stage nodes
            CoefPre pre[4];
            double prod[4], inv[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) prod[n] = node_coeffs_pre<KIND>(M, pt, f + (n + 1) * TAB_FIELDS, pre[n]);
            reciprocal4(prod, inv);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                node_coeffs_fin<KIND>(M, pt, f + (n + 1) * TAB_FIELDS, pre[n], inv[n], ca[n + 1], cb[n + 1], bm[n + 1]);
        } else {
#pragma unroll
            for (int n = 1; n < NN; ++n) node_coeffs<KIND>(M, pt, f + n * TAB_FIELDS, ca[n], cb[n], bm[n]);
        }
        a0 = ca[NN - 1]; b0 = cb[NN - 1]; bm0 = bm[NN - 1];       // unscaled, carried to the next step
        if constexpr (SCHEME == SCHEME_RK8) {
            // fold the step into the node coefficients (shared by all solutions): h a, h^2 b
            const double h2 = h * h;
#pragma unroll
            for (int n = 0; n < NN; ++n) { ca[n] *= h; cb[n] *= h2; bm[n] *= h2; }
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int n = 0; n < NN; ++n) cbs[s][n] = is_cyl_second_order<KIND> ? fma(m2[s], bm[n], cb[n]) : cb[n];
        if constexpr (SCHEME == SCHEME_RK8) {
            rk8_step<NS>(y, yp, ca, cbs);
            const double g = gs[i];
#pragma unroll
            for (int s = 0; s < NS; ++s) yp[s] *= g;
        } else {
            rk4_step<NS>(y, yp, h, ca, cbs);
        }
    }
    if constexpr (SCHEME == SCHEME_RK8 && RANGE) {
        if (iend < M.n_steps) {          // z is in the scale of step iend: back to y'
            const double ih = 1.0 / hs[iend];
#pragma unroll
            for (int s = 0; s < NS; ++s) yp[s] *= ih;
        }
    }
}

// ---- cylinder second-order kinds, RK8: coefficients evaluated already scaled -------------------
// The staged node is {h/r, h^2/r^2, field, h field'} (h = the step the node belongs to), so the
// step-scaled coefficients h a, h^2 b, h^2/r^2 come out of the evaluation directly:
//   density   X = rho - q, Y = rho - p  (q = k^2 beta/w^2, p = k^2 tau/w^2: per point)
//             h a = (h rho') Y/(X Y) - h/r        h^2 b = h^2 k^2 - (h^2 w^2/S) rho^2 X/(X Y)
//   flow      Om = w - k v_z, X = Om^2 - k^2 vA^2, Y = Om^2 - k^2 cT^2
//             h a = -2 k (h v_z') Om Y/(X Y) - h/r   h^2 b = h^2 k^2 - (h^2/s) Om^4 X/(X Y)
// 15 FP64 instructions per node instead of 18 + 3 for the scaling.  The coefficients of the node
// shared with the next step are carried and rescaled by g = h'/h (h a) and g^2 (h^2 b, h^2/r^2).
struct ScaledPoint {
    double q, p, AS;     // density: k^2 beta/w^2, k^2 tau/w^2, w^2/S
    double m2k;          // flow: -2 k
    double inv_s;        // flow: 1/(c^2 + vA^2)
};

template <int KIND>
ESB_HD ScaledPoint make_scaled_point(const DevModel& M, const Point& pt) {
    ScaledPoint sp{};
    if constexpr (KIND == KIND_CYL_DENSITY) {
        // w -> 0: q, p -> 1e30-ish, 1/(X Y) vanishes against the O(1) terms and a = -1/r, b = k^2, the w -> 0
        // limit (the clamp keeps the product of FOUR node denominators inside the double range: reciprocal4)
        const double A = pt.A > 1e-30 ? pt.A : 1e-30;
        sp.q = pt.Kbeta / A;
        sp.p = pt.Ktau / A;
        sp.AS = A / M.S;
    } else if constexpr (KIND == KIND_CYL_FLOW) {
        sp.m2k = -2.0 * pt.k;
        sp.inv_s = 1.0 / M.si;
    }
    return sp;
}

// NFTAB: the node is read from the normal-form table (NF_FIELDS doubles: {-h/(2r) | h^2, h^2/r^2, field,
// h field', ...}) instead of the 4-field pre-scaled one - the normal-form scheme integrates the points
// next to a resonance in these variables (shoot_layer) from its own table.
// (in two halves around the division, like node_coeffs: `pre` returns the product to invert)
template <int KIND>
ESB_HD double node_scaled_pre(const DevModel& M, const Point& pt, const ScaledPoint& sp, const double* f, CoefPre& c) {
    if constexpr (KIND == KIND_SLAB_DENSITY) {
        c.u = f[2] * pt.A;
        c.X = pt.Kalpha - c.u;                  // p1
        c.Y = pt.Ktau - c.u;                    // p3
    } else if constexpr (KIND == KIND_CYL_DENSITY) {
        c.X = f[2] - sp.q;
        c.Y = f[2] - sp.p;
    } else {
        c.Om = fma(-pt.k, f[2], pt.w);
        c.O2 = c.Om * c.Om;
        c.X = fma(-pt.K, M.vAi2, c.O2);
        c.Y = fma(-pt.K, M.cTi2, c.O2);
    }
    return c.X * c.Y;
}

template <int KIND, bool NFTAB>
ESB_HD void node_scaled_fin(const DevModel& M, const Point& pt, const ScaledPoint& sp, double c1, double h2K,
                            const double* f, const CoefPre& c, double inv, double& ha, double& h2b, double& h2bm) {
    if constexpr (KIND == KIND_SLAB_DENSITY) {
        // f = {h^2, -, rho, h rho'} (normal-form table only)
        const double p2 = pt.Kbeta - c.u;
        ha = (pt.AKc * f[3]) * inv;
        h2b = ((c.X * c.X) * p2) * (inv * c1);                // c1 = h^2/S
        h2bm = 0.0;
    } else {
        const double nhinvr = NFTAB ? f[6] + f[6] : -f[0];    // -h/r
        h2bm = f[1];
        if constexpr (KIND == KIND_CYL_DENSITY) {
            const double rho = f[2], hdrho = NFTAB ? f[7] : f[3];
            ha = fma(hdrho * c.Y, inv, nhinvr);
            h2b = fma(-((rho * rho) * c.X) * c1, inv, h2K);       // c1 = h^2 w^2/S
        } else {
            const double hdvz = NFTAB ? f[7] : f[3];
            ha = fma((sp.m2k * hdvz) * c.Om * c.Y, inv, nhinvr);
            h2b = fma(-((c.O2 * c.O2) * c.X) * c1, inv, h2K);     // c1 = h^2/s
        }
    }
}

template <int KIND, bool NFTAB = false>
ESB_HD void node_coeffs_scaled(const DevModel& M, const Point& pt, const ScaledPoint& sp, double c1, double h2K,
                               const double* f, double& ha, double& h2b, double& h2bm) {
    CoefPre c;
    const double prod = node_scaled_pre<KIND>(M, pt, sp, f, c);
    node_scaled_fin<KIND, NFTAB>(M, pt, sp, c1, h2K, f, c, 1.0 / prod, ha, h2b, h2bm);
}

template <int KIND, int NS, bool RANGE = false, bool NFTAB = false>
ESB_HD void integrate_layer_prescaled(const DevModel& M, const Point& pt, const double* __restrict__ tab,
                                      const double (&m2)[NS], double (&y)[NS], double (&yp)[NS], int r0 = 0,
                                      int r1 = 0) {
    // RANGE: steps [r0, r1) of the mesh, else all of it; (y, y') in and out are unscaled
    constexpr int TF = NFTAB ? NF_FIELDS : TAB_FIELDS;
    const int i0 = RANGE ? r0 : 0;
    const int iend = RANGE ? r1 : M.n_steps;
    const double* hs = tab + (size_t)M.n_nodes * TF;
    const double* gs = hs + M.n_steps;
    const double* hs2 = gs + M.n_steps;          // h^2, (h'/h)^2: staged
    const double* gs2 = hs2 + M.n_steps;
    const ScaledPoint sp = make_scaled_point<KIND>(M, pt);
    const double cc = (KIND == KIND_CYL_DENSITY) ? sp.AS : (KIND == KIND_SLAB_DENSITY) ? M.invS : sp.inv_s;
    double ha0, h2b0, h2bm0;
    if (!RANGE || i0 == 0) {
        const double h0 = hs[0], h2 = h0 * h0;
        node_coeffs_scaled<KIND, NFTAB>(M, pt, sp, h2 * cc, h2 * pt.K, tab, ha0, h2b0, h2bm0);
    } else {
        // node 4 i0 is stored in the scale of the step it ends: evaluate it there, rescale like the carry
        const double hp = hs[i0 - 1], h2 = hp * hp, g = gs[i0 - 1], g2 = g * g;
        node_coeffs_scaled<KIND, NFTAB>(M, pt, sp, h2 * cc, h2 * pt.K, tab + (size_t)(i0 * 4) * TF, ha0, h2b0,
                                        h2bm0);
        ha0 *= g; h2b0 *= g2; h2bm0 *= g2;
    }
    {
        const double h0 = hs[i0];
#pragma unroll
        for (int s = 0; s < NS; ++s) yp[s] *= h0;
    }
    for (int i = i0; i < iend; ++i) {
        const double* f = tab + (size_t)(i * 4) * TF;
        const double h2 = hs2[i];
        const double c1 = h2 * cc, h2K = h2 * pt.K;
        double ha[5], h2b[5], h2bm[5], h2bs[NS][5];
        ha[0] = ha0; h2b[0] = h2b0; h2bm[0] = h2bm0;
        {                                        // one reciprocal for the four new stage nodes
            CoefPre pre[4];
            double prod[4], inv[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) prod[n] = node_scaled_pre<KIND>(M, pt, sp, f + (n + 1) * TF, pre[n]);
            reciprocal4(prod, inv);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                node_scaled_fin<KIND, NFTAB>(M, pt, sp, c1, h2K, f + (n + 1) * TF, pre[n], inv[n], ha[n + 1],
                                             h2b[n + 1], h2bm[n + 1]);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int n = 0; n < 5; ++n)
                h2bs[s][n] = is_cyl_second_order<KIND> ? fma(m2[s], h2bm[n], h2b[n]) : h2b[n];
        rk8_step<NS>(y, yp, ha, h2bs);
        const double g = gs[i], g2 = gs2[i];
        ha0 = ha[4] * g; h2b0 = h2b[4] * g2; h2bm0 = h2bm[4] * g2;        // the shared node, in the next step's scale
#pragma unroll
        for (int s = 0; s < NS; ++s) yp[s] *= g;
    }
    if (RANGE && iend < M.n_steps) {     // z is in the scale of step iend: back to y'
        const double ih = 1.0 / hs[iend];
#pragma unroll
        for (int s = 0; s < NS; ++s) yp[s] *= ih;
    }
}

// ---- normal form u'' = q u, Cooper-Verner in Nystrom form (SCHEME_RK8N) ------------------------
// Every second-order kind is (F y')' = g y, i.e. y'' = a y' + b y with a = -F'/F.  The substitution
// u = sqrt|F| y removes the first-derivative term:
//     u'' = q u,    q = b - a'/2 + a^2/4
// and a Runge-Kutta method applied to (u, u')' = (u', q u) never needs the stages of u' on their own
// (tools/gen_rkn_tableau.py): the SAME 8th-order method costs 64 FP64 instructions per solution and
// step instead of 110 (40 + 4 + 5 products with (A A), b^T A, b; four "c_i z" bases shared by the
// stages on the same node; 11 products q_i U_i), with the same four stage nodes.
//   cylinder density  F = r/(rho w^2 - k^2 beta):  X = rho - qh, Y = rho - ph, L = rho'/X
//       q = k^2 + (n^2 - 1/4)/r^2 - (w^2/S) rho^2/Y - rho''/(2X) + 3/4 L^2 - L/(2r)
//   cylinder axial flow  F = r/(rho (Om^2 - k^2 vA^2)):  X = Om^2 - k^2 vA^2, Y = Om^2 - k^2 cT^2,
//       X' = -2 k Om v', X'' = 2 k^2 v'^2 - 2 k Om v'', L = X'/X
//       q = k^2 + (n^2 - 1/4)/r^2 - Om^4/(s Y) - X''/(2X) + 3/4 L^2 - L/(2r)
//   slab density  F = P3/P1, P_j = {alpha, beta, tau} k^2/w^2 - rho, i_j = 1/P_j, a = rho' (i3 - i1)
//       q = (w^2/S) P1 P2 i3 - (i3 - i1)(rho'' + rho'^2 (i3 + i1))/2 + a^2/4
// The staged node holds the step-scaled (k, omega)-independent pieces (model_host.h):
//   cylinder {-h/(2r), h^2/r^2, field, h field', -h^2 field''/2, h^2 field^2}
//   slab     {h^2,     -,       rho,   h rho',   -h^2 rho''/2,   -}
// The end conditions and the matching are stated for (y, y'); nform_edge gives a and the factor of F
// that varies along the layer at a node, so that (y, y') = (u, u' + a u/2)/sqrt|F| at both ends.
struct NPoint {
    double q, p, t;      // density kinds: k^2 {beta, tau, alpha}/w^2
    double AS;           // density kinds: w^2/S
    double dd;           // slab: (alpha - tau) k^2/w^2
    double m2k;          // flow: -2 k
    double inv_s;        // flow: 1/(c^2 + vA^2)
    double K43;          // flow: k^2 / (3/4)  (the staged first derivative carries a factor sqrt(3/4))
};

template <int KIND>
ESB_HD NPoint make_npoint(const DevModel& M, const Point& pt) {
    static_assert(KIND == KIND_CYL_DENSITY || KIND == KIND_CYL_FLOW || KIND == KIND_SLAB_DENSITY,
                  "kinds with a normal form");
    NPoint sp{};
    if constexpr (KIND == KIND_CYL_FLOW) {
        sp.m2k = -2.0 * pt.k;
        sp.inv_s = 1.0 / M.si;
        sp.K43 = pt.K * (4.0 / 3.0);
    } else {
        // w -> 0: the ratios -> 1e30-ish, the reciprocals of their products vanish against the O(1) terms and
        // q = k^2 (+ the azimuthal term), the w -> 0 limit (the clamp keeps the product of FOUR node
        // denominators inside the double range: reciprocal4)
        const double A = pt.A > 1e-30 ? pt.A : 1e-30;
        sp.q = pt.Kbeta / A;
        sp.p = pt.Ktau / A;
        sp.t = pt.Kalpha / A;
        sp.AS = A / M.S;
        sp.dd = sp.t - sp.p;
    }
    return sp;
}

// The two factors whose product a node's coefficients divide by (and, flow kinds, the Doppler-shifted
// frequency they are built from): first half of node_q / node_coeffs_scaled, before the reciprocal.
struct NodeDen { double X, Y, Om, O2; };

template <int KIND>
ESB_HD double node_den(const DevModel& M, const Point& pt, const NPoint& sp, const double* f, NodeDen& d) {
    if constexpr (KIND == KIND_CYL_DENSITY) {
        d.X = f[2] - sp.q;
        d.Y = f[2] - sp.p;
    } else if constexpr (KIND == KIND_CYL_FLOW) {
        d.Om = fma(-pt.k, f[2], pt.w);
        d.O2 = d.Om * d.Om;
        d.X = fma(-pt.K, M.vAi2, d.O2);
        d.Y = fma(-pt.K, M.cTi2, d.O2);
    } else {
        d.X = sp.t - f[2];                 // P1
        d.Y = sp.p - f[2];                 // P3
    }
    return d.X * d.Y;
}

// second half: h^2 q and h^2/r^2 from the factors and inv = 1/(X Y)
template <int KIND>
ESB_HD void node_q_fin(const DevModel& M, const Point& pt, const NPoint& sp, double c1, double h2K, const double* f,
                       const NodeDen& d, double inv, double& h2q, double& h2m) {
    if constexpr (KIND == KIND_CYL_DENSITY) {
        // f[3] = c h rho', f[0] = -h/(2r)/c, c = sqrt(3/4):  3/4 L^2 - L h/(2r) = L'^2 + L' f[0],  L' = c L
        const double iX = inv * d.Y, iY = inv * d.X;
        const double L = f[3] * iX;
        double acc = fma(-(sp.AS * f[5]), iY, h2K);
        acc = fma(f[4], iX, acc);
        acc = fma(L, L, acc);
        h2q = fma(L, f[0], acc);
        h2m = f[1];
    } else if constexpr (KIND == KIND_CYL_FLOW) {
        const double hdv = f[3];                               // c h v'
        const double iX = inv * d.Y, iY = inv * d.X;
        const double kO = sp.m2k * d.Om;                       // -2 k Om
        const double L = (kO * hdv) * iX;                      // c h X'/X
        const double W = fma(sp.K43 * hdv, hdv, -kO * f[4]);   // h^2 X''/2 = k^2 (h v')^2 - k Om h^2 v''
        double acc = fma(-((d.O2 * d.O2) * c1), iY, h2K);
        acc = fma(-W, iX, acc);
        acc = fma(L, L, acc);
        h2q = fma(L, f[0], acc);
        h2m = f[1];
    } else {
        const double hdr = f[3];
        const double P1 = d.X, P3 = d.Y, P2 = sp.q - f[2];
        const double i3 = inv * P1;
        const double dd = inv * sp.dd;                         // i3 - i1
        const double sm = inv * (P1 + P3);                     // i3 + i1
        const double E = hdr * dd;                             // h a
        double acc = ((c1 * P1) * P2) * i3;                    // h^2 b, c1 = h^2 w^2/S
        acc = fma(f[4], dd, acc);
        acc = fma(-0.5 * (E * hdr), sm, acc);
        h2q = fma(0.25 * E, E, acc);
        h2m = 0.0;
    }
}

// h^2 q (without the azimuthal term) and its azimuthal factor h^2/r^2 at one staged node.
// c1 = h^2 w^2/S (density), h^2/s (flow); h2K = h^2 k^2.
template <int KIND>
ESB_HD void node_q(const DevModel& M, const Point& pt, const NPoint& sp, double c1, double h2K, const double* f,
                   double& h2q, double& h2m) {
    NodeDen d;
    const double prod = node_den<KIND>(M, pt, sp, f, d);
    node_q_fin<KIND>(M, pt, sp, c1, h2K, f, d, 1.0 / prod, h2q, h2m);
}

// h a (h = the step the node is stored in) and the varying factor of F at a node:
// F = r/(rho_0 Xf) (cylinder), F = Xf (slab: P3/P1)
template <int KIND>
ESB_HD void nform_edge(const DevModel& M, const Point& pt, const NPoint& sp, const double* f, double& ha, double& Xf) {
    if constexpr (KIND == KIND_CYL_DENSITY) {
        const double X = f[2] - sp.q;
        ha = fma(f[7], 1.0 / X, 2.0 * f[6]);
        Xf = X;
    } else if constexpr (KIND == KIND_CYL_FLOW) {
        const double Om = fma(-pt.k, f[2], pt.w);
        const double X = fma(-pt.K, M.vAi2, Om * Om);
        ha = fma((sp.m2k * Om) * f[7], 1.0 / X, 2.0 * f[6]);
        Xf = X;
    } else {
        const double P1 = sp.t - f[2], P3 = sp.p - f[2];
        ha = f[3] * (sp.dd / (P1 * P3));
        Xf = P3 / P1;
    }
}

// One step for NS solutions in the step-scaled variables (u, z = h u'); q[s][n] = h^2 q_s at the five
// stage nodes {0, (7-sqrt21)/14, 1/2, (7+sqrt21)/14, 1}.  stage -> node as in rk8_generic.
// The tableau coefficients are 64-bit immediates: compiled for 3 resident CTAs per SM (esb.cu
// ESB_GRID_MINB) they stay in uniform registers across the step loop (385 instructions per fused step,
// 290 of them FP64); compiled for 4 the loop re-materialises them with two UMOVs each (527
// instructions, issue-bound: 33.0 instead of 30.5 ms on the bench grid).  Reading them from shared
// memory was tried: nvcc hoists the loads into ~250 registers (2 CTAs/SM, 37.2 ms) or spills them.
template <int NS>
ESB_HD void rkn8_step(double (&u)[NS], double (&z)[NS], const double (&q)[NS][5]) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const double U = u[s], Z = z[s];
        const double(&Q)[5] = q[s];
        const double BH = fma(0.5, Z, U), BM = fma(RKN(c_m), Z, U), BP = fma(RKN(c_p), Z, U), B1 = U + Z;
        const double F1 = Q[0] * U;
        const double F2 = Q[2] * BH;
        const double F3 = Q[2] * fma(RKN(n3_1), F1, BH);
        const double F4 = Q[3] * fma(RKN(n4_2), F2, fma(RKN(n4_1), F1, BP));
        const double F5 = Q[3] * fma(RKN(n5_3), F3, fma(RKN(n5_2), F2, fma(RKN(n5_1), F1, BP)));
        const double F6 = Q[2] * fma(RKN(n6_4), F4, fma(RKN(n6_3), F3, fma(RKN(n6_2), F2, fma(RKN(n6_1), F1, BH))));
        const double F7 = Q[1] * fma(RKN(n7_5), F5, fma(RKN(n7_4), F4, fma(RKN(n7_3), F3, fma(RKN(n7_2), F2, fma(RKN(n7_1), F1, BM)))));
        const double F8 = Q[1] * fma(RKN(n8_6), F6, fma(RKN(n8_5), F5, fma(RKN(n8_4), F4, fma(RKN(n8_3), F3, fma(RKN(n8_1), F1, BM)))));
        const double F9 = Q[2] * fma(RKN(n9_7), F7, fma(RKN(n9_6), F6, fma(RKN(n9_5), F5, fma(RKN(n9_4), F4, fma(RKN(n9_3), F3,
                                 fma(RKN(n9_1), F1, BH))))));
        const double F10 = Q[3] * fma(RKN(n10_8), F8, fma(RKN(n10_7), F7, fma(RKN(n10_6), F6, fma(RKN(n10_5), F5, fma(RKN(n10_4), F4,
                                  fma(RKN(n10_3), F3, fma(RKN(n10_1), F1, BP)))))));
        const double F11 = Q[4] * fma(RKN(n11_9), F9, fma(RKN(n11_8), F8, fma(RKN(n11_7), F7, fma(RKN(n11_6), F6, fma(RKN(n11_5), F5,
                                  fma(RKN(n11_4), F4, fma(RKN(n11_3), F3, B1)))))));
        u[s] = fma(RKN(nb_10), F10, fma(RKN(nb_9), F9, fma(RKN(nb_8), F8, fma(RKN(nb_1), F1, B1))));
        z[s] = fma(RKN(b_11), F11, fma(RKN(b_10), F10, fma(RKN(b_9), F9, fma(RKN(b_8), F8, fma(RKN(b_1), F1, Z)))));
    }
}

// NS solutions of u'' = (q + m2[s] bm) u along the staged mesh; (u, u') in and out unscaled.
// m2[s] = n^2 - 1/4 (cylinder), unused for the slab.  RANGE: steps [r0, r1), else all of the mesh.
template <int KIND, int NS, bool RANGE = false>
ESB_HD void integrate_layer_nform(const DevModel& M, const Point& pt, const NPoint& sp, const double* __restrict__ tab,
                                  const double (&m2)[NS], double (&u)[NS], double (&up)[NS], int r0 = 0,
                                  int r1 = 0) {
    constexpr bool CYL = is_cyl_second_order<KIND>;
    const int i0 = RANGE ? r0 : 0;
    const int iend = RANGE ? r1 : M.n_steps;
    const double* hs = tab + (size_t)M.n_nodes * NF_FIELDS;
    const double* gs = hs + M.n_steps;
    const double* hs2 = gs + M.n_steps;          // h^2, (h'/h)^2: staged, not recomputed every step
    const double* gs2 = hs2 + M.n_steps;
    const double cc = (KIND == KIND_CYL_FLOW) ? sp.inv_s : sp.AS;
    double q0, bm0;
    if (!RANGE || i0 == 0) {
        const double h2 = hs[0] * hs[0];
        node_q<KIND>(M, pt, sp, h2 * cc, h2 * pt.K, tab, q0, bm0);
    } else {
        // node 4 i0 is stored in the scale of the step it ends: evaluate it there, rescale like the carry
        const double hp = hs[i0 - 1], h2 = hp * hp, g = gs[i0 - 1], g2 = g * g;
        node_q<KIND>(M, pt, sp, h2 * cc, h2 * pt.K, tab + (size_t)(i0 * 4) * NF_FIELDS, q0, bm0);
        q0 *= g2; bm0 *= g2;
    }
    {
        const double h0 = hs[i0];
#pragma unroll
        for (int s = 0; s < NS; ++s) up[s] *= h0;
    }
    for (int i = i0; i < iend; ++i) {
        const double* f = tab + (size_t)(i * 4) * NF_FIELDS;
        const double h2 = hs2[i];
        const double c1 = h2 * cc, h2K = h2 * pt.K;
        double q[5], bm[5], qs[NS][5];
        q[0] = q0; bm[0] = bm0;
        {
            NodeDen den[4];
            double prod[4], inv[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) prod[n] = node_den<KIND>(M, pt, sp, f + (n + 1) * NF_FIELDS, den[n]);
            reciprocal4(prod, inv);
#pragma unroll
            for (int n = 0; n < 4; ++n)
                node_q_fin<KIND>(M, pt, sp, c1, h2K, f + (n + 1) * NF_FIELDS, den[n], inv[n], q[n + 1], bm[n + 1]);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int n = 0; n < 5; ++n) qs[s][n] = CYL ? fma(m2[s], bm[n], q[n]) : q[n];
        rkn8_step<NS>(u, up, qs);
        const double g = gs[i], g2 = gs2[i];
        q0 = q[4] * g2; bm0 = bm[4] * g2;                      // the shared node, in the next step's scale
#pragma unroll
        for (int s = 0; s < NS; ++s) up[s] *= g;
    }
    if (RANGE && iend < M.n_steps) {     // z is in the scale of step iend: back to u'
        const double ih = 1.0 / hs[iend];
#pragma unroll
        for (int s = 0; s < NS; ++s) up[s] *= ih;
    }
}

// (a, varying factor of F) at breakpoint `step` of the mesh (node 4 step), a unscaled
template <int KIND>
ESB_HD void nform_end_values(const DevModel& M, const Point& pt, const NPoint& sp, const double* __restrict__ tab,
                             int step, double& a, double& Xf) {
    const double* hs = tab + (size_t)M.n_nodes * NF_FIELDS;
    double ha;
    nform_edge<KIND>(M, pt, sp, tab + (size_t)(step * 4) * NF_FIELDS, ha, Xf);
    a = ha / hs[step > 0 ? step - 1 : 0];
}

// ------------------------------------------------- rotational-flow cylinder ----
// Reference: Twisted_photospheric_nonlinear_flow_kink_fast.py:264-297.  With B_phi = 0, v_z = 0,
// uniform rho and B:  Om = w - m v_phi/r,  a1 = Om^2 - k^2 vA^2,  A2 = (c^2+vA^2)(Om^2 - w_c^2)
//   D  = rho a1 A2                      Q = -a1 rho v_phi^2/r          T = rho v_phi Om
//   C1 = Q Om^2 - 2 m A2 T/r^2          C2 = Om^4 - A2 (m^2/r^2 + k^2)
//   C3 = D (rho a1 + r d/dr(-rho v_phi^2/r^2)) + Q^2 - 4 A2 T^2/r^2
// The reference eliminates xi from  D P' = C3 xi - C1 P,  D (r xi)' = C1 r xi - C2 r P  to get
// P'' = -(F'/F) P' + (g/F) P with F = r D/C3 (sympy-differentiated).  The kernel integrates the
// first-order pair in (P, xi) directly - the same solutions, no coefficient derivatives:
//   P'  = -(C1/D) P + (C3/D) xi          xi' = -(C2/D) P + (C1/D - 1/r) xi
// staged node f = {1/r, v_phi, r d/dr(-rho v_phi^2/r^2), c^2}.
// The staged rotation node holds the (k, omega, m)-independent products, 8 doubles:
//   {1/r, 1/r^2, v_phi/r, c^2 + vA^2, c^2, rho v_phi, rho v_phi^2/r, r d/dr(-rho v_phi^2/r^2)}
constexpr int ROT_FIELDS = 8;

// unscaled pieces of the system matrix: m11 = -C1/D, m12 = C3/D, m21 = -C2/D, m22 = C1/D - 1/r
struct RotCoef { double C1, C2, C3, invD, invr, Dd; };

// INVERT = false leaves invD to the caller (integrate_rotation inverts the four new nodes of a step together)
template <bool INVERT = true>
ESB_HD RotCoef node_rot(const DevModel& M, const Point& p, double m, const double* f) {
    const double invr = f[0], invr2 = f[1], vr = f[2], s = f[3], c2 = f[4], rv = f[5], q0 = f[6], f2 = f[7];
    const double Om = fma(-m, vr, p.w);
    const double O2 = Om * Om;
    const double wA2 = p.K * M.vAi2;
    const double a1 = O2 - wA2;
    const double A2 = fma(O2, s, -wA2 * c2);             // s (Om^2 - w_c^2), w_c^2 = w_A^2 c^2/s
    const double ra1 = M.rho_i * a1;
    const double Dd = ra1 * A2;
    const double Q = -a1 * q0;
    const double T = rv * Om;
    const double TA = (A2 * T) * invr2;
    RotCoef c;
    c.C1 = fma(Q, O2, -2.0 * m * TA);
    c.C2 = fma(O2, O2, -A2 * fma(m * m, invr2, p.K));
    c.C3 = fma(Dd, ra1 + f2, fma(Q, Q, -4.0 * TA * T));
    c.Dd = Dd;
    if (INVERT) c.invD = 1.0 / Dd;
    c.invr = invr;
    return c;
}

// the four entries of the system matrix at one node, multiplied by the step h
ESB_HD void rot_scaled(const RotCoef& c, double h, double& m11, double& m12, double& m21, double& m22) {
    const double hD = h * c.invD;
    m11 = -c.C1 * hD;
    m12 = c.C3 * hD;
    m21 = -c.C2 * hD;
    m22 = fma(c.C1, hD, -h * c.invr);
}

// NS solutions of the (P, xi) system along the staged mesh (axis end -> boundary)
template <int NS, bool RANGE = false>
ESB_HD void integrate_rotation(const DevModel& M, const Point& pt, double m, const double* __restrict__ tab,
                               double (&P)[NS], double (&X)[NS], int r0 = 0, int r1 = 0) {
    const int i0 = RANGE ? r0 : 0;
    const int iend = RANGE ? r1 : M.n_steps;
    const double* hs = tab + (size_t)M.n_nodes * ROT_FIELDS;
    RotCoef c0 = node_rot(M, pt, m, tab + (size_t)(i0 * 4) * ROT_FIELDS);
    for (int i = i0; i < iend; ++i) {
        const double* f = tab + (size_t)(i * 4) * ROT_FIELDS;
        const double h = hs[i];
        double m11[5], m12[5], m21[5], m22[5];
        rot_scaled(c0, h, m11[0], m12[0], m21[0], m22[0]);
        {                                        // one reciprocal for the four new stage nodes
            RotCoef cn[4];
            double prod[4], inv[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                cn[n] = node_rot<false>(M, pt, m, f + (n + 1) * ROT_FIELDS);
                prod[n] = cn[n].Dd;
            }
            reciprocal4(prod, inv);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                cn[n].invD = inv[n];
                rot_scaled(cn[n], h, m11[n + 1], m12[n + 1], m21[n + 1], m22[n + 1]);
            }
            c0 = cn[3];
        }
        const RhsSystem rhs{m11, m12, m21, m22};
        rk8_generic<NS>(P, X, rhs);
    }
}

// --------------------------------------------------------------- exterior ----
ESB_HD double m_e2(const DevModel& M, double K, double A) {
    // Density_cylinder.py:699
    return ((K * M.vAe2 - A) * (K * M.ce2 - A)) / (M.se2 * (K * M.cTe2 - A));
}

// Exact solution at x = -1 of the reference's exterior initial-value problem, slab.
ESB_HD void exterior_slab(const DevModel& M, double k, double me, double& yb, double& ypb) {
    // vx'' = m_e vx   (..._coronal.py:245), from -x0 to -1
    const double L = M.ext_len / k - 1.0;
    if (me < 0.0) {
        // the leaky side (only reached by the opt-in leaky evaluation: the reference skips m_e < 0): the
        // same initial-value problem has the oscillatory solution
        const double q = sqrt(-me);
        double sn, cs;
        sincos(q * L, &sn, &cs);
        yb = fma(M.ic_v, cs, (M.ic_s / q) * sn);
        ypb = fma(-M.ic_v * q, sn, M.ic_s * cs);
        return;
    }
    const double kap = sqrt(me);
    if (kap * L < 1e-8) {
        yb = fma(M.ic_s, L, M.ic_v) + 0.5 * me * L * L * M.ic_v;
        ypb = M.ic_s + me * L * M.ic_v;
    } else {
        const double E = exp(kap * L), Ei = 1.0 / E;
        const double ch = 0.5 * (E + Ei), sh = 0.5 * (E - Ei);
        yb = fma(M.ic_v, ch, (M.ic_s / kap) * sh);
        ypb = fma(M.ic_v * kap, sh, M.ic_s * ch);
    }
}

// Cylinder: P'' + P'/r - (m_e + n^2/r^2) P = 0   (Density_cylinder.py:765), r from -x0 to -1.
// In rho = |r|: P = A I_n(kap rho) + B K_n(kap rho), d/dr = -d/drho.  The Bessel sets at the
// two arguments are computed once (ExtCyl) and serve every azimuthal order.
struct ExtCyl {
    BesselIK B0, B1;
    double kap, z0, ea, eb;
};

ESB_HD void exterior_cyl_prepare(const DevModel& M, double k, double me, int nmax, ExtCyl& E) {
    E.kap = sqrt(me);
    E.z0 = E.kap * (M.ext_len / k);
    bessel_ik_scaled(nmax, E.z0, E.B0);
    bessel_ik_scaled(nmax, E.kap, E.B1);
    E.ea = exp(E.kap - E.z0);
    E.eb = exp(E.z0 - E.kap);
}

ESB_HD void exterior_cyl_order(const DevModel& M, const ExtCyl& E, int n, double& yb, double& ypb) {
    double I0, dI0, K0, dK0, I1, dI1, K1, dK1;
    bessel_order(E.B0, n, E.z0, I0, dI0, K0, dK0);
    bessel_order(E.B1, n, E.kap, I1, dI1, K1, dK1);
    // d/dz at rho0 (z = kap rho); the scripts give dP/dr, and dr = r_sign d(rho)
    const double P0 = M.ic_v, dP0 = M.r_sign * M.ic_s / E.kap;
    // Wronskian I K' - I' K = -1/z.  K0 = e^{z0} K(z0), I0 = e^{-z0} I(z0), hence the e^{+-(z0-z1)}.
    const double As = -E.z0 * (P0 * dK0 - dP0 * K0);
    const double Bs = -E.z0 * (dP0 * I0 - P0 * dI0);
    yb = As * I1 * E.ea + Bs * K1 * E.eb;
    ypb = M.r_sign * E.kap * (As * dI1 * E.ea + Bs * dK1 * E.eb);
}

// ------------------------------------------------- warp-cooperative layer ----
// One warp evaluates ONE point: the ODE is linear, so lane j integrates the two fundamental solutions
// over its own sub-interval of the mesh (steps [j per, (j+1) per)) and the 32 transfer matrices are
// multiplied in order by a shuffle tree.  Twice the arithmetic of the one-solution cylinder scheme,
// 1/32 of its latency: used by the refinement when there are fewer brackets than lanes to fill
// (esb.cu refine_warp_kernel), where the sequential Brent iterations are latency bound.
// T = {T00, T01, T10, T11}: (u, v)_end = T (u, v)_start.
#ifdef __CUDA_ARCH__
template <class F>
__device__ __forceinline__ void warp_transfer(int first, int n_steps, F&& integrate_range, double (&T)[4]) {
    // steps [first, n_steps) shared out over the 32 lanes
    const int lane = threadIdx.x & 31;
    const int per = (n_steps - first + 31) >> 5;
    const int i0 = first + lane * per;
    const int i1 = i0 + per < n_steps ? i0 + per : n_steps;
    double u[2] = {1.0, 0.0}, v[2] = {0.0, 1.0};             // images of (1,0) and (0,1)
    if (i0 < n_steps) integrate_range(i0, i1, u, v);
    T[0] = u[0]; T[1] = u[1]; T[2] = v[0]; T[3] = v[1];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        double B[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) B[q] = __shfl_down_sync(0xffffffffu, T[q], off);
        if ((lane & (2 * off - 1)) == 0) {                    // B (later sub-interval) after T (earlier)
            const double t0 = fma(B[0], T[0], B[1] * T[2]), t1 = fma(B[0], T[1], B[1] * T[3]);
            const double t2 = fma(B[2], T[0], B[3] * T[2]), t3 = fma(B[2], T[1], B[3] * T[3]);
            T[0] = t0; T[1] = t1; T[2] = t2; T[3] = t3;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) T[q] = __shfl_sync(0xffffffffu, T[q], 0);
}
#endif

// ------------------------------------------------------------- shoot_layer ----
// The normal form has a DOUBLE pole where the first-derivative form has a simple one (q contains
// a^2/4): a point whose resonance lies just outside the layer - the analytic continuation of the profile
// reaches the resonant value a little beyond an end of the layer, or the profile comes close to it
// without reaching it - is integrated less accurately in u than in y on the same mesh (measured: 1e-5
// instead of 1e-10 at 0.02 in phase speed above the Alfven continuum, equal beyond ~0.1; the band of 8 % in
// the resonant quantity is ~0.13 in phase speed there, and a narrower one - 4 % - starts to show).  Such points
// (resonant value within NF_BAND, relative, of the range the profile spans in the layer, but outside
// it) take the (y, h y') variables on the same table.  Inside the range (a resonance in the layer: the
// continua, below the noise floor) any form returns noise; they stay on the cheaper one.
#ifndef ESB_NF_BAND
#define ESB_NF_BAND 0.08
#endif
constexpr double NF_BAND = ESB_NF_BAND;

ESB_HD bool in_band(double v, double lo, double hi) {
    return (v > hi && v < hi * (1.0 + NF_BAND)) || (v < lo && v > lo * (1.0 - NF_BAND));
}

template <int KIND>
ESB_HD bool near_resonance(const DevModel& M, const Point& pt, const NPoint& sp) {
    if constexpr (KIND == KIND_CYL_DENSITY) {
        return in_band(sp.q, M.f_lo, M.f_hi);                            // Alfven: rho = k^2 beta/w^2
    } else if constexpr (KIND == KIND_SLAB_DENSITY) {
        return in_band(sp.t, M.f_lo, M.f_hi) || in_band(sp.p, M.f_lo, M.f_hi);   // sound point of F, cusp
    } else {
        // Alfven: (w - k v_z)^2 = k^2 vA^2 with v_z in [f_lo, f_hi]
        const double Oa = fma(-pt.k, M.f_lo, pt.w), Ob = fma(-pt.k, M.f_hi, pt.w);
        const double a2 = Oa * Oa, b2 = Ob * Ob;
        const double hi = fmax(a2, b2), lo = (Oa * Ob <= 0.0) ? 0.0 : fmin(a2, b2);
        return in_band(pt.K * M.vAi2, lo, hi);
    }
}

constexpr int SCHEME_RK8_NTAB = 3;      // internal: the (y, h y') variables read from the normal-form table

// steps [i0, i1) of the mesh by the scheme; FULL: the whole mesh (compile-time bounds)
template <int KIND, int SCHEME, int NS, bool FULL>
ESB_HD void layer_range(const DevModel& M, const Point& pt, const NPoint& sp, const double* __restrict__ tab,
                        const double (&m2)[NS], double (&y)[NS], double (&yp)[NS], int i0, int i1) {
    if constexpr (SCHEME == SCHEME_RK8N) integrate_layer_nform<KIND, NS, !FULL>(M, pt, sp, tab, m2, y, yp, i0, i1);
    else if constexpr (SCHEME == SCHEME_RK8_NTAB)
        integrate_layer_prescaled<KIND, NS, !FULL, true>(M, pt, tab, m2, y, yp, i0, i1);
    else if constexpr (SCHEME == SCHEME_RK8 && is_cyl_second_order<KIND>)
        integrate_layer_prescaled<KIND, NS, !FULL>(M, pt, tab, m2, y, yp, i0, i1);
    else integrate_layer<KIND, SCHEME, NS, !FULL>(M, pt, tab, m2, y, yp, i0, i1);
}

// start vectors -> end of the mesh in the variables of SCHEME.  WARP (all 32 lanes together, same
// arguments): the transfer matrix of the layer is built cooperatively (warp_transfer) and applied to the
// start vectors; all solutions then share m2[0].
template <int KIND, int SCHEME, int NS, bool WARP, bool FULL>
ESB_HD void shoot_vars(const DevModel& M, const Point& pt, const NPoint& sp, const double* __restrict__ tab,
                       const double (&m2)[NS], double (&y)[NS], double (&yp)[NS], int first) {
#ifdef __CUDA_ARCH__
    if constexpr (WARP) {
        const double mm2[2] = {m2[0], m2[0]};
        double T[4];
        warp_transfer(first, M.n_steps, [&](int i0, int i1, double (&u)[2], double (&v)[2]) {
            layer_range<KIND, SCHEME, 2, false>(M, pt, sp, tab, mm2, u, v, i0, i1);
        }, T);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const double y0 = y[s], yp0 = yp[s];
            y[s] = fma(T[0], y0, T[1] * yp0);
            yp[s] = fma(T[2], y0, T[3] * yp0);
        }
    } else
#endif
    layer_range<KIND, SCHEME, NS, FULL>(M, pt, sp, tab, m2, y, yp, first, M.n_steps);
}

// NS solutions given as (y, y') at breakpoint `first` of the mesh -> (y, y') at its end.  m2[s] = the
// square of the azimuthal order of solution s (cylinder).  SCHEME_RK8N integrates u = sqrt|F| y and
// converts at both ends (points next to a resonance: see near_resonance).  FULL: first == 0 known at
// compile time.
// YFORM: every point in the (y, h y') variables.  The lane-per-bracket refinement uses it: its 32
// lanes hold unrelated (k, omega), so the per-point choice would make most warps run BOTH step loops.
template <int KIND, int SCHEME, int NS, bool WARP, bool FULL, bool YFORM = false>
ESB_HD void shoot_layer(const DevModel& M, const Point& pt, const double* __restrict__ tab, const double (&m2)[NS],
                        double (&y)[NS], double (&yp)[NS], int first) {
    if constexpr (SCHEME == SCHEME_RK8N && YFORM) {
        const NPoint sp{};
        shoot_vars<KIND, SCHEME_RK8_NTAB, NS, WARP, FULL>(M, pt, sp, tab, m2, y, yp, first);
    } else if constexpr (SCHEME == SCHEME_RK8N) {
        const NPoint sp = make_npoint<KIND>(M, pt);
        if (near_resonance<KIND>(M, pt, sp)) {
            shoot_vars<KIND, SCHEME_RK8_NTAB, NS, WARP, FULL>(M, pt, sp, tab, m2, y, yp, first);
            return;
        }
        double a0, X0, a_end, XN, mq[NS];
        nform_end_values<KIND>(M, pt, sp, tab, first, a0, X0);
        nform_end_values<KIND>(M, pt, sp, tab, M.n_steps, a_end, XN);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            yp[s] = fma(-0.5 * a0, y[s], yp[s]);          // u' = y' - a y/2 (the common factor sqrt|F| is restored in sf)
            mq[s] = m2[s] - 0.25;
        }
        shoot_vars<KIND, SCHEME_RK8N, NS, WARP, FULL>(M, pt, sp, tab, mq, y, yp, first);
        const double ratio = is_cyl_second_order<KIND> ? (M.r_axis / M.s_start) * (XN / X0) : X0 / XN;
        double sf = sqrt(fabs(ratio));                    // sqrt|F(first)/F(end)|
        if (!(sf <= 1.7e308)) sf = 1.0;                   // a resonance exactly on an end node (inside a continuum)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            yp[s] = fma(0.5 * a_end, y[s], yp[s]) * sf;   // y' = (u' + a u/2)/sqrt|F|
            y[s] *= sf;
        }
    } else {
        const NPoint sp{};
        shoot_vars<KIND, SCHEME, NS, WARP, FULL>(M, pt, sp, tab, m2, y, yp, first);
    }
}

// ------------------------------------------------------------ full point ----
// NM evaluations of the reference's scan-loop body at one (k, omega), one per requested
// mode, sharing everything that does not depend on the mode:
//   cylinder: modes = azimuthal orders; one staged-coefficient evaluation per node and one
//             Bessel set serve all of them, each order integrates its own solution;
//   slab:     sausage and kink are two closures of the same two fundamental solutions.
// den_q[s] = the denominator of int_q[s]: the interior quantity is a ratio whose denominator (the
// boundary value of the integrated solution) passes through zero at the poles of D.  G = D * den_q
// has the roots of D and no such poles; the refinement iterates on G (esb.cu refine_kernel).
// WARP (device only, NM = 1, all 32 lanes of the warp call it with the same arguments): the layer is
// integrated cooperatively (warp_transfer); every lane returns the same values.
// LEAKY (opt-in, esb_dispersion_grid_leaky): points with m_e < 0 are NOT skipped - their exterior is the
// oscillatory solution of the same initial-value problem (J_n, Y_n / cos, sin: bessel_jy.cuh), what the
// reference's odeint would return without its "if m_e < 0: pass".  The default (false) is the reference's rule.
template <int KIND, int SCHEME, int NM, bool WARP = false, bool YFORM = false, bool LEAKY = false>
ESB_HD void eval_point_multi(const DevModel& M, const double* __restrict__ tab, double k, double w,
                             const int (&modes)[NM], double (&ext_q)[NM], double (&int_q)[NM],
                             double (&den_q)[NM]) {
    static_assert(!WARP || NM == 1, "the warp-cooperative evaluation handles one mode");
    const double nanv = nan("");
    const Point pt = make_point(M, k, w);
    // exterior Doppler shift (flow script :207): (w - k U_e)
    const double We = (KIND == KIND_SLAB_FLOW) ? fma(-k, M.U_e, w) : w;
    const double Ae = (KIND == KIND_SLAB_FLOW) ? We * We : pt.A;
    const double me = m_e2(M, pt.K, Ae);
    const bool leaky = LEAKY && me < 0.0;
    if (!(me >= 0.0) && !leaky) {             // "if m_e < 0: pass"  (Density_cylinder.py:760)
#pragma unroll
        for (int s = 0; s < NM; ++s) { ext_q[s] = nanv; int_q[s] = nanv; den_q[s] = nanv; }
        return;
    }
    if constexpr (KIND == KIND_CYL_ROTATION) {
        int nmax = 0;
#pragma unroll
        for (int s = 0; s < NM; ++s) nmax = modes[s] > nmax ? modes[s] : nmax;
        ExtCyl E;
        if (!leaky) exterior_cyl_prepare(M, k, me, nmax, E);
        const double xi_e_const = -1.0 / (M.rho_e * (pt.K * M.vAe2 - pt.A));        // :263
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            double Pb, ypb;
            if (LEAKY && leaky) exterior_cyl_leaky(M.ic_v, M.ic_s, M.r_sign, M.ext_len, k, me, modes[s], Pb, ypb);
            else exterior_cyl_order(M, E, modes[s], Pb, ypb);
            const double xi_e = xi_e_const * ypb;
            // The layer is integrated from the axis end (where the scripts impose their end condition)
            // out to the boundary r = s_start, where P = P_e fixes the scale.
            const double mm = double(modes[s]);
            double xi_b;
            if (modes[s] == 0) {
                // sausage: P'(axis end) = 0  <=>  C3 xi - C1 P = 0 there (sausage script :306): ONE
                // solution, started with (P, xi) proportional to (C3, C1)
                const RotCoef ca = node_rot(M, pt, mm, tab);
                const double nrm = 1.0 / fmax(fabs(ca.C3), fabs(ca.C1));
                double P[1] = {ca.C3 * nrm}, X[1] = {ca.C1 * nrm};
#ifdef __CUDA_ARCH__
                if constexpr (WARP) {
                    double T[4];
                    warp_transfer(0, M.n_steps, [&](int i0, int i1, double (&u)[2], double (&v)[2]) {
                        integrate_rotation<2, true>(M, pt, mm, tab, u, v, i0, i1);
                    }, T);
                    const double p0 = P[0], x0 = X[0];
                    P[0] = fma(T[0], p0, T[1] * x0);
                    X[0] = fma(T[2], p0, T[3] * x0);
                } else
#endif
                integrate_rotation<1>(M, pt, mm, tab, P, X);
                den_q[s] = P[0];
                xi_b = Pb * X[0] / P[0];
            } else {
                // kink (:308): P(axis end) = c = -(B_phi(1)^2 - rho(1) v_phi(1)^2) xi_e(1), xi(axis end) free.
                // ONE solution suffices here too: the system matrix has trace m11 + m22 = -1/r, so the
                // Wronskian of any two solutions obeys r W(r) = const (Abel).  With y = (P, xi) the wanted
                // solution and (P2, X2) the one started from (0, 1) at the axis end,
                //     P X2 - P2 xi = c r_axis / r      =>   xi(1) = (P_e X2(1) - c r_axis/s_start) / P2(1)
                // (the same value the superposition c (P1, X1) + alpha (P2, X2) with P(1) = P_e gives, without
                // integrating (P1, X1)).
                double P[1] = {0.0}, X[1] = {1.0};
#ifdef __CUDA_ARCH__
                if constexpr (WARP) {
                    double T[4];
                    warp_transfer(0, M.n_steps, [&](int i0, int i1, double (&u)[2], double (&v)[2]) {
                        integrate_rotation<2, true>(M, pt, mm, tab, u, v, i0, i1);
                    }, T);
                    P[0] = T[1]; X[0] = T[3];
                } else
#endif
                integrate_rotation<1>(M, pt, mm, tab, P, X);
                const double c = M.rho_vb2 * xi_e;
                den_q[s] = P[0];
                xi_b = fma(Pb, X[0], -c * (M.r_axis / M.s_start)) / P[0];
            }
            ext_q[s] = xi_e;
            int_q[s] = xi_b;      // xi_i(1) = (C1 P + D P')/C3 at r = 1   (:314)
        }
    } else if constexpr (is_cyl_second_order<KIND>) {
        int nmax = 0;
#pragma unroll
        for (int s = 0; s < NM; ++s) nmax = modes[s] > nmax ? modes[s] : nmax;
        ExtCyl E;
        if (!leaky) exterior_cyl_prepare(M, k, me, nmax, E);
        double yb[NM], y[NM], yp[NM], m2[NM];
        // xi_e = -P'/(rho_e (k^2 vA_e^2 - w^2))      (Density_cylinder.py:702,773)
        const double xi_e_const = -1.0 / (M.rho_e * (pt.K * M.vAe2 - pt.A));
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            double ypb;
            if (LEAKY && leaky) exterior_cyl_leaky(M.ic_v, M.ic_s, M.r_sign, M.ext_len, k, me, modes[s], yb[s], ypb);
            else exterior_cyl_order(M, E, modes[s], yb[s], ypb);
            ext_q[s] = xi_e_const * ypb;
            // interior from the axis outwards: sausage P'(axis)=0 (:1084), kink/fluting P(axis)=0 (:787)
            y[s] = modes[s] == 0 ? 1.0 : 0.0;
            yp[s] = modes[s] == 0 ? 0.0 : 1.0;
            m2[s] = double(modes[s]) * double(modes[s]);
        }
        shoot_layer<KIND, SCHEME, NM, WARP, true, YFORM>(M, pt, tab, m2, y, yp, 0);
        double den;
        if constexpr (KIND == KIND_CYL_FLOW) {
            // (C1 P + D P')/C3 at r = -1 with C1 = 0: P'/(rho (Om_b^2 - k^2 vA^2)), Om_b = w - k v_z(-1)
            const double Ob = fma(-k, M.U_b, w);
            den = 1.0 / (M.rho_i * fma(-pt.K, M.vAi2, Ob * Ob));
        } else {
            den = 1.0 / (M.rho_b * pt.A - pt.Kbeta);
        }
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            const double slope = yb[s] * yp[s] / y[s];       // dPi that fsolve finds (:790)
            // xi_i(-1) = (C1 P + D P')/C3 = P'/(rho (w^2 - k^2 vA^2))    (:798)
            int_q[s] = slope * den;
            den_q[s] = y[s];
        }
    } else {
        double yb, ypb;
        exterior_slab(M, k, me, yb, ypb);
        // P_e = p_e_const vx'   (..._coronal.py:221,250 / flow :209,291)
        const double p_e_const = M.rho_e * M.se2 * (pt.K * M.cTe2 - Ae) / (We * (pt.K * M.ce2 - Ae));
        // General layer: the two fundamental solutions at the near boundary, integrated across.
        // Mirror-symmetric layer (every shipped script: x0 = 0): the even and the odd solution,
        // (1, 0) and (0, 1) at the mid-plane, integrated over the far half only - the sausage condition
        // vx(1) = -vx(-1) selects the odd one, the kink condition the even one.  Half the steps.
        double y[2] = {1.0, 0.0}, yp[2] = {0.0, 1.0};
        const double m2[2] = {0.0, 0.0};
        const int first = M.symmetric ? M.n_steps / 2 : 0;
        if (!WARP && NM == 1 && M.symmetric) {
            // one mode of a symmetric layer needs only its own solution (odd: sausage, even: kink)
            const int c = (modes[0] == 0) ? 1 : 0;
            double y1[1] = {c == 0 ? 1.0 : 0.0}, yp1[1] = {c == 0 ? 0.0 : 1.0};
            const double m21[1] = {0.0};
            shoot_layer<KIND, SCHEME, 1, false, false, YFORM>(M, pt, tab, m21, y1, yp1, first);
            y[c] = y1[0]; yp[c] = yp1[0];
        } else {
            shoot_layer<KIND, SCHEME, 2, WARP, false, YFORM>(M, pt, tab, m2, y, yp, first);
        }
        double P_Ti;
        if (KIND == KIND_SLAB_FLOW) {
            // displacement continuity: vx_i(-1) = vx_e(-1) (w - k U(-1))/(w - k U_e)   (flow :290)
            const double Ob = fma(-k, M.U_b, w);
            yb *= Ob / We;
            // P_Ti = rho_i (vA^2+c^2)(k^2 cT^2 - Ob^2)/(Ob (k^2 c^2 - Ob^2))   (flow :223)
            P_Ti = M.rho_i * M.si * (pt.K * M.cTi2 - Ob * Ob) / (Ob * (pt.K * M.ci2 - Ob * Ob));
        } else {
            // P_i(-1) = P_Ti(-1) vx'(-1)   (:234,267)
            const double ub = M.rho_b * pt.A;
            P_Ti = M.S * (pt.Ktau - ub) / (w * (pt.Kalpha - ub));
        }
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            // sausage: vx(1) = -vx(-1) (:259); kink: vx(1) = +vx(-1) (:696)
            double slope;
            if (M.symmetric) {
                // vx = yb Sol(x)/Sol(-1), Sol = odd (sausage) or even (kink): Sol(-1) = -+Sol(1),
                // Sol'(-1) = +-Sol'(1)  ->  vx'(-1) = -yb Sol'(1)/Sol(1)
                const int c = (modes[s] == 0) ? 1 : 0;
                slope = -yb * yp[c] / y[c];
                den_q[s] = y[c];
            } else {
                const double target = (modes[s] == 0) ? -1.0 : 1.0;
                slope = yb * (target - y[0]) / y[1];
                den_q[s] = y[1];
            }
            ext_q[s] = p_e_const * ypb;
            int_q[s] = P_Ti * slope;
        }
    }
}

// --------------------------------------------------------- noise-floor test ----
// true where no resonance of the ODE sits inside the layer (with a relative margin on the resonant
// quantity), i.e. where the integration converges as the mesh is refined: the points the
// discretisation guard (esb.cu guard_kernel) may judge.  Inside the continua (a singular point in the
// layer) ANY step count returns solver noise, as the reference's odeint does.
//   density kinds   rho(x) w^2 = k^2 {beta, tau} (Alfven, cusp; slab also alpha: the sound point of F)
//   flow kinds      (w - k U)^2 = k^2 {vA^2 (cylinder), cT^2, c^2 (slab)}, slab also w = k U
//   rotation        D, C3 or the Doppler-shifted frequency w - m v_phi/r changes sign between staged nodes
//                   (C3 = 0 is singular for the reference's second-order form only, but the reference output
//                   there is noise all the same)
ESB_HD bool outside_range(double v, double lo, double hi, double margin) {
    // lo, hi >= 0 (densities, squared frequencies): a relative margin on either bound
    return v < lo * (1.0 - margin) || v > hi * (1.0 + margin);
}

// What the discretisation guard judges: g = D Y / (|ext Y| + |int Y|), the acceptance test's relative mismatch
// made pole-free (int = N / Y: g = (ext Y - N) / (|ext Y| + |N|) stays regular where Y -> 0) AND projective -
// unchanged when N and Y carry a common factor.  They do: both are built from the solution that dominates
// towards the axis (r^-n), whose amplitude error cancels in int = N / Y (measured, fluting n = 3 at 152 steps:
// N and Y each 2.5e-9 off, D 4e-15); G = D Y itself would report that harmless factor as an error of the sweep.
ESB_HD double guard_deviation(double e0, double i0, double d0, double e, double i, double d) {
    const double g0 = (e0 - i0) * d0 / (fabs(e0 * d0) + fabs(i0 * d0));
    const double g = (e - i) * d / (fabs(e * d) + fabs(i * d));
    return fabs(g0 - g);
}

template <int KIND>
ESB_HD bool resonance_free(const DevModel& M, const Point& pt, double mode, const double* __restrict__ tab,
                           double margin) {
    if constexpr (KIND == KIND_CYL_DENSITY || KIND == KIND_SLAB_DENSITY) {
        const double A = pt.A > 1e-280 ? pt.A : 1e-280;
        bool ok = outside_range(pt.Kbeta / A, M.f_lo, M.f_hi, margin) &&
                  outside_range(pt.Ktau / A, M.f_lo, M.f_hi, margin);
        if (KIND == KIND_SLAB_DENSITY) ok = ok && outside_range(pt.Kalpha / A, M.f_lo, M.f_hi, margin);
        return ok;
    } else if constexpr (KIND == KIND_CYL_FLOW || KIND == KIND_SLAB_FLOW) {
        const double Oa = fma(-pt.k, M.f_lo, pt.w), Ob = fma(-pt.k, M.f_hi, pt.w);
        const double a2 = Oa * Oa, b2 = Ob * Ob;
        const bool through_zero = Oa * Ob <= 0.0;
        const double hi = fmax(a2, b2), lo = through_zero ? 0.0 : fmin(a2, b2);
        bool ok = outside_range(pt.K * M.cTi2, lo, hi, margin);
        if (KIND == KIND_CYL_FLOW) ok = ok && outside_range(pt.K * M.vAi2, lo, hi, margin);
        else ok = ok && outside_range(pt.K * M.ci2, lo, hi, margin) && !through_zero &&
                  fmin(a2, b2) > margin * margin * pt.A;
        return ok;
    } else {
        // D = 0 and C3 = 0 only at a resonance, and the Doppler-shifted frequency Om = w - m v_phi/r must not
        // pass through zero inside the layer either (measured: the deviation between N and 2N steps does not
        // fall with N there): the signs of D, C3 and Om must be the same at every node AND at w (1 - margin),
        // w, w (1 + margin) - a point NEXT to a resonance converges slowly too
        bool pos_d = false, neg_d = false, pos_c = false, neg_c = false, pos_o = false, neg_o = false;
        for (int j = -1; j <= 1; ++j) {
            Point q = pt;
            q.w = pt.w * (1.0 + j * margin);
            q.A = q.w * q.w;
            for (int i = 0; i < M.n_nodes; ++i) {
                const double* f = tab + (size_t)i * ROT_FIELDS;
                const RotCoef c = node_rot(M, q, mode, f);
                const double Om = fma(-mode, f[2], q.w);
                pos_d = pos_d || c.invD > 0.0; neg_d = neg_d || !(c.invD > 0.0);
                pos_c = pos_c || c.C3 > 0.0; neg_c = neg_c || !(c.C3 > 0.0);
                pos_o = pos_o || Om > 0.0; neg_o = neg_o || !(Om > 0.0);
            }
        }
        return !(pos_d && neg_d) && !(pos_c && neg_c) && !(pos_o && neg_o);
    }
}

template <int KIND, int SCHEME, bool WARP = false, bool YFORM = false, bool LEAKY = false>
ESB_HD void eval_point(const DevModel& M, const double* __restrict__ tab, double k, double w, int mode,
                       double& ext_q, double& int_q, double& den_q) {
    const int modes[1] = {mode};
    double e[1], i[1], d[1];
    eval_point_multi<KIND, SCHEME, 1, WARP, YFORM, LEAKY>(M, tab, k, w, modes, e, i, d);
    ext_q = e[0];
    int_q = i[0];
    den_q = d[0];
}

}  // namespace esb
