/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's scan-loop body, discretised with a
 * fixed-step Runge-Kutta integrator so that it is fast enough to check whole
 * grids.  It follows the REFERENCE's structure, not the GPU kernel's:
 *
 *   - the exterior ODE is integrated numerically from x = -3*2*pi/k to x = -1 from
 *     the reference's initial values (Density_cylinder.py:765-770, ..._coronal.py:245-248);
 *     no Bessel functions, no closed forms;
 *   - the interior ODE is integrated FORWARD from the boundary with two fundamental
 *     solutions and the slope is chosen to satisfy the reference's end condition
 *     (what fsolve does on a linear problem: Density_cylinder.py:785-790);
 *   - the coefficients are written from c_i^2(x), vA_i^2(x), cT_i^2(x) and their
 *     derivatives exactly as the reference composes them (F, dF/F, m0 / g/F), with the
 *     profile evaluated analytically at every stage - no tables; where the reference differentiates
 *     F and r C1/C3 with sympy (rotational and axial-flow cylinder) the derivative is taken by
 *     forward-mode automatic differentiation of the same expression tree.
 *
 * The CUDA path uses closed-form exteriors, a backward one-solution integration
 * for the cylinder, algebraically reduced coefficients and a staged table; that the two
 * agree to ~1e-11 is the parity statement.  This file itself is pinned against
 * oracle/reference_path.py (scipy odeint/fsolve, in turn pinned against the executed
 * reference) in tests/test_oracle_pinned.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the
 * shared object built from this file (oracle/_build/liboracle_rk.so).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int kind;            /* 0 slab density, 1 cylinder density, 2 slab sheared flow, 3 cylinder rotation,
                            4 cylinder axial flow v_z(r) = U_e + (U_i0-U_e) exp(-(r-x0)^2/width^2) */
    int n_ext;           /* exterior steps */
    int n_int;           /* interior steps */
    int leaky;           /* 1: do not skip m_e < 0 (test-only extension; 0 = the reference's rule) */
    double c_i0, vA_i0, vA_e, c_e, gamma, rho_i0, rho_A;
    double width, x0;    /* inverted-Gaussian density profile */
    double ic_v, ic_s;   /* exterior initial values */
    double ext_wavelengths;
    double s_start, s_end;
    double U_i0, U_e;    /* kind 2: flow profile U_e + (U_i0-U_e) exp(-(x-x0)^2/width^2); c_i0, vA_i0,
                            rho_i0 are then the uniform interior values */
    double r_sign;       /* cylinder: -1 scripts in r<0 (coronal), +1 scripts in r>0 (photospheric) */
    double v_twist, power; /* kind 3: v_phi = v_twist r^power (uniform rho_i0, vA_i0; r > 0) */
    int profile_kind;      /* density kinds: 0 inverted Gaussian, 1 Epstein (Density_cylinder.py:139-142):
                              (rho_i0 - rho_e)/cosh((x-x0)/width)^8 + rho_e */
    int pad2;
} ork_model;

/* ---- Cooper-Verner 8th order tableau, table driven -------------------- */
#define NSTG 11
static double CA[NSTG][NSTG], CB[NSTG], CC[NSTG];
static int tableau_ready = 0;
static void tableau(void) {
    if (tableau_ready) return;
    const double s = sqrt(21.0);
    memset(CA, 0, sizeof(CA));
    memset(CB, 0, sizeof(CB));
    const double c[NSTG] = {0, .5, .5, (7 + s) / 14, (7 + s) / 14, .5, (7 - s) / 14, (7 - s) / 14, .5,
                            (7 + s) / 14, 1};
    memcpy(CC, c, sizeof(c));
    CA[1][0] = .5;
    CA[2][0] = .25; CA[2][1] = .25;
    CA[3][0] = 1. / 7; CA[3][1] = (-7 - 3 * s) / 98; CA[3][2] = (21 + 5 * s) / 49;
    CA[4][0] = (11 + s) / 84; CA[4][2] = (18 + 4 * s) / 63; CA[4][3] = (21 - s) / 252;
    CA[5][0] = (5 + s) / 48; CA[5][2] = (9 + s) / 36; CA[5][3] = (-231 + 14 * s) / 360; CA[5][4] = (63 - 7 * s) / 80;
    CA[6][0] = (10 - s) / 42; CA[6][2] = (-432 + 92 * s) / 315; CA[6][3] = (633 - 145 * s) / 90;
    CA[6][4] = (-504 + 115 * s) / 70; CA[6][5] = (63 - 13 * s) / 35;
    CA[7][0] = 1. / 14; CA[7][4] = (14 - 3 * s) / 126; CA[7][5] = (13 - 3 * s) / 63; CA[7][6] = 1. / 9;
    CA[8][0] = 1. / 32; CA[8][4] = (91 - 21 * s) / 576; CA[8][5] = 11. / 72; CA[8][6] = (-385 - 75 * s) / 1152;
    CA[8][7] = (63 + 13 * s) / 128;
    CA[9][0] = 1. / 14; CA[9][4] = 1. / 9; CA[9][5] = (-733 - 147 * s) / 2205; CA[9][6] = (515 + 111 * s) / 504;
    CA[9][7] = (-51 - 11 * s) / 56; CA[9][8] = (132 + 28 * s) / 245;
    CA[10][4] = (-42 + 7 * s) / 18; CA[10][5] = (-18 + 28 * s) / 45; CA[10][6] = (-273 - 53 * s) / 72;
    CA[10][7] = (301 + 53 * s) / 72; CA[10][8] = (28 - 28 * s) / 45; CA[10][9] = (49 - 7 * s) / 18;
    CB[0] = 1. / 20; CB[7] = 49. / 180; CB[8] = 16. / 45; CB[9] = 49. / 180; CB[10] = 1. / 20;
    tableau_ready = 1;
}

/* y'' = a(x) y' + b(x) y  for NV independent solutions, one RK8 step */
typedef void (*coef_fn)(const void* ctx, double x, double* a, double* b);

static void rk8_step(coef_fn f, const void* ctx, double x, double h, int nv, double* y, double* yp) {
    double KQ[NSTG][4], KG[NSTG][4];
    for (int i = 0; i < NSTG; ++i) {
        double a, b;
        f(ctx, x + CC[i] * h, &a, &b);
        for (int v = 0; v < nv; ++v) {
            double sp = 0, sq = 0;
            for (int j = 0; j < i; ++j) {
                sp += CA[i][j] * KQ[j][v];
                sq += CA[i][j] * KG[j][v];
            }
            const double P = y[v] + h * sp, Q = yp[v] + h * sq;
            KQ[i][v] = Q;
            KG[i][v] = a * Q + b * P;
        }
    }
    for (int v = 0; v < nv; ++v) {
        double sp = 0, sq = 0;
        for (int i = 0; i < NSTG; ++i) {
            sp += CB[i] * KQ[i][v];
            sq += CB[i] * KG[i][v];
        }
        y[v] += h * sp;
        yp[v] += h * sq;
    }
}

/* ---- equilibrium (Density_cylinder.py:69-221, ..._coronal.py:69-124) ---- */
typedef struct {
    const ork_model* m;
    double k, w, K, A;
    double rho_e, cT_e2, m_e;
    int mode;
} pt_ctx;

static double rho_e_of(const ork_model* m) {
    return m->rho_i0 * (m->c_i0 * m->c_i0 + m->gamma * 0.5 * m->vA_i0 * m->vA_i0) /
           (m->c_e * m->c_e + m->gamma * 0.5 * m->vA_e * m->vA_e);
}

static void profile(const pt_ctx* p, double x, double* rho, double* drho, double* c2, double* dc2,
                    double* vA2, double* dvA2) {
    const ork_model* m = p->m;
    double prof, dprof;
    if (m->profile_kind == 1) {
        const double t = (x - m->x0) / m->width, ch = cosh(t);
        const double ch8 = pow(ch, 8.0);
        prof = (m->rho_i0 - p->rho_e) / ch8 + p->rho_e;
        dprof = (m->rho_i0 - p->rho_e) * (-8.0 / m->width) * sinh(t) / (ch8 * ch);
    } else {
        const double g = exp(-(x - m->x0) * (x - m->x0) / (m->width * m->width));
        prof = p->rho_e + (m->rho_i0 - p->rho_e) * g;
        dprof = (m->rho_i0 - p->rho_e) * g * (-2.0 * (x - m->x0) / (m->width * m->width));
    }
    *rho = m->rho_A * prof;
    *drho = m->rho_A * dprof;
    if (m->kind == 1) { /* B_i = B_0: vA^2 = B_0^2/rho */
        *vA2 = m->vA_i0 * m->vA_i0 * m->rho_i0 / *rho;
        *dvA2 = -*vA2 * *drho / *rho;
    } else {            /* vA_i = vA_i0 sqrt(rho_i0)/sqrt(profile) */
        *vA2 = m->vA_i0 * m->vA_i0 * m->rho_i0 / prof;
        *dvA2 = -*vA2 * dprof / prof;
    }
    const double Cc = p->rho_e * (m->c_e * m->c_e + 0.5 * m->gamma * m->vA_e * m->vA_e);
    *c2 = Cc / *rho - 0.5 * m->gamma * *vA2;
    *dc2 = -Cc * *drho / (*rho * *rho) - 0.5 * m->gamma * *dvA2;
}

static void coef_ext_slab(const void* c, double x, double* a, double* b) {
    (void)x;
    *a = 0.0;
    *b = ((const pt_ctx*)c)->m_e;                       /* dVx_dx_e */
}
static void coef_ext_cyl(const void* c, double r, double* a, double* b) {
    const pt_ctx* p = (const pt_ctx*)c;
    *a = -1.0 / r;
    *b = p->m_e + (double)(p->mode * p->mode) / (r * r); /* dP_dr_e */
}
static void coef_int_slab(const void* c, double x, double* a, double* b) {
    const pt_ctx* p = (const pt_ctx*)c;
    double rho, drho, c2, dc2, vA2, dvA2;
    profile(p, x, &rho, &drho, &c2, &dc2, &vA2, &dvA2);
    const double s = c2 + vA2, ds = dc2 + dvA2;
    const double cT2 = c2 * vA2 / s;
    const double dcT2 = (dc2 * vA2 + c2 * dvA2) / s - cT2 * ds / s;
    const double dlnF = drho / rho + ds / s + p->K * dcT2 / (p->K * cT2 - p->A) - p->K * dc2 / (p->K * c2 - p->A);
    *a = -dlnF;
    *b = (p->K * c2 - p->A) * (p->K * vA2 - p->A) / (s * (p->K * cT2 - p->A));
}
static void coef_int_cyl(const void* c, double r, double* a, double* b) {
    const pt_ctx* p = (const pt_ctx*)c;
    double rho, drho, c2, dc2, vA2, dvA2;
    profile(p, r, &rho, &drho, &c2, &dc2, &vA2, &dvA2);
    const double s = c2 + vA2, cT2 = c2 * vA2 / s;
    const double X = rho * (p->A - p->K * vA2);
    const double dX = drho * (p->A - p->K * vA2) - rho * p->K * dvA2;
    *a = -(1.0 / r - dX / X);
    *b = (double)(p->mode * p->mode) / (r * r) + p->K - p->A * p->A / (s * (p->A - p->K * cT2));
}

/* slab with sheared flow: flow_multiprocessor_coronal.py:211-219,297 */
static void coef_int_flow(const void* c, double x, double* a, double* b) {
    const pt_ctx* p = (const pt_ctx*)c;
    const ork_model* m = p->m;
    const double g = exp(-(x - m->x0) * (x - m->x0) / (m->width * m->width));
    const double t1 = -2.0 * (x - m->x0) / (m->width * m->width);
    const double U = m->U_e + (m->U_i0 - m->U_e) * g;
    const double dU = (m->U_i0 - m->U_e) * g * t1;
    const double ddU = (m->U_i0 - m->U_e) * g * (t1 * t1 - 2.0 / (m->width * m->width));
    const double c2 = m->c_i0 * m->c_i0, vA2 = m->vA_i0 * m->vA_i0, s = c2 + vA2, cT2 = c2 * vA2 / s;
    const double Om = p->w - p->k * U;
    const double m0 = (p->K * c2 - Om * Om) * (p->K * vA2 - Om * Om) / (s * (p->K * cT2 - Om * Om));
    const double t = Om * Om - p->K * cT2;
    const double D = 2.0 * p->k * dU * (t + p->K * p->K * cT2 * c2 / (s * t)) / (Om * (Om * Om - p->K * c2));
    const double coeff = p->k * ddU / Om + p->k * dU * D / Om - m0;
    *a = -D;
    *b = -coeff;
}

/* cylinder with rotational flow: Twisted_photospheric_nonlinear_flow_kink_fast.py:264-297.
 * D, C1, C2, C3 are written as the reference writes them; F = r D/C3 and r C1/C3 are
 * differentiated numerically (8th-order central differences) where the reference uses sympy. */
/* Forward-mode automatic differentiation in r (value, d/dr): the exact derivative of the same expression
 * tree, where the reference calls sympy.diff.  (A finite-difference version of this oracle was wrong by
 * up to 3e-3 at k = 0.05 for the linear rotation law, where C3 comes close to zero and F = r D/C3 varies
 * rapidly; the scipy oracle with sympy derivatives and an independent integration of the first-order
 * system agreed with each other there to 1e-9, and with this version.) */
typedef struct { double v, d; } dual;
static dual dc(double c) { dual r = {c, 0.0}; return r; }
static dual dadd(dual a, dual b) { dual r = {a.v + b.v, a.d + b.d}; return r; }
static dual dsub(dual a, dual b) { dual r = {a.v - b.v, a.d - b.d}; return r; }
static dual dmul(dual a, dual b) { dual r = {a.v * b.v, a.d * b.v + a.v * b.d}; return r; }
static dual ddiv(dual a, dual b) { dual r = {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; return r; }
static dual dscale(double c, dual a) { dual r = {c * a.v, c * a.d}; return r; }
static dual dpow(dual a, double p) { dual r = {pow(a.v, p), p * pow(a.v, p - 1.0) * a.d}; return r; }
static dual dexp(dual a) { const double e = exp(a.v); dual r = {e, e * a.d}; return r; }

typedef struct { dual D, C1, C2, C3; } rot_c;
static rot_c rot_coeffs(const pt_ctx* p, double rv) {
    const ork_model* m = p->m;
    const double rho = m->rho_i0, mm = (double)p->mode;
    const dual r = {rv, 1.0};
    const dual r2 = dmul(r, r);
    const double vA2 = m->vA_i0 * m->vA_i0;
    const double alf2 = p->k * p->k * vA2;                          /* alfven_freq^2 (B_phi = 0) */
    rot_c c;
    if (m->kind == 4) {
        /* Cylinder_method_flow_testing.py:711-746 with v_phi = B_phi = 0 (:190-196): Q = T = 0 */
        const dual arg = dscale(-1.0 / (m->width * m->width), dmul(dsub(r, dc(m->x0)), dsub(r, dc(m->x0))));
        const dual vz = dadd(dc(m->U_e), dscale(m->U_i0 - m->U_e, dexp(arg)));
        const double c2f = m->c_i0 * m->c_i0;
        const dual sh = dsub(dc(p->w), dscale(p->k, vz));           /* shift_freq  :713 */
        const dual s2 = dmul(sh, sh);
        const double cu2 = alf2 * c2f / (c2f + vA2);                /* cusp_freq^2 :719 */
        c.D = dscale(rho * (c2f + vA2), dmul(dsub(s2, dc(alf2)), dsub(s2, dc(cu2))));
        c.C1 = dc(0.0);
        c.C2 = dsub(dmul(s2, s2), dscale(c2f + vA2, dmul(dadd(ddiv(dc(mm * mm), r2), dc(p->k * p->k)),
                                                          dsub(s2, dc(cu2)))));
        c.C3 = dscale(rho, dmul(c.D, dsub(s2, dc(alf2))));
        return c;
    }
    /* Twisted_photospheric_nonlinear_flow_kink_fast.py:105-111, 264-297 */
    const dual vphi = dscale(m->v_twist, dpow(r, m->power));
    const double P0 = m->c_i0 * m->c_i0 * rho / m->gamma;
    const dual Pi = dadd(dscale(rho * m->v_twist * m->v_twist / (2.0 * m->power), dpow(r, 2.0 * m->power)), dc(P0));
    const dual c2 = dscale(m->gamma / rho, Pi);
    const dual s = dadd(c2, dc(vA2));                               /* c^2 + vA^2 */
    const dual shift = dsub(dc(p->w), dscale(mm, ddiv(vphi, r)));
    const dual s2 = dmul(shift, shift);
    const dual cusp2 = ddiv(dscale(alf2, c2), s);
    const dual a1 = dsub(s2, dc(alf2));
    const dual A2 = dmul(s, dsub(s2, cusp2));                       /* (c^2+vA^2)(shift^2 - cusp^2) */
    c.D = dscale(rho, dmul(a1, A2));
    const dual Q = dscale(-rho, ddiv(dmul(a1, dmul(vphi, vphi)), r));
    const dual T = dscale(rho, dmul(vphi, shift));
    c.C1 = dsub(dmul(Q, s2), dscale(2.0 * mm, ddiv(dmul(A2, T), r2)));
    c.C2 = dsub(dmul(s2, s2), dmul(A2, dadd(ddiv(dc(mm * mm), r2), dc(p->k * p->k))));
    /* C3_diff = -rho (v_phi/r)^2 ; r d/dr C3_diff, itself differentiated once more below */
    const dual q = ddiv(vphi, r);
    const dual C3diff = dscale(-rho, dmul(q, q));
    /* r * d(C3diff)/dr as a dual needs the second derivative of C3diff: for v_phi = v r^p it is closed form */
    const double pw = 2.0 * m->power - 2.0;
    const dual rdC3 = dscale(-rho * m->v_twist * m->v_twist * pw, dpow(r, pw));
    (void)C3diff;
    c.C3 = dadd(dmul(c.D, dadd(dscale(rho, a1), rdC3)),
                dsub(dmul(Q, Q), dscale(4.0, ddiv(dmul(A2, dmul(T, T)), r2))));
    return c;
}
static void coef_int_rot(const void* cv, double r, double* a, double* b) {
    const pt_ctx* p = (const pt_ctx*)cv;
    const rot_c c = rot_coeffs(p, r);
    const dual rr = {r, 1.0};
    const dual F = ddiv(dmul(rr, c.D), c.C3);                       /* F = r D/C3  (:288) */
    const dual G = ddiv(dmul(rr, c.C1), c.C3);                      /* r C1/C3 */
    const double g = -G.d - r * (c.C2.v - c.C1.v * c.C1.v / c.C3.v) / c.D.v;   /* (:294) */
    *a = -F.d / F.v;                                                /* dP_dr_i  (:304) */
    *b = g / F.v;
}

static double cluster(double t) {
    const double s = sin(0.5 * M_PI * t);
    return s * s;
}

/* one (k, w): returns 0 and fills ext/int, or 1 when the reference skips the point */
int ork_point(const ork_model* m, int mode, double k, double w, double* ext_q, double* int_q) {
    tableau();
    pt_ctx p;
    p.m = m; p.k = k; p.w = w; p.K = k * k; p.A = w * w; p.mode = mode;
    p.rho_e = rho_e_of(m);
    const double vAe2 = m->vA_e * m->vA_e, ce2 = m->c_e * m->c_e;
    p.cT_e2 = ce2 * vAe2 / (ce2 + vAe2);
    const double We = (m->kind == 2) ? w - k * m->U_e : w;     /* exterior Doppler shift (flow :207) */
    const double Ae = We * We;
    p.m_e = ((p.K * vAe2 - Ae) * (p.K * ce2 - Ae)) / ((vAe2 + ce2) * (p.K * p.cT_e2 - Ae));
    /* m->leaky (test-only extension, default 0 = the reference's rule): the exterior below is integrated
     * numerically, so without the skip it simply returns the oscillatory solution of the same problem */
    if (!(p.m_e >= 0.0) && !(m->leaky == 1 && p.m_e < 0.0)) {
        *ext_q = NAN; *int_q = NAN;
        return 1;
    }
    /* exterior */
    double y[2], yp[2];
    y[0] = m->ic_v; yp[0] = m->ic_s;
    const double rs = (m->kind == 3 || (m->kind == 1 && m->r_sign > 0)) ? 1.0 : -1.0;
    const double x0 = rs * m->ext_wavelengths * 2.0 * M_PI / k;
    if (m->kind == 0 || m->kind == 2) {
        const double h = (-1.0 - x0) / m->n_ext;
        for (int i = 0; i < m->n_ext; ++i) rk8_step(coef_ext_slab, &p, x0 + i * h, h, 1, y, yp);
    } else {
        /* geometric mesh in |r| resolves both the far field and the 1/r^2 term near r=-1 */
        const double L = log(fabs(x0));
        double r = x0;
        for (int i = 0; i < m->n_ext; ++i) {
            const double rn = (i == m->n_ext - 1) ? rs : rs * exp(L * (1.0 - (double)(i + 1) / m->n_ext));
            rk8_step(coef_ext_cyl, &p, r, rn - r, 1, y, yp);
            r = rn;
        }
    }
    const double yb = y[0], ypb = yp[0];
    double rho = 0, drho = 0, c2 = 0, dc2 = 0, vA2 = 0, dvA2 = 0;
    if (m->kind < 2) profile(&p, m->s_start, &rho, &drho, &c2, &dc2, &vA2, &dvA2);
    /* interior: two fundamental solutions forward from the boundary */
    double Y[2] = {1.0, 0.0}, Yp[2] = {0.0, 1.0};
    const int N = m->n_int;
    if (m->kind == 3 || m->kind == 4) {
        double x = m->s_start;
        for (int i = 1; i <= N; ++i) {
            const double xn = (i == N) ? m->s_end : m->s_start + (m->s_end - m->s_start) * cluster((double)i / N);
            rk8_step(coef_int_rot, &p, x, xn - x, 2, Y, Yp);
            x = xn;
        }
        const double xi_e = -ypb / (p.rho_e * (p.K * vAe2 - p.A));
        /* sausage: P'(end) = 0 ; kink: P(end) + (0 - rho v_phi(1)^2) xi_e = 0   (:308) */
        const double rv2 = (m->kind == 4) ? 0.0 : m->rho_i0 * m->v_twist * m->v_twist;
        const double slope = (mode == 0) ? -yb * Yp[0] / Yp[1] : (rv2 * xi_e - yb * Y[0]) / Y[1];
        const rot_c cb = rot_coeffs(&p, m->s_start);
        *ext_q = xi_e;
        *int_q = (cb.C1.v * yb + cb.D.v * slope) / cb.C3.v;       /* inside_xi_solution[0]  (:314) */
    } else if (m->kind != 1) {
        const coef_fn cf = (m->kind == 2) ? coef_int_flow : coef_int_slab;
        const int H = N / 2;
        const double mid = 0.5 * (m->s_start + m->s_end);
        double x = m->s_start;
        for (int i = 1; i <= H; ++i) {
            const double xn = (i == H) ? mid : m->s_start + (mid - m->s_start) * cluster((double)i / H);
            rk8_step(cf, &p, x, xn - x, 2, Y, Yp);
            x = xn;
        }
        for (int i = 1; i <= H; ++i) {
            const double xn = (i == H) ? m->s_end : mid + (m->s_end - mid) * cluster((double)i / H);
            rk8_step(cf, &p, x, xn - x, 2, Y, Yp);
            x = xn;
        }
        /* sausage: vx(1) + vx(-1) = 0, kink: vx(1) - vx(-1) = 0 */
        const double target = (mode == 0) ? -1.0 : 1.0;
        const double p_e_const = p.rho_e * (vAe2 + ce2) * (p.K * p.cT_e2 - Ae) / (We * (p.K * ce2 - Ae));
        double y0 = yb, P_Ti;
        if (m->kind == 2) {
            const double g = exp(-(m->s_start - m->x0) * (m->s_start - m->x0) / (m->width * m->width));
            const double Ob = w - k * (m->U_e + (m->U_i0 - m->U_e) * g);
            const double ci2 = m->c_i0 * m->c_i0, vi2 = m->vA_i0 * m->vA_i0, cTi2 = ci2 * vi2 / (ci2 + vi2);
            y0 = yb * Ob / We;                               /* left_solution scaling (flow :290) */
            P_Ti = m->rho_i0 * (vi2 + ci2) * (p.K * cTi2 - Ob * Ob) / (Ob * (p.K * ci2 - Ob * Ob));
        } else {
            const double cT2 = c2 * vA2 / (c2 + vA2);
            P_Ti = rho * (vA2 + c2) * (p.K * cT2 - p.A) / (w * (p.K * c2 - p.A));
        }
        const double slope = y0 * (target - Y[0]) / Y[1];
        *ext_q = p_e_const * ypb;
        *int_q = P_Ti * slope;
    } else {
        double x = m->s_start;
        for (int i = 1; i <= N; ++i) {
            const double xn = (i == N) ? m->s_end : m->s_start + (m->s_end - m->s_start) * cluster((double)i / N);
            rk8_step(coef_int_cyl, &p, x, xn - x, 2, Y, Yp);
            x = xn;
        }
        /* kink/fluting: P(axis) = 0; sausage: P'(axis) = 0 */
        const double slope = (mode == 0) ? -yb * Yp[0] / Yp[1] : -yb * Y[0] / Y[1];
        *ext_q = -ypb / (p.rho_e * (p.K * vAe2 - p.A));
        *int_q = slope / (rho * (p.A - p.K * vA2));
    }
    return 0;
}

/* whole grid; layout 0 shared w[nw], 1 phase speed, 2 per-k.  Rows are handed to
 * `threads` pthreads (libgomp is not in the image) through an atomic row counter. */
typedef struct {
    const ork_model* m;
    int mode, nk, nw, layout;
    const double *k, *w;
    double *ext, *intq;
    int next;
} grid_job;

static void* grid_worker(void* arg) {
    grid_job* g = (grid_job*)arg;
    for (;;) {
        const int i = __atomic_fetch_add(&g->next, 1, __ATOMIC_RELAXED);
        if (i >= g->nk) break;
        for (int j = 0; j < g->nw; ++j) {
            const double om = g->layout == 0 ? g->w[j] : g->layout == 1 ? g->k[i] * g->w[j]
                                                                          : g->w[(size_t)i * g->nw + j];
            ork_point(g->m, g->mode, g->k[i], om, &g->ext[(size_t)i * g->nw + j],
                      &g->intq[(size_t)i * g->nw + j]);
        }
    }
    return NULL;
}

void ork_grid(const ork_model* m, int mode, const double* k, int nk, const double* w, int nw, int layout,
              double* ext, double* intq, int threads) {
    tableau();
    grid_job g = {m, mode, nk, nw, layout, k, w, ext, intq, 0};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 1; t < threads; ++t) pthread_create(&th[t], NULL, grid_worker, &g);
    grid_worker(&g);
    for (int t = 1; t < threads; ++t) pthread_join(th[t], NULL);
}

/* bisection + secant-free polish: plain bisection to machine precision (independent of
 * the GPU's Brent).  Returns the root; ext/int at the root in out[0..1]. */
double ork_refine(const ork_model* m, int mode, double k, double wlo, double whi, double* out) {
    double e, q;
    ork_point(m, mode, k, wlo, &e, &q);
    double flo = e - q;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (wlo + whi);
        if (mid == wlo || mid == whi) break;
        ork_point(m, mode, k, mid, &e, &q);
        const double fm = e - q;
        if ((fm < 0) == (flo < 0)) { wlo = mid; flo = fm; } else { whi = mid; }
    }
    const double r = 0.5 * (wlo + whi);
    ork_point(m, mode, k, r, &out[0], &out[1]);
    return r;
}
