// Host side of a model upload: the mesh along the layer and the staged table the kernels read.
// Pure host code (no CUDA calls): esb.cu uses it for esb_set_model_fields / esb_mesh_nodes, the
// test harness (tests/host_harness) uses the SAME functions to run the kernels' device code, compiled
// for the host, against the oracle on the CPU test box.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/eigensolver_b200.h"
#include "core.cuh"

namespace esb {

// ---- mesh: stage nodes along the direction of integration -------------------
static inline double cluster(double t) {   // sin^2(pi t/2): clusters nodes at both ends of [0,1]
    const double s = sin(0.5 * M_PI * t);
    return s * s;
}

// mesh = 2: breakpoints with the local step H * min(1, |r|/axis, (edge + d)/edge_width), N steps fix H.
// Cylinder kinds: from the axis end to the boundary, d = distance to the boundary.  Slab kinds: from
// s_start to s_end, no axis term, d = distance to the nearer boundary (symmetric, so the mid-plane is
// a breakpoint for even N).  The node count up to r, t(r) = int dr/h, is accumulated on a fine
// midpoint rule and inverted by linear interpolation.
static inline void graded_breakpoints(const esb_model* m, bool slab, std::vector<double>& out) {
    const int N = m->n_steps, M = 400000;
    const double a = slab ? m->s_start : m->s_end, b = slab ? m->s_end : m->s_start;
    const double ax = (!slab && m->mesh_axis > 0) ? m->mesh_axis : 0.0;
    const double ew = m->mesh_edge_width > 0 ? m->mesh_edge_width : 0.0;
    std::vector<double> t(M + 1);
    t[0] = 0.0;
    const double dr = (b - a) / M;
    for (int j = 0; j < M; ++j) {
        const double r = a + (j + 0.5) * dr;
        double h = 1.0;
        if (ax > 0) h = fmin(h, fabs(r) / ax);
        if (ew > 0) {
            const double d = slab ? fmin(fabs(r - a), fabs(r - b)) : fabs(r - b);
            h = fmin(h, (m->mesh_edge + d) / ew);
        }
        t[j + 1] = t[j] + fabs(dr) / h;
    }
    out.resize(N + 1);
    int j = 0;
    for (int i = 0; i <= N; ++i) {
        const double target = t[M] * double(i) / N;
        while (j < M - 1 && t[j + 1] < target) ++j;
        const double f = (target - t[j]) / (t[j + 1] - t[j]);
        out[i] = a + (j + f) * dr;
    }
    out[0] = a;
    out[N] = b;
    if (slab && N % 2 == 0) {
        // exactly mirror-symmetric about the mid-plane (the accumulated sum above is so only to ~1e-11):
        // a symmetric profile then takes the half-layer path of eval_point_multi
        out[N / 2] = 0.5 * (a + b);
        for (int i = 0; i < N / 2; ++i) out[N - i] = (a + b) - out[i];
    }
}

static inline int build_breakpoints_uncached(const esb_model* m, std::vector<double>& bp) {
    const int N = m->n_steps;
    bp.resize(N + 1);
    const bool cyl = m->kind == ESB_CYLINDER_ROTATION || m->kind == ESB_CYLINDER_DENSITY ||
                     m->kind == ESB_CYLINDER_FLOW;
    if (m->mesh == 2) {
        graded_breakpoints(m, !cyl, bp);                          // cylinder: axis -> boundary
        return ESB_OK;
    }
    if (cyl) {
        // every cylinder kind: from the axis end (s_end) out to the boundary (s_start)
        for (int i = 0; i <= N; ++i) {
            const double t = double(i) / N;
            const double f = m->mesh == 1 ? t : cluster(t);
            bp[i] = m->s_end + (m->s_start - m->s_end) * f;
        }
        bp[0] = m->s_end;
        bp[N] = m->s_start;
    } else {
        // slab: boundary -> mid -> far boundary, clustered at the three of them
        if (N % 2) return ESB_ERR_ARG;
        const int H = N / 2;
        const double mid = 0.5 * (m->s_start + m->s_end);
        for (int i = 0; i <= H; ++i) {
            const double t = double(i) / H;
            const double f = m->mesh == 1 ? t : cluster(t);
            bp[i] = m->s_start + (mid - m->s_start) * f;
            bp[H + i] = mid + (m->s_end - mid) * f;
        }
        bp[0] = m->s_start;
        bp[H] = mid;
        bp[N] = m->s_end;
    }
    return ESB_OK;
}

// The breakpoints depend on the discretisation fields of the model only; a parameter scan calls
// esb_set_model_fields once per equilibrium with the same mesh, so the last mesh is kept.
static inline int build_breakpoints(const esb_model* m, std::vector<double>& bp) {
    struct Key {
        int32_t kind, n_steps, mesh;
        double s_start, s_end, axis, edge, width;
        bool operator==(const Key& o) const {
            return kind == o.kind && n_steps == o.n_steps && mesh == o.mesh && s_start == o.s_start &&
                   s_end == o.s_end && axis == o.axis && edge == o.edge && width == o.width;
        }
    };
    static std::mutex mu;
    static Key last{-1, 0, 0, 0, 0, 0, 0, 0};
    static std::vector<double> last_bp;
    const Key key{m->kind, m->n_steps, m->mesh, m->s_start, m->s_end,
                  m->mesh_axis, m->mesh_edge, m->mesh_edge_width};
    std::lock_guard<std::mutex> lock(mu);
    if (!(key == last)) {
        std::vector<double> fresh;
        const int rc = build_breakpoints_uncached(m, fresh);
        if (rc) return rc;
        last = key;
        last_bp.swap(fresh);
    }
    bp = last_bp;
    return ESB_OK;
}

// the second-order kinds that have the normal-form scheme (ESB_RK8N)
static inline bool kind_has_normal_form(int kind) {
    return kind == ESB_CYLINDER_DENSITY || kind == ESB_CYLINDER_FLOW || kind == ESB_SLAB_DENSITY;
}

static inline int check_model(const esb_model* m) {
    if (!m) return ESB_ERR_ARG;
    if (m->kind < ESB_SLAB_DENSITY || m->kind > ESB_CYLINDER_FLOW) return ESB_ERR_ARG;
    if (m->scheme != ESB_RK4 && m->scheme != ESB_RK8 && m->scheme != ESB_RK8N) return ESB_ERR_ARG;
    if (m->kind == ESB_CYLINDER_ROTATION && m->scheme != ESB_RK8) return ESB_ERR_ARG;
    if (m->scheme == ESB_RK8N && !kind_has_normal_form(m->kind)) return ESB_ERR_ARG;
    if (m->n_steps < 2 || m->n_steps > 8192) return ESB_ERR_ARG;
    if (m->mesh < 0 || m->mesh > 2) return ESB_ERR_ARG;
    if ((m->kind == ESB_SLAB_DENSITY || m->kind == ESB_SLAB_FLOW) && (m->n_steps % 2)) return ESB_ERR_ARG;
    return ESB_OK;
}

static inline const double* stage_fracs(int scheme, int& n) {
    static const double f8[4] = {0.0, C8_M, 0.5, C8_P};
    static const double f4[2] = {0.0, 0.5};
    if (scheme == ESB_RK4) { n = 2; return f4; }
    n = 4;
    return f8;
}

static inline int mesh_size(const esb_model* m) { return m->n_steps * nodes_per_step(m->scheme) + 1; }

static inline int mesh_nodes(const esb_model* m, double* nodes) {
    std::vector<double> bp;
    if (build_breakpoints(m, bp)) return ESB_ERR_ARG;
    int nf;
    const double* fr = stage_fracs(m->scheme, nf);
    const int N = m->n_steps;
    for (int i = 0; i < N; ++i) {
        const double h = bp[i + 1] - bp[i];
        for (int j = 0; j < nf; ++j) nodes[i * nf + j] = bp[i] + fr[j] * h;
    }
    nodes[N * nf] = bp[N];
    return ESB_OK;
}

// number of profile fields esb_set_model_fields expects for (kind, scheme)
static inline int model_n_fields(const esb_model* m) {
    if (m->kind == ESB_SLAB_FLOW || m->kind == ESB_CYLINDER_ROTATION) return 3;
    return m->scheme == ESB_RK8N ? 3 : 2;      // normal form: the second derivative of the profile too
}

// doubles per staged node for (kind, scheme)
static inline int model_tab_fields(const esb_model* m) {
    if (m->kind == ESB_CYLINDER_ROTATION) return ROT_FIELDS;
    return m->scheme == ESB_RK8N ? NF_FIELDS : TAB_FIELDS;
}

// the staged table lives in the shared memory of every CTA
constexpr size_t TABLE_BYTES_MAX = 200 * 1024;

// doubles of the staged table: [n_nodes][fields per node], then h, g, h^2, g^2 per step
static inline size_t model_tab_doubles(const esb_model* m) {
    return (size_t)mesh_size(m) * model_tab_fields(m) + 4 * (size_t)m->n_steps;
}

// the largest n_steps whose table can be staged for this (kind, scheme)
static inline int model_max_steps(const esb_model* m) {
    esb_model t = *m;
    int lo = 0, hi = 1 << 20;                   // table(lo) fits, table(hi) does not
    while (hi - lo > 1) {
        t.n_steps = lo + (hi - lo) / 2;
        if (model_tab_doubles(&t) * sizeof(double) <= TABLE_BYTES_MAX) lo = t.n_steps; else hi = t.n_steps;
    }
    if ((m->kind == ESB_SLAB_DENSITY || m->kind == ESB_SLAB_FLOW) && (lo % 2)) lo -= 1;     // check_model: even
    return lo;
}

struct HostModel {
    DevModel dm;
    std::vector<double> tab;
};

// Everything esb_set_model_fields uploads, built on the host: the device model constants and the
// staged table [n_nodes][fields per node], h[N], g[N] = h[i+1]/h[i] (last: 1/h[N-1]), h^2[N], g^2[N].
// fields[f][node]: density kinds {rho, rho'} (+ rho'' for ESB_RK8N); slab flow {U, U', U''};
// rotation {v_phi, v_phi', c_i^2}; axial flow {v_z, v_z'} (+ v_z'').  boundary[0] = first field at s_start.
static inline int build_host_model(const esb_model* m, const double* const* fields, int32_t n_fields,
                                   int32_t n_nodes, const double* boundary, int32_t n_boundary, HostModel& out,
                                   std::string& err) {
    if (check_model(m) || !fields || !boundary) { err = "bad model"; return ESB_ERR_ARG; }
    if (n_fields != model_n_fields(m) || n_boundary < 1) { err = "wrong number of profile fields"; return ESB_ERR_ARG; }
    for (int f = 0; f < n_fields; ++f)
        if (!fields[f]) { err = "null profile field"; return ESB_ERR_ARG; }
    const int need = mesh_size(m);
    if (n_nodes != need) { err = "n_nodes does not match esb_mesh_size()"; return ESB_ERR_ARG; }
    std::vector<double> nodes(need);
    if (mesh_nodes(m, nodes.data())) { err = "mesh"; return ESB_ERR_ARG; }
    const int N = m->n_steps, nps = nodes_per_step(m->scheme);
    const int tf = model_tab_fields(m);
    std::vector<double>& tab = out.tab;
    tab.assign(model_tab_doubles(m), 0.0);
    // step a node belongs to (the step-end node is stored in the scale of the step it ends and is
    // shared with the next step, which rescales the carried coefficients)
    auto step_h = [&](int i) {
        const int step = i == 0 ? 0 : (i - 1) / nps;
        return nodes[(step + 1) * nps] - nodes[step * nps];
    };
    for (int i = 0; i < need; ++i) {
        double* f = &tab[(size_t)i * tf];
        if (m->scheme == ESB_RK8N) {
            // normal form u'' = q u (core.cuh integrate_layer_nform), everything pre-scaled by the step:
            //   cylinder: {-h/(2r)/c, h^2/r^2, field, c h field', -h^2 field''/2, h^2 field^2 (density),
            //              -h/(2r), h field'},  c = sqrt(3/4): q contains 3/4 L^2 - L/(2r) with L = h field'/X,
            //              and with the first derivative stored times c both terms are single FMAs of L' = c L
            //              (fields 6, 7: the unscaled pair for the (y, h y') variables and the end conversion)
            //   slab:     {h^2, -, rho, h rho', -h^2 rho''/2, -, h^2, h rho'}
            const double h = step_h(i), h2 = h * h;
            const double v = fields[0][i], dv = fields[1][i], ddv = fields[2][i];
            if (m->kind == ESB_SLAB_DENSITY) {
                f[0] = f[6] = h2;
                f[1] = 0.0;
                f[3] = f[7] = h * dv;
            } else {
                const double r = nodes[i];
                f[6] = -0.5 * h / r;
                f[7] = h * dv;
                f[0] = f[6] / NF_SQRT34;
                f[3] = f[7] * NF_SQRT34;
                f[1] = h2 / (r * r);
            }
            f[2] = v;
            f[4] = -0.5 * h2 * ddv;
            f[5] = h2 * v * v;
        } else if (m->kind == ESB_CYLINDER_DENSITY || m->kind == ESB_CYLINDER_FLOW) {
            const double r = nodes[i];
            f[0] = 1.0 / r;
            f[1] = 1.0 / (r * r);
            f[2] = fields[0][i];
            f[3] = fields[1][i];
            if (m->scheme == ESB_RK8) {
                // pre-scaled layout (core.cuh integrate_layer_prescaled): {h/r, h^2/r^2, field, h field'}
                const double h = step_h(i);
                f[0] *= h;
                f[1] *= h * h;
                f[3] *= h;
            }
        } else if (m->kind == ESB_CYLINDER_ROTATION) {
            // fields = {v_phi, v_phi', c^2};  r d/dr(-rho v_phi^2/r^2) = -2 rho v_phi (r v_phi' - v_phi)/r^2
            // staged: the (k, omega, m)-independent products node_rot needs (core.cuh ROT_FIELDS)
            const double r = nodes[i], v = fields[0][i], dv = fields[1][i], c2 = fields[2][i];
            const double rho = m->rho_i0;
            f[0] = 1.0 / r;
            f[1] = 1.0 / (r * r);
            f[2] = v / r;
            f[3] = c2 + m->vA_i0 * m->vA_i0;
            f[4] = c2;
            f[5] = rho * v;
            f[6] = rho * v * v / r;
            f[7] = -2.0 * rho * v * (r * dv - v) / (r * r);
        } else {
            for (int q = 0; q < n_fields; ++q) f[q] = fields[q][i];
        }
    }
    {
        double* hs = &tab[(size_t)need * tf];
        for (int i = 0; i < N; ++i) hs[i] = nodes[(i + 1) * nps] - nodes[i * nps];
        for (int i = 0; i < N; ++i) hs[N + i] = (i + 1 < N) ? hs[i + 1] / hs[i] : 1.0 / hs[i];
        for (int i = 0; i < N; ++i) hs[2 * N + i] = hs[i] * hs[i];             // h^2
        for (int i = 0; i < N; ++i) hs[3 * N + i] = hs[N + i] * hs[N + i];     // (h'/h)^2
    }

    DevModel& d = out.dm;
    memset(&d, 0, sizeof(d));
    d.kind = m->kind;
    d.scheme = m->scheme;
    d.n_steps = N;
    d.n_nodes = need;
    d.vAe2 = m->vA_e * m->vA_e;
    d.ce2 = m->c_e * m->c_e;
    d.se2 = d.vAe2 + d.ce2;
    d.cTe2 = d.ce2 * d.vAe2 / d.se2;
    const double g = m->gamma;
    d.rho_e = m->rho_i0 * (m->c_i0 * m->c_i0 + g * 0.5 * m->vA_i0 * m->vA_i0) /
              (d.ce2 + g * 0.5 * d.vAe2);                    // Density_cylinder.py:80 / flow :56
    d.ic_v = m->ext_ic_value;
    d.ic_s = m->ext_ic_slope;
    d.ext_len = m->ext_wavelengths * 2.0 * M_PI;
    d.s_start = m->s_start;
    d.r_sign = m->r_sign >= 0 ? 1.0 : -1.0;
    d.r_axis = nodes[0];
    d.f_lo = d.f_hi = boundary[0];
    for (int i = 0; i < need; ++i) {
        d.f_lo = fmin(d.f_lo, fields[0][i]);
        d.f_hi = fmax(d.f_hi, fields[0][i]);
    }
    if (m->kind == ESB_SLAB_DENSITY || m->kind == ESB_SLAB_FLOW) {
        // mirror symmetry of the staged profile about the mid-plane: even fields (rho, rho''; U, U'') equal,
        // odd fields (rho'; U') opposite at mirrored nodes (the mesh itself is symmetric for even N)
        bool sym = (N % 2 == 0);
        for (int q = 0; q < n_fields && sym; ++q) {
            const double parity = (q == 1) ? -1.0 : 1.0;
            double scale = 0.0;
            for (int i = 0; i < need; ++i) scale = fmax(scale, fabs(fields[q][i]));
            for (int i = 0; i < need && sym; ++i)
                sym = fabs(fields[q][i] - parity * fields[q][need - 1 - i]) <= 1e-12 * scale;
        }
        for (int i = 0; i < need && sym; ++i)
            sym = fabs((nodes[i] - nodes[0]) - (nodes[need - 1] - nodes[need - 1 - i])) <=
                  1e-12 * fabs(nodes[need - 1] - nodes[0]);
        d.symmetric = sym ? 1 : 0;
    }
    if (m->kind == ESB_SLAB_FLOW || m->kind == ESB_CYLINDER_FLOW) {
        d.ci2 = m->c_i0 * m->c_i0;
        d.vAi2 = m->vA_i0 * m->vA_i0;
        d.si = d.ci2 + d.vAi2;
        d.cTi2 = d.ci2 * d.vAi2 / d.si;
        d.rho_i = m->rho_i0;
        d.U_e = m->U_e;
        d.U_b = boundary[0];
    } else if (m->kind == ESB_CYLINDER_ROTATION) {
        d.vAi2 = m->vA_i0 * m->vA_i0;
        d.rho_i = m->rho_i0;
        d.rho_vb2 = m->rho_i0 * boundary[0] * boundary[0];
    } else {
        // c_i^2 = rho_e (c_e^2 + gamma/2 vA_e^2)/rho - gamma/2 vA_i^2   (Density_cylinder.py:210)
        // cylinder: vA_i^2 = B_0^2/rho = vA_i0^2 rho_i0/rho                (Density_cylinder.py:188-200)
        // slab:     vA_i^2 = vA_i0^2 rho_i0/profile = vA_i0^2 rho_i0 rho_A/rho   (..._coronal.py:117)
        d.beta = m->vA_i0 * m->vA_i0 * m->rho_i0 * (m->kind == ESB_SLAB_DENSITY ? m->rho_A : 1.0);
        d.alpha = d.rho_e * (d.ce2 + 0.5 * g * d.vAe2) - 0.5 * g * d.beta;
        d.S = d.alpha + d.beta;
        d.invS = 1.0 / d.S;
        d.tau = d.alpha * d.beta / d.S;
        d.rho_b = boundary[0];
    }
    if (tab.size() * sizeof(double) > TABLE_BYTES_MAX) {
        err = "n_steps too large for the shared-memory table (200 KB)";
        return ESB_ERR_ARG;
    }
    return ESB_OK;
}

}  // namespace esb
