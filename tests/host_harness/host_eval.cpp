// TEST INFRASTRUCTURE (not product): the kernels' device code (csrc/core.cuh, ESB_HD functions) and
// the model/table builder (csrc/model_host.h) compiled for the HOST with g++, so that the CPU test
// suite can check the exact arithmetic the GPU runs - every kind, scheme and mode - against the
// oracle without a GPU.  Nothing in eigensolver_b200/ loads or links this.
#include <stdint.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../eigensolver_b200/csrc/core.cuh"
#include "../../eigensolver_b200/csrc/model_host.h"

using namespace esb;

template <int KIND, int SCHEME>
static void eval_all(const HostModel& hm, int n_modes, const int32_t* modes, int64_t n, const double* k,
                     const double* w, double* ext, double* intq, double* den) {
    std::atomic<int64_t> next{0};
    auto work = [&]() {
      for (;;) {
        const int64_t p0 = next.fetch_add(64);
        if (p0 >= n) return;
        for (int64_t p = p0; p < n && p < p0 + 64; ++p) {
        if (n_modes == 3) {            // the fused three-mode evaluation of the scan kernel
            const int md[3] = {modes[0], modes[1], modes[2]};
            double e[3], i[3], d[3];
            eval_point_multi<KIND, SCHEME, 3>(hm.dm, hm.tab.data(), k[p], w[p], md, e, i, d);
            for (int s = 0; s < 3; ++s) { ext[s * n + p] = e[s]; intq[s * n + p] = i[s]; den[s * n + p] = d[s]; }
        } else if (n_modes == 2) {
            const int md[2] = {modes[0], modes[1]};
            double e[2], i[2], d[2];
            eval_point_multi<KIND, SCHEME, 2>(hm.dm, hm.tab.data(), k[p], w[p], md, e, i, d);
            for (int s = 0; s < 2; ++s) { ext[s * n + p] = e[s]; intq[s * n + p] = i[s]; den[s * n + p] = d[s]; }
        } else {
            for (int s = 0; s < n_modes; ++s)
                eval_point<KIND, SCHEME>(hm.dm, hm.tab.data(), k[p], w[p], modes[s], ext[s * n + p], intq[s * n + p],
                                         den[s * n + p]);
        }
        }
      }
    };
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

// the opt-in leaky evaluation (eval_point<..., LEAKY = true>), one mode at a time
template <int KIND, int SCHEME>
static void eval_leaky(const HostModel& hm, int n_modes, const int32_t* modes, int64_t n, const double* k,
                       const double* w, double* ext, double* intq, double* den) {
    for (int64_t p = 0; p < n; ++p)
        for (int s = 0; s < n_modes; ++s)
            eval_point<KIND, SCHEME, false, false, true>(hm.dm, hm.tab.data(), k[p], w[p], modes[s], ext[s * n + p],
                                                         intq[s * n + p], den[s * n + p]);
}

template <int KIND>
static int by_scheme(const HostModel& hm, int n_modes, const int32_t* modes, int64_t n, const double* k,
                     const double* w, double* ext, double* intq, double* den) {
    if (hm.dm.scheme == SCHEME_RK8) eval_all<KIND, SCHEME_RK8>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else if (hm.dm.scheme == SCHEME_RK4) {
        if constexpr (KIND == KIND_CYL_ROTATION) return ESB_ERR_ARG;
        else eval_all<KIND, SCHEME_RK4>(hm, n_modes, modes, n, k, w, ext, intq, den);
    } else {
        if constexpr (KIND == KIND_CYL_ROTATION || KIND == KIND_SLAB_FLOW) return ESB_ERR_ARG;
        else eval_all<KIND, SCHEME_RK8N>(hm, n_modes, modes, n, k, w, ext, intq, den);
    }
    return ESB_OK;
}

// (ext, int, den)[mode slot][point] at n arbitrary (k, omega) points; same model arguments as
// esb_set_model_fields.  n_modes = 3 / 2 take the fused evaluation, else one mode at a time.
extern "C" int esbh_eval_points(const esb_model* m, const double* const* fields, int32_t n_fields, int32_t n_nodes,
                                const double* boundary, int32_t n_boundary, int32_t n_modes, const int32_t* modes,
                                int64_t n, const double* k, const double* w, double* ext, double* intq,
                                double* den) {
    HostModel hm;
    std::string err;
    int rc = build_host_model(m, fields, n_fields, n_nodes, boundary, n_boundary, hm, err);
    if (rc) return rc;
    switch (hm.dm.kind) {
        case KIND_SLAB_DENSITY: return by_scheme<KIND_SLAB_DENSITY>(hm, n_modes, modes, n, k, w, ext, intq, den);
        case KIND_CYL_DENSITY: return by_scheme<KIND_CYL_DENSITY>(hm, n_modes, modes, n, k, w, ext, intq, den);
        case KIND_SLAB_FLOW: return by_scheme<KIND_SLAB_FLOW>(hm, n_modes, modes, n, k, w, ext, intq, den);
        case KIND_CYL_ROTATION: return by_scheme<KIND_CYL_ROTATION>(hm, n_modes, modes, n, k, w, ext, intq, den);
        case KIND_CYL_FLOW: return by_scheme<KIND_CYL_FLOW>(hm, n_modes, modes, n, k, w, ext, intq, den);
    }
    return ESB_ERR_ARG;
}

extern "C" int esbh_eval_points_leaky(const esb_model* m, const double* const* fields, int32_t n_fields,
                                      int32_t n_nodes, const double* boundary, int32_t n_boundary, int32_t n_modes,
                                      const int32_t* modes, int64_t n, const double* k, const double* w,
                                      double* ext, double* intq, double* den) {
    HostModel hm;
    std::string err;
    int rc = build_host_model(m, fields, n_fields, n_nodes, boundary, n_boundary, hm, err);
    if (rc) return rc;
    if (hm.dm.kind == KIND_CYL_DENSITY && hm.dm.scheme == SCHEME_RK8N)
        eval_leaky<KIND_CYL_DENSITY, SCHEME_RK8N>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else if (hm.dm.kind == KIND_SLAB_DENSITY && hm.dm.scheme == SCHEME_RK8N)
        eval_leaky<KIND_SLAB_DENSITY, SCHEME_RK8N>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else if (hm.dm.kind == KIND_CYL_ROTATION)
        eval_leaky<KIND_CYL_ROTATION, SCHEME_RK8>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else if (hm.dm.kind == KIND_SLAB_FLOW && hm.dm.scheme == SCHEME_RK8)
        eval_leaky<KIND_SLAB_FLOW, SCHEME_RK8>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else if (hm.dm.kind == KIND_CYL_FLOW && hm.dm.scheme == SCHEME_RK8N)
        eval_leaky<KIND_CYL_FLOW, SCHEME_RK8N>(hm, n_modes, modes, n, k, w, ext, intq, den);
    else
        return ESB_ERR_ARG;
    return ESB_OK;
}

// The discretisation guard's judgement (csrc/esb.cu guard_kernel) at n points: dev[slot][p] = guard_deviation
// between the COARSE model's (ext, int, den) and the FINE model's, NaN where the guard does not judge the point
// (skipped / overflowed / inside a resonant continuum within `margin`).  Both models as in esb_set_model_fields.
template <int KIND, int SCHEME_C, int SCHEME_F>
static void guard_all(const HostModel& hc, const HostModel& hf, int n_modes, const int32_t* modes, int64_t n,
                      const double* k, const double* w, double margin, double* dev) {
    for (int64_t p = 0; p < n; ++p)
        for (int s = 0; s < n_modes; ++s) {
            double e0, i0, d0, e, i, d;
            dev[s * n + p] = nan("");
            eval_point<KIND, SCHEME_C>(hc.dm, hc.tab.data(), k[p], w[p], modes[s], e0, i0, d0);
            if (!(std::isfinite(e0) && std::isfinite(i0) && std::isfinite(d0))) continue;
            const Point pt = make_point(hf.dm, k[p], w[p]);
            if (!resonance_free<KIND>(hf.dm, pt, double(modes[s]), hf.tab.data(), margin)) continue;
            eval_point<KIND, SCHEME_F>(hf.dm, hf.tab.data(), k[p], w[p], modes[s], e, i, d);
            const double v = guard_deviation(e0, i0, d0, e, i, d);
            if (std::isfinite(v)) dev[s * n + p] = v;
        }
}

template <int KIND>
static int guard_by_scheme(const HostModel& hc, const HostModel& hf, int n_modes, const int32_t* modes, int64_t n,
                           const double* k, const double* w, double margin, double* dev) {
    const int sc = hc.dm.scheme, sf = hf.dm.scheme;
    if (sc == SCHEME_RK8 && sf == SCHEME_RK8)
        guard_all<KIND, SCHEME_RK8, SCHEME_RK8>(hc, hf, n_modes, modes, n, k, w, margin, dev);
    else if constexpr (KIND == KIND_CYL_ROTATION || KIND == KIND_SLAB_FLOW)
        return ESB_ERR_ARG;
    else if (sc == SCHEME_RK8N && sf == SCHEME_RK8N)
        guard_all<KIND, SCHEME_RK8N, SCHEME_RK8N>(hc, hf, n_modes, modes, n, k, w, margin, dev);
    else if (sc == SCHEME_RK8N && sf == SCHEME_RK8)
        guard_all<KIND, SCHEME_RK8N, SCHEME_RK8>(hc, hf, n_modes, modes, n, k, w, margin, dev);
    else
        return ESB_ERR_ARG;
    return ESB_OK;
}

extern "C" int esbh_guard_points(const esb_model* mc, const double* const* fields_c, int32_t n_fields_c,
                                 int32_t n_nodes_c, const double* boundary_c, const esb_model* mf,
                                 const double* const* fields_f, int32_t n_fields_f, int32_t n_nodes_f,
                                 const double* boundary_f, int32_t n_boundary, int32_t n_modes, const int32_t* modes, int64_t n, const double* k, const double* w,
                                 double margin, double* dev) {
    HostModel hc, hf;
    std::string err;
    int rc = build_host_model(mc, fields_c, n_fields_c, n_nodes_c, boundary_c, n_boundary, hc, err);
    if (rc) return rc;
    rc = build_host_model(mf, fields_f, n_fields_f, n_nodes_f, boundary_f, n_boundary, hf, err);
    if (rc) return rc;
    if (hc.dm.kind != hf.dm.kind) return ESB_ERR_ARG;
    switch (hc.dm.kind) {
        case KIND_SLAB_DENSITY: return guard_by_scheme<KIND_SLAB_DENSITY>(hc, hf, n_modes, modes, n, k, w, margin, dev);
        case KIND_CYL_DENSITY: return guard_by_scheme<KIND_CYL_DENSITY>(hc, hf, n_modes, modes, n, k, w, margin, dev);
        case KIND_SLAB_FLOW: return guard_by_scheme<KIND_SLAB_FLOW>(hc, hf, n_modes, modes, n, k, w, margin, dev);
        case KIND_CYL_ROTATION: return guard_by_scheme<KIND_CYL_ROTATION>(hc, hf, n_modes, modes, n, k, w, margin, dev);
        case KIND_CYL_FLOW: return guard_by_scheme<KIND_CYL_FLOW>(hc, hf, n_modes, modes, n, k, w, margin, dev);
    }
    return ESB_ERR_ARG;
}
