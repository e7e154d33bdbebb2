// eigensolver_b200: kernels + C ABI (see include/eigensolver_b200.h).
//
// Kernels (all FP64, sm_100a):
//   grid_kernel        one thread per (k, omega), up to 3 modes fused per thread: exterior closed form +
//                      RK shooting across the layer with the profile table staged in shared memory.
//   grid_warp_kernel   one warp per (mode, k, omega) for small grids: the lanes integrate sub-intervals
//                      of the layer and multiply their transfer matrices (core.cuh warp_transfer).
//   bracket_kernel /   one warp per (k-row, omega segment): warp-ballot sign-change detection along
//   scan_kernel        omega, popc prefix -> deterministic, sorted bracket list.
//   refine_kernel      persistent, one lane per bracket, brackets pulled from a queue: Brent iteration
//                      on the pole-free function G = D * Y, warp-level vote (__any_sync) on the pending
//                      flags, poles classified from the scan, acceptance test.
//   refine_warp_kernel the same iteration, one warp per bracket (short bracket lists: latency).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/eigensolver_b200.h"
#include "core.cuh"
#include "model_host.h"

using namespace esb;

// =============================================================== kernels ====
struct GridArgs {
    DevModel M;
    const double* tab;      // global copy of the staged table
    int tab_doubles;
    const double* k;
    const double* w;
    int nk, nw, layout;
    int n_modes;
    int modes[4];
    double* ext;            // [n_modes][nk][nw]
    double* intq;
    double* den;            // optional [n_modes][nk][nw]: denominator of intq (sweep path; see refine_kernel)
    int schedule;           // 0 = by size, 1 = one thread per point, 2 = one warp per point
    unsigned long long* tile_counter;   // work queue head of the persistent scan kernel (zeroed before launch)
    int threads;            // CTA size of the persistent scan (0: ESB_GRID_THREADS, the most the registers allow)
};

__device__ __forceinline__ double omega_at(const double* __restrict__ k, const double* __restrict__ w,
                                           int layout, int nw, int ik, int iw) {
    if (layout == OMEGA_SHARED) return w[iw];
    if (layout == OMEGA_PHASE_SPEED) return k[ik] * w[iw];
    return w[(size_t)ik * nw + iw];
}

__device__ __forceinline__ void stage_table(const double* __restrict__ g, double* s, int n) {
    // n is a multiple of 2 doubles; 16-byte vector copies
    const double2* g2 = reinterpret_cast<const double2*>(g);
    double2* s2 = reinterpret_cast<double2*>(s);
    for (int i = threadIdx.x; i < n / 2; i += blockDim.x) s2[i] = g2[i];
    if ((n & 1) && threadIdx.x == 0) s[n - 1] = g[n - 1];
    __syncthreads();
}

// NM modes are evaluated per thread, sharing the staged coefficients (cylinder) or the whole
// integration (slab); outputs are mode-slot major.
//
// Persistent: ONE CTA per SM stages the table once; its warps pull 32-point tiles (a row index and 32
// consecutive omega) from a global counter until the grid is exhausted.  No CTA turnover - the
// one-CTA-per-128-points launch re-staged the table 79 000 times per bench launch and left 10 % of
// the warp slots idle between a CTA's first and last warp finishing (ncu: 10.8 of 12 warps active).
// Resident warps per SM = what the register budget allows: 12 (<= 168 registers) for the normal-form
// scheme, whose tableau then stays in uniform registers across the step loop (385 instructions per
// fused step, 290 of them FP64; compiled for 128 registers the constants are re-materialised inside
// the loop: 527 instructions, issue-bound: scan 30.5 / 31.6 / 37.2 ms compiled for 3 / 4 / 2 CTAs of
// 128 threads per SM, scripts/gpu_variants.py); 16 (128 registers) for the first-derivative schemes,
// as measured in round 1.
#ifndef ESB_ROT_THREADS
#define ESB_ROT_THREADS 512
#endif
#ifndef ESB_FLOW_THREADS
#define ESB_FLOW_THREADS 512
#endif
#ifndef ESB_GRID_THREADS
#define ESB_GRID_THREADS(KIND, SCHEME) \
    ((SCHEME) == SCHEME_RK8N ? 384 : (KIND) == KIND_CYL_ROTATION ? ESB_ROT_THREADS : (KIND) == KIND_SLAB_FLOW ? ESB_FLOW_THREADS : 512)
#endif
#ifndef ESB_REFINE_MINB
#define ESB_REFINE_MINB(SCHEME) ((SCHEME) == SCHEME_RK8N ? 3 : 4)
#endif
template <int KIND, int SCHEME, int NM>
__global__ void __launch_bounds__(ESB_GRID_THREADS(KIND, SCHEME), 1) grid_kernel(GridArgs g) {
    extern __shared__ __align__(16) double stab[];
    stage_table(g.tab, stab, g.tab_doubles);
    const int lane = threadIdx.x & 31;
    int modes[NM];
#pragma unroll
    for (int s = 0; s < NM; ++s) modes[s] = g.modes[s];
    const size_t plane = (size_t)g.nk * g.nw;
    const int tiles_per_row = (g.nw + 31) >> 5;
    const unsigned long long n_tiles = (unsigned long long)g.nk * tiles_per_row;
    for (;;) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(g.tile_counter, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tiles) break;
        const int ik = (int)(t / tiles_per_row);
        const int iw = (int)(t - (unsigned long long)ik * tiles_per_row) * 32 + lane;
        if (iw >= g.nw) continue;
        const double k = g.k[ik];
        const double w = omega_at(g.k, g.w, g.layout, g.nw, ik, iw);
        double e[NM], i[NM], d[NM];
        eval_point_multi<KIND, SCHEME, NM>(g.M, stab, k, w, modes, e, i, d);
        const size_t o = (size_t)ik * g.nw + iw;
#pragma unroll
        for (int s = 0; s < NM; ++s) {
            // one class of "no value": skipped (m_e < 0) or overflowed (exterior growth beyond
            // the double range, where the reference's odeint fails too) -> NaN in both grids
            const bool fin = isfinite(e[s]) && isfinite(i[s]);
            g.ext[s * plane + o] = fin ? e[s] : nan("");
            g.intq[s * plane + o] = fin ? i[s] : nan("");
            if (g.den) g.den[s * plane + o] = d[s];
        }
    }
}

// One WARP per (mode, k, omega): the warp-cooperative evaluation (core.cuh warp_transfer) for grids too
// small to fill the lanes - a reference-style worker call is one k and ~90 frequencies - where the
// one-thread-per-point kernel lasts as long as one full integration.
template <int KIND, int SCHEME>
__global__ void __launch_bounds__(128) grid_warp_kernel(GridArgs g) {
    extern __shared__ __align__(16) double stab[];
    stage_table(g.tab, stab, g.tab_doubles);
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const size_t plane = (size_t)g.nk * g.nw;
    const size_t total = plane * g.n_modes;
    for (size_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < total; p += n_warps) {
        const int slot = (int)(p / plane);
        const size_t o = p - slot * plane;
        const int ik = (int)(o / g.nw), iw = (int)(o - (size_t)ik * g.nw);
        const double k = g.k[ik];
        const double w = omega_at(g.k, g.w, g.layout, g.nw, ik, iw);
        double e, i, d;
        eval_point<KIND, SCHEME, true>(g.M, stab, k, w, g.modes[slot], e, i, d);
        if (lane == 0) {
            const bool fin = isfinite(e) && isfinite(i);
            g.ext[p] = fin ? e : nan("");
            g.intq[p] = fin ? i : nan("");
            if (g.den) g.den[p] = d;
        }
    }
}

// Opt-in evaluation over the WHOLE grid, leaky side included (core.cuh eval_point_multi<..., LEAKY>): one thread
// per (mode, k, omega), grid-stride.  Not on the reference path (it skips m_e < 0), hence not tuned: the
// throughput kernels above stay free of the extra exterior branch.
template <int KIND, int SCHEME>
__global__ void __launch_bounds__(128) grid_leaky_kernel(GridArgs g) {
    extern __shared__ __align__(16) double stab[];
    stage_table(g.tab, stab, g.tab_doubles);
    const size_t plane = (size_t)g.nk * g.nw;
    const size_t total = plane * g.n_modes;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const int slot = (int)(p / plane);
        const size_t o = p - slot * plane;
        const int ik = (int)(o / g.nw), iw = (int)(o - (size_t)ik * g.nw);
        const double k = g.k[ik];
        const double w = omega_at(g.k, g.w, g.layout, g.nw, ik, iw);
        double e, i, d;
        eval_point<KIND, SCHEME, false, false, true>(g.M, stab, k, w, g.modes[slot], e, i, d);
        const bool fin = isfinite(e) && isfinite(i);
        g.ext[p] = fin ? e : nan("");
        g.intq[p] = fin ? i : nan("");
    }
}

// ---- brackets ------------------------------------------------------------------------------------
__device__ __forceinline__ double mismatch_pct(double e, double i) {
    // the reference's acceptance quantity (Density_cylinder.py:809)
    return fabs(e - i) * 100.0 / fmax(fabs(e), fabs(i));
}

__device__ __forceinline__ bool is_bracket(double d0, double d1) {
    // both evaluated (finite) and strictly opposite signs
    return isfinite(d0) && isfinite(d1) && ((d0 < 0.0 && d1 > 0.0) || (d0 > 0.0 && d1 < 0.0));
}

// One sweep = up to ESB_MAX_MODES mode slots over the same (k, omega) grid.  The bracket passes and the
// refinement of ALL slots are single launches over one descriptor; the only host round trip of a sweep
// is the read-back of the per-slot counts (esb_sweep_resident_multi).
#define ESB_MAX_MODES 4

struct SlotDev {
    const double* gext;     // planes of the scan: end-point values of every bracket
    const double* gint;
    const double* gden;     // denominator of gint at the same points
    int* bk;                // bracket: row, lower and upper omega index (upper = lower + 1 except
    int* bw;                //   across skipped points in the reference rule; upper == lower: a scan
    int* bw2;               //   point the reference accepts as it stands)
    double* omega;
    double* ext;
    double* intq;
    int* accepted;
    int* iters;
    int mode;
    int capacity;           // entries the slot's buffers hold; brackets beyond it are counted, not stored
};

struct SweepDev {
    SlotDev slot[ESB_MAX_MODES];
    int n_slots, nk, nw, nseg;
    int* seg_count;         // [n_slots * nk * nseg]
    int* seg_offset;        // [n_slots * nk * nseg + 1] exclusive prefix over all slots
    int* slot_begin;        // [ESB_MAX_MODES + 1] first global work index of every slot; [n_slots] = total
    int rule;               // ESB_ACCEPT_CONVERGED: adjacent finite points; ESB_ACCEPT_REFERENCE: the script's rule
    double tol_percent;
    int* seg_tmp;           // [items][BRACKET_SEG]: omega indices found by the count pass, segment-local order
};

// One warp per (slot, row, segment of BRACKET_SEG omega intervals); pass 0 counts, pass 1 fills at the
// segment's offset.  Items are ordered (slot, row, segment), so every slot's list is sorted by
// (k index, omega index) whatever the number of segments.
constexpr int BRACKET_SEG = 1024;

__global__ void bracket_kernel(SweepDev sw, int fill) {
    const int lane = threadIdx.x & 31;
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per_slot = sw.nk * sw.nseg;
    if (item >= sw.n_slots * per_slot) return;
    const int sl = item / per_slot;
    const int row = (item - sl * per_slot) / sw.nseg, seg = item - sl * per_slot - row * sw.nseg;
    const SlotDev& S = sw.slot[sl];
    const double* e = S.gext + (size_t)row * sw.nw;
    const double* q = S.gint + (size_t)row * sw.nw;
    const int base = fill ? sw.seg_offset[item] - sw.slot_begin[sl] : 0;
    const int j_begin = seg * BRACKET_SEG;
    const int j_end = min(j_begin + BRACKET_SEG, sw.nw - 1);
    int count = 0;
    for (int j0 = j_begin; j0 < j_end; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < j_end) {
            const double d0 = e[j] - q[j];
            const double d1 = e[j + 1] - q[j + 1];
            hit = is_bracket(d0, d1);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            const int local = count + __popc(ballot & ((1u << lane) - 1u));
            if (fill) {
                const int pos = base + local;
                if (pos < S.capacity) {
                    S.bk[pos] = row;
                    S.bw[pos] = j;
                    if (S.bw2) S.bw2[pos] = j + 1;
                }
            } else if (sw.seg_tmp) {
                sw.seg_tmp[(size_t)item * BRACKET_SEG + local] = j;      // the fill pass need not re-read the planes
            }
        }
        count += __popc(ballot);
    }
    if (!fill && lane == 0) sw.seg_count[item] = count;
}

// Fill pass from the indices the count pass left in seg_tmp: one warp per (slot, row, segment) copies its
// brackets to their place in the sorted list.  Touches the bracket data only (a few MB), where a second
// bracket_kernel pass re-reads the (ext, int) planes (480 MB on the bench sweep: 0.12 ms).
__global__ void bracket_fill_kernel(SweepDev sw) {
    const int lane = threadIdx.x & 31;
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per_slot = sw.nk * sw.nseg;
    if (item >= sw.n_slots * per_slot) return;
    const int sl = item / per_slot;
    const int row = (item - sl * per_slot) / sw.nseg;
    const SlotDev& S = sw.slot[sl];
    const int base = sw.seg_offset[item] - sw.slot_begin[sl];
    const int n = sw.seg_offset[item + 1] - sw.seg_offset[item];
    const int* src = sw.seg_tmp + (size_t)item * BRACKET_SEG;
    for (int t = lane; t < n; t += 32) {
        const int pos = base + t;
        if (pos < S.capacity) {
            const int j = src[t];
            S.bk[pos] = row;
            S.bw[pos] = j;
            if (S.bw2) S.bw2[pos] = j + 1;
        }
    }
}

// The reference's own scan rule, one warp per (slot, row) (Density_cylinder.py:803-821, the same in
// every script).  Walking the EVALUATED points of a row in order (m_e < 0 points are skipped, :760):
//   * a point whose mismatch is below the tolerance is appended to the solutions as it stands and the
//     list of points seen (`all_ws`) is cleared (:809-812);
//   * else, if D changed sign against the previous evaluated point AND more than two points have been
//     seen since the list was last cleared (:814-815), the pair (previous, this) is bisected and the
//     list is cleared (:820).
// The state (points seen since the last clear) is sequential in omega: the warp classifies 32 points at
// a time by ballot and every lane walks the three masks with bit operations.
__global__ void bracket_reference_kernel(SweepDev sw, int fill) {
    const int lane = threadIdx.x & 31;
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= sw.n_slots * sw.nk) return;
    const int sl = item / sw.nk, row = item - sl * sw.nk;
    const SlotDev& S = sw.slot[sl];
    const double* e = S.gext + (size_t)row * sw.nw;
    const double* q = S.gint + (size_t)row * sw.nw;
    const int base = fill ? sw.seg_offset[item] - sw.slot_begin[sl] : 0;
    // the cylinder scripts bisect after MORE THAN TWO points seen (Density_cylinder.py:815), the slab ones
    // after more than one (..._coronal.py:614)
    const int min_seen = sw.rule == ESB_ACCEPT_REFERENCE_SLAB ? 1 : 2;
    int count = 0, seen = 0, prev_j = -1;
    double prev_d = 0.0;                     // xi_diff_check = [0]: the first product is with 0
    for (int j0 = 0; j0 < sw.nw; j0 += 32) {
        const int j = j0 + lane;
        double d = nan("");
        bool ok = false;
        if (j < sw.nw) {
            const double ev = e[j], qv = q[j];
            d = ev - qv;
            ok = mismatch_pct(ev, qv) < sw.tol_percent;
        }
        const unsigned finite_m = __ballot_sync(0xffffffffu, isfinite(d));
        const unsigned accept_m = __ballot_sync(0xffffffffu, ok);
        const unsigned neg_m = __ballot_sync(0xffffffffu, d < 0.0);
        const unsigned zero_m = __ballot_sync(0xffffffffu, d == 0.0);
        for (int b = 0; b < 32; ++b) {
            if (!((finite_m >> b) & 1u)) continue;
            const int jj = j0 + b;
            const int sgn = ((zero_m >> b) & 1u) ? 0 : ((neg_m >> b) & 1u) ? -1 : 1;
            const int psgn = prev_d == 0.0 ? 0 : prev_d < 0.0 ? -1 : 1;
            ++seen;
            int lo = -1, hi = -1;
            if ((accept_m >> b) & 1u) {
                lo = hi = jj;                                    // a solution as it stands
                seen = 0;
            } else if (sgn * psgn < 0 && seen > min_seen) {
                lo = prev_j; hi = jj;
                seen = 0;
            }
            if (lo >= 0) {
                if (fill && lane == 0) {
                    const int pos = base + count;
                    if (pos < S.capacity) { S.bk[pos] = row; S.bw[pos] = lo; S.bw2[pos] = hi; }
                }
                ++count;
            }
            prev_d = (double)sgn;
            prev_j = jj;
        }
    }
    if (!fill && lane == 0) sw.seg_count[item] = count;
}

// exclusive scan of the item counts -> offsets[n + 1]; single block.  Every `slot_items`-th offset is
// also written to slot_begin[] (the first work index of each mode slot; slot_begin[n / slot_items] = total).
__global__ void scan_kernel(const int* __restrict__ counts, int* __restrict__ offsets, int n, int slot_items,
                            int* __restrict__ slot_begin) {
    __shared__ int carry;
    __shared__ int warp_sums[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int v = (i < n) ? counts[i] : 0;
        int x = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += t;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int s = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += t;
            }
            warp_sums[lane] = s;
        }
        __syncthreads();
        const int prefix = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
        if (i < n) {
            offsets[i] = prefix;
            if (slot_begin && slot_items > 0 && i % slot_items == 0) slot_begin[i / slot_items] = prefix;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        offsets[n] = carry;
        if (slot_begin && slot_items > 0) slot_begin[n / slot_items] = carry;
    }
}

// row_offset[row] = seg_offset[row * nseg]  (row = nk gives the total)
__global__ void row_offset_kernel(const int* __restrict__ seg_offset, int nk, int nseg, int* __restrict__ row_offset) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row <= nk) row_offset[row] = seg_offset[(size_t)row * nseg];
}

// ---- refinement ----
struct RefineArgs {
    DevModel M;
    const double* tab;
    int tab_doubles;
    const double* k;
    const double* w;
    int layout;
    SweepDev sw;
    int* counter;           // work queue heads (zeroed before launch): [0] first pass, [1] second pass
};

// ---- Brent's method on G = D * Y, one bracket -------------------------------------------------
// Pole-free iteration.  D = ext - int with int = N/Y, Y the boundary value of the integrated
// interior solution: D changes sign through infinity wherever Y crosses zero (the reference bisects
// such brackets 150 levels deep and then drops them).  The scan stores Y next to (ext, int), so
//   * a bracket whose end points have Y of opposite sign AND int of opposite sign (the numerator
//     N = int * Y keeps its sign) is a pole of D: it is reported at once (omega = the zero of Y by
//     linear interpolation, ext = int = NaN, unaccepted) - no evaluation;
//   * every other bracket is refined on G = D * Y, which has the roots of D and no such poles, so
//     Brent's interpolation steps are not thrown off by a pole next to the root.
// The acceptance test is the reference's, on (ext, int) at the converged root.  Poles of the
// prefactors (e.g. omega = k U for the flow slab) survive in G: a bracket whose best point keeps a
// mismatch above 50 % while |G| grows or the bracket has shrunk to 1e-7 relative is reported,
// unaccepted, without being bisected to machine precision; so is a jump of G (see below).
struct Brent {
    // a = previous iterate, b = best iterate, c = the other end of the bracket; (e?, i?) = (ext, int)
    // there, f? = G there
    double a, b, c, ea, ia, eb, ib, ec, ic, fa, fb, fc, d, e, f0min, f0max, w0, hw0;
    int it;
    bool on_g;      // iterate on G = D * Y (else on D itself)
    double guess;   // first trial point from the scan grid (NaN: none, start with the secant step)
};

// false: the bracket is a pole of D (S.b = its interpolated position)
__device__ __forceinline__ bool brent_init(Brent& S, double a, double b, double ea, double ia, double eb,
                                           double ib, double ya, double yb) {
    S.on_g = true;
    if ((ya < 0.0 && yb > 0.0) || (ya > 0.0 && yb < 0.0)) {
        if ((ia < 0.0 && ib > 0.0) || (ia > 0.0 && ib < 0.0)) {
            // the denominator changes sign and the numerator N = int * Y does not: int, hence D,
            // changes sign through infinity
            const double wp = a - ya * (b - a) / (yb - ya);
            S.b = fmin(fmax(wp, fmin(a, b)), fmax(a, b));
            return false;
        }
        // numerator and denominator vanish together (slab: for even coefficients y2(1) = 0 implies
        // y1(1) = +-1, which is the sausage or the kink target): int = N/Y stays finite, D is smooth
        // and G would have a spurious root at the zero of Y - iterate on D itself
        S.on_g = false;
    }
    S.a = a; S.b = b; S.ea = ea; S.ia = ia; S.eb = eb; S.ib = ib;
    S.fa = (ea - ia) * (S.on_g ? ya : 1.0);
    S.fb = (eb - ib) * (S.on_g ? yb : 1.0);
    S.c = a; S.ec = ea; S.ic = ia; S.fc = S.fa;
    S.d = b - a; S.e = S.d;
    S.f0min = fmin(fabs(S.fa), fabs(S.fb));
    S.f0max = fmax(fabs(S.fa), fabs(S.fb));
    // scales of the bracket as found: they keep the tolerances meaningful for a bracket
    // that straddles omega = 0 (the backward/forward scans of the flow kinds have one per k)
    S.w0 = fmax(fabs(a), fabs(b));
    S.hw0 = 0.5 * fabs(b - a);
    S.it = 0;
    S.guess = nan("");
    return true;
}

// Inverse cubic interpolation through four neighbouring scan points (w_i, G_i), evaluated at G = 0:
// the first trial point of a bracket whose neighbourhood is smooth (G strictly monotone over the four,
// no sign change of Y).  The scan grid is fine against the scale of G, so this lands within ~1e-9 of
// the bracket width of the root and the iteration needs ~4 evaluations instead of ~8.
__device__ __forceinline__ double inverse_cubic_guess(const double (&w)[4], const double (&g)[4]) {
    const double d01 = g[0] - g[1], d12 = g[1] - g[2], d23 = g[2] - g[3];
    if (!((d01 > 0.0 && d12 > 0.0 && d23 > 0.0) || (d01 < 0.0 && d12 < 0.0 && d23 < 0.0))) return nan("");
    double x = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double l = w[i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j != i) l *= g[j] / (g[j] - g[i]);
        x += l;
    }
    return x;
}

// true: evaluate at S.b and hand the values to brent_feed; false: finished, the result is (S.b, S.eb, S.ib)
__device__ __forceinline__ bool brent_next(Brent& S) {
    const double eps = 2.220446049250313e-16;
    if ((S.fb > 0.0 && S.fc > 0.0) || (S.fb < 0.0 && S.fc < 0.0)) {
        S.c = S.a; S.ec = S.ea; S.ic = S.ia; S.fc = S.fa;
        S.d = S.b - S.a; S.e = S.d;
    }
    if (fabs(S.fc) < fabs(S.fb)) {
        S.a = S.b; S.ea = S.eb; S.ia = S.ib; S.fa = S.fb;
        S.b = S.c; S.eb = S.ec; S.ib = S.ic; S.fb = S.fc;
        S.c = S.a; S.ec = S.ea; S.ic = S.ia; S.fc = S.fa;
    }
    const double tol1 = 2.0 * eps * fmax(fabs(S.b), 0.5 * S.w0);
    const double xm = 0.5 * (S.c - S.b);
    const bool converged = fabs(xm) <= tol1 || S.fb == 0.0 || !isfinite(S.fb) || S.it >= 120;
    const bool pole = S.it >= 2 && mismatch_pct(S.eb, S.ib) > 50.0 &&
                      (fabs(S.fb) > 4.0 * S.f0min || fabs(xm) < fmax(1e-7 * fabs(S.b), 1e-3 * S.hw0));
    // a jump: inside a continuum the integration crosses a singular point and G changes sign
    // discontinuously (the reference bisects such brackets 150 levels deep).  Once the bracket
    // has shrunk a million-fold a root would have |G| ~ 1e-6 of its end-point values; a best
    // value still above 1e-4 of them is a discontinuity.  (Were a genuine root ever caught by
    // this, it is already located to 1e-6 of a grid interval, ~1e-10 relative.)
    const bool jump = fabs(xm) < 1e-6 * S.hw0 && fabs(S.fb) > 100.0 * S.f0max * (fabs(xm) / S.hw0);
    if (converged || pole || jump) return false;
    if (S.it == 0 && isfinite(S.guess) && (S.guess - S.b) * (S.c - S.guess) > 0.0 &&
        fabs(S.guess - S.b) > tol1) {
        // first trial from the scan grid; Brent's bookkeeping takes over from the next step on
        S.d = S.guess - S.b;
        S.e = S.d;
    } else if (fabs(S.e) >= tol1 && fabs(S.fa) > fabs(S.fb)) {
        const double s = S.fb / S.fa;
        double p, q;
        if (S.a == S.c) {
            p = 2.0 * xm * s;
            q = 1.0 - s;
        } else {
            const double qq = S.fa / S.fc, rr = S.fb / S.fc;
            p = s * (2.0 * xm * qq * (qq - rr) - (S.b - S.a) * (rr - 1.0));
            q = (qq - 1.0) * (rr - 1.0) * (s - 1.0);
        }
        if (p > 0.0) q = -q;
        p = fabs(p);
        const double m1 = 3.0 * xm * q - fabs(tol1 * q);
        const double m2 = fabs(S.e * q);
        if (2.0 * p < (m1 < m2 ? m1 : m2)) {
            S.e = S.d;
            S.d = p / q;
        } else {
            S.d = xm;
            S.e = S.d;
        }
    } else {
        S.d = xm;
        S.e = S.d;
    }
    S.a = S.b; S.ea = S.eb; S.ia = S.ib; S.fa = S.fb;
    S.b += (fabs(S.d) > tol1) ? S.d : (xm > 0.0 ? tol1 : -tol1);
    return true;
}

__device__ __forceinline__ void brent_feed(Brent& S, double en, double in_, double yn) {
    S.eb = en;
    S.ib = in_;
    S.fb = (en - in_) * (S.on_g ? yn : 1.0);
    ++S.it;
}

// ---- the reference's own bisection (accept rule ESB_ACCEPT_REFERENCE) ---------------------------
// locate_kink (Density_cylinder.py:548-686) evaluates omega = linspace(lo, hi, 3) in order; a point whose
// mismatch is below the tolerance is THE solution (:671-676, no further refinement); otherwise, after the
// third point, a sign change between the middle and the upper point recurses into (middle, upper)
// (:678-684, the list of points seen has three entries only then), up to 150 levels (:558).  A root in
// the LOWER half is therefore not followed - the scripts' tables hold what this rule finds, so the
// drop-in reproduces it.  lo and hi were evaluated by the caller and are outside the band: one new
// evaluation (the middle) per level.
struct RefBisect {
    double a, b, da, db;    // bracket, D at its ends
    bool both_halves;       // the slab scripts test the lower half too (len(loop_ws) > 1, ..._coronal.py:510)
    double mid, em, im;     // last trial and (ext, int) there
    int level;
    bool found;
};

// true: evaluate at S.mid and hand the values to refbisect_feed; false: finished
__device__ __forceinline__ bool refbisect_next(RefBisect& S) {
    if (S.found || S.level > 150 || S.level < 0) return false;
    S.mid = S.a + (S.b - S.a) * 0.5;        // numpy.linspace(a, b, 3)[1]
    return true;
}

__device__ __forceinline__ void refbisect_feed(RefBisect& S, double en, double in_, double tol_percent) {
    S.em = en; S.im = in_;
    const double dm = en - in_;
    if (mismatch_pct(en, in_) < tol_percent) {
        S.found = true;
    } else if (S.both_halves && isfinite(dm) && dm * S.da < 0.0) {
        S.b = S.mid; S.db = dm;
        ++S.level;
    } else if (isfinite(dm) && dm * S.db < 0.0) {
        S.a = S.mid; S.da = dm;
        ++S.level;
    } else {
        S.level = -1 - S.level;             // not followed: the reference stops here without a solution
    }
}

// global work index -> slot and local index (slot_begin[] is ascending)
__device__ __forceinline__ int slot_of(const SweepDev& sw, int tq) {
    int sl = 0;
#pragma unroll
    for (int s = 1; s < ESB_MAX_MODES; ++s)
        if (s < sw.n_slots && tq >= sw.slot_begin[s]) sl = s;
    return sl;
}

struct Pick {
    int sl, t, mode;
    double k;
    bool reference;         // iterate with RefBisect, else Brent
};

enum { PICK_SKIP = 0, PICK_WORK = 1 };

// Bracket tq of the single queue over all mode slots -> slot, local index, point data, iteration state.
// pass 0 takes the brackets without a first trial from the scan grid (inside the continua: up to ~20
// evaluations each), pass 1 the smooth ones (3-4 evaluations): the long ones start first and the queue
// drains on short ones.  Entries finished without an evaluation (poles classified from the scan, scan
// points the reference accepts as they stand, overflow beyond the slot capacity) are written in pass 0.
__device__ __forceinline__ int refine_pickup(const RefineArgs& r, int tq, int pass, bool writer, Pick& P, Brent& S,
                                             RefBisect& B) {
    const SweepDev& sw = r.sw;
    P.sl = slot_of(sw, tq);
    const SlotDev& L = sw.slot[P.sl];
    P.t = tq - sw.slot_begin[P.sl];
    if (P.t >= L.capacity) return PICK_SKIP;
    P.mode = L.mode;
    P.reference = sw.rule != ESB_ACCEPT_CONVERGED;
    const int ik = L.bk[P.t], jlo = L.bw[P.t], jhi = L.bw2[P.t];
    P.k = r.k[ik];
    const size_t o = (size_t)ik * sw.nw + jlo, o2 = (size_t)ik * sw.nw + jhi;
    const double a = omega_at(r.k, r.w, r.layout, sw.nw, ik, jlo);
    if (jhi == jlo) {                       // reference rule: a scan point inside the band
        if (pass == 0 && writer) {
            L.omega[P.t] = a; L.ext[P.t] = L.gext[o]; L.intq[P.t] = L.gint[o];
            L.iters[P.t] = 0; L.accepted[P.t] = 1;
        }
        return PICK_SKIP;
    }
    const double b = omega_at(r.k, r.w, r.layout, sw.nw, ik, jhi);
    if (P.reference) {
        if (pass != 0) return PICK_SKIP;
        B.a = a; B.b = b; B.da = L.gext[o] - L.gint[o]; B.db = L.gext[o2] - L.gint[o2];
        B.both_halves = sw.rule == ESB_ACCEPT_REFERENCE_SLAB;
        B.mid = b; B.em = L.gext[o2]; B.im = L.gint[o2];
        B.level = 0; B.found = false;
        return PICK_WORK;
    }
    if (!brent_init(S, a, b, L.gext[o], L.gint[o], L.gext[o2], L.gint[o2], L.gden[o], L.gden[o2])) {
        if (pass == 0 && writer) {          // a pole of D, classified from the scan
            L.omega[P.t] = S.b; L.ext[P.t] = nan(""); L.intq[P.t] = nan("");
            L.iters[P.t] = 0; L.accepted[P.t] = 0;
        }
        return PICK_SKIP;
    }
    if (S.on_g && jhi == jlo + 1 && jlo >= 1 && jlo + 2 < sw.nw) {
        double w4[4], g4[4];
        bool smooth = true;
        const double y_ref = L.gden[o];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const size_t oq = o + q - 1;
            const double y = L.gden[oq];
            w4[q] = omega_at(r.k, r.w, r.layout, sw.nw, ik, jlo + q - 1);
            g4[q] = (L.gext[oq] - L.gint[oq]) * y;
            smooth = smooth && isfinite(g4[q]) && ((y > 0.0) == (y_ref > 0.0));
        }
        if (smooth) S.guess = inverse_cubic_guess(w4, g4);
    }
    return (isfinite(S.guess) ? 1 : 0) == pass ? PICK_WORK : PICK_SKIP;
}

__device__ __forceinline__ void refine_store(const RefineArgs& r, const Pick& P, const Brent& S, const RefBisect& B) {
    const SlotDev& L = r.sw.slot[P.sl];
    if (P.reference) {
        L.omega[P.t] = B.mid; L.ext[P.t] = B.em; L.intq[P.t] = B.im;
        L.iters[P.t] = B.level < 0 ? -B.level : B.level + 1;          // evaluations used
        L.accepted[P.t] = B.found ? 1 : 0;
        return;
    }
    L.omega[P.t] = S.b; L.ext[P.t] = S.eb; L.intq[P.t] = S.ib;
    L.iters[P.t] = S.it;
    L.accepted[P.t] = (mismatch_pct(S.eb, S.ib) < r.sw.tol_percent) ? 1 : 0;
}

// Persistent kernel, one LANE per bracket: every lane iterates on one bracket at a time and pulls the
// next bracket from a global queue as soon as its own has converged, so a warp never idles behind
// its slowest lane.  The D evaluation (the expensive part) is executed by all 32 lanes together each
// round; __any_sync on the "trial pending" flags ends the loop.  End-point values come from the scan
// grid, not from new evaluations.  The brackets are counted on the device (slot_begin[n_slots]).
template <int KIND, int SCHEME, int MINB>
__global__ void __launch_bounds__(128, MINB) refine_kernel(RefineArgs r) {
    extern __shared__ __align__(16) double stab[];
    stage_table(r.tab, stab, r.tab_doubles);
    const int n_total = r.sw.slot_begin[r.sw.n_slots];
    bool have = false, pending = false, exhausted = false;
    int pass = 0;
    Pick P;
    P.k = 1.0; P.mode = 0; P.reference = false;
    Brent S;
    RefBisect B;
    double wq = 1.0;        // trial frequency
    for (;;) {
        while (!pending && !exhausted) {
            if (!have) {
                const int tq = atomicAdd(r.counter + pass, 1);
                if (tq >= n_total) {
                    if (pass == 0) { pass = 1; continue; }
                    exhausted = true;
                    break;
                }
                if (refine_pickup(r, tq, pass, true, P, S, B) == PICK_SKIP) continue;
                have = true;
            }
            const bool more = P.reference ? refbisect_next(B) : brent_next(S);
            if (!more) {
                refine_store(r, P, S, B);
                have = false;
                continue;
            }
            wq = P.reference ? B.mid : S.b;
            pending = true;
        }
        // (Handing the last few iterating lanes of a warp to the warp-cooperative evaluation once the
        // queue is drained was tried: 3.97 -> 3.88 ms on the bench sweep, and the tables are then no longer
        // bit-reproducible - which lanes switch depends on the order the queue was served in.  Not kept.)
        if (!__any_sync(0xffffffffu, pending)) break;
        double en, in_, yn;
        eval_point<KIND, SCHEME, false, true>(r.M, stab, P.k, wq, P.mode, en, in_, yn);
        if (pending) {
            if (P.reference) refbisect_feed(B, en, in_, r.sw.tol_percent);
            else brent_feed(S, en, in_, yn);
            pending = false;
        }
    }
}

// Persistent kernel, one WARP per bracket: the 32 lanes integrate the layer cooperatively
// (core.cuh warp_transfer), so one evaluation has 1/32 of the latency.  Used when there are fewer
// brackets than lanes to fill, where the lane-per-bracket kernel lasts as long as the sequential
// evaluations of its slowest bracket.  The iteration state is replicated (uniform) across the warp.
template <int KIND, int SCHEME>
__global__ void __launch_bounds__(128) refine_warp_kernel(RefineArgs r) {
    extern __shared__ __align__(16) double stab[];
    stage_table(r.tab, stab, r.tab_doubles);
    const int lane = threadIdx.x & 31;
    const int n_total = r.sw.slot_begin[r.sw.n_slots];
    for (int pass = 0; pass < 2; ++pass) {
        for (;;) {
            int tq = 0;
            if (lane == 0) tq = atomicAdd(r.counter + pass, 1);
            tq = __shfl_sync(0xffffffffu, tq, 0);
            if (tq >= n_total) break;
            Pick P;
            Brent S;
            RefBisect B;
            if (refine_pickup(r, tq, pass, lane == 0, P, S, B) == PICK_SKIP) continue;
            for (;;) {
                const bool more = P.reference ? refbisect_next(B) : brent_next(S);
                if (!more) break;
                double en, in_, yn;
                eval_point<KIND, SCHEME, true>(r.M, stab, P.k, P.reference ? B.mid : S.b, P.mode, en, in_, yn);
                if (P.reference) refbisect_feed(B, en, in_, r.sw.tol_percent);
                else brent_feed(S, en, in_, yn);
            }
            if (lane == 0) refine_store(r, P, S, B);
        }
    }
}

// ---- J_n / Y_n on the device (the leaky side, bessel_jy.cuh) ------------------------------------------
// out[4 i .. 4 i + 3] = {J_n, J_n', Y_n, Y_n'}(x[i])
__global__ void bessel_jy_kernel(int n, const double* __restrict__ x, int count, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    BesselJY b;
    bessel_jy(n, x[i], b);
    double J, dJ, Y, dY;
    bessel_jy_order(b, n, x[i], J, dJ, Y, dY);
    out[4 * i] = J; out[4 * i + 1] = dJ; out[4 * i + 2] = Y; out[4 * i + 3] = dY;
}

// out[2 i], out[2 i + 1] = (P, dP/dr) at |r| = 1 of the leaky exterior at (k[i], w[i]); NaN where m_e >= 0
__global__ void exterior_leaky_kernel(DevModel M, int n, const double* __restrict__ k, const double* __restrict__ w,
                                      int count, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double me = m_e2(M, k[i] * k[i], w[i] * w[i]);
    double yb = nan(""), ypb = nan("");
    if (me < 0.0) exterior_cyl_leaky(M.ic_v, M.ic_s, M.r_sign, M.ext_len, k[i], me, n, yb, ypb);
    out[2 * i] = yb; out[2 * i + 1] = ypb;
}

// ---- accepted modes of all slots packed into one payload (multi-GPU gather) ---------------------------
// rows 1.. of `out` = (global k row, omega, slot) of every accepted mode, slots in order, each slot in table
// order (k index, omega index): deterministic.  Row 0 = (count, entries scanned, capacity exceeded ? 1 : 0).
// Two launches over the same block partition of the concatenated tables: count, then prefix + write.
struct PackArgs {
    const int* accepted[ESB_MAX_MODES];
    const int* k_index[ESB_MAX_MODES];
    const double* omega[ESB_MAX_MODES];
    int begin[ESB_MAX_MODES + 1];      // first concatenated index of every slot; [n_slots] = total
    int n_slots;
    double k_offset, k_stride;
    int* block_count;                  // [gridDim.x]
    double* out;                       // [capacity + 1][3]
    int capacity;
};

constexpr int PACK_THREADS = 256, PACK_PER_BLOCK = 2048;

__global__ void __launch_bounds__(PACK_THREADS) pack_modes_kernel(PackArgs a, int write) {
    __shared__ int warp_sum[PACK_THREADS / 32];
    __shared__ int base_sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = a.begin[a.n_slots];
    const int lo = blockIdx.x * PACK_PER_BLOCK;
    const int hi = lo + PACK_PER_BLOCK < total ? lo + PACK_PER_BLOCK : total;
    int base = 0;
    if (write) {
        if (threadIdx.x == 0) {
            int b = 0, all = 0;
            for (int i = 0; i < (int)gridDim.x; ++i) {
                if (i == (int)blockIdx.x) b = all;
                all += a.block_count[i];
            }
            base_sh = b;
            if (blockIdx.x == 0) {
                a.out[0] = (double)(all < a.capacity ? all : a.capacity);
                a.out[1] = (double)total;
                a.out[2] = all > a.capacity ? 1.0 : 0.0;
            }
        }
        __syncthreads();
        base = base_sh;
    }
    int running = 0;
    for (int c0 = lo; c0 < hi; c0 += PACK_THREADS) {
        const int g = c0 + threadIdx.x;
        int sl = 0, acc = 0;
        if (g < hi) {
#pragma unroll
            for (int q = 1; q < ESB_MAX_MODES; ++q)
                if (q < a.n_slots && g >= a.begin[q]) sl = q;
            acc = a.accepted[sl][g - a.begin[sl]] == 1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if (lane == 0) warp_sum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, chunk = 0;
#pragma unroll
        for (int q = 0; q < PACK_THREADS / 32; ++q) {
            before += q < warp ? warp_sum[q] : 0;
            chunk += warp_sum[q];
        }
        if (write && acc) {
            const int row = base + running + before + __popc(bal & ((1u << lane) - 1u));
            if (row < a.capacity) {
                const int t = g - a.begin[sl];
                double* o = a.out + 3 * (size_t)(row + 1);
                o[0] = fma((double)a.k_index[sl][t], a.k_stride, a.k_offset);
                o[1] = a.omega[sl][t];
                o[2] = (double)sl;
            }
        }
        running += chunk;
        __syncthreads();
    }
    if (!write && threadIdx.x == 0) a.block_count[blockIdx.x] = running;
}

// ---- discretisation guard --------------------------------------------------------------------------
// The integrator is fixed-step where the reference's odeint adapts.  With a guard model set
// (esb_set_guard_fields: the same equilibrium on a finer mesh, normally 2 x n_steps), every sweep
// re-evaluates every stride-th (point, mode) of its scan with the fine table and keeps the worst
// deviation.  8th order: doubling the steps divides the error by ~256, so the deviation IS the
// discretisation error of the sweep at that point.  Judged on g = D Y / (|ext Y| + |int Y|) (Y = the
// denominator of int, stored by the scan; core.cuh guard_deviation): next to a pole of D (Y -> 0) int = N/Y
// amplifies the error of Y without bound while g stays regular, and a common factor of N and Y - which
// cancels in D - cancels in g too.  Points inside the
// resonant continua (resonance_free) are not judged: no step count converges there.
// One lane per sample, one 32-sample tile per warp; runs on a side stream next to the bracket passes.
struct GuardRecord {
    unsigned long long key;        // high 32 bits of the worst deviation (a positive double) | sample id
    unsigned long long n_checked;
    unsigned long long n_above;    // samples above the threshold
};

struct GuardArgs {
    DevModel M;                    // the FINE model
    const double* tab;
    int tab_doubles;
    const double* k;
    const double* w;
    int nk, nw, layout;
    int n_modes;
    int modes[4];
    const double* ext;             // the planes of the sweep being judged
    const double* intq;
    const double* den;
    int stride;
    unsigned n_per_mode;           // samples per mode slot
    double threshold, margin;
    GuardRecord* out;
};

// sample j -> grid point: tiles of 32 CONSECUTIVE omega points (one warp: coalesced reads of the planes and
// lanes that take the same branch of the per-point scheme switch), one tile per 32 * stride points, its
// offset sliding through that window so that a row length that is a multiple of it does not put every tile
// into the same omega columns
__host__ __device__ __forceinline__ size_t guard_point(unsigned j, int stride) {
    const unsigned tile = j >> 5, lane = j & 31u;
    return (size_t)tile * 32u * (unsigned)stride + 32u * (tile % (unsigned)stride) + lane;
}

template <int KIND, int SCHEME>
__global__ void __launch_bounds__(256) guard_kernel(GuardArgs g) {
    extern __shared__ __align__(16) double stab[];
    stage_table(g.tab, stab, g.tab_doubles);
    const int lane = threadIdx.x & 31;
    const size_t plane = (size_t)g.nk * g.nw;
    const unsigned sample = (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long key = 0ULL;
    unsigned checked = 0, above = 0;
    if (sample < g.n_per_mode * (unsigned)g.n_modes) {
        const int slot = (int)(sample / g.n_per_mode);
        const unsigned j = sample - (unsigned)slot * g.n_per_mode;
        const size_t p = guard_point(j, g.stride);
        if (p < plane) {
            const int ik = (int)(p / g.nw), iw = (int)(p - (size_t)ik * g.nw);
            const double e0 = g.ext[slot * plane + p], i0 = g.intq[slot * plane + p], d0 = g.den[slot * plane + p];
            if (isfinite(e0) && isfinite(i0) && isfinite(d0)) {
                const double k = g.k[ik];
                const double w = omega_at(g.k, g.w, g.layout, g.nw, ik, iw);
                const Point pt = make_point(g.M, k, w);
                if (resonance_free<KIND>(g.M, pt, double(g.modes[slot]), stab, g.margin)) {
                    double e, i, d;
                    eval_point<KIND, SCHEME>(g.M, stab, k, w, g.modes[slot], e, i, d);
                    const double dev = guard_deviation(e0, i0, d0, e, i, d);
                    if (isfinite(dev)) {
                        checked = 1;
                        above = dev > g.threshold ? 1 : 0;
                        key = ((unsigned long long)__double_as_longlong(dev) & 0xffffffff00000000ULL) | sample;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xffffffffu, key, off);
        key = o > key ? o : key;
        checked += __shfl_down_sync(0xffffffffu, checked, off);
        above += __shfl_down_sync(0xffffffffu, above, off);
    }
    if (lane == 0 && checked) {
        atomicMax(&g.out->key, key);
        atomicAdd(&g.out->n_checked, (unsigned long long)checked);
        if (above) atomicAdd(&g.out->n_above, (unsigned long long)above);
    }
}

// ---- parameter scans: tables of many equilibria compacted into one ------------------------------
struct ScanOut {
    int *model, *slot, *k_index, *w_index, *accepted, *iters;
    double *omega, *ext, *intq;
};

// counts[t] = entries stored for table t = (model, slot): min(brackets found, capacity)
__global__ void scan_counts_kernel(const int* __restrict__ slot_begin_all, int n_models, int n_modes, int cap,
                                   int* __restrict__ counts, int* __restrict__ found) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_models * n_modes) return;
    const int i = t / n_modes, m = t - i * n_modes;
    const int* sb = slot_begin_all + (size_t)i * (ESB_MAX_MODES + 1);
    const int n = sb[m + 1] - sb[m];
    found[t] = n;
    counts[t] = n < cap ? n : cap;
}

// one CTA per table: its entries to their place in the compact table
__global__ void scan_compact_kernel(const char* __restrict__ base, size_t table_bytes, int cap, int n_modes,
                                    const int* __restrict__ counts, const int* __restrict__ offsets, ScanOut out) {
    const int t = blockIdx.x;
    const int n = counts[t], o = offsets[t];
    const double* d = reinterpret_cast<const double*>(base + (size_t)t * table_bytes);
    const int* q = reinterpret_cast<const int*>(d + 3 * (size_t)cap);
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        out.omega[o + j] = d[j];
        out.ext[o + j] = d[cap + j];
        out.intq[o + j] = d[2 * (size_t)cap + j];
        out.k_index[o + j] = q[j];
        out.w_index[o + j] = q[cap + j];
        out.accepted[o + j] = q[2 * (size_t)cap + j];
        out.iters[o + j] = q[3 * (size_t)cap + j];
        out.model[o + j] = t / n_modes;
        out.slot[o + j] = t - (t / n_modes) * n_modes;
    }
}

// ============================================================ host side ====

struct esb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool model_set = false;
    esb_model model{};
    DevModel dm{};
    double* d_tab = nullptr;
    size_t cap_tab = 0;
    int tab_doubles = 0;
    // scratch for the host-pointer entry points
    double *d_k = nullptr, *d_w = nullptr, *d_ext = nullptr, *d_int = nullptr, *d_den = nullptr;
    size_t cap_k = 0, cap_w = 0, cap_grid = 0, cap_den = 0;
    int *d_rowcount = nullptr, *d_rowoff = nullptr, *d_seg_tmp = nullptr;
    size_t cap_rows = 0, cap_rowoff = 0, cap_seg_tmp = 0;
    // one packed device allocation per slot: [om | e | i] doubles then [bk | bw | acc | it] ints, each
    // `cap` long, so that the whole table moves in ONE copy; `pin` = page-locked host mirror
    struct RootBuf {
        int *bk = nullptr, *bw = nullptr, *acc = nullptr, *it = nullptr, *bw2 = nullptr;
        double *om = nullptr, *e = nullptr, *i = nullptr;
        void* base = nullptr;
        size_t cap = 0;
        int n = 0;
        void* pin = nullptr;
        size_t pin_cap = 0;
    } slots[ESB_MAX_MODES];
    // queue heads: ints [0], [1] = the two passes of the refinement, bytes 8..15 = the scan's tile counter;
    // d_slot_begin[ESB_MAX_MODES + 1] = first work index of every slot (written by scan_kernel), h_counts =
    // its page-locked host copy
    int* d_pack_counts = nullptr;
    size_t cap_pack_counts = 0;
    // esb_scan_models pipeline: the refinement of equilibrium i runs on its own stream beside the scan of
    // equilibrium i + 1 (second set of planes, scan CTAs two thirds of their usual size)
    cudaStream_t refine_stream = nullptr;
    cudaEvent_t ev_fill[2] = {nullptr, nullptr}, ev_refine[2] = {nullptr, nullptr};
    double *d_ext2 = nullptr, *d_int2 = nullptr, *d_den2 = nullptr;
    size_t cap_grid2 = 0;
    int scan_threads = 0;          // CTA size override of the persistent scan kernel (0: default)
    int* d_counter = nullptr;
    int* d_slot_begin = nullptr;
    int* h_counts = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_counts = nullptr, ev_done = nullptr;
    // parameter scans (esb_scan_models): tables of all equilibria, their root tables, the compact result
    double* d_bank = nullptr;
    size_t cap_bank = 0;
    char* d_scan_roots = nullptr;
    size_t cap_scan_roots = 0;
    int* d_scan_ints = nullptr;    // [n_models][ESB_MAX_MODES + 1] slot_begin | counts | found | offsets
    size_t cap_scan_ints = 0;
    int* h_scan_ints = nullptr;    // page-locked: found[n_tables], counts[n_tables], total
    size_t cap_h_scan_ints = 0;
    char* d_compact = nullptr;
    size_t cap_compact = 0;
    char* h_compact = nullptr;     // page-locked mirror handed to the caller
    esb_scan_result scan_dev{};    // the same table, device pointers
    size_t cap_h_compact = 0;
    // discretisation guard (esb_set_guard_fields): fine model + table, side stream, device / pinned record
    bool guard_set = false;
    DevModel g_dm{};
    double* d_gtab = nullptr;
    size_t cap_gtab = 0;
    int g_tab_doubles = 0;
    int guard_stride = 0;
    double guard_threshold = 1e-9;
    cudaStream_t guard_stream = nullptr;
    cudaEvent_t ev_scan = nullptr, ev_guard = nullptr;
    GuardRecord* d_guard = nullptr;
    GuardRecord* h_guard = nullptr;
    bool guard_pending = false;    // a record of the last sweep is (being) written
    int guard_n_modes = 0, guard_nw = 0, guard_last_stride = 0;
    unsigned guard_n_per_mode = 0;
    bool tables_pending = false;   // a sweep's refinement may still be running: consumers wait on ev_done
    int n_sm = 148;
    int accept_rule = ESB_ACCEPT_CONVERGED;
    bool timed = false;
    int64_t launches = 0;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false;
    int schedule = 0;          // 0 = by size, 1 = one lane per point / bracket, 2 = one warp per point / bracket
    int ax_nk = 0, ax_nw = 0, ax_layout = 0;
    std::string err;
};

#define CUDA_TRY(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e);             \
            return ESB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

static int fail(esb_context* c, int code, const char* msg) {
    if (c) c->err = msg;
    return code;
}

template <class T>
static int ensure(esb_context* c, T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return ESB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    CUDA_TRY(c, cudaMalloc((void**)&p, need * sizeof(T)));
    cap = need;
    return ESB_OK;
}

extern "C" int esb_version(void) { return ESB_VERSION; }
extern "C" int esb_sizeof_model(void) { return (int)sizeof(esb_model); }

extern "C" int esb_model_defaults(int32_t kind, esb_model* out) {
    if (!out) return ESB_ERR_ARG;
    memset(out, 0, sizeof(*out));
    out->kind = kind;
    out->scheme = ESB_RK8;
    out->gamma = 5.0 / 3.0;
    out->rho_i0 = 1.0;
    out->rho_A = 1.0;
    out->c_i0 = 1.0;
    out->ext_wavelengths = 3.0;
    out->ext_ic_value = 1e-8;
    out->s_start = -1.0;
    out->r_sign = -1;
    // graded mesh (mesh = 2): measured on the B200 (scripts/gpu_mesh.py) - with these
    // parameters 128 steps reach the accuracy of 256 sin^2-clustered steps; 144 is the default
    out->mesh_axis = 0.16; out->mesh_edge = 0.02; out->mesh_edge_width = 0.10;
    // Default scheme: ESB_RK8N where the kind has a normal form.  Discretisation measured against
    // 1024-step solutions on 40 k x 960 phase speeds (host build of the same code): with
    // 152 steps and mesh_axis 0.13 the normal-form scheme is at least as accurate as ESB_RK8 with its
    // own tuned 144 / 0.16 (the Python host restores those when "rk8" is asked for).
    if (kind == ESB_CYLINDER_DENSITY) {          // Density_cylinder.py:69-72,120,768
        out->vA_i0 = 2.0; out->vA_e = 5.0; out->c_e = 0.5;
        out->ext_ic_slope = 1e-15;
        out->s_end = -0.001;
        out->mesh = 2;
        out->scheme = ESB_RK8N;
        out->mesh_axis = 0.13;
        out->n_steps = 152;
    } else if (kind == ESB_SLAB_DENSITY) {       // ..._coronal.py:69-72,91,247
        out->vA_i0 = 1.2; out->vA_e = 3.0; out->c_e = 0.4;
        out->ext_ic_slope = 1e-8;
        out->s_end = 1.0;
        out->mesh = 2;            // no refinement: a uniform mesh; 176 steps ~ 256 sin^2-clustered ones
        out->mesh_axis = 0.0; out->mesh_edge = 0.0; out->mesh_edge_width = 0.0;
        out->scheme = ESB_RK8N;
        out->n_steps = 176;
    } else if (kind == ESB_SLAB_FLOW) {          // flow_multiprocessor_coronal.py:47-52,72,229
        out->vA_i0 = 1.0; out->c_i0 = 0.3; out->vA_e = 2.5; out->c_e = 0.2;
        out->U_e = 0.0;
        out->ext_ic_slope = 1e-15;
        out->s_end = 1.0;
        // c_i = 0.3 vA_i: shorter interior wavelengths than the density slabs; mild refinement at the
        // boundaries: 320 steps are 2.7x more accurate than 384 sin^2-clustered ones
        out->mesh = 2;
        out->mesh_axis = 0.0; out->mesh_edge = 0.04; out->mesh_edge_width = 0.10;
        out->n_steps = 320;
    } else if (kind == ESB_CYLINDER_ROTATION) {  // Twisted_photospheric_nonlinear_flow_kink_fast.py:73-76,96,302
        out->vA_i0 = 2.0; out->vA_e = 0.5; out->c_e = 1.5;
        out->ext_ic_slope = 1e-8;
        out->r_sign = 1;
        out->s_start = 1.0;
        out->s_end = 0.001;
        out->mesh = 2;            // geometric towards the axis, no boundary refinement
        out->mesh_axis = 0.16; out->mesh_edge = 0.0; out->mesh_edge_width = 0.0;
        // 96 steps: worst deviation from the C oracle 1.4e-10 (99.9 % of the points < 1e-11) over the laws
        // 0.25 r^0.8 / 0.15 r^1.25 / 0.1 r, layer ends 0.001 and 0.01, n = 0..3 (host build of this code,
        // 16 k x 240 phase speeds; 128 steps: 5e-13, 80: 6e-10)
        out->n_steps = 96;
    } else if (kind == ESB_CYLINDER_FLOW) {      // Cylinder_method_flow_testing.py:66-69,120,774
        out->vA_i0 = 2.0; out->vA_e = 5.0; out->c_e = 0.5;
        out->ext_ic_slope = 1e-8;
        out->s_end = -0.001;
        out->mesh = 2;
        out->scheme = ESB_RK8N;
        out->mesh_axis = 0.13;
        out->n_steps = 152;
    } else {
        return ESB_ERR_ARG;
    }
    return ESB_OK;
}

extern "C" int esb_mesh_size(const esb_model* m, int32_t* n_nodes) {
    if (check_model(m) || !n_nodes) return ESB_ERR_ARG;
    *n_nodes = mesh_size(m);
    return ESB_OK;
}

extern "C" int esb_mesh_nodes(const esb_model* m, double* nodes) {
    if (check_model(m) || !nodes) return ESB_ERR_ARG;
    return mesh_nodes(m, nodes);
}

extern "C" int esb_model_n_fields(const esb_model* m, int32_t* n_fields) {
    if (check_model(m) || !n_fields) return ESB_ERR_ARG;
    *n_fields = model_n_fields(m);
    return ESB_OK;
}

extern "C" int esb_model_max_steps(const esb_model* m, int32_t* max_steps) {
    if (!m || !max_steps) return ESB_ERR_ARG;
    esb_model t = *m;
    t.n_steps = 8;                              // any valid count: the answer depends on (kind, scheme) only
    if (check_model(&t)) return ESB_ERR_ARG;
    *max_steps = model_max_steps(&t);
    return ESB_OK;
}

extern "C" int esb_create(int32_t device, esb_context** out) {
    if (!out) return ESB_ERR_ARG;
    *out = nullptr;
    esb_context* c = new (std::nothrow) esb_context();
    if (!c) return ESB_ERR_ALLOC;
    c->device = device;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device >= n) {
        // no CPU fallback: the context cannot exist without a device
        fprintf(stderr, "eigensolver_b200: no usable CUDA device (%s)\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        delete c;
        return ESB_ERR_CUDA;
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&c->stream) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
        cudaMalloc((void**)&c->d_counter, 16) != cudaSuccess ||
        cudaMalloc((void**)&c->d_slot_begin, (ESB_MAX_MODES + 1) * sizeof(int)) != cudaSuccess ||
        cudaHostAlloc((void**)&c->h_counts, (ESB_MAX_MODES + 1) * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        esb_destroy(c);
        return ESB_ERR_CUDA;
    }
    *out = c;
    return ESB_OK;
}

extern "C" int esb_destroy(esb_context* c) {
    if (!c) return ESB_OK;
    cudaSetDevice(c->device);
    void* ptrs[] = {c->d_tab, c->d_k, c->d_w, c->d_ext, c->d_int, c->d_den, c->d_rowcount, c->d_rowoff, c->d_counter,
                    c->d_slot_begin, c->d_bank, c->d_scan_roots, c->d_scan_ints, c->d_compact};
    if (c->h_scan_ints) cudaFreeHost(c->h_scan_ints);
    if (c->h_compact) cudaFreeHost(c->h_compact);
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_counts) cudaFreeHost(c->h_counts);
    if (c->d_gtab) cudaFree(c->d_gtab);
    if (c->d_pack_counts) cudaFree(c->d_pack_counts);
    if (c->d_ext2) cudaFree(c->d_ext2);
    if (c->d_int2) cudaFree(c->d_int2);
    if (c->d_den2) cudaFree(c->d_den2);
    for (int b = 0; b < 2; ++b) {
        if (c->ev_fill[b]) cudaEventDestroy(c->ev_fill[b]);
        if (c->ev_refine[b]) cudaEventDestroy(c->ev_refine[b]);
    }
    if (c->refine_stream) cudaStreamDestroy(c->refine_stream);
    if (c->d_seg_tmp) cudaFree(c->d_seg_tmp);
    if (c->d_guard) cudaFree(c->d_guard);
    if (c->h_guard) cudaFreeHost(c->h_guard);
    if (c->ev_scan) cudaEventDestroy(c->ev_scan);
    if (c->ev_guard) cudaEventDestroy(c->ev_guard);
    if (c->guard_stream) cudaStreamDestroy(c->guard_stream);
    if (c->ev_counts) cudaEventDestroy(c->ev_counts);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    for (auto& sl : c->slots) {
        if (sl.base) cudaFree(sl.base);
        if (sl.pin) cudaFreeHost(sl.pin);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return ESB_OK;
}

extern "C" const char* esb_last_error(const esb_context* c) { return c ? c->err.c_str() : "null context"; }
extern "C" int64_t esb_launch_count(const esb_context* c) { return c ? c->launches : 0; }

extern "C" double esb_last_kernel_ms(const esb_context* c) {
    if (!c || !c->timed) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0;
    return ms;
}

// The model is built completely on the host (model_host.h) and committed to the context only after
// the upload has succeeded: a failed call leaves the previous model intact.
extern "C" int esb_set_model_fields(esb_context* c, const esb_model* m, const double* const* fields,
                                    int32_t n_fields, int32_t n_nodes, const double* boundary,
                                    int32_t n_boundary) {
    if (!c) return ESB_ERR_ARG;
    HostModel hm;
    std::string err;
    int rc = build_host_model(m, fields, n_fields, n_nodes, boundary, n_boundary, hm, err);
    if (rc) return fail(c, rc, err.c_str());
    CUDA_TRY(c, cudaSetDevice(c->device));
    // a parameter scan re-uploads a table of the same size: keep the allocation.  The copy goes on the
    // sweep stream, so it is ordered after the kernels of the previous equilibrium that still read it.
    cudaStream_t s = c->use_user_stream ? c->user_stream : c->stream;
    if (hm.tab.size() > c->cap_tab || !c->d_tab) {
        // the old table may still be read by kernels in flight
        CUDA_TRY(c, cudaStreamSynchronize(s));
        c->model_set = false;
        if ((rc = ensure(c, c->d_tab, c->cap_tab, hm.tab.size()))) return rc;
    }
    if (cudaMemcpyAsync(c->d_tab, hm.tab.data(), hm.tab.size() * sizeof(double), cudaMemcpyHostToDevice, s) !=
            cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {      // `tab` is a pageable buffer
        c->model_set = false;                                                // the old table is overwritten
        return fail(c, ESB_ERR_CUDA, "table upload failed");
    }
    c->dm = hm.dm;
    c->tab_doubles = (int)hm.tab.size();
    c->model = *m;
    c->model_set = true;
    c->guard_set = false;          // the guard belongs to the previous equilibrium
    c->guard_pending = false;
    return ESB_OK;
}

extern "C" int esb_set_model(esb_context* c, const esb_model* m, const double* rho, const double* drho,
                             int32_t n_nodes, double rho_boundary) {
    if (!c) return ESB_ERR_ARG;
    if (!m || (m->kind != ESB_SLAB_DENSITY && m->kind != ESB_CYLINDER_DENSITY) || m->scheme == ESB_RK8N)
        return fail(c, ESB_ERR_ARG, "esb_set_model is for the density kinds with ESB_RK4 / ESB_RK8");
    const double* fields[2] = {rho, drho};
    return esb_set_model_fields(c, m, fields, 2, n_nodes, &rho_boundary, 1);
}

// ---- launches ---------------------------------------------------------------
constexpr int TABLE_SMEM_MAX = (int)TABLE_BYTES_MAX;  // model_host.h refuses larger tables
constexpr int MAX_DEVICES = 64;

// Function attributes are set once per (kernel instantiation, device): the staged table is read-only
// broadcast data, so shared memory gets the whole carve-out and the dynamic limit is the table maximum.
template <class K>
static cudaError_t configure_once(K kernel, bool (&done)[MAX_DEVICES]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < MAX_DEVICES && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_SMEM_MAX);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < MAX_DEVICES) done[dev] = true;
    return cudaSuccess;
}

template <int KIND, int SCHEME, int NM>
static cudaError_t launch_grid_nm(const GridArgs& g, cudaStream_t s, int n_sm) {
    const size_t smem = (size_t)g.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(grid_kernel<KIND, SCHEME, NM>, done);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(g.tile_counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    // fewer warps than the launch bound leave registers for a refinement CTA of the previous equilibrium
    // beside it (esb_scan_models)
    const int threads = (g.threads > 0 && g.threads < ESB_GRID_THREADS(KIND, SCHEME)) ? g.threads
                                                                                    : ESB_GRID_THREADS(KIND, SCHEME);
    const long long n_tiles = (long long)g.nk * ((g.nw + 31) / 32);
    long long blocks = (n_tiles + threads / 32 - 1) / (threads / 32);
    // one CTA per SM; a grid with fewer tiles than warp slots spreads its tiles over all SMs
    if (n_tiles >= n_sm) blocks = n_sm;
    grid_kernel<KIND, SCHEME, NM><<<(int)blocks, threads, smem, s>>>(g);
    return cudaGetLastError();
}

constexpr size_t GRID_WARP_MAX_POINTS = 8192;     // (mode, k, omega) triples; above it one thread per point

template <int KIND, int SCHEME>
static cudaError_t launch_grid_warp(const GridArgs& g, cudaStream_t s, int n_sm) {
    const size_t smem = (size_t)g.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(grid_warp_kernel<KIND, SCHEME>, done);
    if (e != cudaSuccess) return e;
    const size_t total = (size_t)g.nk * g.nw * g.n_modes;
    int blocks = (int)((total + 3) / 4);           // 4 warps = 4 points per CTA at a time
    if (blocks > n_sm * 4) blocks = n_sm * 4;
    grid_warp_kernel<KIND, SCHEME><<<blocks, 128, smem, s>>>(g);
    return cudaGetLastError();
}

template <int KIND, int SCHEME>
static cudaError_t launch_grid(const GridArgs& g, cudaStream_t s, int n_sm) {
    if (g.schedule == 2 || (g.schedule == 0 && (size_t)g.nk * g.nw * g.n_modes <= GRID_WARP_MAX_POINTS))
        return launch_grid_warp<KIND, SCHEME>(g, s, n_sm);
    switch (g.n_modes) {
        case 1: return launch_grid_nm<KIND, SCHEME, 1>(g, s, n_sm);
        case 2: return launch_grid_nm<KIND, SCHEME, 2>(g, s, n_sm);
        case 3: return launch_grid_nm<KIND, SCHEME, 3>(g, s, n_sm);
        case 4: return launch_grid_nm<KIND, SCHEME, 4>(g, s, n_sm);
        default: return cudaErrorInvalidValue;
    }
}

// Persistent refinement launches: as many CTAs as are resident at once (n_total = host copy of the
// bracket count, for the launch size only - the kernels read the count from the device).
template <int KIND, int SCHEME, int MINB>
static cudaError_t launch_refine_b(const RefineArgs& r, cudaStream_t s, int n_total, int n_sm) {
    const size_t smem = (size_t)r.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(refine_kernel<KIND, SCHEME, MINB>, done);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_kernel<KIND, SCHEME, MINB>, 128, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int blocks = (n_total + 127) / 128;
    if (blocks > n_sm * per_sm) blocks = n_sm * per_sm;
    if (blocks < 1) blocks = 1;
    refine_kernel<KIND, SCHEME, MINB><<<blocks, 128, smem, s>>>(r);
    return cudaGetLastError();
}

template <int KIND, int SCHEME>
static cudaError_t launch_refine_warp(const RefineArgs& r, cudaStream_t s, int n_total, int n_sm) {
    const size_t smem = (size_t)r.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(refine_warp_kernel<KIND, SCHEME>, done);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, refine_warp_kernel<KIND, SCHEME>, 128, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int blocks = (n_total + 3) / 4;               // 4 warps = 4 brackets per CTA at a time
    if (blocks > n_sm * per_sm) blocks = n_sm * per_sm;
    if (blocks < 1) blocks = 1;
    refine_warp_kernel<KIND, SCHEME><<<blocks, 128, smem, s>>>(r);
    return cudaGetLastError();
}

// Resident CTAs per SM the lane-per-bracket kernel is compiled for (ESB_REFINE_MINB): 4 for the
// first-derivative schemes (5..8 spill the evaluation loop and are 5-30 % slower, round 1), 3 for the
// normal-form scheme (no spills at <= 168 registers; brackets + refinement of the bench sweep 4.9 ms
// against 5.2 ms at 4 and 5.0 ms at 2, scripts/gpu_variants.py)
template <int KIND, int SCHEME>
static cudaError_t launch_refine(const RefineArgs& r, cudaStream_t s, int n_total, int n_sm) {
    return launch_refine_b<KIND, SCHEME, ESB_REFINE_MINB(SCHEME)>(r, s, n_total, n_sm);
}

// (kind, scheme) of the uploaded model -> template instantiation.  The rotational kind exists for
// RK8 only (check_model enforces it).
template <class F>
static cudaError_t dispatch_kind(int kind, int scheme, F&& f) {
    using std::integral_constant;
#define ESB_KIND_CASE(K)                                                                           \
    case K:                                                                                        \
        return scheme == SCHEME_RK8 ? f(integral_constant<int, K>{}, integral_constant<int, SCHEME_RK8>{}) \
                                    : f(integral_constant<int, K>{}, integral_constant<int, SCHEME_RK4>{});
#define ESB_KIND_CASE_N(K)                                                                         \
    case K:                                                                                        \
        if (scheme == SCHEME_RK8N) return f(integral_constant<int, K>{}, integral_constant<int, SCHEME_RK8N>{}); \
        return scheme == SCHEME_RK8 ? f(integral_constant<int, K>{}, integral_constant<int, SCHEME_RK8>{}) \
                                    : f(integral_constant<int, K>{}, integral_constant<int, SCHEME_RK4>{});
    switch (kind) {
        ESB_KIND_CASE_N(KIND_SLAB_DENSITY)
        ESB_KIND_CASE_N(KIND_CYL_DENSITY)
        ESB_KIND_CASE(KIND_SLAB_FLOW)
        ESB_KIND_CASE_N(KIND_CYL_FLOW)
        case KIND_CYL_ROTATION:
            return f(integral_constant<int, KIND_CYL_ROTATION>{}, integral_constant<int, SCHEME_RK8>{});
        default:
            return cudaErrorInvalidValue;
    }
#undef ESB_KIND_CASE
#undef ESB_KIND_CASE_N
}


// ---- discretisation guard: set-up, launch, report ---------------------------------------------
extern "C" int esb_set_guard_fields(esb_context* c, const esb_model* fine, const double* const* fields,
                                    int32_t n_fields, int32_t n_nodes, const double* boundary,
                                    int32_t n_boundary, int32_t stride, double threshold) {
    if (!c) return ESB_ERR_ARG;
    if (stride == 0 || stride < ESB_GUARD_AUTO) {                       // switch the guard off
        c->guard_set = false;
        c->guard_stride = 0;
        return ESB_OK;
    }
    if (!c->model_set) return fail(c, ESB_ERR_ARG, "set the model before its guard");
    if (!fine || fine->kind != c->model.kind) return fail(c, ESB_ERR_ARG, "guard model of another kind");
    if (!(threshold > 0.0)) return fail(c, ESB_ERR_ARG, "guard threshold");
    HostModel hm;
    std::string err;
    int rc = build_host_model(fine, fields, n_fields, n_nodes, boundary, n_boundary, hm, err);
    if (rc) return fail(c, rc, err.c_str());
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->guard_stream) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->guard_stream, cudaStreamNonBlocking));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_scan, cudaEventDisableTiming));
        CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_guard, cudaEventDisableTiming));
        CUDA_TRY(c, cudaMalloc((void**)&c->d_guard, sizeof(GuardRecord)));
        CUDA_TRY(c, cudaHostAlloc((void**)&c->h_guard, sizeof(GuardRecord), cudaHostAllocDefault));
    }
    // the old table may still be read by a guard kernel in flight
    CUDA_TRY(c, cudaStreamSynchronize(c->guard_stream));
    c->guard_set = false;
    if ((rc = ensure(c, c->d_gtab, c->cap_gtab, hm.tab.size()))) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_gtab, hm.tab.data(), hm.tab.size() * sizeof(double), cudaMemcpyHostToDevice,
                                c->guard_stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->guard_stream));       // `tab` is a pageable buffer
    c->g_dm = hm.dm;
    c->g_tab_doubles = (int)hm.tab.size();
    c->guard_stride = stride;
    c->guard_threshold = threshold;
    c->guard_set = true;
    c->guard_pending = false;
    return ESB_OK;
}

template <int KIND, int SCHEME>
static cudaError_t launch_guard(const GuardArgs& g, cudaStream_t s) {
    const size_t smem = (size_t)g.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(guard_kernel<KIND, SCHEME>, done);
    if (e != cudaSuccess) return e;
    const unsigned total = g.n_per_mode * (unsigned)g.n_modes;
    guard_kernel<KIND, SCHEME><<<(total + 255) / 256, 256, smem, s>>>(g);
    return cudaGetLastError();
}

// after the scan of a sweep has been enqueued on `s`: the guard pass on the side stream
static int guard_launch(esb_context* c, int n_modes, const int32_t* modes, int nk, int nw, int layout,
                        cudaStream_t s) {
    const size_t plane = (size_t)nk * nw;
    GuardArgs g;
    g.M = c->g_dm;
    g.tab = c->d_gtab;
    g.tab_doubles = c->g_tab_doubles;
    g.k = c->d_k; g.w = c->d_w; g.nk = nk; g.nw = nw; g.layout = layout;
    g.n_modes = n_modes;
    for (int i = 0; i < 4; ++i) g.modes[i] = i < n_modes ? modes[i] : 0;
    g.ext = c->d_ext; g.intq = c->d_int; g.den = c->d_den;
    // ESB_GUARD_AUTO: about 32 k samples per sweep whatever its size, at least every 64th point
    int stride = c->guard_stride;
    if (stride == ESB_GUARD_AUTO) {
        stride = 64;
        while (stride < 4096 && plane * n_modes / stride > 32768) stride *= 2;
    }
    g.stride = stride;
    g.n_per_mode = 32u * (unsigned)((plane + 32 * (size_t)stride - 1) / (32 * (size_t)stride));
    g.threshold = c->guard_threshold;
    g.margin = 0.15;      // in the resonant (squared) quantity: ~7 % in phase speed (the parity tests: 0.02 absolute)
    g.out = c->d_guard;
    CUDA_TRY(c, cudaEventRecord(c->ev_scan, s));
    CUDA_TRY(c, cudaStreamWaitEvent(c->guard_stream, c->ev_scan, 0));
    CUDA_TRY(c, cudaMemsetAsync(c->d_guard, 0, sizeof(GuardRecord), c->guard_stream));
    const cudaError_t e = dispatch_kind(c->g_dm.kind, c->g_dm.scheme, [&](auto kind, auto scheme) {
        return launch_guard<decltype(kind)::value, decltype(scheme)::value>(g, c->guard_stream);
    });
    CUDA_TRY(c, e);
    c->launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_guard, c->d_guard, sizeof(GuardRecord), cudaMemcpyDeviceToHost, c->guard_stream));
    CUDA_TRY(c, cudaEventRecord(c->ev_guard, c->guard_stream));
    c->guard_pending = true;
    c->guard_n_modes = n_modes;
    c->guard_nw = nw;
    c->guard_n_per_mode = g.n_per_mode;
    c->guard_last_stride = stride;
    return ESB_OK;
}

extern "C" int esb_guard_result(esb_context* c, esb_guard_report* out) {
    if (!c || !out) return ESB_ERR_ARG;
    memset(out, 0, sizeof(*out));
    out->slot = out->k_index = out->w_index = -1;
    out->threshold = c->guard_threshold;
    out->stride = c->guard_set ? c->guard_stride : 0;
    if (!c->guard_set || !c->guard_pending) return ESB_OK;        // no guard, or no sweep since it was set
    out->stride = c->guard_last_stride;
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaEventSynchronize(c->ev_guard));
    const GuardRecord r = *c->h_guard;
    out->n_checked = (int64_t)r.n_checked;
    out->n_above = (int64_t)r.n_above;
    if (r.n_checked) {
        const unsigned long long bits = r.key & 0xffffffff00000000ULL;
        double worst;
        memcpy(&worst, &bits, sizeof(worst));
        out->worst = worst;
        const unsigned sample = (unsigned)(r.key & 0xffffffffULL);
        const int slot = (int)(sample / c->guard_n_per_mode);
        const size_t p = guard_point(sample - (unsigned)slot * c->guard_n_per_mode, c->guard_last_stride);
        out->slot = slot;
        out->k_index = (int32_t)(p / c->guard_nw);
        out->w_index = (int32_t)(p - (size_t)out->k_index * c->guard_nw);
    }
    return ESB_OK;
}

static int check_mode(const esb_context* c, int mode) {
    if (c->model.kind == ESB_SLAB_DENSITY || c->model.kind == ESB_SLAB_FLOW)
        return (mode == 0 || mode == 1) ? 0 : -1;
    return (mode >= 0 && mode <= ESB_MAX_ORDER) ? 0 : -1;
}

static int check_modes(const esb_context* c, int n_modes, const int32_t* modes) {
    if (n_modes < 1 || n_modes > ESB_MAX_MODES || !modes) return -1;
    for (int i = 0; i < n_modes; ++i)
        if (check_mode(c, modes[i])) return -1;
    return 0;
}

// one fused launch: every requested mode at every (k, omega); outputs mode-slot major
static int grid_dev_multi(esb_context* c, int n_modes, const int32_t* modes, const double* d_k, int nk,
                          const double* d_w, int nw, int layout, double* d_ext, double* d_int,
                          cudaStream_t s, double* d_den = nullptr) {
    if (!c->model_set) return fail(c, ESB_ERR_ARG, "model not set");
    if (nk <= 0 || nw <= 0 || !d_k || !d_w || !d_ext || !d_int || layout < 0 || layout > 2 ||
        check_modes(c, n_modes, modes))
        return fail(c, ESB_ERR_ARG, "bad grid arguments");
    GridArgs g;
    g.M = c->dm;
    g.tab = c->d_tab;
    g.tab_doubles = c->tab_doubles;
    g.k = d_k; g.w = d_w; g.nk = nk; g.nw = nw; g.layout = layout;
    g.n_modes = n_modes;
    for (int i = 0; i < 4; ++i) g.modes[i] = i < n_modes ? modes[i] : 0;
    g.ext = d_ext; g.intq = d_int; g.den = d_den;
    g.schedule = c->schedule;
    g.threads = c->scan_threads;
    g.tile_counter = reinterpret_cast<unsigned long long*>(c->d_counter + 2);
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaEventRecord(c->ev0, s));
    const cudaError_t e = dispatch_kind(c->dm.kind, c->dm.scheme, [&](auto kind, auto scheme) {
        return launch_grid<decltype(kind)::value, decltype(scheme)::value>(g, s, c->n_sm);
    });
    CUDA_TRY(c, e);
    CUDA_TRY(c, cudaEventRecord(c->ev1, s));
    c->timed = true;
    c->launches += 1;
    return ESB_OK;
}

static cudaStream_t cur_stream(esb_context* c) { return c->use_user_stream ? c->user_stream : c->stream; }

extern "C" int esb_dispersion_grid_dev(esb_context* c, int32_t mode, const double* d_k, int32_t nk,
                                       const double* d_w, int32_t nw, int32_t layout, double* d_ext,
                                       double* d_int, void* stream) {
    if (!c) return ESB_ERR_ARG;
    return grid_dev_multi(c, 1, &mode, d_k, nk, d_w, nw, layout, d_ext, d_int,
                          stream ? (cudaStream_t)stream : cur_stream(c));
}

static int bracket_segments(int nw) {
    const int n = (nw - 1 + BRACKET_SEG - 1) / BRACKET_SEG;
    return n > 0 ? n : 1;
}

// count pass + scan over all slots of `sw` (segment offsets in c->d_rowoff, slot_begin in c->d_slot_begin),
// then the asynchronous copy of slot_begin[] to the page-locked c->h_counts, marked by c->ev_counts
static int brackets_count(esb_context* c, SweepDev& sw, cudaStream_t s, int* d_slot_begin = nullptr) {
    const bool to_host = d_slot_begin == nullptr;      // the sweep reads the counts; a batch keeps them on the device
    if (!d_slot_begin) d_slot_begin = c->d_slot_begin;
    const bool ref = sw.rule != ESB_ACCEPT_CONVERGED;
    sw.nseg = ref ? 1 : bracket_segments(sw.nw);
    const size_t per_slot = (size_t)sw.nk * sw.nseg, items = per_slot * sw.n_slots;
    int rc;
    if ((rc = ensure(c, c->d_rowcount, c->cap_rows, items + 1))) return rc;
    if ((rc = ensure(c, c->d_rowoff, c->cap_rowoff, items + 1))) return rc;
    sw.seg_count = c->d_rowcount;
    sw.seg_offset = c->d_rowoff;
    sw.slot_begin = d_slot_begin;
    sw.seg_tmp = nullptr;
    if (!ref) {
        if ((rc = ensure(c, c->d_seg_tmp, c->cap_seg_tmp, items * BRACKET_SEG))) return rc;
        sw.seg_tmp = c->d_seg_tmp;
    }
    const int threads = 128, per_block = threads / 32;
    const int blocks = (int)((items + per_block - 1) / per_block);
    if (ref) bracket_reference_kernel<<<blocks, threads, 0, s>>>(sw, 0);
    else bracket_kernel<<<blocks, threads, 0, s>>>(sw, 0);
    CUDA_TRY(c, cudaGetLastError());
    scan_kernel<<<1, 1024, 0, s>>>(c->d_rowcount, c->d_rowoff, (int)items, (int)per_slot, d_slot_begin);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 2;
    if (to_host) {
        CUDA_TRY(c, cudaMemcpyAsync(c->h_counts, d_slot_begin, (sw.n_slots + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaEventRecord(c->ev_counts, s));
    }
    return ESB_OK;
}

static int brackets_fill(esb_context* c, const SweepDev& sw, cudaStream_t s) {
    const size_t items = (size_t)sw.nk * sw.nseg * sw.n_slots;
    const int threads = 128, per_block = threads / 32;
    const int blocks = (int)((items + per_block - 1) / per_block);
    if (sw.rule != ESB_ACCEPT_CONVERGED) bracket_reference_kernel<<<blocks, threads, 0, s>>>(sw, 1);
    else if (sw.seg_tmp) bracket_fill_kernel<<<blocks, threads, 0, s>>>(sw);
    else bracket_kernel<<<blocks, threads, 0, s>>>(sw, 1);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 1;
    return ESB_OK;
}

extern "C" int esb_brackets_dev(esb_context* c, const double* d_ext, const double* d_int, int32_t nk,
                                int32_t nw, int32_t* d_row_offset, int32_t* d_bk, int32_t* d_bw,
                                int32_t capacity, int32_t* n_host, void* stream) {
    if (!c || !d_ext || !d_int || !d_row_offset || nk <= 0 || nw <= 0) return ESB_ERR_ARG;
    cudaStream_t s = stream ? (cudaStream_t)stream : (c->use_user_stream ? c->user_stream : c->stream);
    CUDA_TRY(c, cudaSetDevice(c->device));
    SweepDev sw{};
    sw.n_slots = 1; sw.nk = nk; sw.nw = nw;
    sw.rule = ESB_ACCEPT_CONVERGED;
    sw.slot[0].gext = d_ext; sw.slot[0].gint = d_int;
    sw.slot[0].bk = d_bk; sw.slot[0].bw = d_bw; sw.slot[0].bw2 = nullptr;
    sw.slot[0].capacity = (d_bk && d_bw && capacity > 0) ? capacity : 0;
    int rc;
    if ((rc = brackets_count(c, sw, s))) return rc;
    row_offset_kernel<<<(nk + 1 + 255) / 256, 256, 0, s>>>(c->d_rowoff, nk, sw.nseg, d_row_offset);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 1;
    CUDA_TRY(c, cudaEventSynchronize(c->ev_counts));
    const int total = c->h_counts[1];
    if (n_host) *n_host = total;
    if (sw.slot[0].capacity > 0 && total > 0)
        if ((rc = brackets_fill(c, sw, s))) return rc;
    return total > capacity && d_bk ? ESB_ERR_CAPACITY : ESB_OK;
}

static int ensure_grid(esb_context* c, size_t n) {
    if (c->d_ext && c->d_int && c->cap_grid >= n) return ESB_OK;
    if (c->d_ext) cudaFree(c->d_ext);
    if (c->d_int) cudaFree(c->d_int);
    c->d_ext = c->d_int = nullptr;
    c->cap_grid = 0;
    CUDA_TRY(c, cudaMalloc((void**)&c->d_ext, n * sizeof(double)));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_int, n * sizeof(double)));
    c->cap_grid = n;
    return ESB_OK;
}

extern "C" int esb_upload_axes(esb_context* c, const double* k, int32_t nk, const double* w, int32_t nw,
                               int32_t layout);

static size_t w_len(int layout, int nk, int nw) {
    return layout == OMEGA_PER_K ? (size_t)nk * nw : (size_t)nw;
}

extern "C" int esb_dispersion_grid_multi(esb_context* c, int32_t n_modes, const int32_t* modes,
                                         const double* k, int32_t nk, const double* w, int32_t nw,
                                         int32_t layout, double* ext, double* intq) {
    if (!c) return ESB_ERR_ARG;
    if (!c->model_set) return fail(c, ESB_ERR_ARG, "model not set");
    if (!k || !w || !ext || !intq || nk <= 0 || nw <= 0 || layout < 0 || layout > 2 ||
        check_modes(c, n_modes, modes))
        return fail(c, ESB_ERR_ARG, "bad grid arguments");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = esb_upload_axes(c, k, nk, w, nw, layout))) return rc;
    const size_t n = (size_t)nk * nw * n_modes;
    if ((rc = ensure_grid(c, n))) return rc;
    cudaStream_t s = cur_stream(c);
    if ((rc = grid_dev_multi(c, n_modes, modes, c->d_k, nk, c->d_w, nw, layout, c->d_ext, c->d_int, s)))
        return rc;
    CUDA_TRY(c, cudaMemcpyAsync(ext, c->d_ext, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(intq, c->d_int, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    return ESB_OK;
}

template <int KIND, int SCHEME>
static cudaError_t launch_grid_leaky(const GridArgs& g, cudaStream_t s, int n_sm) {
    const size_t smem = (size_t)g.tab_doubles * sizeof(double);
    static bool done[MAX_DEVICES] = {};
    cudaError_t e = configure_once(grid_leaky_kernel<KIND, SCHEME>, done);
    if (e != cudaSuccess) return e;
    const size_t total = (size_t)g.nk * g.nw * g.n_modes;
    size_t blocks = (total + 127) / 128;
    if (blocks > (size_t)n_sm * 8) blocks = (size_t)n_sm * 8;
    grid_leaky_kernel<KIND, SCHEME><<<(int)blocks, 128, smem, s>>>(g);
    return cudaGetLastError();
}

// D over the whole grid, the leaky side (m_e < 0) included; same arguments and layout as
// esb_dispersion_grid_multi.  Where m_e >= 0 the values are those of the regular evaluation.
extern "C" int esb_dispersion_grid_leaky(esb_context* c, int32_t n_modes, const int32_t* modes, const double* k,
                                         int32_t nk, const double* w, int32_t nw, int32_t layout, double* ext,
                                         double* intq) {
    if (!c) return ESB_ERR_ARG;
    if (!c->model_set) return fail(c, ESB_ERR_ARG, "model not set");
    if (!k || !w || !ext || !intq || nk <= 0 || nw <= 0 || layout < 0 || layout > 2 ||
        check_modes(c, n_modes, modes))
        return fail(c, ESB_ERR_ARG, "bad grid arguments");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = esb_upload_axes(c, k, nk, w, nw, layout))) return rc;
    const size_t n = (size_t)nk * nw * n_modes;
    if ((rc = ensure_grid(c, n))) return rc;
    cudaStream_t s = cur_stream(c);
    GridArgs g{};
    g.M = c->dm;
    g.tab = c->d_tab;
    g.tab_doubles = c->tab_doubles;
    g.k = c->d_k; g.w = c->d_w; g.nk = nk; g.nw = nw; g.layout = layout;
    g.n_modes = n_modes;
    for (int i = 0; i < 4; ++i) g.modes[i] = i < n_modes ? modes[i] : 0;
    g.ext = c->d_ext; g.intq = c->d_int; g.den = nullptr;
    const cudaError_t e = dispatch_kind(c->dm.kind, c->dm.scheme, [&](auto kind, auto scheme) {
        return launch_grid_leaky<decltype(kind)::value, decltype(scheme)::value>(g, s, c->n_sm);
    });
    CUDA_TRY(c, e);
    c->launches += 1;
    CUDA_TRY(c, cudaMemcpyAsync(ext, c->d_ext, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(intq, c->d_int, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    return ESB_OK;
}

extern "C" int esb_dispersion_grid(esb_context* c, int32_t mode, const double* k, int32_t nk,
                                   const double* w, int32_t nw, int32_t layout, double* ext, double* intq) {
    return esb_dispersion_grid_multi(c, 1, &mode, k, nk, w, nw, layout, ext, intq);
}

extern "C" int esb_set_schedule(esb_context* c, int32_t mode) {
    if (!c || mode < 0 || mode > 2) return ESB_ERR_ARG;
    c->schedule = mode;
    return ESB_OK;
}

extern "C" int esb_set_stream(esb_context* c, void* stream) {
    if (!c) return ESB_ERR_ARG;
    c->user_stream = (cudaStream_t)stream;
    c->use_user_stream = true;
    return ESB_OK;
}

extern "C" int esb_upload_axes(esb_context* c, const double* k, int32_t nk, const double* w, int32_t nw,
                               int32_t layout) {
    if (!c) return ESB_ERR_ARG;
    if (!k || !w || nk <= 0 || nw <= 0 || layout < 0 || layout > 2) return fail(c, ESB_ERR_ARG, "bad axes");
    CUDA_TRY(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c, c->d_k, c->cap_k, (size_t)nk))) return rc;
    if ((rc = ensure(c, c->d_w, c->cap_w, w_len(layout, nk, nw)))) return rc;
    cudaStream_t s = cur_stream(c);
    CUDA_TRY(c, cudaMemcpyAsync(c->d_k, k, (size_t)nk * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_w, w, w_len(layout, nk, nw) * sizeof(double), cudaMemcpyHostToDevice, s));
    c->ax_nk = nk; c->ax_nw = nw; c->ax_layout = layout;
    return ESB_OK;
}

static size_t slot_bytes(size_t cap) { return cap * (3 * sizeof(double) + 5 * sizeof(int)); }

static void slot_carve(esb_context::RootBuf& sl, void* base, size_t cap, esb_roots* out) {
    double* d = (double*)base;
    int* q = (int*)(d + 3 * cap);
    if (out) {
        out->omega = d; out->ext = d + cap; out->intq = d + 2 * cap;
        out->k_index = q; out->w_index = q + cap; out->accepted = q + 2 * cap; out->iterations = q + 3 * cap;
    } else {
        sl.om = d; sl.e = d + cap; sl.i = d + 2 * cap;
        sl.bk = q; sl.bw = q + cap; sl.acc = q + 2 * cap; sl.it = q + 3 * cap; sl.bw2 = q + 4 * cap;
    }
}

static int ensure_slot(esb_context* c, esb_context::RootBuf& sl, size_t total) {
    if (sl.cap >= total && sl.base) return ESB_OK;
    if (sl.base) cudaFree(sl.base);
    void* pin = sl.pin;
    const size_t pin_cap = sl.pin_cap;
    sl = esb_context::RootBuf();
    sl.pin = pin;
    sl.pin_cap = pin_cap;
    const size_t cap = ((total + total / 4 + 64) + 1) & ~(size_t)1;      // even: the int block stays 8-byte aligned
    CUDA_TRY(c, cudaMalloc(&sl.base, slot_bytes(cap)));
    sl.cap = cap;
    slot_carve(sl, sl.base, cap, nullptr);
    return ESB_OK;
}

static void slots_to_dev(esb_context* c, SweepDev& sw, int n_modes, const int32_t* modes, size_t plane) {
    for (int m = 0; m < n_modes; ++m) {
        esb_context::RootBuf& sl = c->slots[m];
        SlotDev& q = sw.slot[m];
        q.gext = c->d_ext + m * plane; q.gint = c->d_int + m * plane; q.gden = c->d_den + m * plane;
        q.bk = sl.bk; q.bw = sl.bw; q.bw2 = sl.bw2;
        q.omega = sl.om; q.ext = sl.e; q.intq = sl.i; q.accepted = sl.acc; q.iters = sl.it;
        q.mode = modes[m];
        q.capacity = (int)sl.cap;
    }
}

// scan (ONE fused launch for all modes) -> brackets of all modes (one count, one scan, one fill
// launch) -> refinement of all of them (one persistent launch), on the axes already resident in HBM.
// The root tables stay on the device, one slot per mode (esb_download_roots_slot / esb_roots_pinned /
// esb_roots_device hand them out).  One host wait per sweep: the per-slot bracket counts, copied to
// page-locked memory while the fill pass runs; the refinement is still in flight when the call returns
// (every table accessor orders itself after it).
extern "C" int esb_sweep_resident_multi(esb_context* c, int32_t n_modes, const int32_t* modes,
                                        double tol_percent, int32_t* n_roots, int32_t* n_brackets) {
    if (!c) return ESB_ERR_ARG;
    if (!c->model_set) return fail(c, ESB_ERR_ARG, "model not set");
    if (c->ax_nk <= 0 || c->ax_nw <= 1) return fail(c, ESB_ERR_ARG, "axes not uploaded (need nw >= 2)");
    if (check_modes(c, n_modes, modes)) return fail(c, ESB_ERR_ARG, "bad modes");
    const int nk = c->ax_nk, nw = c->ax_nw, layout = c->ax_layout;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = cur_stream(c);
    int rc;
    const size_t plane = (size_t)nk * nw;
    for (int m = 0; m < ESB_MAX_MODES; ++m) c->slots[m].n = 0;      // slots of an earlier, wider sweep are void
    if ((rc = ensure_grid(c, plane * n_modes))) return rc;
    if ((rc = ensure(c, c->d_den, c->cap_den, plane * n_modes))) return rc;
    if ((rc = grid_dev_multi(c, n_modes, modes, c->d_k, nk, c->d_w, nw, layout, c->d_ext, c->d_int, s,
                             c->d_den)))
        return rc;
    // (worker-sized sweeps - the warp-per-point regime - are latency bound and are not sampled)
    c->guard_pending = false;
    if (c->guard_set && plane * n_modes > GRID_WARP_MAX_POINTS &&
        (rc = guard_launch(c, n_modes, modes, nk, nw, layout, s)))
        return rc;
    // first sweep of a context: room for one bracket per 64 grid points (grown on demand below)
    for (int m = 0; m < n_modes; ++m)
        if (!c->slots[m].base && (rc = ensure_slot(c, c->slots[m], std::max<size_t>(1024, plane / 64)))) return rc;
    SweepDev sw{};
    sw.n_slots = n_modes; sw.nk = nk; sw.nw = nw;
    sw.rule = c->accept_rule;
    sw.tol_percent = tol_percent;
    slots_to_dev(c, sw, n_modes, modes, plane);
    if ((rc = brackets_count(c, sw, s))) return rc;
    if ((rc = brackets_fill(c, sw, s))) return rc;                  // runs while the host waits for the counts
    CUDA_TRY(c, cudaEventSynchronize(c->ev_counts));
    int n_total = c->h_counts[n_modes];
    bool grown = false;
    for (int m = 0; m < n_modes; ++m) {
        const int total = c->h_counts[m + 1] - c->h_counts[m];
        if (n_brackets) n_brackets[m] = total;
        if (n_roots) n_roots[m] = total;
        if ((size_t)total > c->slots[m].cap) {
            if (!grown) CUDA_TRY(c, cudaStreamSynchronize(s));       // the fill pass still writes the old buffers
            grown = true;
            if ((rc = ensure_slot(c, c->slots[m], (size_t)total))) return rc;
        }
        c->slots[m].n = total;
    }
    if (grown) {                                                    // rare: a table outgrew its slot
        slots_to_dev(c, sw, n_modes, modes, plane);
        if ((rc = brackets_fill(c, sw, s))) return rc;
    }
    if (n_total > 0) {
        CUDA_TRY(c, cudaMemsetAsync(c->d_counter, 0, 2 * sizeof(int), s));
        RefineArgs r;
        r.M = c->dm;
        r.tab = c->d_tab;
        r.tab_doubles = c->tab_doubles;
        r.k = c->d_k; r.w = c->d_w; r.layout = layout;
        r.sw = sw;
        r.counter = c->d_counter;
        // ONE persistent launch refines the brackets of every mode (single work queue):
        // one warp per bracket (1/32 of the evaluation latency) or one lane per bracket.  Measured
        // crossovers (scripts/gpu_refine_modes.py, profiles/r01o_refine_modes.log): the cylinder
        // second-order kinds integrate ONE solution per lane but two per warp lane, so the warp kernel
        // wins only while the lanes cannot be filled (< ~24 k brackets); the two-solution kinds do the
        // same arithmetic either way and the warp kernel has a 32x shorter tail: ahead up to > 100 k
        // (rotation, whose per-lane exterior Bessel work is replicated 32 times: ~100 k)
        const int kind = c->dm.kind;
        const int limit = (kind == KIND_CYL_DENSITY || kind == KIND_CYL_FLOW) ? 24000
                          : kind == KIND_CYL_ROTATION ? 100000 : 250000;
        const bool warp_path = c->schedule == 2 || (c->schedule == 0 && n_total <= limit);
        const cudaError_t e = dispatch_kind(c->dm.kind, c->dm.scheme, [&](auto kind, auto scheme) {
            if (warp_path)
                return launch_refine_warp<decltype(kind)::value, decltype(scheme)::value>(r, s, n_total, c->n_sm);
            return launch_refine<decltype(kind)::value, decltype(scheme)::value>(r, s, n_total, c->n_sm);
        });
        CUDA_TRY(c, e);
        c->launches += 1;
    }
    // the next sweep overwrites the planes the guard pass reads: the sweep is complete when both are
    if (c->guard_pending) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_guard, 0));
    CUDA_TRY(c, cudaEventRecord(c->ev_done, s));
    c->tables_pending = true;
    return ESB_OK;
}

// device pointers of the compact table of the last esb_scan_models (valid until the next scan)
extern "C" int esb_scan_device(esb_context* c, esb_scan_result* out) {
    if (!c || !out) return ESB_ERR_ARG;
    *out = c->scan_dev;
    return ESB_OK;
}

template <class T>
static int ensure_pinned(esb_context* c, T*& p, size_t& cap, size_t need) {
    if (need <= cap && p) return ESB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    CUDA_TRY(c, cudaHostAlloc((void**)&p, need * sizeof(T), cudaHostAllocDefault));
    cap = need;
    return ESB_OK;
}

// A parameter scan as ONE batched job (BASELINE configs[4]; the reference has no scan driver - a user
// edits the speeds / profile constants of a script and reruns it): n_models equilibria of the same kind
// and discretisation over the axes resident in HBM.  All tables are built on the host and uploaded in
// one copy; per equilibrium the five launches of a sweep (scan, count, prefix, fill, refine) are
// enqueued back to back with NO host synchronisation - the bracket counts stay on the device, the
// refinement reads them there, every root table has room for `capacity_per_table` entries; at the end
// the tables are compacted on the device into one (model, slot, k_index, w_index, omega, ext, int,
// accepted, iterations) table and copied to page-locked memory in one piece.
extern "C" int esb_scan_models(esb_context* c, int32_t n_models, const esb_model* models,
                               const double* const* fields, int32_t n_fields, int32_t n_nodes,
                               const double* boundary, int32_t n_modes, const int32_t* modes, double tol_percent,
                               int32_t capacity_per_table, int32_t download, int32_t* n_brackets,
                               esb_scan_result* out) {
    if (!c) return ESB_ERR_ARG;
    if (n_models < 1 || !models || !fields || !boundary || !out) return fail(c, ESB_ERR_ARG, "bad scan arguments");
    if (c->ax_nk <= 0 || c->ax_nw <= 1) return fail(c, ESB_ERR_ARG, "axes not uploaded (need nw >= 2)");
    const int nk = c->ax_nk, nw = c->ax_nw, layout = c->ax_layout;
    const size_t plane = (size_t)nk * nw;
    // ---- host: every model's constants and staged table
    std::vector<HostModel> hm(n_models);
    std::string err;
    for (int i = 0; i < n_models; ++i) {
        int rc = build_host_model(&models[i], fields + (size_t)i * n_fields, n_fields, n_nodes, boundary + i, 1,
                                  hm[i], err);
        if (rc) return fail(c, rc, err.c_str());
        if (hm[i].dm.kind != hm[0].dm.kind || hm[i].dm.scheme != hm[0].dm.scheme ||
            hm[i].tab.size() != hm[0].tab.size())
            return fail(c, ESB_ERR_ARG, "the models of a scan share kind, scheme and mesh size");
    }
    const esb_model model_keep = c->model;
    c->model = models[0];                      // check_modes reads the kind
    const int bad_modes = check_modes(c, n_modes, modes);
    c->model = model_keep;
    if (bad_modes) return fail(c, ESB_ERR_ARG, "bad modes");
    const size_t tab_doubles = hm[0].tab.size();
    const int n_tables = n_models * n_modes;
    // even: every table of the packed allocation then starts 8-byte aligned
    const int cap = ((capacity_per_table > 0 ? capacity_per_table
                                             : (int)std::max<size_t>(4096, plane / 24)) + 1) & ~1;
    const size_t table_bytes = slot_bytes((size_t)cap);
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = cur_stream(c);
    int rc;
    if ((rc = ensure(c, c->d_bank, c->cap_bank, tab_doubles * n_models))) return rc;
    if ((rc = ensure(c, c->d_scan_roots, c->cap_scan_roots, table_bytes * n_tables))) return rc;
    const size_t n_ints = (size_t)n_models * (ESB_MAX_MODES + 1) + 3 * (size_t)n_tables + 2;
    if ((rc = ensure(c, c->d_scan_ints, c->cap_scan_ints, n_ints))) return rc;
    if ((rc = ensure_pinned(c, c->h_scan_ints, c->cap_h_scan_ints, 2 * (size_t)n_tables + 2))) return rc;
    if ((rc = ensure_grid(c, plane * n_modes))) return rc;
    if ((rc = ensure(c, c->d_den, c->cap_den, plane * n_modes))) return rc;
    int* d_sb = c->d_scan_ints;
    int* d_counts = d_sb + (size_t)n_models * (ESB_MAX_MODES + 1);
    int* d_found = d_counts + n_tables;
    int* d_offsets = d_found + n_tables;       // [n_tables + 1]
    {
        // one upload of all tables (pageable staging: the copy is synchronous with respect to the host buffer)
        std::vector<double> bank(tab_doubles * n_models);
        for (int i = 0; i < n_models; ++i) memcpy(&bank[tab_doubles * i], hm[i].tab.data(), tab_doubles * sizeof(double));
        CUDA_TRY(c, cudaMemcpyAsync(c->d_bank, bank.data(), bank.size() * sizeof(double), cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaStreamSynchronize(s));
    }
    const int kind = hm[0].dm.kind;
    const double* tab_keep = c->d_tab;
    const int tabd_keep = c->tab_doubles;
    const DevModel dm_keep = c->dm;
    const bool set_keep = c->model_set;
    // Pipeline: the refinement of equilibrium i (latency bound at the end of its queue; on a small grid for
    // most of its duration) runs on its own stream beside the scan of equilibrium i + 1.  The persistent scan
    // kernel normally takes the whole register file of an SM; launched with two thirds of its warps it
    // leaves room for one refinement CTA per SM.  Two sets of planes alternate; the root tables and the
    // bracket counts are per equilibrium anyway.
    static const bool overlap_env = [] { const char* e = getenv("ESB_SCAN_OVERLAP"); return !e || atoi(e) != 0; }();
    // Measured (scripts/gpu_scan_breakdown.py, one rank's share of the configs[4] job on 1 / 2 / 4 / 8 GPUs):
    // cylinder density family 541 / 284 / 157 / 94 ms without, 549 / 286 / 149 / 84 ms with the pipeline - the
    // normal-form scan loses more from running on 8 instead of 12 warps than a throughput-bound refinement
    // gains, so it is pipelined only while the refinement is latency bound (at most ~2 brackets per lane of
    // one launch); slab flow family 678 / 311 / 159 / 84 -> 661 / 309 / 157 / 82 ms: always.
    const bool overlap = overlap_env && n_models >= 2 &&
                         (hm[0].dm.scheme != SCHEME_RK8N || plane * n_modes / 64 <= 131072);
    if (overlap) {
        if (!c->refine_stream) {
            CUDA_TRY(c, cudaStreamCreateWithFlags(&c->refine_stream, cudaStreamNonBlocking));
            for (int b = 0; b < 2; ++b) {
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_fill[b], cudaEventDisableTiming));
                CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_refine[b], cudaEventDisableTiming));
            }
        }
        if (c->cap_grid2 < plane * n_modes) {
            CUDA_TRY(c, cudaStreamSynchronize(c->refine_stream));
            if (c->d_ext2) cudaFree(c->d_ext2);
            if (c->d_int2) cudaFree(c->d_int2);
            if (c->d_den2) cudaFree(c->d_den2);
            c->d_ext2 = c->d_int2 = c->d_den2 = nullptr;
            c->cap_grid2 = 0;
            CUDA_TRY(c, cudaMalloc((void**)&c->d_ext2, plane * n_modes * sizeof(double)));
            CUDA_TRY(c, cudaMalloc((void**)&c->d_int2, plane * n_modes * sizeof(double)));
            CUDA_TRY(c, cudaMalloc((void**)&c->d_den2, plane * n_modes * sizeof(double)));
            c->cap_grid2 = plane * n_modes;
        }
    }
    const int threads_keep = c->scan_threads;
    if (overlap) c->scan_threads = hm[0].dm.scheme == SCHEME_RK8N ? 256 : 384;
    bool refined[2] = {false, false};
    for (int i = 0; i < n_models; ++i) {
        const int b = overlap ? (i & 1) : 0;
        double* p_ext = b ? c->d_ext2 : c->d_ext;
        double* p_int = b ? c->d_int2 : c->d_int;
        double* p_den = b ? c->d_den2 : c->d_den;
        cudaStream_t rs = overlap ? c->refine_stream : s;
        // the context's launch helpers read (dm, d_tab): point them at equilibrium i of the bank
        c->dm = hm[i].dm;
        c->d_tab = c->d_bank + tab_doubles * i;
        c->tab_doubles = (int)tab_doubles;
        c->model_set = true;
        // these planes were read by the refinement of equilibrium i - 2
        if (overlap && refined[b] && cudaStreamWaitEvent(s, c->ev_refine[b], 0) != cudaSuccess) { rc = ESB_ERR_CUDA; break; }
        rc = grid_dev_multi(c, n_modes, modes, c->d_k, nk, c->d_w, nw, layout, p_ext, p_int, s, p_den);
        if (!rc) {
            SweepDev sw{};
            sw.n_slots = n_modes; sw.nk = nk; sw.nw = nw;
            sw.rule = c->accept_rule;
            sw.tol_percent = tol_percent;
            for (int m = 0; m < n_modes; ++m) {
                esb_context::RootBuf rb;
                slot_carve(rb, c->d_scan_roots + table_bytes * ((size_t)i * n_modes + m), (size_t)cap, nullptr);
                SlotDev& q = sw.slot[m];
                q.gext = p_ext + m * plane; q.gint = p_int + m * plane; q.gden = p_den + m * plane;
                q.bk = rb.bk; q.bw = rb.bw; q.bw2 = rb.bw2;
                q.omega = rb.om; q.ext = rb.e; q.intq = rb.i; q.accepted = rb.acc; q.iters = rb.it;
                q.mode = modes[m];
                q.capacity = cap;
            }
            rc = brackets_count(c, sw, s, d_sb + (size_t)i * (ESB_MAX_MODES + 1));
            if (!rc) rc = brackets_fill(c, sw, s);
            if (!rc && overlap) {
                if (cudaEventRecord(c->ev_fill[b], s) != cudaSuccess ||
                    cudaStreamWaitEvent(rs, c->ev_fill[b], 0) != cudaSuccess)
                    rc = ESB_ERR_CUDA;
            }
            if (!rc && cudaMemsetAsync(c->d_counter, 0, 2 * sizeof(int), rs) != cudaSuccess) rc = ESB_ERR_CUDA;
            if (!rc) {
                RefineArgs r;
                r.M = c->dm;
                r.tab = c->d_tab;
                r.tab_doubles = c->tab_doubles;
                r.k = c->d_k; r.w = c->d_w; r.layout = layout;
                r.sw = sw;
                r.counter = c->d_counter;
                // the bracket count is on the device only: a full persistent launch, the schedule from an
                // estimate (one bracket per 64 grid points: the bench sweep has one per 107) against the
                // measured crossovers of esb_sweep_resident_multi
                const size_t n_est = plane * n_modes / 64;
                const size_t limit = (kind == KIND_CYL_DENSITY || kind == KIND_CYL_FLOW) ? 24000
                                     : kind == KIND_CYL_ROTATION ? 100000 : 250000;
                const bool warp_path = c->schedule == 2 || (c->schedule == 0 && n_est <= limit);
                const int n_launch = (int)std::min<size_t>(plane, (size_t)1 << 30);
                const cudaError_t e = dispatch_kind(kind, hm[0].dm.scheme, [&](auto kd, auto scheme) {
                    if (warp_path)
                        return launch_refine_warp<decltype(kd)::value, decltype(scheme)::value>(r, rs, n_launch, c->n_sm);
                    return launch_refine<decltype(kd)::value, decltype(scheme)::value>(r, rs, n_launch, c->n_sm);
                });
                if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = ESB_ERR_CUDA; }
                c->launches += 1;
                if (!rc && overlap) {
                    if (cudaEventRecord(c->ev_refine[b], rs) != cudaSuccess) rc = ESB_ERR_CUDA;
                    refined[b] = true;
                }
            }
        }
        if (rc) break;
    }
    c->scan_threads = threads_keep;
    if (overlap)                                 // the compaction below reads every root table
        for (int b = 0; b < 2; ++b)
            if (refined[b] && cudaStreamWaitEvent(s, c->ev_refine[b], 0) != cudaSuccess && !rc) rc = ESB_ERR_CUDA;
    // the context keeps the model it had before the scan
    c->dm = dm_keep;
    c->d_tab = const_cast<double*>(tab_keep);
    c->tab_doubles = tabd_keep;
    c->model_set = set_keep;
    for (int m = 0; m < ESB_MAX_MODES; ++m) c->slots[m].n = 0;
    if (rc) return rc;
    // ---- counts of all tables -> host (the only wait on the scan itself)
    scan_counts_kernel<<<(n_tables + 127) / 128, 128, 0, s>>>(d_sb, n_models, n_modes, cap, d_counts, d_found);
    CUDA_TRY(c, cudaGetLastError());
    scan_kernel<<<1, 1024, 0, s>>>(d_counts, d_offsets, n_tables, 0, nullptr);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 2;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scan_ints, d_found, (size_t)n_tables * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scan_ints + n_tables, d_offsets, ((size_t)n_tables + 1) * sizeof(int),
                                cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    bool overflow = false;
    for (int t = 0; t < n_tables; ++t) {
        if (n_brackets) n_brackets[t] = c->h_scan_ints[t];
        overflow = overflow || c->h_scan_ints[t] > cap;
    }
    const size_t total = (size_t)c->h_scan_ints[2 * n_tables];
    // ---- compact table: [omega | ext | int] doubles, [model | slot | k | w | accepted | iterations] ints
    const size_t entry = 3 * sizeof(double) + 6 * sizeof(int);
    const size_t cap_e = (total + 1) & ~(size_t)1;
    if ((rc = ensure(c, c->d_compact, c->cap_compact, cap_e * entry + 64))) return rc;
    if (download && (rc = ensure_pinned(c, c->h_compact, c->cap_h_compact, cap_e * entry + 64))) return rc;
    auto carve = [&](char* base, esb_scan_result& r) {
        double* d = (double*)base;
        int* q = (int*)(d + 3 * cap_e);
        r.omega = d; r.ext = d + cap_e; r.intq = d + 2 * cap_e;
        r.model = q; r.slot = q + cap_e; r.k_index = q + 2 * cap_e; r.w_index = q + 3 * cap_e;
        r.accepted = q + 4 * cap_e; r.iterations = q + 5 * cap_e;
    };
    esb_scan_result dv{}, hv{};
    carve(c->d_compact, dv);
    if (download) carve(c->h_compact, hv);
    dv.n_entries = (int32_t)total;
    c->scan_dev = dv;
    if (total > 0) {
        ScanOut so{dv.model, dv.slot, dv.k_index, dv.w_index, dv.accepted, dv.iterations, dv.omega, dv.ext, dv.intq};
        scan_compact_kernel<<<n_tables, 256, 0, s>>>(c->d_scan_roots, table_bytes, cap, n_modes, d_counts, d_offsets, so);
        CUDA_TRY(c, cudaGetLastError());
        c->launches += 1;
        if (download) {
            CUDA_TRY(c, cudaMemcpyAsync(c->h_compact, c->d_compact, cap_e * entry, cudaMemcpyDeviceToHost, s));
            CUDA_TRY(c, cudaStreamSynchronize(s));
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_done, s));
    c->tables_pending = true;
    hv.n_entries = (int32_t)total;
    *out = hv;
    if (overflow) return fail(c, ESB_ERR_CAPACITY, "capacity_per_table too small (n_brackets holds the sizes found)");
    return ESB_OK;
}

extern "C" int esb_sweep_resident(esb_context* c, int32_t mode, double tol_percent, int32_t* n_roots,
                                  int32_t* n_brackets) {
    return esb_sweep_resident_multi(c, 1, &mode, tol_percent, n_roots, n_brackets);
}

extern "C" int esb_download_roots_slot(esb_context* c, int32_t slot, esb_roots* out, int32_t max_roots) {
    if (!c || !out || slot < 0 || slot >= ESB_MAX_MODES) return ESB_ERR_ARG;
    const esb_context::RootBuf& sl = c->slots[slot];
    const size_t nb = (size_t)sl.n;
    if ((int64_t)nb > (int64_t)max_roots) return fail(c, ESB_ERR_CAPACITY, "max_roots too small");
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = cur_stream(c);
    if (c->tables_pending) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_done, 0));
    if (nb) {
        if (out->k_index) CUDA_TRY(c, cudaMemcpyAsync(out->k_index, sl.bk, nb * 4, cudaMemcpyDeviceToHost, s));
        if (out->w_index) CUDA_TRY(c, cudaMemcpyAsync(out->w_index, sl.bw, nb * 4, cudaMemcpyDeviceToHost, s));
        if (out->omega) CUDA_TRY(c, cudaMemcpyAsync(out->omega, sl.om, nb * 8, cudaMemcpyDeviceToHost, s));
        if (out->ext) CUDA_TRY(c, cudaMemcpyAsync(out->ext, sl.e, nb * 8, cudaMemcpyDeviceToHost, s));
        if (out->intq) CUDA_TRY(c, cudaMemcpyAsync(out->intq, sl.i, nb * 8, cudaMemcpyDeviceToHost, s));
        if (out->accepted) CUDA_TRY(c, cudaMemcpyAsync(out->accepted, sl.acc, nb * 4, cudaMemcpyDeviceToHost, s));
        if (out->iterations) CUDA_TRY(c, cudaMemcpyAsync(out->iterations, sl.it, nb * 4, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(c, cudaStreamSynchronize(s));
    return ESB_OK;
}

extern "C" int esb_roots_pinned(esb_context* c, int32_t slot, esb_roots* out, int32_t* n_roots) {
    if (!c || !out || slot < 0 || slot >= ESB_MAX_MODES) return ESB_ERR_ARG;
    esb_context::RootBuf& sl = c->slots[slot];
    memset(out, 0, sizeof(*out));
    if (n_roots) *n_roots = sl.n;
    if (sl.n == 0) return ESB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (sl.pin_cap < sl.cap) {
        if (sl.pin) cudaFreeHost(sl.pin);
        sl.pin = nullptr;
        sl.pin_cap = 0;
        CUDA_TRY(c, cudaHostAlloc(&sl.pin, slot_bytes(sl.cap), cudaHostAllocDefault));
        sl.pin_cap = sl.cap;
    }
    // the pinned mirror has the layout of a table of capacity pin_cap; copy the 7 used prefixes
    esb_roots h;
    slot_carve(sl, sl.pin, sl.pin_cap, &h);
    cudaStream_t s = cur_stream(c);
    if (c->tables_pending) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_done, 0));
    const size_t nb = (size_t)sl.n;
    CUDA_TRY(c, cudaMemcpyAsync(h.omega, sl.om, nb * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.ext, sl.e, nb * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.intq, sl.i, nb * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.k_index, sl.bk, nb * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.w_index, sl.bw, nb * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.accepted, sl.acc, nb * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(h.iterations, sl.it, nb * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    *out = h;
    return ESB_OK;
}

extern "C" int esb_download_roots(esb_context* c, esb_roots* out, int32_t max_roots) {
    return esb_download_roots_slot(c, 0, out, max_roots);
}

extern "C" int esb_roots_device(esb_context* c, int32_t slot, esb_roots* out, int32_t* n_roots) {
    if (!c || !out || slot < 0 || slot >= ESB_MAX_MODES) return ESB_ERR_ARG;
    const esb_context::RootBuf& sl = c->slots[slot];
    out->k_index = sl.bk; out->w_index = sl.bw; out->omega = sl.om; out->ext = sl.e;
    out->intq = sl.i; out->accepted = sl.acc; out->iterations = sl.it;
    if (n_roots) *n_roots = sl.n;
    return ESB_OK;
}


// Accepted modes of the first n_slots mode slots of the last sweep, packed on the device into the caller's
// buffer d_out[(capacity + 1) * 3] (doubles): row 0 = (rows written, table entries scanned, 1 if more modes
// than `capacity` were found), rows 1.. = (k_offset + k_stride * k_index, omega, slot) in (slot, k index,
// omega index) order.  Asynchronous on the context's stream, ordered after the sweep; consumer_stream (a
// cudaStream_t, may be NULL) is made to wait for it.  The payload of a multi-GPU gather (one fixed-size
// all-gather, no count exchange, no host synchronisation).
extern "C" int esb_pack_modes_dev(esb_context* c, int32_t n_slots, double k_offset, double k_stride, double* d_out,
                                  int32_t capacity, void* consumer_stream) {
    if (!c || !d_out || n_slots < 1 || n_slots > ESB_MAX_MODES || capacity < 0) return ESB_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = cur_stream(c);
    PackArgs a{};
    a.n_slots = n_slots;
    a.begin[0] = 0;
    for (int m = 0; m < n_slots; ++m) {
        const esb_context::RootBuf& sl = c->slots[m];
        a.accepted[m] = sl.acc; a.k_index[m] = sl.bk; a.omega[m] = sl.om;
        a.begin[m + 1] = a.begin[m] + sl.n;
    }
    a.k_offset = k_offset; a.k_stride = k_stride;
    a.out = d_out; a.capacity = capacity;
    const int total = a.begin[n_slots];
    const int blocks = total > 0 ? (total + PACK_PER_BLOCK - 1) / PACK_PER_BLOCK : 1;
    int rc;
    if ((rc = ensure(c, c->d_pack_counts, c->cap_pack_counts, (size_t)blocks))) return rc;
    a.block_count = c->d_pack_counts;
    pack_modes_kernel<<<blocks, PACK_THREADS, 0, s>>>(a, 0);
    CUDA_TRY(c, cudaGetLastError());
    pack_modes_kernel<<<blocks, PACK_THREADS, 0, s>>>(a, 1);
    CUDA_TRY(c, cudaGetLastError());
    c->launches += 2;
    if (consumer_stream && (cudaStream_t)consumer_stream != s) {      // the reader's stream waits for the payload
        CUDA_TRY(c, cudaEventRecord(c->ev_counts, s));
        CUDA_TRY(c, cudaStreamWaitEvent((cudaStream_t)consumer_stream, c->ev_counts, 0));
    }
    return ESB_OK;
}

// The sweep returns with its refinement still in flight.  stream = a cudaStream_t that will read the
// tables handed out by esb_roots_device: it is made to wait (on the device) for the sweep; NULL: the
// calling host thread waits instead.
extern "C" int esb_tables_wait(esb_context* c, void* stream) {
    if (!c) return ESB_ERR_ARG;
    if (!c->tables_pending) return ESB_OK;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (stream) {
        CUDA_TRY(c, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_done, 0));
    } else {
        CUDA_TRY(c, cudaEventSynchronize(c->ev_done));
        c->tables_pending = false;
    }
    return ESB_OK;
}

extern "C" int esb_set_accept_rule(esb_context* c, int32_t rule) {
    if (!c || rule < ESB_ACCEPT_CONVERGED || rule > ESB_ACCEPT_REFERENCE_SLAB) return ESB_ERR_ARG;
    c->accept_rule = rule;
    return ESB_OK;
}

extern "C" int esb_find_roots(esb_context* c, int32_t mode, const double* k, int32_t nk, const double* w,
                              int32_t nw, int32_t layout, double tol_percent, int32_t max_roots,
                              esb_roots* out, int32_t* n_roots, int32_t* n_brackets) {
    if (!c) return ESB_ERR_ARG;
    if (!out || !n_roots || max_roots < 0) return fail(c, ESB_ERR_ARG, "bad arguments");
    int rc;
    if ((rc = esb_upload_axes(c, k, nk, w, nw, layout))) return rc;
    if ((rc = esb_sweep_resident(c, mode, tol_percent, n_roots, n_brackets))) return rc;
    return esb_download_roots(c, out, max_roots);
}

// ---- FP64 pipe peak: the roofline denominator bench.py reports against ----
__global__ void dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" int esb_fp64_peak(esb_context* c, double* tflops) {
    if (!c || !tflops) return ESB_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CUDA_TRY(c, cudaGetDeviceProperties(&prop, c->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double* d = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaStream_t s = cur_stream(c);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(c->ev0, s);
        dfma_peak_kernel<<<blocks, threads, 0, s>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(c->ev1, s);
        cudaError_t e = cudaEventSynchronize(c->ev1);
        if (e != cudaSuccess) { cudaFree(d); CUDA_TRY(c, e); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    c->timed = false;
    c->launches += 4;
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) * 1e-12;
    return ESB_OK;
}

// Host-side self test of the integrator the kernels use (same rk8_step / rk4_step code, compiled
// for the host): y'' = sin(t) y' - (1 + t^2) y, y(0) = 1, y'(0) = 0.3, uniform steps over [0, T].
extern "C" int esb_rk_selftest(int32_t scheme, int32_t n_steps, double T, double out[2]) {
    if ((scheme != ESB_RK4 && scheme != ESB_RK8) || n_steps < 1 || !out) return ESB_ERR_ARG;
    double y[1] = {1.0}, yp[1] = {0.3};
    const double h = T / n_steps;
    for (int i = 0; i < n_steps; ++i) {
        const double t0 = i * h;
        if (scheme == ESB_RK8) {
            const double c[5] = {0.0, C8_M, 0.5, C8_P, 1.0};
            double ha[5], h2b[1][5];
            for (int n = 0; n < 5; ++n) {
                const double t = t0 + c[n] * h;
                ha[n] = h * sin(t);
                h2b[0][n] = -h * h * (1.0 + t * t);
            }
            double z[1] = {h * yp[0]};           // step-scaled slope, as in integrate_layer
            rk8_step<1>(y, z, ha, h2b);
            yp[0] = z[0] / h;
        } else {
            const double c[3] = {0.0, 0.5, 1.0};
            double a[3], b[1][3];
            for (int n = 0; n < 3; ++n) {
                const double t = t0 + c[n] * h;
                a[n] = sin(t);
                b[0][n] = -(1.0 + t * t);
            }
            rk4_step<1>(y, yp, h, a, b);
        }
    }
    out[0] = y[0];
    out[1] = yp[0];
    return ESB_OK;
}

extern "C" int esb_bessel_jy(int32_t n, double x, double out[4]) {
    if (n < 0 || n > ESB_MAX_ORDER || !(x > 0.0) || !out) return ESB_ERR_ARG;
    BesselJY b;
    bessel_jy(n, x, b);
    bessel_jy_order(b, n, x, out[0], out[1], out[2], out[3]);
    return ESB_OK;
}

// host arrays in and out, evaluated by the DEVICE build of the evaluators
static int dev_map(esb_context* c, const double* a, const double* b, int32_t count, int n_out, double* out,
                   const std::function<cudaError_t(const double*, const double*, double*, cudaStream_t)>& launch) {
    if (!c || !a || !out || count <= 0) return ESB_ERR_ARG;
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = cur_stream(c);
    double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    int rc = ESB_OK;
    auto done = [&](int code) {
        cudaFree(d_a); cudaFree(d_b); cudaFree(d_o);
        return code;
    };
    if (cudaMalloc((void**)&d_a, sizeof(double) * count) != cudaSuccess ||
        (b && cudaMalloc((void**)&d_b, sizeof(double) * count) != cudaSuccess) ||
        cudaMalloc((void**)&d_o, sizeof(double) * count * n_out) != cudaSuccess)
        return done(fail(c, ESB_ERR_ALLOC, "device allocation"));
    if (cudaMemcpyAsync(d_a, a, sizeof(double) * count, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        (b && cudaMemcpyAsync(d_b, b, sizeof(double) * count, cudaMemcpyHostToDevice, s) != cudaSuccess))
        return done(fail(c, ESB_ERR_CUDA, "upload"));
    if (launch(d_a, d_b, d_o, s) != cudaSuccess) return done(fail(c, ESB_ERR_CUDA, "launch"));
    c->launches += 1;
    if (cudaMemcpyAsync(out, d_o, sizeof(double) * count * n_out, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
        return done(fail(c, ESB_ERR_CUDA, "download"));
    return done(rc);
}

extern "C" int esb_bessel_jy_dev(esb_context* c, int32_t n, const double* x, int32_t count, double* out) {
    if (n < 0 || n > ESB_MAX_ORDER) return ESB_ERR_ARG;
    return dev_map(c, x, nullptr, count, 4, out, [&](const double* dx, const double*, double* d_o, cudaStream_t s) {
        bessel_jy_kernel<<<(count + 127) / 128, 128, 0, s>>>(n, dx, count, d_o);
        return cudaGetLastError();
    });
}

// (P, dP/dr) at |r| = 1 of the exterior solution where m_e < 0, for the exterior medium and the initial
// values of `m`: host build, and the same through the device build
extern "C" int esb_exterior_leaky(const esb_model* m, int32_t n, double k, double w, double out[2]) {
    if (check_model(m) || n < 0 || n > ESB_MAX_ORDER || !(k > 0.0) || !out) return ESB_ERR_ARG;
    const double vAe2 = m->vA_e * m->vA_e, ce2 = m->c_e * m->c_e, se2 = vAe2 + ce2, cTe2 = ce2 * vAe2 / se2;
    const double K = k * k, A = w * w;
    const double me = ((K * vAe2 - A) * (K * ce2 - A)) / (se2 * (K * cTe2 - A));
    out[0] = out[1] = nan("");
    if (!(me < 0.0)) return ESB_OK;           // not on the leaky side: the regular exterior applies
    exterior_cyl_leaky(m->ext_ic_value, m->ext_ic_slope, m->r_sign >= 0 ? 1.0 : -1.0,
                       m->ext_wavelengths * 2.0 * M_PI, k, me, n, out[0], out[1]);
    return ESB_OK;
}

extern "C" int esb_exterior_leaky_dev(esb_context* c, int32_t n, const double* k, const double* w, int32_t count,
                                      double* out) {
    if (!c || !c->model_set) return ESB_ERR_ARG;
    if (n < 0 || n > ESB_MAX_ORDER || !w) return ESB_ERR_ARG;
    const DevModel M = c->dm;
    return dev_map(c, k, w, count, 2, out, [&](const double* dk, const double* dw, double* d_o, cudaStream_t s) {
        exterior_leaky_kernel<<<(count + 127) / 128, 128, 0, s>>>(M, n, dk, dw, count, d_o);
        return cudaGetLastError();
    });
}

extern "C" int esb_bessel_ik_scaled(int32_t n, double z, double out[4]) {
    if (n < 0 || n > ESB_MAX_ORDER || !(z > 0.0) || !out) return ESB_ERR_ARG;
    BesselIK b;
    bessel_ik_scaled(n, z, b);
    bessel_order(b, n, z, out[0], out[1], out[2], out[3]);
    return ESB_OK;
}
