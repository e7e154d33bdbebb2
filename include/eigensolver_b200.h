/* eigensolver_b200 - C ABI of the B200 dispersion-function hot path.
 *
 * Drop-in boundary for the shooting solvers of samuelskirvin/EIGENSOLVER.  Every
 * reference solver script evaluates, for each (k, omega) of a scan,
 *
 *     D(omega, k) = exterior matched quantity - interior matched quantity
 *
 * by an exterior ODE integration, an interior shooting integration (odeint +
 * fsolve) and a subtraction, then bisects sign changes of D along omega:
 *
 *   Cylinder/Non-uniform density/Coronal/solvers/Density_cylinder.py
 *       kink()    :546-824   (scan loop 694-821, bisection locate_kink 548-686)
 *       sausage() :847-1122  (scan loop 990-1119)
 *   Slab/Non uniform density/Coronal/Solvers/multiprocessor_Inhomogeneous_method_coronal.py
 *       sausage() :461-600, kink() :640-790
 *
 * This library replaces exactly that: esb_dispersion_grid() is the body of the
 * scan loop over a whole (k, omega) grid, esb_find_roots() is scan + bracket +
 * refinement + the reference's acceptance test.  Plain pointers and sizes only;
 * all "host" entry points take HOST pointers and do their own H2D/D2H copies, the
 * "_dev" entry points take DEVICE pointers (data already resident in HBM).
 *
 * Threading: one esb_context per host thread / per GPU.  All calls return 0 on
 * success, a negative esb_status otherwise; esb_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute call fails with
 * ESB_ERR_CUDA.
 */
#ifndef EIGENSOLVER_B200_H
#define EIGENSOLVER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESB_VERSION 130   /* 1.1: esb_model grew the mesh_* fields; pinned tables, schedules
                             1.2: ESB_RK8N (normal-form Nystrom scheme, needs the profile's second derivative);
                                  up to 4 fused modes; esb_set_accept_rule; esb_tables_wait
                             1.3: esb_set_guard_fields / esb_guard_result (discretisation guard);
                                  esb_bessel_jy[_dev], esb_exterior_leaky[_dev] (J_n, Y_n: the leaky side); esb_pack_modes_dev;
                                  esb_model_max_steps; the guard judges the projective mismatch (see esb_set_guard_fields) */

typedef struct esb_context esb_context;

enum esb_status {
    ESB_OK = 0,
    ESB_ERR_ARG = -1,      /* bad argument / model not set */
    ESB_ERR_CUDA = -2,     /* CUDA runtime error (no device, launch failure ...) */
    ESB_ERR_CAPACITY = -3, /* more roots than max_roots; n_roots holds the needed size */
    ESB_ERR_ALLOC = -4
};

/* Solver family = which reference script the model restates. */
enum esb_model_kind {
    ESB_SLAB_DENSITY = 0,     /* slab, rho(x)   : ...Inhomogeneous_method_coronal.py        */
    ESB_CYLINDER_DENSITY = 1, /* cylinder rho(r): Density_cylinder.py                       */
    ESB_SLAB_FLOW = 2,        /* slab, sheared flow U(x): flow_multiprocessor_coronal.py    */
    ESB_CYLINDER_ROTATION = 3,/* cylinder, rotational flow v_phi(r): Twisted_photospheric_*.py
                                 (uniform rho_i0, vA_i0; written in r > 0; RK8 only)         */
    ESB_CYLINDER_FLOW = 4     /* cylinder, axial flow v_z(r): Cylinder_method_flow_testing.py
                                 (uniform rho_i0, c_i0, vA_i0 inside; written in r < 0)      */
};

/* Fixed-step integrator used across the layer. */
enum esb_scheme {
    ESB_RK4 = 0, /* classical 4-stage, stage nodes c = 0, 1/2, 1                              */
    ESB_RK8 = 1, /* Cooper-Verner 11-stage 8th order, nodes c = 0, (7-/+sqrt21)/14, 1/2, 1   */
    ESB_RK8N = 2 /* the same method in Nystrom form on the normal form u'' = q u (u = sqrt|F| y) of
                    the second-order kinds (cylinder density / axial flow, slab density): 64 instead of
                    110 FP64 instructions per solution and step.  The default of those kinds; the
                    profile's SECOND derivative is a third field of esb_set_model_fields            */
};

/* How the omega axis is given. */
enum esb_omega_layout {
    ESB_OMEGA_SHARED = 0,      /* w[nw]:      omega_ij = w[j]              (same for every k) */
    ESB_OMEGA_PHASE_SPEED = 1, /* w[nw]:      omega_ij = k[i] * w[j]       (w = omega/k grid,
                                  what the reference's driver builds: Density_cylinder.py:1145) */
    ESB_OMEGA_PER_K = 2        /* w[nk*nw]:   omega_ij = w[i*nw + j]                          */
};

/* Equilibrium + discretisation.  Speeds as in Density_cylinder.py:69-80. */
typedef struct esb_model {
    int32_t kind;          /* esb_model_kind */
    int32_t scheme;        /* esb_scheme */
    int32_t n_steps;       /* integration steps across the layer */
    int32_t mesh;          /* 0 = sin^2 clustering at the layer ends, 1 = uniform, 2 = graded
                              (target step size, see mesh_axis/mesh_edge below) */
    double c_i0, vA_i0, vA_e, c_e, gamma, rho_i0;
    double rho_A;          /* amplitude multiplying the density profile (reference rho_A) */
    double ext_ic_value;   /* exterior initial values at x = -3*2*pi/k: (1e-8,            */
    double ext_ic_slope;   /*   1e-8) slab :247, (1e-8, 1e-15) cylinder :768              */
    double ext_wavelengths;/* exterior domain = ext_wavelengths*2*pi/k  (reference: 3)     */
    double s_start, s_end; /* layer: boundary and far end: (-1, 1) slab, (-1, -0.001) cyl  */
    double U_e;            /* ESB_SLAB_FLOW: exterior flow speed (flow script :52); there c_i0,
                              vA_i0, rho_i0 are the uniform interior c_i, vA_i, rho_i.
                              ESB_CYLINDER_FLOW: unused by the kernels (that script's exterior
                              is at rest in its own frame: m_e and xi_e use omega unshifted,
                              Cylinder_method_flow_testing.py:706-709)                         */
    int32_t r_sign;        /* cylinder: -1 = script written in r<0 (coronal, default), +1 = r>0
                              (photospheric: s_start=1, s_end=0.001, slope given as dP/dr)     */
    int32_t reserved;
    /* mesh = 2: the local step is H * min(1, |r|/mesh_axis, (mesh_edge + |r - s_start|)/mesh_edge_width)
     * with H fixed by n_steps: geometric towards the axis (where the 1/r, m^2/r^2 coefficients
     * need h ~ r), refined towards the boundary (where a resonance just outside the layer makes
     * the coefficients vary fastest), uniform in between.  Slab kinds: no axis term, refined
     * towards both boundaries. */
    double mesh_axis, mesh_edge, mesh_edge_width;
} esb_model;

/* sizeof(esb_model) as compiled into the library: a binding checks its own struct against it. */
int esb_sizeof_model(void);

/* Defaults of the reference scripts for `kind` (coronal parameter set). */
int esb_model_defaults(int32_t kind, esb_model* out);

/* Number of profile sample nodes the chosen scheme needs, and their positions.
 * The caller samples its density profile (any function) at these nodes:
 * this is what replaces the reference's sympy `profile(x)` + lambdify.
 * Nodes are ordered along the direction of integration. */
int esb_mesh_size(const esb_model* m, int32_t* n_nodes);
int esb_mesh_nodes(const esb_model* m, double* nodes /* [n_nodes] */);

int esb_create(int32_t device, esb_context** out);
int esb_destroy(esb_context* ctx);
const char* esb_last_error(const esb_context* ctx);

/* Upload the model: rho[n_nodes], drho[n_nodes] = profile and its derivative at
 * esb_mesh_nodes(); rho_b = rho at the boundary s_start. */
int esb_set_model(esb_context* ctx, const esb_model* m, const double* rho, const double* drho,
                  int32_t n_nodes, double rho_boundary);

/* Generic form: fields[f][n_nodes] sampled at esb_mesh_nodes().
 *   density kinds : fields = {rho, rho'} (+ rho'' with ESB_RK8N),   boundary = {rho(s_start)}
 *   ESB_SLAB_FLOW : fields = {U, U', U''},      boundary = {U(s_start)}
 *   ESB_CYLINDER_ROTATION : fields = {v_phi, v_phi', c_i^2}, boundary = {v_phi(s_start)}
 *   ESB_CYLINDER_FLOW : fields = {v_z, v_z'} (+ v_z'' with ESB_RK8N), boundary = {v_z(s_start)}
 * esb_model_n_fields() returns the count for a model.  A failed call leaves the previous model set. */
int esb_model_n_fields(const esb_model* m, int32_t* n_fields);
/* The staged table ([esb_mesh_size() nodes][products per node] + 4 doubles per step) lives in the shared memory
 * of every CTA (200 KB): the largest n_steps esb_set_model_fields accepts for this model's (kind, scheme) -
 * 710 for ESB_RK8N and the rotation kind, 1279 for the other first-derivative tables.  Host code, no GPU. */
int esb_model_max_steps(const esb_model* m, int32_t* max_steps);
int esb_set_model_fields(esb_context* ctx, const esb_model* m, const double* const* fields,
                         int32_t n_fields, int32_t n_nodes, const double* boundary,
                         int32_t n_boundary);

/* D over a grid.  mode: slab 0 = sausage, 1 = kink; cylinder = azimuthal order
 * 0 (sausage), 1 (kink), 2, 3 (fluting).  ext/intq: [nk*nw] row-major, NaN where the
 * reference skips the point (m_e < 0) or where the exterior solution overflows the double
 * range (there the reference's odeint fails as well).  D = ext - intq. */
int esb_dispersion_grid(esb_context* ctx, int32_t mode, const double* k, int32_t nk,
                        const double* w, int32_t nw, int32_t omega_layout, double* ext,
                        double* intq);

/* Root table: scan + sign-change brackets along omega + Brent refinement + the
 * reference's acceptance test |ext-int|*100/max(|ext|,|int|) < tol_percent
 * (Density_cylinder.py:809).  Outputs hold up to max_roots entries, sorted by
 * (k index, omega index of the bracket); *n_roots = number found. */
typedef struct esb_roots {
    int32_t* k_index;     /* [max_roots] row of k[]                                   */
    int32_t* w_index;     /* [max_roots] bracket = (w_index, w_index+1)                */
    double* omega;        /* [max_roots] refined root                                  */
    double* ext;          /* [max_roots] exterior quantity at the root                 */
    double* intq;         /* [max_roots] interior quantity at the root                 */
    int32_t* accepted;    /* [max_roots] 1 = passes acceptance test (a mode), 0 = pole */
    int32_t* iterations;  /* [max_roots] Brent iterations used (0: a pole recognised from the
                             scan itself - the denominator of the interior quantity changes sign
                             across the bracket; omega = its interpolated zero, ext = int = NaN) */
} esb_roots;

int esb_find_roots(esb_context* ctx, int32_t mode, const double* k, int32_t nk, const double* w,
                   int32_t nw, int32_t omega_layout, double tol_percent, int32_t max_roots,
                   esb_roots* out, int32_t* n_roots, int32_t* n_brackets);

/* The same pipeline in three steps, for callers that keep the axes resident in HBM
 * (bench.py's device-resident timing, multi-GPU gathers):
 *   esb_upload_axes     H2D copy of k[] and w[] into context-owned buffers
 *   esb_sweep_resident  grid -> brackets -> refinement, root table left on the device
 *   esb_download_roots  D2H copy of the root table (synchronises the stream)
 *   esb_roots_device    device pointers of the root table (valid until the next sweep; see esb_tables_wait) */
int esb_upload_axes(esb_context* ctx, const double* k, int32_t nk, const double* w, int32_t nw,
                    int32_t omega_layout);
int esb_sweep_resident(esb_context* ctx, int32_t mode, double tol_percent, int32_t* n_roots,
                       int32_t* n_brackets);
int esb_download_roots(esb_context* ctx, esb_roots* out, int32_t max_roots);
int esb_roots_device(esb_context* ctx, int32_t slot, esb_roots* out, int32_t* n_roots);
/* A sweep returns with its refinement still in flight on the context's stream.  The download calls order
 * themselves after it; a consumer of esb_roots_device pointers calls esb_tables_wait first: `stream` (a
 * cudaStream_t) is made to wait on the device, NULL blocks the calling host thread until the tables are
 * complete. */
int esb_tables_wait(esb_context* ctx, void* stream);

/* Discretisation guard.  The integrator is fixed-step where the reference's odeint adapts (Density_cylinder.py
 * :768-790): a profile sharper than the shipped ones silently loses digits.  esb_set_guard_fields takes the
 * SAME equilibrium on a finer mesh - `fine` = the model with (normally) 2 x n_steps, the fields sampled at ITS
 * esb_mesh_nodes() - and from then on every esb_sweep_resident[_multi] re-evaluates every stride-th
 * (grid point, mode) of its scan with the fine table, on a side stream next to the bracket passes (~2/stride of
 * the scan's arithmetic).  esb_guard_result waits for that pass and reports the worst deviation - judged on
 * g = D Y / (|ext Y| + |int Y|), Y = the denominator of int: the acceptance test's relative mismatch, regular at
 * the poles of D and unchanged by a common factor of int's numerator and denominator, which cancels in D (the
 * amplitude error of the solution that dominates towards the axis: 2.5e-9 for the fluting order n = 3 at the
 * default steps, which G = D Y alone - the measure of the first 1.3 builds - reported although D is converged to
 * 1e-13) - over the sampled points outside the resonant
 * continua, where it occurred, and how many samples exceeded `threshold`.  8th order: halving the step divides
 * the error by ~256, so the value is (to 0.4 %) the discretisation error of the sweep itself.  The samples are
 * tiles of 32 consecutive omega points, one tile per 32 * stride grid points; stride = ESB_GUARD_AUTO lets every
 * sweep pick the stride that judges about 32 k samples (64 <= stride <= 4096: 0.5 % of a 3e7-point sweep, 3 % of
 * a 1e5-point one); stride = 0 switches the guard off; it must be set again after every
 * esb_set_model[_fields].  Sweeps of at most 8192
 * (point, mode) pairs - the latency-bound worker-sized calls - and esb_scan_models are not sampled
 * (n_checked = 0). */
#define ESB_GUARD_AUTO (-1)
typedef struct esb_guard_report {
    double worst;        /* largest deviation over the judged samples (0 if none) */
    double threshold;
    int32_t slot, k_index, w_index;   /* where (-1: none) */
    int32_t stride;      /* the stride the last sweep used; 0: no guard set */
    int64_t n_checked;   /* samples judged (evaluated, outside the continua) */
    int64_t n_above;     /* ... of them above the threshold */
} esb_guard_report;
int esb_set_guard_fields(esb_context* ctx, const esb_model* fine, const double* const* fields, int32_t n_fields,
                         int32_t n_nodes, const double* boundary, int32_t n_boundary, int32_t stride,
                         double threshold);
int esb_guard_result(esb_context* ctx, esb_guard_report* out);

/* Accepted modes of mode slots 0..n_slots-1 of the last sweep, packed ON THE DEVICE into the caller's device
 * buffer d_out[(capacity + 1) * 3] (doubles): row 0 = (rows written, table entries scanned, 1 if more than
 * `capacity` modes were found), rows 1.. = (k_offset + k_stride * k_index, omega, slot), ordered by (slot,
 * k index, omega index).  Asynchronous on the context's stream, after the sweep; `consumer_stream` (a
 * cudaStream_t that will read d_out, may be NULL) is made to wait for it.  This is the payload of the
 * multi-GPU gather (what the reference's parent process collects from its workers' queues,
 * Density_cylinder.py:1150-1170): one fixed-size all-gather, no count exchange, no host synchronisation. */
int esb_pack_modes_dev(esb_context* ctx, int32_t n_slots, double k_offset, double k_stride, double* d_out,
                       int32_t capacity, void* consumer_stream);

/* What a sweep reports (esb_set_accept_rule; default ESB_ACCEPT_CONVERGED).
 *   ESB_ACCEPT_CONVERGED  sign changes of D between ADJACENT evaluated grid points, every bracket refined to
 *                         2 eps |omega| (Brent on the pole-free G = D Y), accepted = the reference's test
 *                         at the converged root.  One entry per bracket.
 *   ESB_ACCEPT_REFERENCE  the scripts' own rule, point for point (Density_cylinder.py:803-821 and
 *                         locate_kink :548-686): walking the evaluated points of a row in order, (i) a grid
 *                         point whose mismatch is below tol_percent is a solution as it stands (w_index ==
 *                         its index, iterations 0) and restarts the count of points seen; (ii) a sign change
 *                         against the previous EVALUATED point (skipped m_e < 0 points in between do not
 *                         matter) is bisected only if more than two points have been seen since the last
 *                         restart; (iii) the bisection evaluates the middle of the pair, takes the first
 *                         point inside the band as THE solution (no further refinement) and otherwise
 *                         follows a sign change in the UPPER half only, as the scripts' recursion does
 *                         (accepted = 0 where it gives up).  This reproduces the point sets the scripts
 *                         pickle; omega is then a dyadic point of the frequency grid, not a converged root.
 *   ESB_ACCEPT_REFERENCE_SLAB  the slab scripts' variant of the same rule (..._method_coronal.py:606-624,
 *                         :502-518): a sign change is bisected after more than ONE point seen, and the
 *                         bisection follows the lower half too. */
enum esb_accept_rule { ESB_ACCEPT_CONVERGED = 0, ESB_ACCEPT_REFERENCE = 1, ESB_ACCEPT_REFERENCE_SLAB = 2 };
int esb_set_accept_rule(esb_context* ctx, int32_t rule);

/* Several modes in ONE fused scan (n_modes <= 4): cylinder orders share the staged
 * coefficient evaluation and the Bessel sets, the slab's sausage and kink share the whole
 * integration.  Grids are mode-slot major: ext[(slot*nk + i)*nw + j].  The sweep keeps one
 * root table per mode slot on the device. */
int esb_dispersion_grid_multi(esb_context* ctx, int32_t n_modes, const int32_t* modes, const double* k,
                              int32_t nk, const double* w, int32_t nw, int32_t omega_layout,
                              double* ext, double* intq);
int esb_sweep_resident_multi(esb_context* ctx, int32_t n_modes, const int32_t* modes,
                             double tol_percent, int32_t* n_roots /* [n_modes] */,
                             int32_t* n_brackets /* [n_modes] */);
int esb_download_roots_slot(esb_context* ctx, int32_t slot, esb_roots* out, int32_t max_roots);

/* A parameter scan as ONE batched job (BASELINE configs[4]): n_models equilibria of the same kind, scheme
 * and mesh size, swept over the axes uploaded with esb_upload_axes.  fields[i * n_fields + f] = field f of
 * model i at ITS esb_mesh_nodes(), boundary[i] = its first field at s_start.  Every (model, mode) root
 * table has room for capacity_per_table entries (0: one per 24 grid points); n_brackets[i * n_modes + m]
 * receives the sizes found.  No host synchronisation between the equilibria; the result is ONE compact
 * table in page-locked memory owned by the context (valid until the next scan or esb_destroy), ordered by
 * (model, mode slot, k index, omega index).  ESB_ERR_CAPACITY: some table outgrew its room (the result
 * holds the first capacity_per_table entries of it).  Where it pays (always for the first-derivative kinds;
 * for the normal-form kinds while the refinement of one grid is latency bound) the refinement of equilibrium
 * i runs on a second stream beside the scan of equilibrium i + 1, on a second set of planes; the results do not
 * depend on it (environment ESB_SCAN_OVERLAP=0 switches it off). */
typedef struct esb_scan_result {
    int32_t n_entries;
    int32_t* model;       /* index into models[]                       */
    int32_t* slot;        /* index into modes[]                        */
    int32_t* k_index;
    int32_t* w_index;
    int32_t* accepted;
    int32_t* iterations;
    double* omega;
    double* ext;
    double* intq;
} esb_scan_result;
int esb_scan_models(esb_context* ctx, int32_t n_models, const esb_model* models, const double* const* fields,
                    int32_t n_fields, int32_t n_nodes, const double* boundary, int32_t n_modes,
                    const int32_t* modes, double tol_percent, int32_t capacity_per_table, int32_t download,
                    int32_t* n_brackets /* [n_models * n_modes] */, esb_scan_result* out);
/* download = 0 leaves the compact table on the device (out->n_entries is set, its pointers are NULL):
 * esb_scan_device hands out the device pointers (multi-GPU gathers read them in place; order the consumer
 * with esb_tables_wait). */
int esb_scan_device(esb_context* ctx, esb_scan_result* out);

/* D2H copy of the root table of `slot` into page-locked host buffers OWNED BY THE CONTEXT; *out
 * receives their addresses, *n_roots the entry count.  One packed copy at full PCIe rate, no
 * pageable staging.  The buffers stay valid until the next esb_roots_pinned call for the same slot
 * or esb_destroy.  Synchronises the stream. */
int esb_roots_pinned(esb_context* ctx, int32_t slot, esb_roots* out, int32_t* n_roots);

/* Schedule of the scan and refinement kernels: 0 (default) = chosen by size, 1 = one lane per
 * (k, omega) point / per bracket (throughput: every lane integrates its own point), 2 = one warp per
 * point / per bracket (latency: the 32 lanes integrate sub-intervals of the layer and multiply their
 * transfer matrices).  A schedule is bit-reproducible; the two agree to rounding (~1e-14 relative),
 * so force one of them when grids of different sizes are to be compared bit for bit. */
int esb_set_schedule(esb_context* ctx, int32_t mode);

/* Run every launch and copy of this context on `stream` (a cudaStream_t, e.g. the
 * caller's torch stream) instead of the context's own stream. */
int esb_set_stream(esb_context* ctx, void* stream);

/* Measured FP64 FMA throughput of the device (independent DFMA chains, TFLOP/s):
 * the roofline denominator bench.py quotes the grid kernel against. */
int esb_fp64_peak(esb_context* ctx, double* tflops);

/* Device-resident variants: all pointers are device pointers, `stream` is a
 * cudaStream_t (0 = default stream).  Asynchronous; no host synchronisation. */
int esb_dispersion_grid_dev(esb_context* ctx, int32_t mode, const double* d_k, int32_t nk,
                            const double* d_w, int32_t nw, int32_t omega_layout, double* d_ext,
                            double* d_intq, void* stream);

/* Brackets of a device-resident D grid: d_row_count[nk] and, sorted, d_bracket_w[*]
 * (caller passes capacity; returns total via host int after a stream sync). */
int esb_brackets_dev(esb_context* ctx, const double* d_ext, const double* d_intq, int32_t nk,
                     int32_t nw, int32_t* d_row_offset /* [nk+1] */, int32_t* d_bracket_k,
                     int32_t* d_bracket_w, int32_t capacity, int32_t* n_brackets_host,
                     void* stream);

/* Host-side helper (no GPU needed): scaled modified Bessel functions used by the
 * exterior solution, out = {e^-z I_n, d/dz, e^z K_n, d/dz}.  For unit tests. */
int esb_bessel_ik_scaled(int32_t n, double z, double out[4]);

/* Bessel functions J_n, Y_n (n <= 3): the oscillatory (leaky, m_e < 0) side of the exterior solution, which the
 * reference's scan loop skips (Density_cylinder.py:760) and the sweeps therefore report as "no value".
 *   esb_bessel_jy            host build: out = {J_n, J_n', Y_n, Y_n'}(x), x > 0
 *   esb_bessel_jy_dev        the same evaluators run on the device: x[count] in, out[4 * count]
 *   esb_exterior_leaky       closed form of the reference's own exterior initial-value problem
 *                            (:765-770) where m_e < 0: out = (P, dP/dr) at |r| = 1 for azimuthal order n,
 *                            the exterior medium and initial values of `m` (NaN where m_e >= 0)
 *   esb_exterior_leaky_dev   the same on the device for k[count], w[count] and the context's model
 *   esb_dispersion_grid_leaky  OPT-IN: D over the whole grid with the leaky side included - where m_e < 0 the
 *                            exterior is that closed form (slab: cos / sin) and the interior shooting and the
 *                            matching are unchanged, i.e. what the reference's scan loop would evaluate without
 *                            its "if m_e < 0: pass"; where m_e >= 0 the values of esb_dispersion_grid_multi.
 *                            Same arguments and layout as esb_dispersion_grid_multi.  The sweeps, brackets and
 *                            root tables keep the reference's skip rule. */
int esb_bessel_jy(int32_t n, double x, double out[4]);
int esb_bessel_jy_dev(esb_context* ctx, int32_t n, const double* x, int32_t count, double* out);
int esb_exterior_leaky(const esb_model* m, int32_t n, double k, double w, double out[2]);
int esb_exterior_leaky_dev(esb_context* ctx, int32_t n, const double* k, const double* w, int32_t count,
                           double* out);
int esb_dispersion_grid_leaky(esb_context* ctx, int32_t n_modes, const int32_t* modes, const double* k,
                              int32_t nk, const double* w, int32_t nw, int32_t omega_layout, double* ext,
                              double* intq);

/* Host-side helper (no GPU needed): integrates y'' = sin(t) y' - (1+t^2) y, y(0)=1, y'(0)=0.3
 * over [0,T] in n_steps uniform steps with the SAME step functions the kernels use; out =
 * {y(T), y'(T)}.  tests/test_tableau.py checks the 8th / 4th order of convergence with it. */
int esb_rk_selftest(int32_t scheme, int32_t n_steps, double T, double out[2]);

/* Timing hook for bench.py: device time (ms) of the last grid kernel launch
 * measured with CUDA events on the launching stream; <0 if none. */
double esb_last_kernel_ms(const esb_context* ctx);
/* Number of kernels this context has launched so far. */
int64_t esb_launch_count(const esb_context* ctx);

int esb_version(void);

#ifdef __cplusplus
}
#endif
#endif
