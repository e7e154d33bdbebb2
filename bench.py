#!/usr/bin/env python
"""Benchmark of the dispersion-function hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over the workload BASELINE.json quotes the metric
on (configs[1]): the cylinder with non-uniform density, modes n = 0, 1, 2, on a
1000 k x 10000 omega grid - 3e7 evaluations of D(omega,k) followed by bracket detection
and root refinement.  With N > 1 every rank sweeps its own 1000 wavenumbers (rows r, r+N, ...) of
an N*1000 k grid (weak scaling, no data-path collective) and the modes found are gathered
with NCCL inside the timed region.

  value  whole-job D evaluations per second, axes already resident in HBM
  e2e    the same through the public host API: pinned host k/omega in, the root tables of the
         three modes out into page-locked host memory (N > 1: gathered modes on rank 0)
  roofline.bound = "fp64": the kernel is an FP64-pipe kernel (no tensor cores, 24 B of HBM
         traffic per 2.4e4 flops), so the bound is the FP64 FMA rate, measured in the same
         process with a DFMA-chain kernel (esb_fp64_peak); frac = algorithmic flops (fma = 2) against it,
         pipe_frac = FP64 instructions issued against its instruction rate; traffic = ncu DRAM bytes per launch.
  configs / strong_scaling / guard / roots_per_sec_regular: BASELINE configs[0], [2], [3] at full size, the
         fixed-size configs[4] job, the discretisation guard's report of the timed sweeps, brackets outside
         the resonant continua.
  cpu_baseline  the oracle port of the reference's scipy path (odeint + fsolve) timed on a
         bounded sample with all host cores.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NK, NW = 1000, 10000
MODES = (0, 1, 2)
K_RANGE = (0.01, 4.5)          # Density_cylinder.py:1126  wavenumber = linspace(0.01, 4.5, ...)
W_RANGE = (0.5, 5.0)           # c_e .. vA_e : the phase-speed window in which m_e >= 0
N_STEPS = 152                  # the cylinder kind's default: normal-form scheme, graded mesh, see DESIGN.md
# algorithmic FP64 flops (fma = 2, mul/add/div = 1), see DESIGN.md.  Per step of the Cooper-Verner method
# in Nystrom form on u'' = q u (scheme "rk8n"), step-scaled variables (u, h u'):
#   shared by the modes evaluated together: 4 node evaluations of h^2 q x 17 flop + 6 (step scalings) = 74
#   per mode: 52 FMAs (3 stage bases, 40 + 4 + 5 tableau products) + 1 add + 11 products q_i U_i
#             + 5 FMAs (n^2 - 1/4) h^2/r^2 + 1 (rescaling h u') = 127
# one D evaluation alone: 201 flop/step; three fused: (74 + 3*127)/3 = 152 flop/step/eval
# (round 1, first-derivative form "rk8": 292 and 244 flop/step, 144 steps)
FLOPS_PER_EVAL = 201 * N_STEPS + 700
FLOPS_FUSED_LAUNCH = ((74 + 3 * 127) * N_STEPS + 3 * 500) * NK * NW
# FP64-pipe instructions of the same launch, from the SASS of the step loop (tools/sass_loop_count.py:
# 275 per step for three modes: 195 DFMA, 68 DMUL, 11 DADD, 1 MUFU.RCP64H - one reciprocal serves the four
# new stage nodes of a step) - the pipe-utilisation view of the same roofline
FP64_INSTR_FUSED_LAUNCH = (275 * N_STEPS + 3 * 300) * NK * NW
WORKLOAD = "cylinder non-uniform density, n=0,1,2, 1000 k x 10000 omega per GPU"


# ----------------------------------------------------------------- CPU legs ----
def _cpu_init():
    import warnings
    warnings.filterwarnings("ignore")


def _cpu_eval(args):
    """One D evaluation through the oracle port of the reference's scipy path (coefficients in closed form:
    the per-point sympy / lambdify rebuild of Density_cylinder.py:705-757 hoisted out)."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import reference_path as rp
    mode, k, w = args
    prof = rp.GaussianDensity(rp.CYL_CORONAL, width=0.95, const_B=True)
    e, i = rp.dispersion(rp.CylinderDensity(prof, mode), k, w)      # scipy defaults, fsolve
    return e - i


def _cpu_eval_unhoisted(args):
    """The same evaluation with the reference's own cost structure: sympy re-derives and lambdifies the six
    coefficient functions at every (k, omega), as the script's scan loop does."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import reference_path as rp
    mode, k, w = args
    prof = rp.GaussianDensity(rp.CYL_CORONAL, width=0.95, const_B=True)
    e, i = rp.dispersion(rp.CylinderDensityPerPoint(prof, mode), k, w)
    return e - i


def cpu_sample(nk, nw, seed=0):
    """A bounded sample of the same workload: nk wavenumbers x nw phase speeds x 3 modes."""
    rng = np.random.default_rng(seed)
    ks = rng.uniform(K_RANGE[0], K_RANGE[1], nk)
    Ws = rng.uniform(W_RANGE[0], W_RANGE[1], nw)
    return [(m, k, k * W) for m in MODES for k in ks for W in Ws]


def time_cpu(pool, cores, nk, nw, seed, fn=_cpu_eval):
    pts = cpu_sample(nk, nw, seed)
    t = time.perf_counter()
    pool.map(fn, pts, chunksize=max(1, len(pts) // (cores * 8)))
    dt = time.perf_counter() - t
    return len(pts), dt


def unhoisted_rate(pool, cores):
    """evaluations/s of the per-point-sympy variant on a small sample (~10 s on all cores)"""
    time_cpu(pool, cores, 1, cores, 2, _cpu_eval_unhoisted)          # sympy import, caches
    nk, nw = 4, max(4, 4 * cores)
    n, dt = time_cpu(pool, cores, nk, nw, 7, _cpu_eval_unhoisted)
    return {"value": n / dt, "unit": "evals/s",
            "sample": "%d k x %d omega x 3 modes (%.0f s): coefficients re-derived with sympy and lambdified at "
                      "every point, as Density_cylinder.py:705-757 does" % (nk, nw, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nk, nw = 8, max(8, 4 * cores)           # ~4 s of work on all cores per step
    with mp.get_context("fork").Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(args.warmup):
            time_cpu(pool, cores, 2, cores, 1)
        tot_n = tot_t = 0.0
        for s in range(args.steps):
            n, dt = time_cpu(pool, cores, nk, nw, 100 + s)
            tot_n += n
            tot_t += dt
        unhoisted = unhoisted_rate(pool, cores)
    value = tot_n / tot_t
    sample = "%d k x %d omega x 3 modes per step (uniform random in the workload's k/omega box)" % (nk, nw)
    line = {
        "impl": "reference", "metric": "dispersion_evals_per_sec", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "sampled": "the CPU arm times a uniform random SAMPLE of the workload's (k, omega) box per step "
                              "(the full 3e7-point grid takes ~27 h on these cores); rates are per evaluation"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample,
                         "unhoisted": unhoisted},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle/reference_path.py = the reference's numpy/scipy path (odeint + fsolve at scipy "
                "defaults) with the sympy/lambdify coefficient rebuild hoisted out (value: the faster, conservative "
                "rate); cpu_baseline.unhoisted = the same with the rebuild per point, the reference's own cost "
                "structure; the reference scripts themselves need /root/reference, matplotlib and numpy<1.18 and "
                "cannot travel",
    }
    print(json.dumps(line))


# ----------------------------------------------------------------- GPU arm ----
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused grid kernel on this
    workload, from the committed `ncu --set full` capture (profiles/ncu_traffic.json names the
    report it was read from); None if no capture of this configuration is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            rec = json.load(fh)
        if rec.get("nk") == NK and rec.get("nw") == NW and rec.get("modes") == list(MODES) and \
                rec.get("n_steps") == N_STEPS:
            return float(rec["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML in-process (every 5 ms), the
    `nvidia-smi` query of the profiling recipe as a fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm = []
        self.sm_max = None
        self.seen = set()
        self.source = None
        self.stop = threading.Event()

    def _nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        uuid = None
        try:      # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES, NVML indexes the machine
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis else self.index
        except Exception:
            phys = self.index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.source = "nvml"
        while not self.stop.is_set():
            self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.seen.update(n for n, b in bits.items() if r & b)
            self.stop.wait(0.005)

    def _smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.source = "nvidia-smi"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                r = [x.strip() for x in out.split(",")]
                if r and r[0].isdigit():
                    self.sm.append(int(r[0]))
                    self.sm_max = int(r[1]) if r[1].isdigit() else self.sm_max
                    self.seen.update(n for j, n in enumerate(self.NAMES) if len(r) > 2 + j and r[2 + j] == "Active")
            except Exception:
                pass
            self.stop.wait(0.1)

    def run(self):
        try:
            self._nvml()
        except Exception:
            self._smi()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max,
                "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(sm), "source": self.source}


def cylinder_continua(medium, profile, margin=0.02):
    """Phase-speed intervals of the bench equilibrium in which an Alfven or cusp resonance sits inside
    the layer (the noise floor): [min, max] of vA(r) and cT(r) over the layer, +- margin."""
    r = np.linspace(-1.0, -0.001, 4001)
    rho = profile(medium, r)[0]
    beta = medium.vA_i0**2 * medium.rho_i0
    alpha = medium.rho_e * (medium.c_e**2 + 0.5 * medium.gamma * medium.vA_e**2) - 0.5 * medium.gamma * beta
    vA2, c2 = beta / rho, alpha / rho
    cT2 = c2 * vA2 / (c2 + vA2)
    return [(np.sqrt(v.min()) - margin, np.sqrt(v.max()) + margin) for v in (vA2, cT2)]


def other_configs(esb, local):
    """BASELINE configs[0], [2], [3] at full size on one GPU (scan + brackets + refinement, axes resident):
    evaluations per second, time of the scan kernel and its share - the driver-run counterpart of
    profiles/r01u_configs_throughput.jsonl."""
    import torch
    out = []
    cases = [
        ("configs[0] slab non-uniform density, sausage+kink, 200 k x 2000 omega", "slab_density", {}, [0, 1],
         np.linspace(0.001, 0.75, 200), np.linspace(0.41, 2.95, 2000), 20),
        ("configs[2] slab sheared flow U0(x), backward+forward branches, 2000 k x 20000 omega", "slab_flow",
         dict(medium=esb.FlowMedium(U_i0=0.35), profile=esb.GaussianFlow(1.0)), [0, 1],
         np.linspace(0.01, 4.5, 2000), np.linspace(-2.7, 2.7, 20000), 3),
        ("configs[3] cylinder rotational flow Omega(r), n = 0..3, 2000 k x 20000 omega", "cylinder_rotation",
         dict(profile=esb.PowerLawRotation(0.15, 1.25), s_end=0.01), [0, 1, 2, 3],
         np.linspace(0.25, 4.0, 2000), np.linspace(0.40, 1.6, 20000), 2),
    ]
    for name, kind, kw, modes, k, W, steps in cases:
        with esb.DispersionSolver(kind, device=local, **kw) as s:
            s.upload_axes(k, W)
            s.sweep_resident_multi(modes)
            s.lib.esb_tables_wait(s.ctx, None)
            torch.cuda.synchronize()
            t = time.perf_counter()
            kms = []
            for _ in range(steps):
                ns = s.sweep_resident_multi(modes)
                kms.append(s.last_kernel_ms())
            s.lib.esb_tables_wait(s.ctx, None)
            dt = (time.perf_counter() - t) / steps
            evals = len(modes) * k.size * W.size
            out.append({"config": name, "scheme": s.spec.scheme, "n_steps": int(s.model.n_steps), "modes": modes,
                        "evals_per_step": evals, "ms_per_step": 1e3 * dt, "evals_per_sec": evals / dt,
                        "kernel_ms": float(np.mean(kms)), "kernel_share": float(np.mean(kms)) / (1e3 * dt),
                        "brackets": int(sum(ns))})
    return out


SCAN_CONTRASTS = 20
SCAN_AMPLITUDES = 20


def strong_scaling_job(esb, local, dev, rank, world, steps=2):
    """BASELINE configs[4] as stated: a parameter scan of 1e9 D evaluations - 20 density contrasts x
    (n = 0, 1, 2) of the cylinder and 20 flow amplitudes x (sausage, kink) of the slab, each on a
    1000 k x 10000 omega grid - as a FIXED-SIZE job split over the GPUs: every rank sweeps all 40
    equilibria on its strided share of the wavenumbers (eigensolver_b200.scan), the accepted modes are
    gathered.  Host tables in, compact root tables in page-locked memory out, gather included."""
    import torch
    import torch.distributed as dist
    from eigensolver_b200.scan import density_flow_grid, gather_scan_modes_device, parameter_scan
    dens, flow = density_flow_grid(np.linspace(0.1, 0.4, SCAN_CONTRASTS), np.linspace(0.05, 0.9, SCAN_AMPLITUDES))
    k = np.linspace(K_RANGE[0], K_RANGE[1], NK)
    Wd = np.linspace(W_RANGE[0], W_RANGE[1], NW)
    Wf = np.linspace(-2.7, 2.7, NW)
    evals = (len(dens) * 3 + len(flow) * 2) * NK * NW

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with esb.DispersionSolver("cylinder_density", device=local) as sd, \
            esb.DispersionSolver("slab_flow", device=local) as sf:
        host = {}

        def job():
            """host tables in -> (model, mode, k row, omega) of every accepted mode in page-locked memory on
            rank 0 (N > 1: gathered over NCCL from the device buffers)"""
            n = 0
            for solver, pts, W, modes in ((sd, dens, Wd, [0, 1, 2]), (sf, flow, Wf, [0, 1])):
                if world > 1:
                    res = parameter_scan(solver, pts, k, W, modes, rank=rank, world=world, download=False)
                    g = gather_scan_modes_device(solver, res, dev)
                    if rank == 0:
                        if host.get("cap", 0) < g.shape[0]:
                            host["cap"] = int(g.shape[0] * 1.25) + 1024
                            host["buf"] = torch.empty((host["cap"], 4), dtype=torch.float64).pin_memory()
                        host["buf"][: g.shape[0]].copy_(g, non_blocking=True)
                        torch.cuda.current_stream(dev).synchronize()
                    n += g.shape[0]
                else:
                    res = parameter_scan(solver, pts, k, W, modes)
                    n += int(np.asarray(res.table["accepted"]).sum())
            return n
        job()                                   # warm-up: allocations, capacities
        sync()
        t = time.perf_counter()
        for _ in range(steps):
            n_modes = job()
        sync()
        dt = (time.perf_counter() - t) / steps
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt[0])
    return {"workload": "configs[4]: %d density contrasts x n=0,1,2 (cylinder) + %d flow amplitudes x sausage,kink "
                        "(slab), 1000 k x 10000 omega each; every rank sweeps all equilibria on k[rank::N]"
                        % (SCAN_CONTRASTS, SCAN_AMPLITUDES),
            "evals": evals, "seconds": dt, "evals_per_sec": evals / dt, "n_gpus": world, "steps": steps,
            "modes_gathered": int(n_modes), "scaling": "strong"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import eigensolver_b200 as esb
    from eigensolver_b200.distributed import gather_modes_device, shard_k

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: a world*NK wavenumber grid over K_RANGE, rank r owns rows r, r+world, ... (strided:
    # the number of modes grows with k, contiguous slabs would leave the step time to the busiest rank)
    k_all = np.linspace(K_RANGE[0], K_RANGE[1], NK * world)
    k, k_off, k_stride = shard_k(k_all, rank, world, layout="strided")
    W = np.linspace(W_RANGE[0], W_RANGE[1], NW)
    # pinned host staging for the end-to-end leg
    k_pin = torch.from_numpy(k.copy()).pin_memory()
    W_pin = torch.from_numpy(W.copy()).pin_memory()

    solver = esb.DispersionSolver("cylinder_density", n_steps=N_STEPS, scheme="rk8n", device=local)
    stream = torch.cuda.current_stream(dev)
    solver.set_stream(stream.cuda_stream)
    fp64_peak = solver.fp64_peak_tflops()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    _pin = {}

    def pinned_out(n):
        """page-locked host buffer for the gathered table (grown geometrically, reused across steps)"""
        if _pin.get("cap", 0) < n:
            _pin["cap"] = int(n * 1.25) + 1024
            _pin["buf"] = torch.empty((_pin["cap"], 3), dtype=torch.float64).pin_memory()
        return _pin["buf"][:n]

    def step_resident():
        """axes resident in HBM; root tables stay on the device until the gather."""
        ns = solver.sweep_resident_multi(MODES)
        kms = [solver.last_kernel_ms()]
        if world > 1:       # NCCL gather of the modes of all slots straight from the device buffers
            gather_modes_device(solver, len(MODES), k_off, dev, k_stride=k_stride)
        return kms, ns

    def step_e2e():
        """Host buffers in, root tables on the host out.  One GPU: the public host call
        (find_roots_multi = esb_upload_axes + esb_sweep_resident_multi + esb_roots_pinned: pinned k/omega
        in, the full root tables of the three modes out into page-locked buffers).
        Several GPUs: every rank uploads its pinned k/omega, sweeps, the modes (what the reference's
        sol_ks / sol_omegas hold) are gathered over NCCL and rank 0 copies the global table to the host."""
        h2d = k_pin.numel() * 8 + W_pin.numel() * 8
        if world == 1:
            tables = solver.find_roots_multi(MODES, k_pin.numpy(), W_pin.numpy(), pinned=True)
            d2h = sum(len(t.omega) * (8 * 3 + 4 * 4) for t in tables)
            return h2d, d2h, sum(int(t.accepted.sum()) for t in tables), sum(len(t.omega) for t in tables)
        solver.upload_axes(k_pin.numpy(), W_pin.numpy())
        ns = solver.sweep_resident_multi(MODES)
        g = gather_modes_device(solver, len(MODES), k_off, dev, k_stride=k_stride)
        d2h = 0
        if rank == 0:       # the parent process of the reference: it alone holds the collected lists
            host_buf = pinned_out(g.shape[0])
            host_buf.copy_(g, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            d2h = host_buf.numel() * 8
        return h2d, d2h, g.shape[0], sum(ns)

    # ---- device-resident timing ----
    solver.upload_axes(k, W)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = solver.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ms = []
    for _ in range(args.steps):
        kms, tables = step_resident()
        kernel_ms += kms
    ev1.record(stream)
    barrier()
    launches = solver.launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    # ---- end-to-end timing ----
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        h2d, d2h, n_modes, n_brackets = step_e2e()
    e1.record(stream)
    barrier()
    ms_e2e_dev = e0.elapsed_time(e1)
    ms_e2e = max(ms_e2e_dev, 1e3 * (time.perf_counter() - t_wall))   # host-side work counts too
    sampler.stop.set()
    sampler.join()

    # ---- brackets of rank 0 outside the continua (the region parity is claimed for), after the timing
    n_regular = None
    if rank == 0:
        iv = cylinder_continua(solver.medium, solver.profile)
        solver.upload_axes(k, W)
        ns = solver.sweep_resident_multi(MODES)
        n_regular = 0
        for slot, nn in enumerate(ns):
            tab = solver.download_roots(nn, slot)
            ok = np.ones(nn, bool)
            for lo, hi in iv:
                for Wp in (W[tab.w_index], W[tab.w_index + 1]):
                    ok &= (Wp < lo) | (Wp > hi)
            n_regular += int(ok.sum())
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    evals_per_step = len(MODES) * NK * NW * world
    value = evals_per_step * args.steps / (ms * 1e-3)
    e2e = evals_per_step * args.steps / (ms_e2e * 1e-3)

    guard = solver.guard_report()          # of the last sweep: the discretisation error the run carried
    solver.close()
    strong = None if args.no_extras else strong_scaling_job(esb, local, dev, rank, world)
    configs = other_configs(esb, local) if (world == 1 and not args.no_extras) else None
    if rank == 0:
        kms = float(np.mean(kernel_ms))                      # one fused launch = 3 modes x NK*NW evals
        achieved = FLOPS_FUSED_LAUNCH / (kms * 1e-3) * 1e-12
        nominal = 148 * 64 * 2 * 1.965e9 * 1e-12
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            with mp.get_context("fork").Pool(cores, initializer=_cpu_init) as pool:
                time_cpu(pool, cores, 2, cores, 1)
                n, dt = time_cpu(pool, cores, 16, max(8, 8 * cores), 5)      # ~15 s on all cores
                unhoisted = unhoisted_rate(pool, cores)
            cpu = {"value": n / dt, "unit": "evals/s", "cores": cores, "kind": "port", "unhoisted": unhoisted,
                   "sample": "16 k x %d omega x 3 modes, uniform random in the workload's box (%.0f s); oracle/"
                             "reference_path.py (scipy odeint + fsolve, the reference's algorithm)"
                             % (max(8, 8 * cores), dt)}
        line = {
            "metric": "dispersion_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "modes": list(MODES), "nk_per_gpu": NK, "nw": NW, "k_shards": "strided",
                       "n_steps": N_STEPS, "mesh": "graded", "scheme": "rk8n", "profile": "inverted Gaussian, width 0.95",
                       "l2": "working set 720 MB of (ext, int, Y) written per step > 126 MB L2; inputs are 88 KB"},
            "roots_per_sec": n_brackets * world * args.steps / (ms * 1e-3),
            "roots_per_sec_regular": n_regular * world * args.steps / (ms * 1e-3),
            "brackets_regular_rank0": n_regular,
            "modes_found": n_modes, "brackets_rank0": n_brackets,
            "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": ncu_traffic(),
                         "kernel": "grid_kernel<cylinder,rk8n,3 modes fused>", "kernel_ms": kms,
                         "flops_per_launch": FLOPS_FUSED_LAUNCH,
                         "fp64_instr_per_launch": FP64_INSTR_FUSED_LAUNCH,
                         "pipe_frac": FP64_INSTR_FUSED_LAUNCH / (kms * 1e-3) / (fp64_peak * 1e12 / 2.0),
                         "peak_source": "esb_fp64_peak: DFMA-chain kernel measured in this process "
                                        "(nominal 148 SM x 64 FMA/clk x 1.965 GHz = %.1f TFLOP/s)" % nominal,
                         "kernel_share_of_step": kms / (ms / args.steps)},
            "clocks": sampler.summary(),
            "guard": {"worst_deviation": guard["worst"], "samples_judged": guard["n_checked"],
                      "above_1e-9": guard["n_above"], "stride": guard["stride"],
                      "what": "built-in discretisation guard, inside the timed sweeps: (point, mode) samples "
                              "re-evaluated at 2 x n_steps on a side stream, outside the resonant continua"},
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if strong:
            line["strong_scaling"] = strong
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the configs[0,2,3] record and the configs[4] strong-scaling job")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
