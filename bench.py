#!/usr/bin/env python
"""bench.py -- 3D-region cell-steps/s (PISO+MULES) of the B200-native hot path, BASELINE.json's metric.

A "step" is one 3D-region time step (region3d/solve3d.H): nAlphaSubCycles MULES-limited alpha1 sub-cycles,
interface.correct(), UEqn assembly (momentumPredictor off), nCorrectors x (pd assembly + b200PCG/DIC solve + flux +
reconstruct), continuity errors, p = pd + rho*gh -- all device-resident, through the C-ABI of libb200foam.so.

Workloads
  --scaling weak (default)  N = 1: BASELINE.json configs[1], 3D-only hex channel with weir, 200x50x100 minus the weir block = 985 000
                            cells; N > 1: the same 985k-cell slab per GPU (global mesh (200*N)x50x100, `simple` x-slabs).
  --scaling strong          BASELINE.json configs[2]: ONE 400x100x200 channel with weir (7.88 M cells) decomposed over the N GPUs.

Every leg times the SAME time steps from the SAME state: fresh fields -> correctPhi -> W warm-up steps -> checkpoint -> K timed steps
(the window [W, W+K)).  `value` (device-resident), `e2e` (host buffers, restarted from the checkpoint), the cpu_baseline leg and
`--impl reference` (the CPU oracle, which replays correctPhi and the W warm-up steps itself) all run that window, so
pcg_iterations_per_step is the same work everywhere (the N-way CPU run has rank-local DIC like foam-extend under mpirun, hence its own
iteration counts; us_per_pcg_iteration is reported beside every throughput).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling weak|strong]      # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]                        # the reference's CPU algorithm on the host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version banner to stdout)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

METRIC = "3D-region cell-steps/s (PISO+MULES)"
UNIT = "cell-steps/s"
WEAK = dict(dims=(200, 50, 100), lengths=(20.0, 5.0, 10.0), name="BASELINE configs[1]")
STRONG = dict(dims=(400, 100, 200), lengths=(40.0, 10.0, 20.0), name="BASELINE configs[2]")
WEIR = (0.45, 0.5, 0.3)
DELTA_T = 0.002         # max Courant number 0.21 -> 0.34 over the timed window at 985k cells (mean 0.004): the flow accelerates to 3-4 m/s
                        # over the weir, so the survey's "Co ~ 0.2"  is met in the max norm.  Measured on the device
                        # (b200_region3d_get_courant; profiles/bench_r2a_dt.md): deltaT 0.005 gives max Co 0.83 -> 1.03, beyond what explicit
                        # MULES is run at.  Every line carries the Courant numbers of its own window.


# ----------------------------------------------------------------------------------------------- the synthetic case
def channel_fields(names, C, Cfb, patch_slices, lengths):
    """sharp wavy interface, water moving at 1 m/s; inlet fixedValue, outlet/atmosphere fixed pd."""
    lx, ly, lz = lengths
    alpha = np.where(C[:, 2] < 0.5 * lz + 0.05 * lz * np.sin(2 * np.pi * C[:, 0] / lx), 1.0, 0.0)
    U = np.zeros((C.shape[0], 3)); U[:, 0] = alpha
    pd = np.zeros(C.shape[0])
    FV, ZG = 0, 1
    tA, tU, tP = [], [], []
    for n in names:
        if n == "inlet": tA.append(FV); tU.append(FV); tP.append(ZG)
        elif n in ("outlet", "atmosphere"): tA.append(ZG); tU.append(ZG); tP.append(FV)
        elif n.startswith("procBoundary"): tA.append(3); tU.append(3); tP.append(3)
        else: tA.append(ZG); tU.append(FV); tP.append(ZG)
    B = Cfb.shape[0]
    rvA = np.zeros(B); rvU = np.zeros((B, 3))
    if "inlet" in patch_slices:
        sl = patch_slices["inlet"]
        rvA[sl] = np.where(Cfb[sl, 2] < 0.5 * lz, 1.0, 0.0)
        rvU[sl, 0] = rvA[sl]
    return dict(alpha=alpha, U=U, pd=pd, tA=tA, tU=tU, tP=tP, rvA=rvA, rvU=rvU)


def set_controls(c, delta_t=DELTA_T):
    c.deltaT = delta_t; c.nCorr = 3; c.nNonOrthCorr = 0; c.nAlphaCorr = 1; c.nAlphaSubCycles = 2; c.cAlpha = 1.0
    c.rho1 = 1000.0; c.rho2 = 1.0; c.nu1 = 1e-6; c.nu2 = 1.48e-5; c.sigma = 0.07
    c.g[0] = 0.0; c.g[1] = 0.0; c.g[2] = -9.81
    c.pdTol = 1e-7; c.pdRelTol = 0.05; c.pdFinalTol = 1e-7; c.pdFinalRelTol = 0.0; c.pcorrTol = 1e-7; c.pcorrRelTol = 0.0
    c.maxIter = 1000; c.minIter = 0; c.adjustPhi = 1; c.mulesIncludeOwn = 0; c.divUScheme = 2
    c.momentumPredictor = 0; c.nonOrthCorrected = 0; c.pdRefCell = -1; c.pdRefValue = 0.0
    return c


def problem(args, world):
    """global mesh of this run: (dims, lengths, weir, decomposition, description)"""
    if args.scaling == "strong":
        dims, lengths = STRONG["dims"], STRONG["lengths"]
        weir = WEIR
        desc = "3D hex channel with weir, %dx%dx%d minus weir = 7 880 000 cells decomposed over the GPUs (%s)" % (dims + (STRONG["name"],))
    else:
        dims = (WEAK["dims"][0] * world, WEAK["dims"][1], WEAK["dims"][2])
        lengths = (WEAK["lengths"][0] * world, WEAK["lengths"][1], WEAK["lengths"][2])
        weir = (WEIR[0] / world, WEIR[0] / world + (WEIR[1] - WEIR[0]) / world, WEIR[2]) if world > 1 else WEIR
        desc = "3D-only hex channel with weir, %dx%dx%d minus weir = 985000 cells per GPU (%s)" % (WEAK["dims"] + (WEAK["name"],))
    return dims, lengths, weir, desc


def decomposition(args, dims, world):
    """`simple` decomposition (n_x n_y n_z) of decomposeParDict: x-slabs unless `--decomp a,b,c` says otherwise.  Weak scaling: each rank keeps
    the 200x50x100 block.  Strong scaling: x-slabs too -- measured at N=8 on 400x100x200 (profiles/bench_r2c_n8.json against
    bench_r2b_strong_n8.json): (4,1,2) = 100^3 blocks has the shorter wavefront chain (298 against 348) and a 7% shorter iteration (280 us
    against 299 us), but the rank-local DIC of the extra z cut costs ~9% more iterations, so the step is no faster."""
    if args.decomp:
        d = tuple(int(v) for v in args.decomp.split(","))
        if len(d) != 3 or d[0] * d[1] * d[2] != world:
            raise SystemExit("--decomp %s does not multiply to %d ranks" % (args.decomp, world))
        return d
    return (world, 1, 1)


def polyhedral_leg(lib, dims=(300, 180, 180), lengths=(30.0, 18.0, 18.0)):
    """BASELINE configs[3] on one GPU: per-kernel CUDA-event times of Amul and the DIC precondition inside one PCG solve of a random SPD
    system on the polyhedral box (what `solver b200PCG;` does on irregular lduAddressing).  Parity of this path: tests/test_polyhedral.py,
    tests/test_irregular_ldu.py and tests/test_atsize_region.py (2.56 M cells, against the oracle)."""
    from shallowinterfoam_b200 import capi
    t0 = time.time()
    box = lib.polymesh_box(*dims, *lengths, jitter=(0.25, 12345))
    pm = box.agglomerate(capi.checkerboard_groups(*dims)); box.release()
    N, F = pm.nCells, pm.nInternalFaces
    ldu = pm.ldu()
    build_s = time.time() - t0
    rng = np.random.default_rng(1)
    upper = -rng.uniform(0.5, 1.5, F)
    diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += rng.uniform(0.01, 0.1, N)
    b = rng.uniform(-1, 1, N)
    ldu.pcg_solve(diag, upper, np.zeros(N), b, tol=1e-9, maxIter=5)          # warm-up (allocations)
    lib.set_option("pcg_kernel_timers", 1); lib.timers_reset()
    x, perf, hist = ldu.pcg_solve(diag, upper, np.zeros(N), b, tol=1e-9, maxIter=400)
    tm = lib.timers()
    lib.set_option("pcg_kernel_timers", 0)
    it = max(1, perf.nIterations)
    res = b - ldu.amul(diag, upper, x)                                        # true residual of the solution, by the product's own Amul
    peak = measured_peak()[0]
    out = {"workload": "polyhedral box from %dx%dx%d hexes (2/3 of the 2x2x2 blocks merged into 24-face cells, vertices jittered by 0.25 of the spacing): "
                       "%d cells, %d internal faces (BASELINE configs[3])" % (dims + (N, F)),
           "cells": N, "faces_per_cell": 2.0 * F / N, "wavefronts": int(ldu.renumbering()[2]), "sweep_tiles": int(len(ldu.tiling()[2]) - 1),
           "setup_s": build_s, "pcg_iterations": int(perf.nIterations), "converged": bool(perf.converged), "pcg_solve_ms": tm["pcg_solve"][0],
           "true_residual_L1_over_b_L1": float(np.abs(res).sum() / np.abs(b).sum())}
    for k, t_ms, alg in (("amul", tm["amul"][0], 8 * (3 * N + F) + 8 * F), ("dic_precondition", tm["dic_fwd"][0] + tm["dic_bwd"][0], 48 * N + 32 * F)):
        t = t_ms * 1e-3 / it
        out[k] = {"us": 1e6 * t, "algorithmic_bytes": alg, "GBps": alg / t / 1e9, "frac_of_measured_peak": alg / t / 1e9 / peak}
    ldu.release(); pm.release()
    return out


def workload(args, desc):
    return {"workload": desc, "deltaT": args.delta_t, "Co_target": "max Co ~0.2-0.35 (see `courant`)", "nCorrectors": 3, "nAlphaSubCycles": 2, "nAlphaCorr": 1,
            "pd_solver": "b200PCG+DIC tol 1e-7 relTol 0.05 (pdFinal relTol 0)", "momentumPredictor": "no", "turbulence": "laminar",
            "window": "fresh fields -> correctPhi -> %d warm-up steps -> the %d timed steps (every leg and both arms time these same steps)"
                      % (args.warmup, args.steps)}


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device=0):
        self.rows = []; self.proc = None; self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- the CPU arm (oracle)
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_window(args, dims, lengths, weir, n_proc, warm, steps, budget_s):
    """The reference's CPU algorithm (oracle/: restatement of foam-extend-3.1, which is not installable here) on the same window:
    correctPhi -> `warm` steps -> `steps` timed steps, decomposed `n_proc` ways (one thread per `simple` sub-domain, rank-local DIC,
    halo swaps and rank-ordered global sums: what `mpirun -np n_proc shallowInterFoam -parallel` does).  The timed steps are cut short
    when the projected time exceeds budget_s; the sample says how many ran."""
    from oracle import cpu_run
    return cpu_run.run_window(dims, lengths, weir, channel_fields, lambda c: set_controls(c, args.delta_t), n_proc, warm, steps, budget_s)


def run_reference(args):
    """bench.py --impl reference: the reference's CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    dims, lengths, weir, desc = problem(args, world)
    n_proc = args.cpu_procs or host_cores()
    r = cpu_window(args, dims, lengths, weir, n_proc, args.warmup, args.steps, budget_s=args.cpu_budget or 420.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(workload(args, desc), cells_total=r["cells"]),
            "pcg_iterations_per_step": r["pcg_iterations_per_step"], "us_per_pcg_iteration": r["us_per_pcg_iteration"],
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the N = 1 workload
    (profiles/ncu_traffic.json: {"workload": ..., "kernels": {name: bytes}}); None when the file is absent"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None


def pinned(n):
    try:
        import torch
        return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        return np.empty(n)


def preflight(lib, rank, world, dist):
    """untimed: the decomposed small case on these N GPUs against the N-rank / serial CPU oracle, over NVLink peer memory (default) and
    over NCCL (p2p_*=0).  The oracle is the CHECKER here; nothing of it runs in a timed region."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from multirank_worker import decomposed_checks
    out = {}
    for tag, opts in (("peer_memory", {"p2p_halo": 1, "p2p_allreduce": 1}), ("nccl", {"p2p_halo": 0, "p2p_allreduce": 0})):
        for k, v in opts.items():
            lib.set_option(k, v)
        try:
            res = decomposed_checks(lib, rank, world)
            res["ok"] = True
        except AssertionError as e:
            res = {"ok": False, "error": str(e)[:300]}
        ok = [res["ok"]]
        if dist:
            import torch
            t = torch.tensor([1 if res["ok"] else 0]); dist.all_reduce(t, op=dist.ReduceOp.MIN); ok = [bool(t.item())]
        res["ok_all_ranks"] = ok[0]
        out[tag] = res
    lib.set_option("p2p_halo", 1); lib.set_option("p2p_allreduce", 1)
    summary = {"amul_bitwise": all(o.get("amul_bitwise", False) for o in out.values()),
               "pcg_hist_1e-8": max(o.get("pcg_hist_rel", 1.0) for o in out.values()),
               "fields_linf": max(o.get("fields_linf", 1.0) for o in out.values()),
               "ok": all(o["ok_all_ranks"] for o in out.values()), "transports": out,
               "case": "16x4x8 weir channel over %d ranks: Amul vs N-rank oracle, 2 consecutive PCG solves, correctPhi + 2 region steps vs the serial oracle" % world}
    return summary


def build_case(lib, args, world, rank, mixed_alpha_patch=None):
    """this rank's sub-mesh of the global channel on the device + the 3D region with the synthetic fields.  mixed_alpha_patch: name of a
    patch whose alpha1 condition becomes `mixed` (the sif* coupling conditions are mixed BCs evaluated by the host)"""
    from shallowinterfoam_b200 import capi
    dims, lengths, weir, desc = problem(args, world)
    gpm = lib.polymesh_box(*dims, *lengths, weir=weir)
    if world > 1:
        proc = gpm.decompose_simple(decomposition(args, dims, world))
        pm = gpm.submesh(proc, world, rank)
        gpm.release()
    else:
        pm = gpm
    g = pm.geometry()
    F = pm.nInternalFaces
    slices = {n: pm.patch_slice(n) for n in pm.patchNames}
    cs = channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, lengths)
    if mixed_alpha_patch in pm.patchNames:
        cs["tA"][pm.patchNames.index(mixed_alpha_patch)] = 2
    dm = pm.device_mesh()
    ctl = set_controls(capi.Region3dControls(), args.delta_t)
    reg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    return pm, dm, reg, cs, ctl, dims, lengths, desc


def run_ours(args):
    try:
        import torch  # noqa: F401  (first, so libb200foam.so binds to the NCCL torch bundles; pinned host buffers; torch.distributed plumbing)
    except Exception:
        pass
    from shallowinterfoam_b200 import capi
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = capi.load()
    for o in args.opt:          # tuning knobs of the library (b200_set_option), e.g. --opt p2p_allreduce=0
        k, v = o.split("="); lib.set_option(k, float(v))
    dist = None
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("gloo")      # plumbing only: broadcasts the ncclUniqueId and reduces the timings
        ids = [lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    lib.init(local, rank, world, nccl_id)

    def barrier():
        if dist:
            dist.barrier()

    def reduce_max(vals):
        if not dist:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    parity = None
    if world > 1 and not args.no_preflight:
        parity = preflight(lib, rank, world, dist)
        if not parity["ok"]:
            if rank == 0:
                emit({"metric": METRIC, "error": "multi-GPU parity pre-flight failed", "parity_check": parity})
            sys.exit(3)

    # ---- mesh + region: this rank's part of the global channel
    pm, dm, reg, cs, ctl, dims, lengths, desc = build_case(lib, args, world, rank)
    F = pm.nInternalFaces
    reg.correct_phi(); reg.clear_log()
    N = pm.nCells
    total_cells = N
    if dist:
        import torch
        t = torch.tensor([N], dtype=torch.int64); dist.all_reduce(t); total_cells = int(t.item())
    K = args.steps

    # ---- W warm-up steps, then the checkpoint every leg restarts from
    for _ in range(args.warmup):
        reg.step()
    ck = reg.fields()
    courant0 = reg.courant()

    # ---- value: device-resident steps, inputs already in HBM.  Timed on the device with CUDA events on the launching
    #      stream (library timer "step" = events around every step), max over ranks.
    reg.clear_log()
    sampler = ClockSampler(local); sampler.start()
    barrier(); lib.timers_reset(); l0 = lib.launches(); t0 = time.time()
    reg.step(K)
    tm = lib.timers(); wall = time.time() - t0; launches = lib.launches() - l0
    barrier()
    clocks = sampler.stop()
    dev_s = tm["step"][0] * 1e-3
    iters_value = [p[2] for p in reg.perfs()]
    courant1 = reg.courant()
    dev_s, wall = reduce_max([dev_s, wall])
    value = total_cells * K / dev_s

    # ---- e2e: the SAME K steps from the checkpoint through b200_region3d_step_host with HOST buffers (pinned), H2D + D2H inside the timed region
    nF = F + pm.nBoundaryFaces
    hb = [pinned(N), pinned(3 * N), pinned(N), pinned(nF), pinned(N)]
    reg.restore(ck["alpha1"], ck["U"], ck["pd"], ck["phi"])
    hb[0][:] = ck["alpha1"]; hb[1][:] = ck["U"].reshape(-1); hb[2][:] = ck["pd"]; hb[3][:] = ck["phi"]; hb[4][:] = ck["p"]
    reg.clear_log()
    barrier(); t0 = time.time()
    for _ in range(K):
        reg.step_host(*hb)
    e2e_s = time.time() - t0
    iters_e2e = [p[2] for p in reg.perfs()]
    barrier()
    e2e_s, = reduce_max([e2e_s])
    e2e = {"value": total_cells * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * (5 * N + nF) * world,
           "d2h_bytes_per_step": 8 * (6 * N + nF) * world, "steps": K, "pcg_iterations_per_step": sum(iters_e2e) / K,
           "same_iterations_as_value_leg": iters_e2e == iters_value,
           "api": "b200_region3d_step_host (host buffers in, host buffers out), restarted from the same checkpoint as `value`"}

    # ---- roofline of the dominant kernel class, measured live with CUDA events (separate pass: first step of the window again,
    #      per-kernel timers on) -- and, at N = 1, the plug-point replay of that step (e2e_plugin)
    reg.restore(ck["alpha1"], ck["U"], ck["pd"], ck["phi"])
    lib.set_option("pcg_kernel_timers", 1)
    lib.timers_reset()
    reg.clear_log()
    reg.step(1)
    tk = lib.timers()
    lib.set_option("pcg_kernel_timers", 0)
    peak, peak_src = measured_peak()
    n_solves = ctl.nCorr * (ctl.nNonOrthCorr + 1)
    pass_perfs = reg.perfs()[-n_solves:]
    n_iter = max(1, sum(p[2] for p in pass_perfs))

    def per_iter(name):      # per-iteration kernels: averaged over the iterations actually performed (launches queued past convergence return at once)
        ms, calls = tk.get(name, (0.0, 0))
        return (ms / n_iter * 1e-3) if calls else None

    def per_launch(name):
        ms, calls = tk.get(name, (0.0, 0))
        return (ms / calls * 1e-3) if calls else None

    alg = {"dic_precondition": 48 * N + 32 * F, "amul": 8 * (3 * N + F) + 8 * F, "mules": 424 * N + 288 * F}
    t_fwd, t_bwd, t_amul, t_mules = per_iter("dic_fwd"), per_iter("dic_bwd"), per_iter("amul"), per_launch("mules")
    step_ms = tk["step"][0]
    traffic = ncu_traffic() if (world == 1 and args.scaling == "weak") else None
    tr = (traffic or {}).get("kernels", {})
    kernels = {}

    def add(name, t, share_ms, note=None, alg_bytes=None):
        k = {"us": 1e6 * t, "share_of_step": share_ms / step_ms}
        if alg_bytes:
            k["GBps"] = alg_bytes / t / 1e9
            k["frac_of_measured_peak"] = k["GBps"] / peak
        if name in tr:     # real DRAM traffic of this kernel class per launch (ncu) over the live duration
            k["dram_bytes_ncu"] = tr[name]; k["dram_GBps"] = tr[name] / t / 1e9
        if note:
            k["note"] = note
        kernels[name] = k

    if t_fwd and t_bwd:
        add("dic_precondition", t_fwd + t_bwd, tk["dic_fwd"][0] + tk["dic_bwd"][0],
            "k_dic_fwd (+ fused x/rA update and residual sum) + k_dic_bwd (+ fused wArA sum), per PCG iteration", alg["dic_precondition"])
    if t_amul:
        add("amul", t_amul, tk["amul"][0], None, alg["amul"])
    if t_mules:
        add("mules_explicitSolve", t_mules, tk["mules"][0],
            "GBps counts upstream's 14-pass byte structure (424N+288F); the 6 fused kernels move about half of it: see dram_GBps", alg["mules"])
    for name in ("pcg_vec", "allreduce", "halo"):     # the rest of a PCG iteration (multi-rank: scalar all-reduces incl. the wait for the slowest rank; halo push + interface update)
        t = per_iter(name)
        if t:
            add(name, t, tk[name][0])
    dom = "dic_precondition" if "dic_precondition" in kernels else "amul"
    ach = kernels[dom]["GBps"] if dom in kernels else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "frac_of_8TBps_nominal": (ach / 8000.0) if ach else None, "peak_source": peak_src,
                "traffic": tr.get(dom), "traffic_source": (traffic or {}).get("source"),
                "algorithmic_bytes_per_launch": alg[dom], "kernels": kernels,
                "pcg_iterations_in_pass": n_iter,
                "note": "DIC sweeps are bound by the dependency chain (wavefronts -> tile hops), not by bandwidth -- see DESIGN.md section 4.1"}

    # ---- e2e_plugin (N = 1): what `solver b200PCG;` + the MULES replacement deliver to an unmodified shallowInterFoam -- every linear
    #      system and MULES call of that step replayed through the two plug points from PAGEABLE host arrays (H2D/D2H inside)
    e2e_plugin = None
    if world == 1:
        reg.restore(ck["alpha1"], ck["U"], ck["pd"], ck["phi"])      # the same step once more, keeping host copies of what the plug points see
        reg.capture(True)
        reg.step(1)
        solves, mules = reg.captured_solves(), reg.captured_mules()
        reg.capture(False)
        ldu = dm.ldu
        bcA = dm.bc(cs["tA"], 1, cs["rvA"])
        t_solve, it_solve, t_mul = [], [], []
        import ctypes as C
        dpt = C.POINTER(C.c_double)

        def ptr(a):
            return a.ctypes.data_as(dpt)

        def plug_pcg(s, timed=True):
            # exactly the adapter's call: b200_pcg_solve(ldu, diag, upper, NULL, NULL, x, b, ctl, perf, NULL, 0) on pageable arrays
            x = s["x0"].copy(); ctl_ = capi.SolverControls(s["tol"], s["relTol"], ctl.minIter, ctl.maxIter); perf = capi.SolverPerf()
            t0 = time.time()
            lib.check(lib.c.b200_pcg_solve(ldu.h, ptr(s["diag"]), ptr(s["upper"]), None, None, ptr(x), ptr(s["source"]), C.byref(ctl_), C.byref(perf), None, C.c_int32(0)))
            return time.time() - t0, int(perf.nIterations)

        def plug_mules(q):
            psi = q["psi"].copy(); phiPsi = q["phiPsi"].copy()
            t0 = time.time()
            lib.check(lib.c.b200_mules_explicit_solve(dm.h, bcA.ref, ptr(psi), ptr(q["psib"]), ptr(q["psi0"]), ptr(q["phi"]), ptr(phiPsi),
                                                      C.c_double(q["deltaT"]), C.c_double(1.0), C.c_double(0.0), None))
            return time.time() - t0

        if solves:      # one untimed call of each plug point: first-use allocations of the host-pointer staging buffers
            plug_pcg(solves[0])
        if mules:
            plug_mules(mules[0])
        for s in solves:      # every call three times, the median counts (host-side copies on a shared VM are noisy)
            r3 = sorted(plug_pcg(s) for _ in range(3))
            t_solve.append(r3[1][0]); it_solve.append(r3[1][1])
        for q in mules:
            t_mul.append(sorted(plug_mules(q) for _ in range(3))[1])
        plug_s = sum(t_solve) + sum(t_mul)
        resident_ms = tk["pcg_solve"][0] + tk["mules"][0]
        e2e_plugin = {"ms_per_step_in_plugin_calls": 1e3 * plug_s, "pcg_solve_ms": [1e3 * t for t in t_solve], "pcg_iterations": it_solve,
                      "same_iterations_as_resident": it_solve == [p[2] for p in pass_perfs],
                      "mules_ms": [1e3 * t for t in t_mul], "resident_ms_same_calls": resident_ms,
                      "h2d_bytes_per_step": len(solves) * 8 * (3 * N + F) + len(mules) * 8 * (2 * N + pm.nBoundaryFaces + 2 * nF),
                      "d2h_bytes_per_step": len(solves) * 8 * N + len(mules) * 8 * (N + nF),
                      "cell_steps_per_s_of_the_plugged_part": N / plug_s,
                      "timing": "median of 3 calls each", "host_copy_threads": int(float(dict(o.split("=", 1) for o in args.opt).get("host_copy_threads", 4))),
                      "api": "b200_pcg_solve x%d + b200_mules_explicit_solve x%d of one time step, pageable host arrays in foam-extend order "
                             "(what foam/b200PCG + foam/b200MULES call); the rest of the step stays in foam-extend on the host and is not timed"
                             % (len(solves), len(mules))}

    # ---- e2e_coupled: the same K steps with a HOST boundary-condition callback every step, the way the sif* coupling conditions work
    #      [REF shallowInterFoam.C:109-128; sifAlpha1FvPatchScalarField.C:431-435]: alpha1 on the outlet patch is a `mixed` condition
    #      whose refValue / refGrad / valueFraction the host evaluates from the patch-internal alpha1 and U it reads back (KBs, not fields).
    #      valueFraction = 0 and refGrad = 0 keep the physics of the zeroGradient outlet, so the iteration counts stay comparable.
    reg2 = dm.region3d(ctl, dm.bc([2 if n == "outlet" else t for n, t in zip(pm.patchNames, cs["tA"])], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]),
                       dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    reg2.restore(ck["alpha1"], ck["U"], ck["pd"], ck["phi"])
    p_out = pm.patchNames.index("outlet"); n_out = int(pm.patchSize[p_out])
    reg2.clear_log()
    barrier(); t0 = time.time(); t_bc = 0.0
    for _ in range(K):
        tb = time.time()
        if n_out:
            a_int = reg2.patch_internal(0, p_out); u_int = reg2.patch_internal(1, p_out).reshape(-1, 3)
            ref_val = np.clip(a_int, 0.0, 1.0) * (u_int[:, 0] >= 0)                  # (a stand-in for the sif formulae: any host function of them)
            reg2.set_patch_mixed(0, p_out, ref_val, np.zeros(n_out), np.zeros(n_out))
        t_bc += time.time() - tb
        reg2.step(1)
    cpl_s = time.time() - t0
    iters_cpl = [p[2] for p in reg2.perfs()]
    barrier()
    cpl_s, t_bc = reduce_max([cpl_s, t_bc])
    e2e_coupled = {"value": total_cells * K / cpl_s, "unit": UNIT, "steps": K, "host_bc_ms_per_step": 1e3 * t_bc / K,
                   "d2h_bytes_per_step": 8 * 4 * n_out, "h2d_bytes_per_step": 8 * 3 * n_out, "pcg_iterations_per_step": sum(iters_cpl) / K,
                   "same_iterations_as_value_leg": iters_cpl == iters_value,
                   "api": "b200_region3d_get_patch_internal + b200_region3d_set_patch_mixed on the outlet patch every step, then b200_region3d_step(1)"}
    reg2.release()

    # ---- strong-scaling leg (default run only): BASELINE configs[2], ONE 7.88 M-cell channel decomposed over the same N GPUs
    strong = None
    if args.scaling == "weak" and not args.no_strong_leg:
        reg.release(); dm.release(); pm.release()
        sargs = argparse.Namespace(**vars(args)); sargs.scaling = "strong"
        spm, sdm, sreg, scs, sctl, sdims, slengths, sdesc = build_case(lib, sargs, world, rank)
        sreg.correct_phi()
        for _ in range(2):
            sreg.step()
        sreg.clear_log()
        barrier(); lib.timers_reset()
        sreg.step(3)
        stm = lib.timers(); barrier()
        s_dev, = reduce_max([stm["step"][0] * 1e-3])
        s_iters = sum(p[2] for p in sreg.perfs())
        s_cells = spm.nCells
        if dist:
            import torch
            t = torch.tensor([s_cells], dtype=torch.int64); dist.all_reduce(t); s_cells = int(t.item())
        strong = {"workload": sdesc, "cells_total": s_cells, "value": s_cells * 3 / s_dev, "unit": UNIT, "steps": 3, "warmup": 2,
                  "ms_per_step": 1e3 * s_dev / 3, "pcg_iterations_per_step": s_iters / 3, "us_per_pcg_iteration": 1e6 * s_dev / max(s_iters, 1),
                  "scaling": "strong", "decomposition": "simple %s" % (decomposition(sargs, sdims, world),) if world > 1 else "none"}
        sreg.release(); sdm.release(); spm.release()

    # ---- polyhedral leg (default run, N = 1): BASELINE configs[3], a 4.05 M-cell polyhedral box (hexes + 24-face super-cells, jittered
    #      vertices): the LDU kernels on irregular addressing -- Amul, DIC precondition (chunked sweep tiles), one PCG solve
    poly = None
    if world == 1 and args.scaling == "weak" and not args.no_poly_leg:
        poly = polyhedral_leg(lib)

    if rank != 0:
        return
    its = sum(iters_value)
    desc_cfg = dict(workload(args, desc), cells_total=total_cells, l2="inputs (~1 GB of fields+mesh per GPU) exceed the 126 MB L2")
    if world > 1:
        desc_cfg["decomposition"] = "simple %s" % (decomposition(args, dims, world),)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": desc_cfg,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "pcg_iterations_per_step": its / max(K, 1), "us_per_pcg_iteration": 1e6 * dev_s / max(its, 1),
            "wall_ms_per_step": 1e3 * wall / K,
            "courant": {"mean_start": courant0[0], "max_start": courant0[1], "mean_end": courant1[0], "max_end": courant1[1],
                        "api": "b200_region3d_get_courant (device reduction)"}}
    line["e2e_coupled"] = e2e_coupled
    if strong:
        line["strong_scaling_leg"] = strong
    if poly is not None:
        line["polyhedral_leg"] = poly
    if e2e_plugin:
        line["e2e_plugin"] = e2e_plugin
    if parity:
        line["parity_check"] = parity
    if world == 1 and not args.no_cpu_baseline:
        # the N-way decomposed oracle ("port") on the box's host cores, same window, bounded to ~30 s of timed CPU work
        dims1, lengths1, weir1, _ = problem(args, 1)
        n_proc = args.cpu_procs or host_cores()
        r = cpu_window(args, dims1, lengths1, weir1, n_proc, args.warmup, K, budget_s=args.cpu_budget or 45.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                "pcg_iterations_per_step": r["pcg_iterations_per_step"], "us_per_pcg_iteration": r["us_per_pcg_iteration"]}
    emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: file descriptor 1 is pointed at stderr for the rest of the process (NCCL
    prints its version banner with printf, torch.distributed and CUDA libraries may chat too), the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n"); out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--decomp", default="", help="`simple` decomposition a,b,c (default: x-slabs)")
    ap.add_argument("--delta-t", dest="delta_t", type=float, default=DELTA_T)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-preflight", action="store_true")
    ap.add_argument("--no-poly-leg", action="store_true", help="skip the configs[3] polyhedral leg of the default run")
    ap.add_argument("--no-strong-leg", action="store_true", help="skip the configs[2] strong-scaling leg of the default (weak) run")
    ap.add_argument("--cpu-procs", type=int, default=0, help="sub-domains (threads) of the CPU arm; default = host cores")
    ap.add_argument("--cpu-budget", type=float, default=0.0, help="seconds of timed CPU work before the CPU arm stops early")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
