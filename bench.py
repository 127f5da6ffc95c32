#!/usr/bin/env python
"""bench.py -- 3D-region cell-steps/s (PISO+MULES) of the B200-native hot path, BASELINE.json's metric.

A "step" is one 3D-region time step (region3d/solve3d.H): nAlphaSubCycles MULES-limited alpha1 sub-cycles,
interface.correct(), UEqn assembly (momentumPredictor off), nCorrectors x (pd assembly + b200PCG/DIC solve + flux +
reconstruct), continuity errors, p = pd + rho*gh -- all device-resident, through the C-ABI of libb200foam.so.

N = 1 workload (BASELINE.json configs[1]): 3D-only hex channel with weir, 200x50x100 minus the weir block = 985 000
cells, laminar, PCG+DIC (pd: tol 1e-7 relTol 0.05; pdFinal: relTol 0), nCorrectors 3, nAlphaSubCycles 2.
N > 1: weak scaling -- the same 985k-cell slab per GPU (global mesh (200*N)x50x100, `simple` x-slab decomposition),
halo swaps + scalar all-reduces over NCCL.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU oracle on the host cores (reference arm)
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
DIMS = (200, 50, 100)
LENGTHS = (20.0, 5.0, 10.0)
WEIR = (0.45, 0.5, 0.3)
DELTA_T = 0.002


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


def set_controls(c):
    c.deltaT = DELTA_T; c.nCorr = 3; c.nNonOrthCorr = 0; c.nAlphaCorr = 1; c.nAlphaSubCycles = 2; c.cAlpha = 1.0
    c.rho1 = 1000.0; c.rho2 = 1.0; c.nu1 = 1e-6; c.nu2 = 1.48e-5; c.sigma = 0.07
    c.g[0] = 0.0; c.g[1] = 0.0; c.g[2] = -9.81
    c.pdTol = 1e-7; c.pdRelTol = 0.05; c.pdFinalTol = 1e-7; c.pdFinalRelTol = 0.0; c.pcorrTol = 1e-7; c.pcorrRelTol = 0.0
    c.maxIter = 1000; c.minIter = 0; c.adjustPhi = 1; c.mulesIncludeOwn = 0; c.divUScheme = 2
    return c


WORKLOAD = {"workload": "3D-only hex channel with weir, %dx%dx%d minus weir = 985000 cells per GPU (BASELINE configs[1])" % DIMS,
            "deltaT": DELTA_T, "nCorrectors": 3, "nAlphaSubCycles": 2, "nAlphaCorr": 1,
            "pd_solver": "b200PCG+DIC tol 1e-7 relTol 0.05 (pdFinal relTol 0)", "momentumPredictor": "no", "turbulence": "laminar"}


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


# ----------------------------------------------------------------------------------------------- reference arm (CPU oracle)
def build_oracle_region(dims, lengths, weir, n_threads=1):
    from oracle import foam_oracle as fo, mesh_oracle as mo
    mesh = mo.block_mesh_box(*dims, *lengths)
    if weir:
        mesh = mo.subset_mesh(mesh, mo.weir_keep_mask(*dims, *weir))
    g = mo.geometry(mesh)
    m = fo.Mesh(mesh["owner"], mesh["neighbour"], mesh["patches"], g, mesh["nCells"])
    names = [p[0] for p in mesh["patches"]]
    slices = {n: m.patch_slice(n) for n in names}
    cs = channel_fields(names, g["C"], g["Cf"][m.F:], slices, lengths)
    ctl = set_controls(fo.OrcControls())
    r = fo.Region(m, ctl, fo.BC(m, cs["tA"], 1, cs["rvA"]), fo.BC(m, cs["tU"], 3, cs["rvU"]), fo.BC(m, cs["tP"], 1),
                  cs["alpha"], cs["U"], cs["pd"])
    r.correct_phi(); r.clear_log()
    return m, r


def run_reference(args):
    """bench.py --impl reference: the reference's CPU algorithm (oracle port of foam-extend-3.1: /root/reference is only the
    40-line call scripts and foam-extend is not installable here -- DESIGN.md section 5) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3)); warm = min(args.warmup, 1)   # bounded sample: ~12 s of CPU per 985k-cell step
    m, r = build_oracle_region(DIMS, LENGTHS, WEIR)
    for _ in range(warm):
        r.step()
    r.clear_log()
    t0 = time.time()
    for _ in range(steps):
        r.step()
    dt = time.time() - t0
    val = m.N * steps / dt
    iters = sum(p[2] for p in r.perfs())
    sample = "%d warm-up + %d timed steps of the same 985000-cell workload (driver asked %d/%d; bounded to keep the run short), " \
             "%d PCG iterations, serial oracle" % (warm, steps, args.warmup, args.steps, iters)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": WORKLOAD,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------- our arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def pinned(n):
    try:
        import torch
        return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        return np.empty(n)


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

    # ---- mesh: this rank's slab of the (200*N) x 50 x 100 channel
    dims = (DIMS[0] * world, DIMS[1], DIMS[2]); lengths = (LENGTHS[0] * world, LENGTHS[1], LENGTHS[2])
    weir = (WEIR[0] / world, WEIR[0] / world + (WEIR[1] - WEIR[0]) / world, WEIR[2]) if world > 1 else WEIR
    gpm = lib.polymesh_box(*dims, *lengths, weir=weir)
    if world > 1:
        proc = gpm.decompose_simple((world, 1, 1))
        pm = gpm.submesh(proc, world, rank)
    else:
        pm = gpm
    g = pm.geometry()
    F = pm.nInternalFaces
    slices = {n: pm.patch_slice(n) for n in pm.patchNames}
    cs = channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, lengths)
    dm = pm.device_mesh()
    ctl = set_controls(capi.Region3dControls())
    reg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    reg.correct_phi(); reg.clear_log()
    N = pm.nCells
    total_cells = N
    if dist:
        import torch
        t = torch.tensor([N], dtype=torch.int64); dist.all_reduce(t); total_cells = int(t.item())

    def barrier():
        if dist:
            dist.barrier()

    # ---- value: device-resident steps, inputs already in HBM.  Timed on the device with CUDA events on the launching
    #      stream (library timer "step" = events around every step), max over ranks.
    for _ in range(args.warmup):
        reg.step()
    reg.clear_log()
    sampler = ClockSampler(local); sampler.start()
    barrier(); lib.timers_reset(); l0 = lib.launches(); t0 = time.time()
    reg.step(args.steps)
    tm = lib.timers(); wall = time.time() - t0; launches = lib.launches() - l0
    barrier()
    clocks = sampler.stop()
    dev_s = tm["step"][0] * 1e-3
    iters = [p[2] for p in reg.perfs()]
    if dist:
        import torch
        t = torch.tensor([dev_s, wall], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dev_s, wall = float(t[0]), float(t[1])
    value = total_cells * args.steps / dev_s

    # ---- e2e: the same step through b200_region3d_step_host with HOST buffers (pinned), H2D + D2H inside the timed region
    f = reg.fields()
    nF = F + pm.nBoundaryFaces
    hb = [pinned(N), pinned(3 * N), pinned(N), pinned(nF), pinned(N)]
    hb[0][:] = f["alpha1"]; hb[1][:] = f["U"].reshape(-1); hb[2][:] = f["pd"]; hb[3][:] = f["phi"]; hb[4][:] = f["p"]
    e2e_steps = max(2, min(args.steps, 5))
    reg.step_host(*hb)
    reg.clear_log()
    barrier(); t0 = time.time()
    for _ in range(e2e_steps):
        reg.step_host(*hb)
    e2e_s = time.time() - t0
    e2e_iters = sum(p[2] for p in reg.perfs()) / e2e_steps   # (later time steps than the `value` leg: the flow has settled a little)
    barrier()
    if dist:
        import torch
        t = torch.tensor([e2e_s], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t[0])
    e2e = {"value": total_cells * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * (5 * N + nF) * world,
           "d2h_bytes_per_step": 8 * (6 * N + nF) * world, "steps": e2e_steps, "pcg_iterations_per_step": e2e_iters,
           "api": "b200_region3d_step_host (host buffers in, host buffers out)"}

    # ---- roofline of the dominant kernel class, measured live with CUDA events (separate pass, per-kernel timers on)
    lib.set_option("pcg_kernel_timers", 1)
    lib.timers_reset()
    reg.clear_log()
    reg.step(1)
    tk = lib.timers()
    lib.set_option("pcg_kernel_timers", 0)
    peak, peak_src = measured_peak()

    # average duration of the launches that did work: the solver enqueues iterations in batches of 4, the ones queued past
    # convergence return at once, so per-iteration kernels are averaged over the iterations actually performed
    n_iter = max(1, sum(p[2] for p in reg.perfs()[-(ctl.nCorr * (ctl.nNonOrthCorr + 1)):]))

    def per_iter(name):
        ms, calls = tk.get(name, (0.0, 0))
        return (ms / n_iter * 1e-3) if calls else None

    def per_launch(name):
        ms, calls = tk.get(name, (0.0, 0))
        return (ms / calls * 1e-3) if calls else None

    alg = {"dic_precondition": 48 * N + 32 * F, "amul": 8 * (3 * N + F) + 8 * F, "mules": 424 * N + 288 * F}
    t_fwd, t_bwd, t_amul, t_mules = per_iter("dic_fwd"), per_iter("dic_bwd"), per_iter("amul"), per_launch("mules")
    step_ms = tk["step"][0]
    kernels = {}
    if t_fwd and t_bwd:
        kernels["dic_precondition"] = {"us": 1e6 * (t_fwd + t_bwd), "GBps": alg["dic_precondition"] / (t_fwd + t_bwd) / 1e9,
                                       "share_of_step": (tk["dic_fwd"][0] + tk["dic_bwd"][0]) / step_ms,
                                       "note": "k_dic_fwd (+ fused x/rA update and residual sum) + k_dic_bwd (+ fused wArA sum), per PCG iteration"}
    if t_amul:
        kernels["amul"] = {"us": 1e6 * t_amul, "GBps": alg["amul"] / t_amul / 1e9, "share_of_step": tk["amul"][0] / step_ms}
    if t_mules:
        kernels["mules_explicitSolve"] = {"us": 1e6 * t_mules, "GBps": alg["mules"] / t_mules / 1e9, "share_of_step": tk["mules"][0] / step_ms}
    for name in ("pcg_vec", "allreduce", "halo"):     # the rest of a PCG iteration (multi-rank: scalar all-reduces incl. the wait for the slowest rank; halo push + interface update)
        t = per_iter(name)
        if t:
            kernels[name] = {"us": 1e6 * t, "share_of_step": tk[name][0] / step_ms}
    dom = "dic_precondition" if "dic_precondition" in kernels else "amul"
    ach = kernels[dom]["GBps"] if dom in kernels else None
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
    # (profiles/r1f_ncu_full_raw.csv: k_dic_fwd 98.9 + 4.6 MB, k_dic_bwd 75.0 + 1.7 MB, k_amul 63.1 + 1.4 MB); only valid for N = 1
    ncu_traffic = {"dic_precondition": 180.2e6, "amul": 64.5e6}
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "frac_of_8TBps_nominal": (ach / 8000.0) if ach else None, "peak_source": peak_src,
                "traffic": ncu_traffic.get(dom) if world == 1 else None,
                "algorithmic_bytes_per_launch": alg[dom], "kernels": kernels,
                "pcg_iterations_in_pass": n_iter,
                "note": "DIC sweeps are bound by the dependency chain (348 wavefronts -> 35 tile hops of ~2.4 us), not by bandwidth; the "
                        "per-iteration working set (~180 MB: matrix, sweep operand blocks, 8 vectors) exceeds the 126 MB L2 -- see DESIGN.md"}

    if rank != 0:
        return
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(WORKLOAD, cells_total=total_cells, l2="inputs (~1 GB of fields+mesh per GPU) exceed the 126 MB L2"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "pcg_iterations_per_step": sum(iters) / max(args.steps, 1), "wall_ms_per_step": 1e3 * wall / args.steps}
    if world == 1 and not args.no_cpu_baseline:
        # the oracle ("port") on one host core, on a bounded sample of the same workload: 1 step (about 12 s of CPU)
        t0 = time.time()
        m, r = build_oracle_region(DIMS, LENGTHS, WEIR)
        t1 = time.time(); r.step(); dt = time.time() - t1
        line["cpu_baseline"] = {"value": m.N / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "1 step of the same 985000-cell workload after correctPhi (%.1f s; set-up %.1f s untimed), serial oracle" % (dt, t1 - t0)}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (repeatable)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
