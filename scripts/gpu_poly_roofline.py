"""BASELINE configs[3] at size: polyhedral box (hex box with 2/3 of its 2x2x2 blocks agglomerated into 24-face super-cells, jittered
vertices) -- Amul / DIC / PCG through the C-ABI: bitwise parity against the CPU oracle at full size and CUDA-event times."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import foam_oracle as fo, mesh_oracle as mo
from cases import oracle_mesh_of, random_spd
from shallowinterfoam_b200 import capi
import bench
lib = capi.load(); lib.init(0, 0, 1)
dims = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (240, 160, 160)
L = (24.0, 16.0, 16.0)
t0 = time.time()
box = lib.polymesh_box(*dims, *L, jitter=(0.25, 12345))
pm = box.agglomerate(mo.block_groups(*dims)); box.release()
N, F = pm.nCells, pm.nInternalFaces
m = oracle_mesh_of(pm)
dm = pm.device_mesh()      # with geometry: the sweep plan may classify faces by their normals
ldu = dm.ldu
band, rows_max, tile_start = ldu.tiling()
nfaces = np.bincount(pm.owner, minlength=N) + np.bincount(pm.neighbour, minlength=N)
out = {"mesh": "polyhedral box from %dx%dx%d hexes: %d cells, %d internal faces (%.2f per cell), faces per cell %d..%d" % (dims + (N, F, F / N, nfaces.min(), nfaces.max())),
       "build_s": time.time() - t0, "sweep_tiles": int(len(tile_start) - 1), "mean_rows_per_tile": float(N / max(1, len(tile_start) - 1))}
diag, upper = random_spd(m, seed=1)
rng = np.random.default_rng(2); psi = rng.uniform(0, 1, N); b = rng.uniform(-1, 1, N)
out["amul_bitwise"] = bool(np.array_equal(fo.amul(m, diag, upper, upper, psi), ldu.amul(diag, upper, psi)))
rD = ldu.dic_calc_rd(diag, upper)
out["rD_bitwise"] = bool(np.array_equal(fo.dic_calc_rd(m, diag, upper), rD))
out["precondition_bitwise"] = bool(np.array_equal(fo.dic_precondition(m, rD, upper, b), ldu.dic_precondition(rD, upper, b)))
rowCell, level, nLevels = ldu.renumbering()
out["wavefronts"] = int(nLevels)
lib.set_option("pcg_kernel_timers", 1)
lib.timers_reset()
x, perf, hist = ldu.pcg_solve(diag, upper, np.zeros(N), b, tol=1e-9, maxIter=400)
tm = lib.timers()
it = max(1, perf.nIterations)
peak = bench.measured_peak()[0]
alg = {"amul": 8 * (3 * N + F) + 8 * F, "dic_precondition": 48 * N + 32 * F}
t_amul = tm["amul"][0] / it * 1e-3; t_dic = (tm["dic_fwd"][0] + tm["dic_bwd"][0]) / it * 1e-3
out["pcg_iterations"] = int(perf.nIterations); out["final_residual"] = float(perf.finalResidual)
for k, t in (("amul", t_amul), ("dic_precondition", t_dic)):
    out[k] = {"us": 1e6 * t, "GBps": alg[k] / t / 1e9, "frac_of_measured_peak": alg[k] / t / 1e9 / peak}
t0 = time.time(); xo, po, ho = fo.pcg_solve(m, diag, upper, np.zeros(N), b, tol=1e-9, maxIter=400); out["oracle_pcg_s"] = time.time() - t0
out["pcg_iterations_equal"] = bool(po.nIterations == perf.nIterations)
out["pcg_history_1e-8"] = bool(np.allclose(hist, ho, rtol=1e-8, atol=1e-13 * ho[0]))
out["gpu_pcg_solve_ms"] = tm["pcg_solve"][0]
print(json.dumps(out))
