"""GPU probe: per-kernel-class device times of Amul / DIC sweeps / PCG iteration on box meshes (random SPD matrix)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi

lib = capi.load(); lib.init(0, 0, 1)
sizes = [(50, 10, 20), (200, 50, 100)] + ([(400, 100, 200)] if "--big" in sys.argv else [])
for dims in sizes:
    t0 = time.time()
    pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
    ldu = pm.ldu()
    N, F = pm.nCells, pm.nInternalFaces
    rowCell, level, nLevels = ldu.renumbering()
    print("mesh %s N=%d F=%d levels=%d build %.1fs" % (dims, N, F, nLevels, time.time() - t0), flush=True)
    rng = np.random.default_rng(0)
    upper = -rng.uniform(0.5, 1.5, F)
    diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
    b = rng.uniform(-1, 1, N)
    rD = ldu.dic_calc_rd(diag, upper)
    for rep in range(2):
        lib.timers_reset()
        for i in range(10):
            ldu.amul(diag, upper, b)
        for i in range(10):
            ldu.dic_precondition(rD, upper, b)
        t0 = time.time()
        x, perf, hist = ldu.pcg_solve(diag, upper, np.zeros(N), b, tol=1e-30, maxIter=100)
        wall = time.time() - t0
        tm = lib.timers()
    amul_b = 8 * (3 * N + F) + 8 * F
    dic_b = 48 * N + 32 * F
    print("  amul     %.1f us  -> %.0f GB/s algorithmic" % (1e3 * tm["amul"][0] / tm["amul"][1], amul_b / (tm["amul"][0] / tm["amul"][1] * 1e-3) / 1e9))
    f = tm["dic_fwd"][0] / tm["dic_fwd"][1]; bw = tm["dic_bwd"][0] / tm["dic_bwd"][1]
    print("  dic fwd  %.1f us  bwd %.1f us -> %.0f GB/s algorithmic ; %.2f us/level fwd" % (1e3 * f, 1e3 * bw, dic_b / ((f + bw) * 1e-3) / 1e9, 1e3 * f / nLevels))
    print("  pcg      %d iters, solve %.2f ms device (%.1f us/iter), wall %.2f ms incl. H2D/D2H" % (perf.nIterations, tm["pcg_solve"][0], 1e3 * tm["pcg_solve"][0] / max(perf.nIterations, 1), 1e3 * wall))
    print("  residual %.3e -> %.3e" % (perf.initialResidual, perf.finalResidual), flush=True)
    ldu.release(); pm.release()
