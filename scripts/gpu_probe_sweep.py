"""Scan the CTA count of the data-flow DIC sweeps (per-level latency vs polling pressure)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
for dims in [(200, 50, 100), (400, 100, 200)]:
    pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
    ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
    _, _, nLevels = ldu.renumbering()
    rng = np.random.default_rng(0)
    upper = -rng.uniform(0.5, 1.5, F)
    diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
    b = rng.uniform(-1, 1, N)
    rD = ldu.dic_calc_rd(diag, upper)
    ref = ldu.dic_precondition(rD, upper, b)
    print("mesh", dims, "N", N, "levels", nLevels, "rows/level", N / nLevels, flush=True)
    for g in [0, 8, 16, 32, 64, 148, 296, 592]:
        lib.set_option("sweep_grid", g)
        lib.timers_reset()
        for i in range(6):
            out = ldu.dic_precondition(rD, upper, b)
        assert np.array_equal(out, ref)
        tm = lib.timers()
        f = tm["dic_fwd"][0] / tm["dic_fwd"][1]; bw = tm["dic_bwd"][0] / tm["dic_bwd"][1]
        print("  grid %4d: fwd %.1f us bwd %.1f us  (%.2f us/level)" % (g, 1e3 * f, 1e3 * bw, 1e3 * f / nLevels), flush=True)
    lib.set_option("sweep_grid", 0)
    ldu.release(); pm.release()
