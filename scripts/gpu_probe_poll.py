"""Probe: DIC sweep time vs number of resident CTAs and poll back-off (L2 polling contention)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
dims = (200, 50, 100)
pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
rD = ldu.dic_calc_rd(diag, upper)
for grid, ns in [(0, 0), (0, 20), (0, 100), (0, 300), (148, 0), (148, 100), (100, 0), (200, 0)]:
    if True:
        lib.set_option("sweep_grid", grid); lib.set_option("sweep_poll_ns", ns)
        ldu.dic_precondition(rD, upper, b)
        lib.timers_reset()
        for i in range(5):
            ldu.dic_precondition(rD, upper, b)
        tm = lib.timers()
        print("grid %3d pollNs %4d: fwd %6.1f bwd %6.1f us" % (grid, ns, 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
