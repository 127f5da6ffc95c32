"""A/B of two builds of the library on the same box: stand-alone DIC sweep times on the 985k-cell channel (argv: library paths)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
path = sys.argv[1]
lib = capi.B200Lib(path); lib.init(0, 0, 1)
dims = (200, 50, 100)
pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
rD = ldu.dic_calc_rd(diag, upper)
for rep in range(3):
    lib.timers_reset()
    for i in range(20):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    print("%s: fwd %6.1f bwd %6.1f us" % (os.path.basename(path), 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
