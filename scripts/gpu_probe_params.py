"""DIC sweep time vs plan parameters (band levels, lanes, tile rows) on the 985k-cell weir channel: every variant must reproduce the
first variant's preconditioner result bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
dims = (200, 50, 100)
pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
ref = None
configs = [(16, 256, 4096), (16, 192, 3072), (16, 128, 2048), (12, 256, 3072), (14, 256, 3584), (16, 224, 3584), (10, 256, 2560), (8, 256, 2048), (16, 256, 8192), (16, 160, 2560)]
for band, lanes, rows in configs:
    lib.set_option("sweep_band_levels", band); lib.set_option("sweep_threads", lanes); lib.set_option("sweep_tile_rows", rows)
    ldu = pm.ldu()
    rD = ldu.dic_calc_rd(diag, upper)
    w = ldu.dic_precondition(rD, upper, b)
    if ref is None:
        ref = w
    ok = np.array_equal(w, ref)
    lib.timers_reset()
    for i in range(20):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    _, _, ts = ldu.tiling()
    print("band %2d lanes %3d rows %4d: tiles %4d  fwd %6.1f bwd %6.1f us  bitwise %s" % (band, lanes, rows, len(ts) - 1,
          1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1], ok), flush=True)
    ldu.release()
