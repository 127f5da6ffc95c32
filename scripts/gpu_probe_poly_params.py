"""Polyhedral box: stand-alone DIC sweep times for a few plan / polling parameters (options are read when the ldu is created)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import mesh_oracle as mo
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
dims = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (240, 160, 160)
box = lib.polymesh_box(*dims, dims[0] / 10.0, dims[1] / 10.0, dims[2] / 10.0, jitter=(0.25, 12345))
pm = box.agglomerate(mo.block_groups(*dims)); box.release()
N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
ref = None
for band, rows, poll in ((16, 4096, 0), (16, 4096, 100), (16, 4096, 300), (24, 4096, 0), (32, 4096, 0), (12, 4096, 0), (16, 700, 0)):
    lib.set_option("sweep_band_levels", band); lib.set_option("sweep_tile_rows", rows); lib.set_option("sweep_poll_ns", poll)
    ldu = pm.ldu()
    rD = ldu.dic_calc_rd(diag, upper)
    w = ldu.dic_precondition(rD, upper, b)
    if ref is None: ref = w
    assert np.array_equal(w, ref)
    lib.timers_reset()
    for i in range(10):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    bl, tr, ts = ldu.tiling()
    print("band %d (effective %d) tile_rows %d poll_ns %d: tiles %d  fwd %6.1f bwd %6.1f us" % (band, bl, rows, poll, len(ts) - 1, 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
    ldu.release()
