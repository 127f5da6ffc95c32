"""Per-tile timeline of the forward DIC sweep (%globaltimer stamps): where does a tile spend its time?"""
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
ref = ldu.dic_precondition(rD, upper, b)
for wk in (0,):
    assert np.array_equal(ldu.dic_precondition(rD, upper, b), ref)
    lib.timers_reset()
    for i in range(10):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    print("run %d: fwd %6.1f bwd %6.1f us" % (wk, 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
lib.set_option("sweep_trace", 1)
ldu.dic_precondition(rD, upper, b)
tr, tl, ts = ldu.sweep_trace()
lib.set_option("sweep_trace", 0)
tr = tr.astype(np.int64); t0 = tr[:, 0].min(); tr -= t0
print("tiles %d, sweep span %.1f us" % (tr.shape[0], tr[:, 4].max() / 1e3))
print("per-tile mean (us): start->halo in (poll + staging of the next tile) %.2f, steps %.2f, hand-over + epilogue + publish %.2f" % (
    (tr[:, 2] - tr[:, 0]).mean() / 1e3, (tr[:, 3] - tr[:, 2]).mean() / 1e3, (tr[:, 4] - tr[:, 3]).mean() / 1e3))
rows = np.diff(ts)
ch = ldu.tile_chain
print("chains %d" % (ch.max() + 1))
for c in (0, 5, 20, 40, 60):
    if c > ch.max():
        continue
    tiles = np.nonzero(ch == c)[0]
    tiles = tiles[np.argsort(tr[tiles, 0])]
    print("chain %d: (tile, level, rows | start, t0 at barrier, halo in, steps done, t0 at final barrier, published)" % c)
    for t in tiles:
        print("  %4d L%2d %5d | %7.2f %7.2f %7.2f %7.2f %7.2f %7.2f" % (t, tl[t], rows[t], tr[t, 0] / 1e3, tr[t, 1] / 1e3, tr[t, 2] / 1e3, tr[t, 3] / 1e3, tr[t, 5] / 1e3, tr[t, 4] / 1e3))
