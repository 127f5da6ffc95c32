"""Per-tile timeline of the forward DIC sweep (%globaltimer stamps): where does a tile spend its time?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
dims = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (200, 50, 100)
pm = lib.polymesh_box(*dims, dims[0] / 10.0, dims[1] / 10.0, dims[2] / 10.0, weir=(0.45, 0.5, 0.3))
ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
rD = ldu.dic_calc_rd(diag, upper)
ref = ldu.dic_precondition(rD, upper, b)
for ro, pub in ((0, 1), (1, 1), (1, 2)):     # operands in shared memory (round 1) / registers (round 2) / registers + hybrid publish
    lib.set_option("sweep_reg_ops", ro); lib.set_option("sweep_publish", pub)
    assert np.array_equal(ldu.dic_precondition(rD, upper, b), ref)
    lib.timers_reset()
    for i in range(10):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    print("sweep_reg_ops %d publish %d: fwd %6.1f bwd %6.1f us" % (ro, pub, 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
lib.set_option("sweep_reg_ops", 1); lib.set_option("sweep_publish", int(sys.argv[1]) if len(sys.argv) > 1 else 1)
lib.set_option("sweep_trace", 1)
ldu.dic_precondition(rD, upper, b)
tr, tl, ts = ldu.sweep_trace()
lib.set_option("sweep_trace", 0)
tr = tr.astype(np.int64); t0 = tr[:, 0].min(); tr -= t0
print("tiles %d, sweep span %.1f us" % (tr.shape[0], tr[:, 3].max() / 1e3))
print("per-tile mean (us): ticket->staged %.2f, staged->halo ready %.2f, levels %.2f" % (
    (tr[:, 1] - tr[:, 0]).mean() / 1e3, (tr[:, 2] - tr[:, 1]).mean() / 1e3, (tr[:, 3] - tr[:, 2]).mean() / 1e3))
rows = np.diff(ts)
# per-tile lag: halo ready - latest "steps done" / "published" among the tiles it depends on
rowCell, level, _ = ldu.renumbering()
cellRow = np.empty(N, dtype=np.int64); cellRow[rowCell] = np.arange(N)
tile_of_row = np.searchsorted(ts, np.arange(N), side="right") - 1
ta, tb = tile_of_row[cellRow[pm.owner[:F]]], tile_of_row[cellRow[pm.neighbour]]
m = ta != tb
dep_done = np.zeros(tr.shape[0], dtype=np.int64); dep_pub = np.zeros(tr.shape[0], dtype=np.int64)
np.maximum.at(dep_done, tb[m], tr[ta[m], 3]); np.maximum.at(dep_pub, tb[m], tr[ta[m], 4])
has = np.zeros(tr.shape[0], dtype=bool); has[tb[m]] = True
lag_done = (tr[has, 2] - dep_done[has]) / 1e3; lag_pub = (tr[has, 2] - dep_pub[has]) / 1e3
print("publish pass (us): mean %.2f max %.2f" % ((tr[:, 4] - tr[:, 3]).mean() / 1e3, (tr[:, 4] - tr[:, 3]).max() / 1e3))
print("ready - last dep steps-done (us): min %.2f median %.2f mean %.2f max %.2f" % (lag_done.min(), np.median(lag_done), lag_done.mean(), lag_done.max()))
print("ready - last dep published  (us): min %.2f median %.2f mean %.2f max %.2f" % (lag_pub.min(), np.median(lag_pub), lag_pub.mean(), lag_pub.max()))
late = tr[has, 1] > dep_pub[has]
print("tiles whose staging finished after their deps were published: %d of %d" % (late.sum(), has.sum()))
print("level: nTiles  ticket(min)  staged(max)  ready(min..max)  done(max)  levels_us(mean)  rows(mean)")
for l in range(tl.max() + 1):
    m = tl == l
    print("%3d: %3d  %8.1f  %8.1f  %8.1f..%8.1f  %8.1f  %6.2f  %6.0f" % (l, m.sum(), tr[m, 0].min() / 1e3, tr[m, 1].max() / 1e3, tr[m, 2].min() / 1e3,
          tr[m, 2].max() / 1e3, tr[m, 3].max() / 1e3, (tr[m, 3] - tr[m, 2]).mean() / 1e3, rows[m].mean()))
