"""Per-tile timeline of the forward DIC sweep on the polyhedral box (chunked plan): %globaltimer stamps per tile."""
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
ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
rng = np.random.default_rng(0)
upper = -rng.uniform(0.5, 1.5, F)
diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
b = rng.uniform(-1, 1, N)
rD = ldu.dic_calc_rd(diag, upper)
ref = ldu.dic_precondition(rD, upper, b)
for ro in (0, 1):
    lib.set_option("sweep_reg_ops", ro)
    assert np.array_equal(ldu.dic_precondition(rD, upper, b), ref)
    lib.timers_reset()
    for i in range(10):
        ldu.dic_precondition(rD, upper, b)
    tm = lib.timers()
    print("sweep_reg_ops %d: fwd %6.1f bwd %6.1f us" % (ro, 1e3 * tm["dic_fwd"][0] / tm["dic_fwd"][1], 1e3 * tm["dic_bwd"][0] / tm["dic_bwd"][1]), flush=True)
lib.set_option("sweep_reg_ops", 1)
lib.set_option("sweep_trace", 1)
ldu.dic_precondition(rD, upper, b)
tr, tl, ts = ldu.sweep_trace()
lib.set_option("sweep_trace", 0)
tr = tr.astype(np.int64); t0 = tr[:, 0].min(); tr -= t0
rows = np.diff(ts)
print("cells %d tiles %d, tile levels %d, sweep span %.1f us; rows/tile mean %.0f" % (N, tr.shape[0], tl.max() + 1, tr[:, 3].max() / 1e3, rows.mean()))
print("per-tile mean (us): ticket->staged %.2f, staged->halo ready %.2f, steps %.2f, publish %.2f" % (
    (tr[:, 1] - tr[:, 0]).mean() / 1e3, (tr[:, 2] - tr[:, 1]).mean() / 1e3, (tr[:, 3] - tr[:, 2]).mean() / 1e3, (tr[:, 4] - tr[:, 3]).mean() / 1e3))
big = rows > 600
print("tiles with > 600 rows: %d, steps time mean %.2f us" % (big.sum(), (tr[big, 3] - tr[big, 2]).mean() / 1e3))
rowCell, level, _ = ldu.renumbering()
cellRow = np.empty(N, dtype=np.int64); cellRow[rowCell] = np.arange(N)
tile_of_row = np.searchsorted(ts, np.arange(N), side="right") - 1
ta, tb = tile_of_row[cellRow[pm.owner[:F]]], tile_of_row[cellRow[pm.neighbour]]
m = ta != tb
dep_done = np.zeros(tr.shape[0], dtype=np.int64)
np.maximum.at(dep_done, tb[m], tr[ta[m], 3])
has = np.zeros(tr.shape[0], dtype=bool); has[tb[m]] = True
lag = (tr[has, 2] - dep_done[has]) / 1e3
print("halo ready - last dep steps-done (us): min %.2f median %.2f mean %.2f max %.2f" % (lag.min(), np.median(lag), lag.mean(), lag.max()))
# critical path reconstruction: walk back from the tile that finished last through the dependency that finished last
pred_last = np.full(tr.shape[0], -1, dtype=np.int64)
order = np.argsort(tr[ta[m], 3], kind="stable")
pred_last[tb[m][order]] = ta[m][order]          # the last assignment wins = the latest-finishing dependency
t = int(np.argmax(tr[:, 3])); chain = []
while t >= 0:
    chain.append(t); t = int(pred_last[t])
chain = chain[::-1]
st = np.array([(tr[c, 3] - tr[c, 2]) for c in chain]) / 1e3
hop = np.array([(tr[chain[i + 1], 2] - tr[chain[i], 3]) for i in range(len(chain) - 1)]) / 1e3
print("critical chain: %d tiles, steps time total %.1f us (mean %.2f), hops total %.1f us (mean %.2f, max %.2f)" % (len(chain), st.sum(), st.mean(), hop.sum(), hop.mean() if len(hop) else 0, hop.max() if len(hop) else 0))
dd = np.array([tr[chain[i], 3] for i in range(len(chain) - 1)])
nx = np.array(chain[1:])
print("chain tiles relative to their last dependency's steps-done (us): ticket %.2f (late for %d of %d), staged %.2f (late for %d), ready %.2f" % (
    ((tr[nx, 0] - dd) / 1e3).mean(), ((tr[nx, 0] - dd) > 0).sum(), len(nx), ((tr[nx, 1] - dd) / 1e3).mean(), ((tr[nx, 1] - dd) > 0).sum(), ((tr[nx, 2] - dd) / 1e3).mean()))
early = (tr[nx, 1] - dd) < 0
print("hop of the chain tiles staged BEFORE their dependency finished: mean %.2f us (%d tiles); staged after: mean %.2f us (%d tiles)" % (
    ((tr[nx[early], 2] - dd[early]) / 1e3).mean(), early.sum(), ((tr[nx[~early], 2] - dd[~early]) / 1e3).mean() if (~early).any() else 0.0, (~early).sum()))
print("first tile of the chain became ready at %.1f us" % (tr[chain[0], 2] / 1e3))
