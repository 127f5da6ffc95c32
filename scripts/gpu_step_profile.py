"""Per-kernel-class CUDA-event breakdown of one bench step (985k-cell weir channel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
for a in sys.argv[1:]:
    k, v = a.split("="); lib.set_option(k, float(v))
DIMS, LENGTHS = bench.WEAK["dims"], bench.WEAK["lengths"]
pm = lib.polymesh_box(*DIMS, *LENGTHS, weir=bench.WEIR)
g = pm.geometry(); F = pm.nInternalFaces
slices = {n: pm.patch_slice(n) for n in pm.patchNames}
cs = bench.channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, LENGTHS)
dm = pm.device_mesh()
ctl = bench.set_controls(capi.Region3dControls())
reg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
reg.correct_phi(); reg.clear_log()
for _ in range(3):
    reg.step()
reg.clear_log()
lib.timers_reset(); t0 = time.time(); reg.step(3); wall = (time.time() - t0) / 3
tm = lib.timers()
iters = [p[2] for p in reg.perfs()]
print("plain: step %.2f ms (wall %.2f), pcg_solve %.2f ms, %d solves, %.1f iterations/step" % (tm["step"][0] / 3, 1e3 * wall, tm["pcg_solve"][0] / 3, len(iters) / 3, sum(iters) / 3))
lib.set_option("pcg_kernel_timers", 1)
reg.clear_log(); lib.timers_reset(); reg.step(2)
tm = lib.timers()
it = sum(p[2] for p in reg.perfs()) / 2
print("with per-kernel timers: step %.2f ms, %.1f iterations/step" % (tm["step"][0] / 2, it))
for k, (ms, calls) in sorted(tm.items(), key=lambda kv: -kv[1][0]):
    if calls:
        print("  %-12s %9.3f ms/step  %6d calls/step  %8.1f us/call" % (k, ms / 2, calls // 2, 1e3 * ms / calls))
