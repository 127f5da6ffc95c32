"""HBM-bound kernel rooflines on the 8M-cell channel (BASELINE configs[2] on one GPU: nothing fits in the 126 MB L2):
algorithmic bytes / CUDA-event time of Amul, DIC precondition and MULES explicitSolve."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
dims = (400, 100, 200); lengths = (40.0, 10.0, 20.0)
pm = lib.polymesh_box(*dims, *lengths, weir=bench.WEIR)
N, F = pm.nCells, pm.nInternalFaces
peak = bench.measured_peak()[0]
g = pm.geometry()
slices = {n: pm.patch_slice(n) for n in pm.patchNames}
cs = bench.channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, lengths)
dm = pm.device_mesh()
ctl = bench.set_controls(capi.Region3dControls())
reg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
reg.correct_phi(); reg.clear_log()
reg.step(1)
reg.clear_log()
lib.set_option("pcg_kernel_timers", 1)
lib.timers_reset(); reg.step(2)
tm = lib.timers()
it = sum(p[2] for p in reg.perfs())
alg = {"amul": 8 * (3 * N + F) + 8 * F, "dic_precondition": 48 * N + 32 * F, "mules": 424 * N + 288 * F}
t_amul = tm["amul"][0] / it * 1e-3
t_dic = (tm["dic_fwd"][0] + tm["dic_bwd"][0]) / it * 1e-3
t_mules = tm["mules"][0] / tm["mules"][1] * 1e-3
out = {"mesh": "%dx%dx%d minus weir = %d cells" % (dims + (N,)), "pcg_iterations": it, "step_ms": tm["step"][0] / 2,
       "cell_steps_per_s": N * 2 / (tm["step"][0] * 1e-3), "peak_GBps_measured_copy": peak}
for k, t in (("amul", t_amul), ("dic_precondition", t_dic), ("mules", t_mules)):
    out[k] = {"us": 1e6 * t, "GBps": alg[k] / t / 1e9, "frac_of_measured_peak": alg[k] / t / 1e9 / peak, "frac_of_8TBps": alg[k] / t / 8e12}
print(json.dumps(out))
