"""Host <-> device field transfers from pageable memory through the C-ABI: b200_region3d_set_fields (63 MB in) and b200_region3d_get_fields
(80 MB out) at 985k cells, for a few values of the option host_copy_threads (1 = plain cudaMemcpyAsync from pageable memory)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from shallowinterfoam_b200 import capi
from cases import channel_case, default_controls
lib = capi.load(); lib.init(0, 0, 1)
L = (20.0, 5.0, 10.0)
pm = lib.polymesh_box(200, 50, 100, *L, weir=(0.45, 0.5, 0.3))
cs = channel_case(pm, *L)
dm = pm.device_mesh()
ctl = default_controls(capi.Region3dControls, 0.002)
rg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
N, F, B = pm.nCells, pm.nInternalFaces, pm.nBoundaryFaces
f = rg.fields()
a, U, pd, phi = f["alpha1"].copy(), f["U"].copy().reshape(-1), f["pd"].copy(), f["phi"].copy()
oa, oU, opd, ophi, op, orho = (np.zeros(N), np.zeros(3 * N), np.zeros(N), np.zeros(F + B), np.zeros(N), np.zeros(N))
d = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
bytes_in = 8 * (5 * N + F + B); bytes_out = 8 * (7 * N + F + B)
for thr in (1, 2, 4, 8, 1, 4):
    lib.set_option("host_copy_threads", thr)
    for rep in range(2):
        t0 = time.time()
        for i in range(5):
            lib.check(lib.c.b200_region3d_get_fields(rg.h, d(oa), d(oU), d(opd), d(ophi), d(op), d(orho)))
        t_out = (time.time() - t0) / 5
        t0 = time.time()
        for i in range(5):
            lib.check(lib.c.b200_region3d_set_fields(rg.h, d(a), d(U), d(pd), d(phi)))
            lib.check(lib.c.b200_region3d_get_fields(rg.h, d(oa), d(oU), d(opd), d(ophi), d(op), d(orho)))
        t_in = (time.time() - t0) / 5 - t_out
    assert np.array_equal(oa, a) and np.array_equal(oU, U) and np.array_equal(ophi, phi)
    print("host_copy_threads %d: set_fields %.2f ms = %.1f GB/s, get_fields %.2f ms = %.1f GB/s" % (thr, 1e3 * t_in, bytes_in / t_in / 1e9, 1e3 * t_out, bytes_out / t_out / 1e9), flush=True)
