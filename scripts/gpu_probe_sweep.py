"""Scan the DIC sweep-plan parameters (band levels x tile rows x smem budget) on the 1M / 8M weir meshes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shallowinterfoam_b200 import capi
lib = capi.load(); lib.init(0, 0, 1)
sizes = [(200, 50, 100)] + ([(400, 100, 200)] if "--big" in sys.argv else [])
combos = [(16, 2048, 110, 128), (16, 4096, 220, 256), (8, 1024, 72, 128), (12, 1536, 110, 128), (24, 3072, 220, 128), (32, 4096, 220, 128),
          (16, 1024, 72, 64), (8, 512, 36, 64), (16, 1536, 110, 96), (20, 2560, 150, 128), (32, 2048, 110, 64)]
if "--quick" in sys.argv:
    combos = combos[:3]
for dims in sizes:
    ref = None
    for band, rows, kb, lanes in combos:
        lib.set_option("sweep_band_levels", band); lib.set_option("sweep_tile_rows", rows); lib.set_option("sweep_smem_bytes", kb * 1024)
        lib.set_option("sweep_threads", lanes)
        t0 = time.time()
        pm = lib.polymesh_box(*dims, 20.0, 5.0, 10.0, weir=(0.45, 0.5, 0.3))
        ldu = pm.ldu(); N, F = pm.nCells, pm.nInternalFaces
        tb = time.time() - t0
        _, _, nLevels = ldu.renumbering(); _, _, ts = ldu.tiling()
        rng = np.random.default_rng(0)
        upper = -rng.uniform(0.5, 1.5, F)
        diag = np.zeros(N); np.add.at(diag, pm.owner[:F], -upper); np.add.at(diag, pm.neighbour, -upper); diag += 1e-4
        b = rng.uniform(-1, 1, N)
        rD = ldu.dic_calc_rd(diag, upper)
        out = ldu.dic_precondition(rD, upper, b)
        if ref is None:
            ref = out
        assert np.array_equal(out, ref), "tiling changed the DIC result"
        line = "mesh %s levels %d band %2d rows %4d smem %3dK lanes %3d tiles %5d (build %.1fs):" % (dims, nLevels, band, rows, kb, lanes, ts.size - 1, tb)
        for cps in (0,):
            lib.set_option("sweep_ctas_per_sm", cps)
            lib.timers_reset()
            for i in range(5):
                ldu.dic_precondition(rD, upper, b)
            ldu.dic_calc_rd(diag, upper)
            tm = lib.timers()
            f = tm["dic_fwd"][0] / tm["dic_fwd"][1]; bw = tm["dic_bwd"][0] / tm["dic_bwd"][1]; rd = tm["dic_calc_rD"][0] / tm["dic_calc_rD"][1]
            line += "  fwd %6.1f bwd %6.1f us (coeffs+rd avg %6.1f)" % (1e3 * f, 1e3 * bw, 1e3 * rd)
        print(line, flush=True)
        lib.timers_reset()
        for i in range(5):
            ldu.amul(diag, upper, b)
        tm = lib.timers()
        print("      amul %.1f us" % (1e3 * tm["amul"][0] / tm["amul"][1]), flush=True)
        ldu.release(); pm.release()
