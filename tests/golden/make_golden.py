"""Generates the committed fixtures of tests/golden/ from the CPU ORACLE (oracle/, the restatement of foam-extend-3.1).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

The reference ships no golden vectors and foam-extend-3.1 cannot be built here, so these are REGRESSION
fixtures: they freeze what the oracle computes today (a later edit of the oracle that changes a bit anywhere fails
tests/test_golden.py), and they give the CUDA path a fixed target that does not depend on rebuilding the oracle.  They do not pin
the oracle to foam-extend -- the independent pins are the known-answer tests in tests/test_golden.py (tridiagonal systems, where
DIC is the exact factorisation and upstream's PCG must converge in one iteration to numpy's direct solution; hand-computed rD) and
`python -m shallowinterfoam_b200.foamcase check` against a real foam-extend-3.1 run.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

LDU_CASE = dict(dims=(12, 5, 7), lengths=(5.0, 1.0, 2.0), weir=(0.45, 0.6, 0.3))
REGION_CASE = dict(dims=(20, 6, 10), lengths=(5.0, 1.0, 2.0), weir=(0.45, 0.55, 0.3), steps=2)


def oracle_mesh(case):
    from oracle import foam_oracle as fo, mesh_oracle as mo
    mesh = mo.block_mesh_box(*case["dims"], *case["lengths"])
    if case["weir"]:
        mesh = mo.subset_mesh(mesh, mo.weir_keep_mask(*case["dims"], *case["weir"]))
    g = mo.geometry(mesh)
    return mesh, g, fo.Mesh(mesh["owner"], mesh["neighbour"], mesh["patches"], g, mesh["nCells"])


def ldu_golden():
    from oracle import foam_oracle as fo
    from cases import random_spd
    mesh, g, m = oracle_mesh(LDU_CASE)
    diag, upper = random_spd(m, seed=1)
    rng = np.random.default_rng(2)
    psi = rng.uniform(0, 1, m.N); b = rng.uniform(-1, 1, m.N)
    rD = fo.dic_calc_rd(m, diag, upper)
    x, perf, hist = fo.pcg_solve(m, diag, upper, np.zeros(m.N), b, tol=1e-10, relTol=0.0, maxIter=200)
    return dict(owner=mesh["owner"], neighbour=mesh["neighbour"], diag=diag, upper=upper, psi=psi, b=b,
                amul=fo.amul(m, diag, upper, upper, psi), sumA=fo.sumA(m, diag, upper, upper), rD=rD,
                precond=fo.dic_precondition(m, rD, upper, b), x=x, hist=hist,
                perf=np.array([perf.initialResidual, perf.finalResidual, perf.nIterations]))


def region_golden():
    import bench
    from oracle import foam_oracle as fo
    from cases import default_controls
    mesh, g, m = oracle_mesh(REGION_CASE)
    names = [p[0] for p in mesh["patches"]]
    slices = {n: m.patch_slice(n) for n in names}
    cs = bench.channel_fields(names, g["C"], g["Cf"][m.F:], slices, REGION_CASE["lengths"])
    ctl = default_controls(fo.OrcControls)
    r = fo.Region(m, ctl, fo.BC(m, cs["tA"], 1, cs["rvA"]), fo.BC(m, cs["tU"], 3, cs["rvU"]), fo.BC(m, cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    r.correct_phi(); r.clear_log()
    r.step(REGION_CASE["steps"])
    f = r.fields()
    out = {k: np.array(f[k]) for k in ("alpha1", "U", "pd", "phi", "p")}
    out["perfs"] = np.array([[p[0], p[1], p[2]] for p in r.perfs()])
    out["owner"] = mesh["owner"]; out["neighbour"] = mesh["neighbour"]
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "ldu_12x5x7_weir.npz"), **ldu_golden())
    np.savez_compressed(os.path.join(HERE, "region_20x6x10_weir.npz"), **region_golden())
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
