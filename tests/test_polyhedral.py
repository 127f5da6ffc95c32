"""BASELINE configs[3] (polyhedral box, irregular LDU addressing): hex box with a checkerboard of 2x2x2 blocks agglomerated into
single polyhedral cells (6 .. 24 faces per cell, several faces between the same two cells) and jittered vertices.
Mesh bit-exact against the numpy oracle; LDU kernels bitwise, region steps to the bars of tests/test_region.py."""
import numpy as np
import pytest
from oracle import mesh_oracle as mo
from cases import oracle_mesh_of, channel_case, default_controls, random_spd

L = (5.0, 1.0, 2.0)


def make(lib, dims, jitter=True):
    pm = lib.polymesh_box(*dims, *L, jitter=(0.25, 12345) if jitter else None)
    return pm.agglomerate(mo.block_groups(*dims))


@pytest.mark.parametrize("dims", [(6, 4, 4), (9, 5, 6)])
def test_polyhedral_mesh_bit_exact(dims):
    from shallowinterfoam_b200 import capi
    lib = capi.load()   # host tools need no device
    pm = make(lib, dims)
    ref = mo.block_mesh_box(*dims, *L)
    ref = mo.jitter_points(ref, 0.25, L[0] / dims[0], L[1] / dims[1], L[2] / dims[2], 12345)
    ref = mo.agglomerate(ref, mo.block_groups(*dims))
    assert np.array_equal(capi.checkerboard_groups(*dims), mo.block_groups(*dims))   # the product-side helper bench.py uses
    assert pm.nCells == ref["nCells"] < dims[0] * dims[1] * dims[2]
    assert np.array_equal(pm.points, ref["points"]) and np.array_equal(pm.faces, ref["faces"])
    assert np.array_equal(pm.owner, ref["owner"]) and np.array_equal(pm.neighbour, ref["neighbour"])
    assert pm.patches == ref["patches"]
    F = pm.nInternalFaces
    assert np.all(pm.owner[:F] < pm.neighbour)
    key = pm.owner[:F].astype(np.int64) * pm.nCells + pm.neighbour
    assert np.all(np.diff(key) >= 0) and np.any(np.diff(key) == 0)          # upper-triangular order, with repeated cell pairs
    nfaces = np.bincount(pm.owner, minlength=pm.nCells) + np.bincount(pm.neighbour, minlength=pm.nCells)
    assert nfaces.min() == 6 and nfaces.max() == 24                         # hexes and 2x2x2 super-cells
    g, gr = pm.geometry(), mo.geometry(ref)
    for k in g:
        assert np.array_equal(g[k], gr[k]), k
    box = lib.polymesh_box(*dims, *L).geometry()
    assert abs(g["V"].sum() - box["V"].sum()) < 1e-12 * box["V"].sum() and np.all(g["V"] > 0)
    s = np.zeros((pm.nCells, 3)); np.add.at(s, pm.owner, g["Sf"]); np.add.at(s, pm.neighbour, -g["Sf"][:F])
    assert np.abs(s).max() < 1e-14                                          # closed polyhedra


def test_polyhedral_ldu_kernels_bitwise_and_pcg(lib, oracle):
    pm = make(lib, (10, 6, 8))
    m = oracle_mesh_of(pm)
    ldu = pm.ldu()
    diag, upper = random_spd(m, seed=11)
    rng = np.random.default_rng(12)
    psi = rng.uniform(0, 1, m.N); b = rng.uniform(-1, 1, m.N)
    assert np.array_equal(oracle.amul(m, diag, upper, upper, psi), ldu.amul(diag, upper, psi))
    assert np.array_equal(oracle.sumA(m, diag, upper, upper), ldu.sumA(diag, upper))
    rD = ldu.dic_calc_rd(diag, upper)
    assert np.array_equal(oracle.dic_calc_rd(m, diag, upper), rD)
    assert np.array_equal(oracle.dic_precondition(m, rD, upper, b), ldu.dic_precondition(rD, upper, b))
    xo, po, ho = oracle.pcg_solve(m, diag, upper, np.zeros(m.N), b, tol=1e-10, maxIter=300)
    xg, pg, hg = ldu.pcg_solve(diag, upper, np.zeros(m.N), b, tol=1e-10, maxIter=300)
    assert pg.nIterations == po.nIterations and np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    assert np.abs(xg - xo).max() <= 1e-10 * np.abs(xo).max()
    ldu.release()


def test_polyhedral_region_steps(lib, oracle):
    from shallowinterfoam_b200 import capi
    pm = make(lib, (12, 4, 8))
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    dm = pm.device_mesh()
    co = default_controls(oracle.OrcControls); cg = default_controls(capi.Region3dControls)
    ro = oracle.Region(m, co, oracle.BC(m, cs["tA"], 1, cs["rvA"]), oracle.BC(m, cs["tU"], 3, cs["rvU"]), oracle.BC(m, cs["tP"], 1),
                       cs["alpha"], cs["U"], cs["pd"])
    rg = dm.region3d(cg, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    ro.correct_phi(); rg.correct_phi()
    ro.clear_log(); rg.clear_log()
    for s in range(2):
        ro.step(); rg.step()
    fo_, fg = ro.fields(), rg.fields()
    for k in ("alpha1", "U", "pd", "phi", "p", "rho"):
        assert np.abs(fo_[k] - fg[k]).max() <= 1e-10 * max(np.abs(fo_[k]).max(), 1e-300), k
    assert [p[2] for p in ro.perfs()] == [p[2] for p in rg.perfs()]
    ho, hg = ro.res_hist(), rg.res_hist()
    assert ho.shape == hg.shape and np.allclose(ho, hg, rtol=1e-8, atol=1e-13)
    rg.release(); dm.release()
