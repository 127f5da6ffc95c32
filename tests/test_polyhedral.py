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


@pytest.mark.parametrize("band,rows,smem,lanes", [(16, 4096, 225000, 256), (2, 64, 110 * 1024, 32), (32, 4096, 225000, 128), (16, 2048, 30000, 64),
                                                  (8, 96, 110 * 1024, 64), (64, 8192, 225000, 256)])
def test_polyhedral_tiling_variants_bitwise(oracle, band, rows, smem, lanes):
    """(emulated build only: this is a test of the plan builder's branches; on the GPU the chunked kernels run in test_irregular_ldu.py,
    the two tests above and tests/test_atsize_region.py)
    chunked sweep plans (rows with up to 17 sources applied as chains of <= 3-source chunks) under different tilings: tiny tiles (almost
    every source is a polled halo value), bands high enough for tiles of more than 16 steps (shared-memory operand path with chunk flags), a
    tiny shared-memory budget (pieces halved laterally or by level), one tile per band.  DIC / DILU sweeps, calcReciprocalD and PCG stay
    bitwise / 1e-8 against the oracle, and the plan keeps its invariants."""
    from conftest import get_lib
    lib = get_lib("emu")
    lib.set_option("sweep_band_levels", band); lib.set_option("sweep_tile_rows", rows); lib.set_option("sweep_smem_bytes", smem)
    lib.set_option("sweep_threads", lanes)
    try:
        pm = make(lib, (14, 10, 12))
        m = oracle_mesh_of(pm)
        ldu = pm.ldu()
        F = pm.nInternalFaces
        assert max(np.bincount(pm.owner[:F], minlength=m.N).max(), np.bincount(pm.neighbour, minlength=m.N).max()) > 3   # a chunked plan
        rowCell, level, _ = ldu.renumbering()
        bl, tr, ts = ldu.tiling()
        assert sorted(rowCell.tolist()) == list(range(m.N)) and ts[0] == 0 and ts[-1] == m.N and np.all(np.diff(ts) > 0)
        inv = np.empty(m.N, dtype=np.int64); inv[rowCell] = np.arange(m.N)
        tile_of_row = np.searchsorted(ts, np.arange(m.N), side="right") - 1
        assert np.all(tile_of_row[inv[pm.owner[:F]]] <= tile_of_row[inv[pm.neighbour]])        # ticket order is topological
        diag, upper = random_spd(m, seed=31)
        rng = np.random.default_rng(32)
        rA = rng.uniform(-1, 1, m.N); b = rng.uniform(-1, 1, m.N)
        rD = oracle.dic_calc_rd(m, diag, upper)
        assert np.array_equal(rD, ldu.dic_calc_rd(diag, upper))
        w = oracle.dic_precondition(m, rD, upper, rA)
        for _ in range(2):
            assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
        lower = -rng.uniform(0.5, 1.5, m.F)
        rDl = oracle.dilu_calc_rd(m, diag, upper, lower)
        assert np.array_equal(rDl, ldu.dilu_calc_rd(diag, upper, lower))
        for tr_ in (False, True):
            assert np.array_equal(oracle.dilu_precondition(m, rDl, upper, lower, rA, tr_), ldu.dilu_precondition(rDl, upper, lower, rA, tr_))
        xo, po, ho = oracle.pcg_solve(m, diag, upper, np.zeros(m.N), b, tol=1e-10, maxIter=300)
        xg, pg, hg = ldu.pcg_solve(diag, upper, np.zeros(m.N), b, tol=1e-10, maxIter=300)
        assert pg.nIterations == po.nIterations and np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
        ldu.release()
    finally:
        for k, v in (("sweep_band_levels", 16), ("sweep_tile_rows", 4096), ("sweep_smem_bytes", 225000), ("sweep_threads", 256)):
            lib.set_option(k, v)
