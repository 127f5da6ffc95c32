"""a1-a4: Amul, sumA, DIC and PCG through the C-ABI vs the CPU oracle.
Amul / sumA / calcReciprocalD / precondition: BIT-EXACT (row-gather keeps upstream's per-cell summation order; no FMA).
PCG: reductions are tree sums on the device vs sequential sums on the CPU => residual history to 1e-8 relative
(north_star tolerance), solution to 1e-10 of its max."""
import numpy as np
import pytest
from cases import oracle_mesh_of, random_spd

MESHES = [((12, 5, 7), (0.45, 0.6, 0.3)), ((16, 8, 10), None), ((33, 3, 5), None), ((1, 1, 1), None), ((2, 1, 1), None)]


def make(lib, dims, weir):
    pm = lib.polymesh_box(*dims, 5.0, 1.0, 2.0, weir=weir)
    return pm, pm.ldu()


@pytest.mark.parametrize("dims,weir", MESHES)
def test_renumbering_and_addressing_bit_exact(lib, oracle, dims, weir):
    from oracle import mesh_oracle as mo
    pm, ldu = make(lib, dims, weir)
    F = pm.nInternalFaces
    rowCell, level, nLevels = ldu.renumbering()
    band_levels, tile_rows, tile_start = ldu.tiling()
    plan = mo.sweep_plan(pm.owner[:F], pm.neighbour, pm.nCells, band_levels, tile_rows)
    lev, perm, inv, ts = plan["level"], plan["rowCell"], plan["cellRow"], plan["tileStart"]
    assert np.array_equal(level, lev) and np.array_equal(rowCell, perm) and np.array_equal(tile_start, ts)
    # ticket order is a topological order of the tile graph: every dependency goes to the same or a later tile,
    # and to a strictly higher tile level when it leaves the tile
    tile_of_row = np.searchsorted(tile_start, np.arange(pm.nCells), side="right") - 1
    if F:
        ta, tb = tile_of_row[inv[pm.owner[:F]]], tile_of_row[inv[pm.neighbour]]
        assert np.all(ta <= tb)
        assert np.all(plan["tileLevel"][ta[ta != tb]] < plan["tileLevel"][tb[ta != tb]])
        # the potentials never decrease along a dependency (what makes the tile graph acyclic on any mesh)
        for k in ("level", "potJ", "potK"):
            assert np.all(plan[k][pm.owner[:F]] <= plan[k][pm.neighbour])
    if weir is None and min(dims) > 1:
        i, j, k = np.unravel_index(np.arange(pm.nCells), dims[::-1])[::-1]
        assert np.array_equal(plan["potJ"], j) and np.array_equal(plan["potK"], k)   # stride classes recover (j, k) of a hex block
    assert sorted(rowCell.tolist()) == list(range(pm.nCells))          # a permutation
    assert nLevels == (lev.max() + 1 if pm.nCells else 0)
    if weir is None:
        assert nLevels == sum(dims) - 2                               # hyperplanes i+j+k of a box
    # level schedule validity: every lower neighbour sits in an earlier level
    assert np.all(lev[pm.owner[:F]] < lev[pm.neighbour])
    os_, lo, ls = ldu.addressing()
    r = mo.ldu_addressing(pm.owner[:F], pm.neighbour, pm.nCells)
    assert np.array_equal(os_, r[0]) and np.array_equal(lo, r[1]) and np.array_equal(ls, r[2])
    ldu.release()


@pytest.mark.parametrize("dims,weir", MESHES)
def test_amul_sumA_dic_bitwise(lib, oracle, dims, weir):
    pm, ldu = make(lib, dims, weir)
    m = oracle_mesh_of(pm)
    diag, upper = random_spd(m, seed=3)
    rng = np.random.default_rng(7)
    psi = rng.uniform(0, 1, m.N)
    assert np.array_equal(oracle.amul(m, diag, upper, upper, psi), ldu.amul(diag, upper, psi))
    lower = -rng.uniform(0.5, 1.5, m.F)
    assert np.array_equal(oracle.amul(m, diag, upper, lower, psi), ldu.amul(diag, upper, psi, lower=lower))
    assert np.array_equal(oracle.sumA(m, diag, upper, lower), ldu.sumA(diag, upper, lower=lower))
    rD = oracle.dic_calc_rd(m, diag, upper)
    assert np.array_equal(rD, ldu.dic_calc_rd(diag, upper))
    rA = rng.uniform(-1, 1, m.N)
    w = oracle.dic_precondition(m, rD, upper, rA)
    assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
    # repeated application re-arms the data-flow sentinels correctly
    assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
    ldu.release()


def test_amul_linearity_and_symmetry(lib, oracle):
    pm, ldu = make(lib, (16, 8, 10), None)
    m = oracle_mesh_of(pm)
    diag, upper = random_spd(m, seed=5)
    rng = np.random.default_rng(11)
    x, y = rng.normal(size=m.N), rng.normal(size=m.N)
    Ax, Ay = ldu.amul(diag, upper, x), ldu.amul(diag, upper, y)
    assert abs(np.dot(y, Ax) - np.dot(x, Ay)) < 1e-10 * abs(np.dot(y, Ax))
    assert np.allclose(ldu.amul(diag, upper, 2.0 * x + y), 2.0 * Ax + Ay, rtol=1e-13, atol=1e-12)
    ldu.release()


@pytest.mark.parametrize("dims,weir,tol,relTol", [((12, 5, 7), (0.45, 0.6, 0.3), 1e-10, 0.0), ((16, 8, 10), None, 1e-7, 0.05),
                                                  ((16, 8, 10), None, 1e-9, 0.0), ((2, 1, 1), None, 1e-12, 0.0)])
def test_pcg_matches_oracle(lib, oracle, dims, weir, tol, relTol):
    pm, ldu = make(lib, dims, weir)
    m = oracle_mesh_of(pm)
    diag, upper = random_spd(m, seed=9, shift=(1e-4, 1e-3))
    rng = np.random.default_rng(13)
    b = rng.uniform(-1, 1, m.N); x0 = rng.uniform(-0.1, 0.1, m.N)
    xo, po, ho = oracle.pcg_solve(m, diag, upper, x0, b, tol=tol, relTol=relTol, maxIter=500)
    xg, pg, hg = ldu.pcg_solve(diag, upper, x0, b, tol=tol, relTol=relTol, maxIter=500)
    assert pg.nIterations == po.nIterations
    assert pg.converged == po.converged == 1 and pg.singular == 0
    # per-iteration residual history: 1e-8 relative (north_star), with the FP64 round-off floor of a sum of
    # O(1) terms (1e-13 of the initial residual) once the residual has dropped many orders of magnitude
    assert np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    assert abs(pg.initialResidual - po.initialResidual) <= 1e-12 * po.initialResidual
    assert np.abs(xg - xo).max() <= 1e-10 * max(np.abs(xo).max(), 1e-300)
    # true residual check, independent of both implementations
    r = b - oracle.amul(m, diag, upper, upper, xg)
    assert np.abs(r).sum() <= 10 * max(tol, relTol * 1.0) * (np.abs(b).sum() + np.abs(oracle.amul(m, diag, upper, upper, x0)).sum() + 1)
    ldu.release()


def test_pcg_controls(lib, oracle):
    pm, ldu = make(lib, (16, 8, 10), None)
    m = oracle_mesh_of(pm)
    diag, upper = random_spd(m, seed=2)
    b = np.ones(m.N); x0 = np.zeros(m.N)
    for kw in (dict(tol=1e-30, maxIter=7), dict(tol=1.0, minIter=3), dict(tol=1.0, maxIter=0), dict(tol=1e-6, maxIter=1000)):
        xo, po, ho = oracle.pcg_solve(m, diag, upper, x0, b, **kw)
        xg, pg, hg = ldu.pcg_solve(diag, upper, x0, b, **kw)
        assert (pg.nIterations, pg.converged) == (po.nIterations, po.converged), kw
        assert np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    # already-converged initial guess: zero iterations, x untouched
    xg, pg, hg = ldu.pcg_solve(diag, upper, xo, b, tol=1e-3)
    assert pg.nIterations == 0 and np.array_equal(xg, xo)
    ldu.release()


def test_pcg_rejects_asymmetric(lib):
    from shallowinterfoam_b200.capi import B200Error
    pm, ldu = make(lib, (4, 3, 2), None)
    F = pm.nInternalFaces
    with pytest.raises(B200Error) as e:
        ldu.pcg_solve(np.ones(pm.nCells), -np.ones(F) * 0.1, np.zeros(pm.nCells), np.ones(pm.nCells), lower=-np.ones(F) * 0.2)
    assert e.value.code == -5
    ldu.release()


def test_bad_addressing_is_rejected(lib):
    from shallowinterfoam_b200.capi import B200Error
    with pytest.raises(B200Error):
        lib.ldu_create(3, [1, 0], [2, 1])      # not sorted by lower
    with pytest.raises(B200Error):
        lib.ldu_create(3, [0, 2], [1, 1])      # lower >= upper


@pytest.mark.parametrize("band,rows,smem,lanes", [(1, 8, 110 * 1024, 32), (3, 64, 110 * 1024, 32), (5, 96, 110 * 1024, 64), (64, 4096, 200 * 1024, 128),
                                                  (16, 2048, 24000, 256)])
def test_dic_tiling_variants_bitwise(lib, oracle, band, rows, smem, lanes):
    """small tiles: almost every dependency is a polled halo value; huge tiles keep everything in one CTA; a tiny shared-memory
    budget forces the piece-halving path"""
    from oracle import mesh_oracle as mo
    lib.set_option("sweep_band_levels", band); lib.set_option("sweep_tile_rows", rows); lib.set_option("sweep_smem_bytes", smem)
    lib.set_option("sweep_threads", lanes)
    try:
        pm, ldu = make(lib, (14, 9, 11), (0.45, 0.6, 0.3))
        m = oracle_mesh_of(pm)
        F = pm.nInternalFaces
        bl, tr, ts = ldu.tiling()
        assert (bl, tr) == (band, rows)
        plan = mo.sweep_plan(pm.owner[:F], pm.neighbour, pm.nCells, band, rows, smem, lanes)
        rowCell, level, _ = ldu.renumbering()
        assert np.array_equal(rowCell, plan["rowCell"]) and np.array_equal(ts, plan["tileStart"])
        diag, upper = random_spd(m, seed=21)
        rng = np.random.default_rng(22)
        rA = rng.uniform(-1, 1, m.N)
        rD = oracle.dic_calc_rd(m, diag, upper)
        assert np.array_equal(rD, ldu.dic_calc_rd(diag, upper))
        w = oracle.dic_precondition(m, rD, upper, rA)
        for _ in range(2):
            assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
        b = rng.uniform(-1, 1, m.N)
        xo, po, ho = oracle.pcg_solve(m, diag, upper, np.zeros(m.N), b, tol=1e-9, maxIter=300)
        xg, pg, hg = ldu.pcg_solve(diag, upper, np.zeros(m.N), b, tol=1e-9, maxIter=300)
        assert pg.nIterations == po.nIterations and np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
        # a converged solve leaves one of the ready-flag vectors (y / z) populated; the next entry point must re-arm them
        # (stale halo values would be taken for published ones): precondition and a second solve on the same handle
        assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
        xg2, pg2, hg2 = ldu.pcg_solve(diag, upper, np.zeros(m.N), b, tol=1e-9, maxIter=300)
        assert pg2.nIterations == pg.nIterations and np.array_equal(hg2, hg) and np.array_equal(xg2, xg)
        ldu.release()
    finally:
        lib.set_option("sweep_band_levels", 16); lib.set_option("sweep_tile_rows", 4096); lib.set_option("sweep_smem_bytes", 225000)
        lib.set_option("sweep_threads", 256)


def test_registry_key_reuse_with_new_addressing_rebuilds(lib, oracle):
    """the device cache is keyed by the address of the lduAddressing; when the mesh changes under the same address the entry is
    validated by a hash of lower/upper/faceCells and rebuilt, and b200_ldu_invalidate drops it explicitly"""
    import ctypes as C
    key = 0x5151
    pm1 = lib.polymesh_box(6, 4, 3, 1.0, 1.0, 1.0)
    pm2 = lib.polymesh_box(4, 6, 3, 1.0, 1.0, 1.0)          # same cell and face counts, different addressing
    assert pm1.nCells == pm2.nCells and pm1.nInternalFaces == pm2.nInternalFaces
    l1 = pm1.ldu(key=key)
    l1b = pm1.ldu(key=key)
    assert l1.h.value == l1b.h.value                         # cache hit
    l2 = pm2.ldu(key=key)                                    # same key, other addressing: rebuilt, not trusted
    m2 = oracle_mesh_of(pm2)
    diag, upper = random_spd(m2, seed=3)
    psi = np.random.default_rng(1).uniform(0, 1, m2.N)
    assert np.array_equal(oracle.amul(m2, diag, upper, upper, psi), l2.amul(diag, upper, psi))
    lib.check(lib.c.b200_ldu_invalidate(C.c_void_p(key)))
    l3 = pm2.ldu(key=key)
    assert np.array_equal(oracle.amul(m2, diag, upper, upper, psi), l3.amul(diag, upper, psi))
    l3.release()


@pytest.mark.parametrize("publish,tail", [(0, 4), (1, 4), (2, 4)])
def test_dic_publish_modes_bitwise(lib, oracle, publish, tail):
    """how a sweep tile's rows reach global memory (one TMA bulk store per tile / a store in the step that computes the row / hybrid:
    early steps by one asynchronous bulk store, the last `tail` steps in the loop) never changes a bit of the result"""
    lib.set_option("sweep_publish", publish); lib.set_option("sweep_bulk_tail", tail)
    try:
        pm, ldu = make(lib, (18, 9, 11), (0.45, 0.6, 0.3))
        m = oracle_mesh_of(pm)
        diag, upper = random_spd(m, seed=31)
        rng = np.random.default_rng(32)
        rA = rng.uniform(-1, 1, m.N); b = rng.uniform(-1, 1, m.N)
        rD = oracle.dic_calc_rd(m, diag, upper)
        w = oracle.dic_precondition(m, rD, upper, rA)
        for _ in range(2):
            assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
        xo, po, ho = oracle.pcg_solve(m, diag, upper, np.zeros(m.N), b, tol=1e-9, maxIter=300)
        xg, pg, hg = ldu.pcg_solve(diag, upper, np.zeros(m.N), b, tol=1e-9, maxIter=300)
        assert pg.nIterations == po.nIterations and np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
        ldu.release()
    finally:
        lib.set_option("sweep_publish", -1); lib.set_option("sweep_bulk_tail", 4)
