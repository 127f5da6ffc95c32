"""BASELINE configs[3] in miniature: irregular (polyhedral-like) LDU addressing -- rows with 0..12+ lower / upper neighbours,
long-range faces, no stride structure.  Exercises the chunked rows (more than 3 sources) and the isolated rows of the DIC sweep tiles, the plan's
monotone potentials on an arbitrary DAG, and Amul / sumA / DIC / PCG parity (bitwise / 1e-8) against the oracle."""
import numpy as np
import pytest
from oracle import foam_oracle as fo
from oracle import mesh_oracle as mo


def random_ldu(n, seed, max_deg=6, window=40, far=0.1):
    """upper-triangular-ordered random addressing: faces sorted by (lower, upper), lower < upper, no duplicates"""
    rng = np.random.default_rng(seed)
    lo, up = [], []
    for c in range(n - 1):
        k = int(rng.integers(0, max_deg + 1))
        cand = set()
        for _ in range(k):
            if rng.uniform() < far:
                u = int(rng.integers(c + 1, n))
            else:
                u = int(c + 1 + rng.integers(0, window))
            if u < n:
                cand.add(u)
        for u in sorted(cand):
            lo.append(c); up.append(u)
    return np.asarray(lo, dtype=np.int32), np.asarray(up, dtype=np.int32)


def oracle_ldu_mesh(n, lower, upper):
    F = lower.size
    z3 = np.zeros((F, 3)); z1 = np.zeros(F)
    geom = dict(Sf=z3, magSf=z1, Cf=z3, C=np.zeros((n, 3)), V=np.ones(n), weights=z1, deltaCoeffs=z1)
    return fo.Mesh(lower, upper, [], geom, n)


def spd(m, seed):
    rng = np.random.default_rng(seed)
    upper = -rng.uniform(0.5, 1.5, m.F)
    diag = np.zeros(m.N)
    np.add.at(diag, m.owner[:m.F], -upper); np.add.at(diag, m.neighbour, -upper)
    diag += rng.uniform(0.05, 0.2, m.N)
    return diag, upper


@pytest.mark.parametrize("n,seed,max_deg,window", [(300, 1, 6, 40), (1500, 2, 12, 200), (400, 3, 3, 5), (64, 4, 2, 64)])
def test_irregular_ldu_parity(lib, oracle, n, seed, max_deg, window):
    lower, upper_addr = random_ldu(n, seed, max_deg, window)
    ldu = lib.ldu_create(n, lower, upper_addr)
    m = oracle_ldu_mesh(n, lower, upper_addr)
    # the plan: bit-exact renumbering, monotone potentials, topological ticket order
    rowCell, level, nLevels = ldu.renumbering()
    bl, tr, ts = ldu.tiling()
    plan = mo.sweep_plan(lower, upper_addr, n, bl, tr)
    assert np.array_equal(level, plan["level"]) and np.all(level[lower] < level[upper_addr])
    if not plan["chunked"]:   # (rows with more than 3 sources are applied as chains of chunks; that plan is checked by its properties)
        assert np.array_equal(rowCell, plan["rowCell"]) and np.array_equal(ts, plan["tileStart"])
        for k in ("potJ", "potK"):
            assert np.all(plan[k][lower] <= plan[k][upper_addr])
    assert sorted(rowCell.tolist()) == list(range(n)) and ts[0] == 0 and ts[-1] == n and np.all(np.diff(ts) > 0)
    inv = np.empty(n, dtype=np.int64); inv[rowCell] = np.arange(n)
    tile_of_row = np.searchsorted(ts, np.arange(n), side="right") - 1
    assert np.all(tile_of_row[inv[lower]] <= tile_of_row[inv[upper_addr]])
    # arithmetic
    diag, upper = spd(m, seed + 10)
    rng = np.random.default_rng(seed + 20)
    psi = rng.uniform(-1, 1, n)
    assert np.array_equal(fo.amul(m, diag, upper, upper, psi), ldu.amul(diag, upper, psi))
    lowc = -rng.uniform(0.5, 1.5, m.F)
    assert np.array_equal(fo.amul(m, diag, upper, lowc, psi), ldu.amul(diag, upper, psi, lower=lowc))
    assert np.array_equal(fo.sumA(m, diag, upper, lowc), ldu.sumA(diag, upper, lower=lowc))
    rD = fo.dic_calc_rd(m, diag, upper)
    assert np.array_equal(rD, ldu.dic_calc_rd(diag, upper))
    rA = rng.uniform(-1, 1, n)
    w = fo.dic_precondition(m, rD, upper, rA)
    for _ in range(2):
        assert np.array_equal(w, ldu.dic_precondition(rD, upper, rA))
    b = rng.uniform(-1, 1, n)
    xo, po, ho = fo.pcg_solve(m, diag, upper, np.zeros(n), b, tol=1e-10, maxIter=400)
    xg, pg, hg = ldu.pcg_solve(diag, upper, np.zeros(n), b, tol=1e-10, maxIter=400)
    assert pg.nIterations == po.nIterations and pg.converged == 1
    assert np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    assert np.abs(xg - xo).max() <= 1e-10 * np.abs(xo).max()
    ldu.release()


def test_irregular_small_tiles(lib, oracle):
    """tiny tiles on an irregular graph: halo-heavy tiles, oversized classes cut into pieces, many tile levels"""
    lib.set_option("sweep_band_levels", 2); lib.set_option("sweep_tile_rows", 24); lib.set_option("sweep_threads", 32)
    try:
        n = 700
        lower, upper_addr = random_ldu(n, 7, 8, 60)
        ldu = lib.ldu_create(n, lower, upper_addr)
        m = oracle_ldu_mesh(n, lower, upper_addr)
        diag, upper = spd(m, 8)
        rD = fo.dic_calc_rd(m, diag, upper)
        assert np.array_equal(rD, ldu.dic_calc_rd(diag, upper))
        rA = np.random.default_rng(9).uniform(-1, 1, n)
        assert np.array_equal(fo.dic_precondition(m, rD, upper, rA), ldu.dic_precondition(rD, upper, rA))
        ldu.release()
    finally:
        lib.set_option("sweep_band_levels", 16); lib.set_option("sweep_tile_rows", 4096); lib.set_option("sweep_threads", 256)
