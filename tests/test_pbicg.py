"""f1: b200PBiCG + DILU (the momentum predictor's solver, [REF region3d/UEqn.H:19-34]) through the C-ABI vs the oracle restatement of
[UPSTREAM DILUPreconditioner.C, PBiCG.C].  calcReciprocalD / precondition / preconditionT: BIT-EXACT (the closed-tile sweeps apply each
row's faces in upstream's order: losort order forward, face order backward, and the reverse for the transposed system).  PBiCG: identical
iteration counts, residual history to 1e-8 relative, solution to 1e-10; plus an independent true-residual check and a known answer: on a
1-D chain DILU is the exact LU factorisation of the (asymmetric) tridiagonal matrix, so PBiCG converges in ONE iteration to numpy's
direct solution."""
import numpy as np
import pytest
from cases import oracle_mesh_of

MESHES = [((12, 5, 7), (0.45, 0.6, 0.3)), ((16, 8, 10), None), ((33, 3, 5), None), ((2, 1, 1), None)]


def random_asym(m, seed=1, shift=(0.05, 0.2)):
    """convection-diffusion like: lower = -(d + c), upper = -(d - c), diagonally dominant"""
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.5, 1.5, m.F); c = rng.uniform(-0.4, 0.4, m.F)
    lower = -(d + c); upper = -(d - c)
    diag = np.zeros(m.N)
    np.add.at(diag, m.owner[:m.F], -lower); np.add.at(diag, m.neighbour, -upper)
    diag += rng.uniform(shift[0], shift[1], m.N)
    return diag, upper, lower


@pytest.mark.parametrize("dims,weir", MESHES)
def test_dilu_bitwise(lib, oracle, dims, weir):
    pm = lib.polymesh_box(*dims, 5.0, 1.0, 2.0, weir=weir)
    ldu = pm.ldu(); m = oracle_mesh_of(pm)
    diag, upper, lower = random_asym(m, seed=3)
    rng = np.random.default_rng(7)
    rD = oracle.dilu_calc_rd(m, diag, upper, lower)
    assert np.array_equal(rD, ldu.dilu_calc_rd(diag, upper, lower))
    rA = rng.uniform(-1, 1, m.N)
    for tr in (False, True):
        w = oracle.dilu_precondition(m, rD, upper, lower, rA, tr)
        assert np.array_equal(w, ldu.dilu_precondition(rD, upper, lower, rA, tr))
        assert np.array_equal(w, ldu.dilu_precondition(rD, upper, lower, rA, tr))      # re-armed ready flags
    # with lower == upper DILU degenerates to DIC
    assert np.array_equal(ldu.dilu_calc_rd(diag, upper, upper), ldu.dic_calc_rd(diag, upper))
    assert np.array_equal(ldu.dilu_precondition(rD, upper, upper, rA), ldu.dic_precondition(rD, upper, rA))
    # M^-T is the transpose of M^-1:  y.(M^-1 x) == x.(M^-T y)
    x, y = rng.normal(size=m.N), rng.normal(size=m.N)
    a = np.dot(y, ldu.dilu_precondition(rD, upper, lower, x, False)); b = np.dot(x, ldu.dilu_precondition(rD, upper, lower, y, True))
    assert abs(a - b) <= 1e-11 * max(abs(a), 1e-300)
    ldu.release()


@pytest.mark.parametrize("dims,weir,tol,relTol", [((12, 5, 7), (0.45, 0.6, 0.3), 1e-10, 0.0), ((16, 8, 10), None, 1e-6, 0.1), ((2, 1, 1), None, 1e-12, 0.0)])
def test_pbicg_matches_oracle(lib, oracle, dims, weir, tol, relTol):
    pm = lib.polymesh_box(*dims, 5.0, 1.0, 2.0, weir=weir)
    ldu = pm.ldu(); m = oracle_mesh_of(pm)
    diag, upper, lower = random_asym(m, seed=9)
    rng = np.random.default_rng(13)
    b = rng.uniform(-1, 1, m.N); x0 = rng.uniform(-0.1, 0.1, m.N)
    xo, po, ho = oracle.pbicg_solve(m, diag, upper, lower, x0, b, tol=tol, relTol=relTol, maxIter=300)
    xg, pg, hg = ldu.pbicg_solve(diag, upper, lower, x0, b, tol=tol, relTol=relTol, maxIter=300)
    assert pg.nIterations == po.nIterations and pg.converged == po.converged == 1 and pg.singular == 0
    assert np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    assert np.abs(xg - xo).max() <= 1e-10 * max(np.abs(xo).max(), 1e-300)
    r = b - oracle.amul(m, diag, upper, lower, xg)                      # true residual, independent of both solvers
    assert np.abs(r).sum() <= 10 * max(tol, relTol) * (np.abs(b).sum() + np.abs(oracle.amul(m, diag, upper, lower, x0)).sum() + 1)
    # controls: maxIter, minIter, converged initial guess
    for kw in (dict(tol=1e-30, maxIter=5), dict(tol=1.0, minIter=2), dict(tol=1.0, maxIter=0)):
        _, p1, h1 = oracle.pbicg_solve(m, diag, upper, lower, x0, b, **kw)
        _, p2, h2 = ldu.pbicg_solve(diag, upper, lower, x0, b, **kw)
        assert (p1.nIterations, p1.converged) == (p2.nIterations, p2.converged), kw
        assert np.allclose(h2, h1, rtol=1e-8, atol=1e-13 * h1[0])
    # a PCG solve on the same handle afterwards still works (the sweeps share the ready-flag vectors and operand blocks)
    ds, us = np.abs(diag) + 1.0, -np.abs(upper)
    np.add.at(ds, m.owner[:m.F], -us); np.add.at(ds, m.neighbour, -us)
    xs, ps, hs = oracle.pcg_solve(m, ds, us, x0, b, tol=1e-9)
    xq, pq, hq = ldu.pcg_solve(ds, us, x0, b, tol=1e-9)
    assert pq.nIterations == ps.nIterations and np.allclose(hq, hs, rtol=1e-8, atol=1e-13 * hs[0])
    ldu.release()


def test_pbicg_known_answer_on_a_chain(lib, oracle):
    """1-D chain: the LDU matrix is tridiagonal, DILU is its exact LU factorisation (no fill-in), so upstream's PBiCG reaches the direct
    solution after ONE iteration -- a known answer that does not come from the oracle (numpy.linalg.solve)."""
    n = 40
    pm = lib.polymesh_box(n, 1, 1, 4.0, 1.0, 1.0)
    ldu = pm.ldu(); m = oracle_mesh_of(pm)
    rng = np.random.default_rng(2)
    lower = -rng.uniform(0.6, 1.4, m.F); upper = -rng.uniform(0.2, 1.0, m.F)
    diag = rng.uniform(3.0, 4.0, m.N)
    A = np.diag(diag)
    for f in range(m.F):
        A[m.neighbour[f], m.owner[f]] = lower[f]; A[m.owner[f], m.neighbour[f]] = upper[f]
    b = rng.uniform(-1, 1, m.N)
    x_direct = np.linalg.solve(A, b)
    for solve in (lambda: oracle.pbicg_solve(m, diag, upper, lower, np.zeros(n), b, tol=1e-12, maxIter=50),
                  lambda: ldu.pbicg_solve(diag, upper, lower, np.zeros(n), b, tol=1e-12, maxIter=50)):
        x, p, h = solve()
        assert p.nIterations == 1 and p.converged == 1
        assert np.abs(x - x_direct).max() <= 1e-13 * np.abs(x_direct).max()
    # rD recurrence by hand: rD_0 = 1/d_0 ; rD_i = 1/(d_i - u_{i-1} l_{i-1} rD_{i-1})
    rd = np.empty(n); rd[0] = 1.0 / diag[0]
    for i in range(1, n):
        rd[i] = 1.0 / (diag[i] - upper[i - 1] * lower[i - 1] / (1.0 / rd[i - 1]))
    assert np.allclose(ldu.dilu_calc_rd(diag, upper, lower), rd, rtol=1e-14, atol=0)
    ldu.release()
