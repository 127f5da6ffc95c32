"""a5-a14: MULES, limited face fluxes, pressure assembly and the fvc building blocks through the C-ABI vs the oracle.
Everything here is a row/face gather in upstream's summation order without FMA contraction => BIT-EXACT, except
alpha_stats (a tree reduction: 1e-14 relative)."""
import numpy as np
import pytest
from cases import oracle_mesh_of, channel_case

FV, ZG, MX = 0, 1, 2
MESHES = [((12, 5, 7), (0.45, 0.6, 0.3)), ((16, 8, 10), None), ((3, 1, 2), None), ((1, 1, 1), None)]


def setup(lib, oracle, dims, weir, seed=0):
    L = (5.0, 1.0, 2.0)
    pm = lib.polymesh_box(*dims, *L, weir=weir)
    m = oracle_mesh_of(pm)
    dm = pm.device_mesh()
    rng = np.random.default_rng(seed)
    P = pm.nPatches
    types = [[FV, ZG, MX][(p + seed) % 3] for p in range(P)]
    rv = rng.uniform(0, 1, m.B); rg = rng.uniform(-1, 1, m.B); vf = rng.uniform(0, 1, m.B)
    bo = oracle.BC(m, types, 1, rv, rg, vf)
    bg = dm.bc(types, 1, rv, rg, vf)
    return pm, m, dm, rng, bo, bg


@pytest.mark.parametrize("dims,weir", MESHES)
def test_building_blocks_bitwise(lib, oracle, dims, weir):
    pm, m, dm, rng, bo, bg = setup(lib, oracle, dims, weir, 1)
    vf = rng.uniform(0, 1, m.N)
    vfb = oracle.bc_evaluate(m, bo, vf)
    assert np.array_equal(oracle.interpolate(m, vf, vfb), dm.interpolate(vf, vfb))
    V = rng.normal(size=(m.N, 3)); Vb = rng.normal(size=(m.B, 3))
    assert np.array_equal(oracle.interpolate(m, V, Vb, 3), dm.interpolate(V, Vb, 3))
    assert np.array_equal(oracle.sngrad(m, bo, vf, vfb), dm.sngrad(bg, vf, vfb))
    ssf = rng.normal(size=m.F + m.B)
    assert np.array_equal(oracle.surface_integrate(m, ssf), dm.surface_integrate(ssf))
    go, gbo = oracle.grad_gauss(m, bo, vf, vfb)
    gg, gbg = dm.grad_gauss(bg, vf, vfb)
    assert np.array_equal(go, gg) and np.array_equal(gbo, gbg)
    if m.N > 1:
        ro = oracle.reconstruct(m, oracle.reconstruct_tensor(m), ssf)
        assert np.array_equal(ro, dm.reconstruct(ssf))
    st_o, st_g = oracle.alpha_stats(m, vf, vfb), dm.alpha_stats(vf, vfb)
    assert np.allclose(st_o, st_g, rtol=1e-14, atol=0) and st_o[1] == st_g[1] and st_o[2] == st_g[2]
    dm.release()


def test_gradient_of_linear_field_is_exact(lib, oracle):
    pm, m, dm, rng, bo, bg = setup(lib, oracle, (10, 6, 8), None)
    g = pm.geometry()
    a = np.array([0.3, -1.2, 2.0])
    vf = g["C"] @ a
    vfb = g["Cf"][m.F:] @ a
    types = [FV] * pm.nPatches
    grad, _ = dm.grad_gauss(dm.bc(types, 1, vfb), vf, vfb)
    assert np.abs(grad - a).max() < 1e-12
    # reconstruct of a uniform-vector flux returns the vector
    ssf = g["Sf"] @ a
    assert np.abs(dm.reconstruct(ssf) - a).max() < 1e-12
    dm.release()


@pytest.mark.parametrize("dims,weir", MESHES)
@pytest.mark.parametrize("scheme", [0, 1, 2, 3])
def test_limited_fluxes_bitwise(lib, oracle, dims, weir, scheme):
    pm, m, dm, rng, bo, bg = setup(lib, oracle, dims, weir, 2)
    vf = np.clip(rng.normal(0.5, 0.6, m.N), -0.1, 1.1)     # includes values outside [0,1] and flat regions
    vf[: m.N // 3] = 1.0
    vfb = oracle.bc_evaluate(m, bo, vf)
    flux = rng.normal(size=m.F + m.B)
    flux[::7] = 0.0
    grad = oracle.grad_gauss(m, bo, vf, vfb)[0] if scheme == 0 else None
    fo_ = oracle.flux_limited(m, scheme, flux, vf, vfb, grad)
    fg = dm.flux_limited(scheme, bg, flux, vf, vfb)
    assert np.array_equal(fo_, fg)
    dm.release()


@pytest.mark.parametrize("dims,weir", MESHES)
def test_mules_bitwise_bounded_conservative(lib, oracle, dims, weir):
    pm, m, dm, rng, bo, bg = setup(lib, oracle, dims, weir, 3)
    g = pm.geometry()
    psi0 = np.where(g["C"][:, 0] < 2.5, 1.0, 0.0) * rng.uniform(0.9, 1.0, m.N)
    psi = psi0.copy()
    types = [ZG] * pm.nPatches
    bo = oracle.BC(m, types, 1); bg = dm.bc(types, 1)
    psib = oracle.bc_evaluate(m, bo, psi)
    # a divergence-free-ish flux: uniform velocity (1, 0.2, 0.1) & Sf
    U = np.array([1.0, 0.2, 0.1])
    phi = g["Sf"] @ U
    dt = 0.2 * (5.0 / dims[0]) / 1.0
    grad = oracle.grad_gauss(m, bo, psi, psib)[0]
    phiPsi = oracle.flux_limited(m, 0, phi, psi, psib, grad)
    po, fo_, lo = oracle.mules_explicit_solve(m, bo, psi, psib, psi0, phi, phiPsi, dt)
    pg, fg, lg = dm.mules_explicit_solve(bg, psi, psib, psi0, phi, phiPsi, dt)
    assert np.array_equal(lo, lg) and np.array_equal(fo_, fg) and np.array_equal(po, pg)
    assert lg.min() >= 0.0 and lg.max() <= 1.0
    if m.N > 10:
        assert pg.min() > -1e-6 and pg.max() < 1 + 1e-6                      # MULES boundedness
        # conservation: d(sum psi V) = -dt * sum of boundary fluxes
        lhs = ((pg - psi0) * g["V"]).sum()
        rhs = -dt * fg[m.F:].sum()
        assert abs(lhs - rhs) < 1e-13 * max(1.0, abs(rhs))
    # limiter alone, 0..3 iterations
    phiBD = np.where(phi[:m.F] >= 0, psi[pm.owner[:m.F]], psi[pm.neighbour]) * phi[:m.F]
    phiBD = np.concatenate([phiBD, phi[m.F:] * psib])
    for n in (0, 1, 2, 3, 4):
        assert np.array_equal(oracle.mules_limiter(m, bo, psi, psib, psi0, phiBD, phiPsi - phiBD, dt, nIter=n),
                              dm.mules_limiter(bg, psi, psib, psi0, phiBD, phiPsi - phiBD, dt, nIter=n)), n
    dm.release()


@pytest.mark.parametrize("dims,weir", MESHES)
def test_pressure_assembly_bitwise(lib, oracle, dims, weir):
    pm, m, dm, rng, bo, bg = setup(lib, oracle, dims, weir, 4)
    gammaf = rng.uniform(1e-3, 1.0, m.F + m.B)
    pd = rng.normal(size=m.N)
    pb = oracle.bc_evaluate(m, bo, pd)
    phi = rng.normal(size=m.F + m.B)
    o = oracle.laplacian_assemble(m, bo, gammaf, pb, phi)
    g = dm.laplacian_assemble(bg, gammaf, pb, phi)
    for a, b in zip(o, g):
        assert np.array_equal(a, b)
    o2 = oracle.laplacian_assemble(m, bo, gammaf, pb, None)
    g2 = dm.laplacian_assemble(bg, gammaf, pb, None)
    assert np.array_equal(o2[2], g2[2]) and np.all(g2[2] == 0)
    diag, upper, source, ic, bcv = o
    assert np.array_equal(oracle.matrix_flux(m, upper, upper, ic, bcv, pd), dm.matrix_flux(upper, upper, ic, bcv, pd))
    lower = rng.normal(size=m.F)
    assert np.array_equal(oracle.matrix_flux(m, upper, lower, ic, bcv, pd), dm.matrix_flux(upper, lower, ic, bcv, pd))
    # the assembled operator is symmetric negative semi-definite with zero row sums away from fixed-value patches
    if m.F:
        rowsum = diag.copy(); np.add.at(rowsum, pm.owner[:m.F], upper); np.add.at(rowsum, pm.neighbour, upper)
        assert np.abs(rowsum).max() < 1e-12 * np.abs(diag).max()
    dm.release()


def test_laplacian_of_quadratic_on_uniform_box(lib, oracle):
    """analytic check: div(grad(x^2)) = 2 on a uniform hex box away from the boundary"""
    pm = lib.polymesh_box(12, 6, 6, 6.0, 3.0, 3.0)
    dm = pm.device_mesh()
    g = pm.geometry()
    psi = g["C"][:, 0] ** 2
    psib = g["Cf"][pm.nInternalFaces:, 0] ** 2
    bc = dm.bc([FV] * pm.nPatches, 1, psib)
    diag, upper, source, ic, bcv = dm.laplacian_assemble(bc, np.ones(pm.nInternalFaces + pm.nBoundaryFaces), psib)
    lap = dm.ldu.amul(diag, upper, psi) / g["V"]
    i = np.rint(g["C"][:, 0] / 0.5 - 0.5)
    interior = (i > 0) & (i < 11)
    wall = np.zeros(pm.nCells, dtype=bool); wall[pm.owner[pm.nInternalFaces:]] = True
    assert np.abs(lap[~wall] - 2.0).max() < 1e-10
    dm.release()
