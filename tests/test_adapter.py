"""(b) the foam-extend adapters themselves -- foam/b200PCG/b200PCG.C, foam/b200PBiCG/b200PBiCG.C, foam/b200MULES/b200MULES.C -- compiled
UNCHANGED against foam/miniFoam.H (a stand-in that declares the foam-extend-3.1 names they use; foam-extend cannot be installed here)
and driven the way foam-extend drives them: lduMatrix::solver::New(dict)->solve(x, b) through the run-time selection tables, and
b200::MULES::explicitSolve(psi, phi, phiPsi, 1, 0) on vol/surface fields.  CPU: the library builds, exports its entry points and FAILS
LOUDLY without a GPU.  GPU: results against the oracle, and the error paths (wrong preconditioner, wrong table, non-processor interface)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from cases import oracle_mesh_of, random_spd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FOAM = os.path.join(ROOT, "shallowinterfoam_b200", "foam")
dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


@pytest.fixture(scope="module")
def mini():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "shallowinterfoam_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", FOAM], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(FOAM, "libb200Foam_mini.so"))
    lib.mini_last_error.restype = C.c_char_p
    return lib


def solve(mini, solver, precon, pm, diag, upper, lower, x0, b, tol=1e-9, relTol=0.0, maxIter=500, ifaces=()):
    F = pm.nInternalFaces
    lo = np.ascontiguousarray(pm.owner[:F], dtype=np.int32); up = np.ascontiguousarray(pm.neighbour, dtype=np.int32)
    x = np.array(x0, dtype=np.float64); perf = np.zeros(5); name = C.create_string_buffer(64)
    sizes = np.array([len(f[0]) for f in ifaces], dtype=np.int32); ranks = np.array([f[1] for f in ifaces], dtype=np.int32)
    fcs = np.concatenate([np.asarray(f[0], dtype=np.int32) for f in ifaces]) if ifaces else np.zeros(1, dtype=np.int32)
    co = np.concatenate([np.asarray(f[2], dtype=np.float64) for f in ifaces]) if ifaces else np.zeros(1)
    rc = mini.mini_solve(solver.encode(), precon.encode(), C.c_int(pm.nCells), C.c_int(F), _i(lo), _i(up), _d(np.ascontiguousarray(diag)),
                         _d(np.ascontiguousarray(upper)), _d(None if lower is None else np.ascontiguousarray(lower)), C.c_int(len(ifaces)),
                         _i(sizes), _i(fcs), _i(ranks), _d(co), _d(co), C.c_double(tol), C.c_double(relTol), C.c_int(maxIter), _d(x),
                         _d(np.ascontiguousarray(b)), _d(perf), name)
    return rc, x, perf, name.value.decode(), mini.mini_last_error().decode()


def test_adapters_compile_against_the_stand_in_and_fail_loudly_without_a_gpu(mini):
    for sym in ("mini_solve", "mini_mules", "mini_last_error"):
        assert hasattr(mini, sym)
    # the three adapter translation units are in the library, registered in the run-time selection tables by their static objects
    out = subprocess.check_output(["nm", "-C", "--defined-only", os.path.join(FOAM, "libb200Foam_mini.so")], text=True)
    for sym in ("Foam::b200PCG::solve", "Foam::b200PBiCG::solve", "Foam::b200::MULES::explicitSolve", "Foam::b200PCG::typeName", "Foam::b200PBiCG::typeName"):
        assert sym in out, sym
    from shallowinterfoam_b200 import capi
    lib = capi.load()
    if lib.device_count() >= 1:
        pytest.skip("a GPU is present: the no-device error path cannot be shown here")
    pm = lib.polymesh_box(4, 3, 2, 1.0, 1.0, 1.0)
    F = pm.nInternalFaces
    rc, x, perf, name, err = solve(mini, "b200PCG", "DIC", pm, np.ones(pm.nCells) * 7, -np.ones(F), None, np.zeros(pm.nCells), np.ones(pm.nCells))
    assert rc == 1 and "no CUDA device" in err          # no CPU fallback behind the plug point
    rc, *_, err = solve(mini, "PCG", "DIC", pm, np.ones(pm.nCells) * 7, -np.ones(F), None, np.zeros(pm.nCells), np.ones(pm.nCells))
    assert rc == 1 and "Unknown symmetric matrix solver PCG" in err   # only the plug-ins are in the stand-in's tables


@pytest.mark.gpu
def test_b200PCG_and_b200PBiCG_through_the_selection_tables(mini, oracle):
    from conftest import get_lib
    lib = get_lib("gpu")
    pm = lib.polymesh_box(14, 6, 8, 5.0, 1.0, 2.0, weir=(0.45, 0.6, 0.3))
    m = oracle_mesh_of(pm)
    diag, upper = random_spd(m, seed=9, shift=(1e-4, 1e-3))
    rng = np.random.default_rng(13)
    b = rng.uniform(-1, 1, m.N); x0 = rng.uniform(-0.1, 0.1, m.N)
    # `solver b200PCG; preconditioner DIC;`
    rc, x, perf, name, err = solve(mini, "b200PCG", "DIC", pm, diag, upper, None, x0, b, tol=1e-9)
    assert rc == 0, err
    xo, po, ho = oracle.pcg_solve(m, diag, upper, x0, b, tol=1e-9, maxIter=500)
    assert name == "b200DICPCG" and int(perf[2]) == po.nIterations and perf[3] == 1 and perf[4] == 0
    assert abs(perf[0] - po.initialResidual) <= 1e-8 * po.initialResidual and abs(perf[1] - po.finalResidual) <= 1e-8 * po.finalResidual + 1e-13 * po.initialResidual
    assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
    # twice through the same lduAddressing: the second solve hits the library's registry (the solver object is a per-solve temporary)
    rc, x2, perf2, _, err = solve(mini, "b200PCG", "DIC", pm, diag, upper, None, x0, b, tol=1e-9)
    assert rc == 0 and np.array_equal(x2, x) and np.array_equal(perf2, perf)
    # `solver b200PBiCG; preconditioner DILU;` for an asymmetric matrix
    lower = upper * rng.uniform(0.6, 1.4, m.F)
    dg = diag + np.abs(upper).max()
    rc, xa, pa, name, err = solve(mini, "b200PBiCG", "DILU", pm, dg, upper, lower, x0, b, tol=1e-9)
    assert rc == 0, err
    xb, pb, hb = oracle.pbicg_solve(m, dg, upper, lower, x0, b, tol=1e-9, maxIter=500)
    assert name == "b200DILUPBiCG" and int(pa[2]) == pb.nIterations and np.abs(xa - xb).max() <= 1e-10 * np.abs(xb).max()
    # error paths: what a mis-typed fvSolution must produce
    rc, *_, err = solve(mini, "b200PCG", "GAMG", pm, diag, upper, None, x0, b)
    assert rc == 1 and "preconditioner GAMG is not available" in err
    rc, *_, err = solve(mini, "b200PCG", "", pm, diag, upper, None, x0, b)
    assert rc == 1 and "preconditioner" in err
    rc, *_, err = solve(mini, "b200PCG", "DIC", pm, dg, upper, lower, x0, b)
    assert rc == 1 and "Unknown asymmetric matrix solver b200PCG" in err          # PCG is registered in the symMatrix table only
    rc, *_, err = solve(mini, "b200PBiCG", "DIC", pm, dg, upper, lower, x0, b)
    assert rc == 1 and "implements DILU only" in err
    F = pm.nInternalFaces
    fc = pm.owner[F:F + 5]
    rc, *_, err = solve(mini, "b200PCG", "DIC", pm, diag, upper, None, x0, b, ifaces=[(fc, -1, np.zeros(5))])
    assert rc == 1 and "only processor interfaces" in err


@pytest.mark.gpu
def test_b200MULES_explicitSolve_through_the_adapter(mini, oracle):
    from conftest import get_lib
    from cases import channel_case
    lib = get_lib("gpu")
    L = (5.0, 1.0, 2.0)
    pm = lib.polymesh_box(16, 6, 8, *L, weir=(0.45, 0.6, 0.3))
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    g = pm.geometry(); F = pm.nInternalFaces; B = pm.nBoundaryFaces
    bcO = oracle.BC(m, cs["tA"], 1, cs["rvA"])
    rng = np.random.default_rng(4)
    psi0 = np.clip(cs["alpha"] + rng.uniform(-0.2, 0.2, m.N) * (cs["alpha"] > 0) * (cs["alpha"] < 1) + rng.uniform(0, 0.3, m.N) * 0.0, 0, 1)
    psi0 = np.clip(0.5 + 0.5 * np.sin(3 * g["C"][:, 0]) * np.cos(2 * g["C"][:, 2]), 0, 1)
    psib = oracle.bc_evaluate(m, bcO, psi0)
    Uf = np.zeros((F + B, 3)); Uf[:, 0] = 1.0; Uf[:, 2] = 0.3 * np.sin(g["Cf"][:, 0])
    phi = np.einsum("ij,ij->i", Uf, g["Sf"])
    phi[F:][~np.isin(np.arange(B), np.r_[pm.patch_slice("inlet"), pm.patch_slice("outlet"), pm.patch_slice("atmosphere")])] = 0.0
    phiHO = oracle.flux_limited(m, 3, phi, psi0, psib)          # linear (unbounded) high-order flux
    dt = 0.02
    po, fo_, lam = oracle.mules_explicit_solve(m, bcO, psi0, psib, psi0, phi, phiHO, dt)
    psi = psi0.copy(); phiPsi = phiHO.copy(); n = C.c_int()
    pstart = np.ascontiguousarray(pm.patchStart, dtype=np.int32); psize = np.ascontiguousarray(pm.patchSize, dtype=np.int32)
    prank = np.full(pm.nPatches, -1, dtype=np.int32)
    rc = mini.mini_mules(C.c_int(m.N), C.c_int(F), C.c_int(pm.nPatches), _i(pstart), _i(psize), _i(prank),
                         _i(np.ascontiguousarray(pm.owner, dtype=np.int32)), _i(np.ascontiguousarray(pm.neighbour, dtype=np.int32)),
                         _d(np.ascontiguousarray(g["Sf"])), _d(g["magSf"]), _d(np.ascontiguousarray(g["Cf"])), _d(np.ascontiguousarray(g["C"])),
                         _d(g["V"]), _d(g["weights"]), _d(g["deltaCoeffs"]), C.c_double(dt), _d(psi), _d(psib), _d(psi0), _d(phi), _d(phiPsi),
                         C.byref(n))
    assert rc == 0, mini.mini_last_error().decode()
    assert n.value == 2                                        # psi.correctBoundaryConditions() before and after: the BC dispatch stays on the host
    assert np.array_equal(psi, po) and np.array_equal(phiPsi, fo_)     # bitwise the oracle's MULES
    assert psi.min() >= -1e-12 and psi.max() <= 1 + 1e-12 and np.abs(psi - psi0).max() > 1e-3
