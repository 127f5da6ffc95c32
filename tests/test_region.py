"""The device-resident 3D-region step (region3d/solve3d.H replay) vs the oracle step driver, N steps on identical inputs.
Tolerances (north_star): per-solve residual history 1e-8 relative (with the FP64 round-off floor 1e-13*initial);
identical PCG iteration counts; fields: L-inf <= 1e-10 * max|field| after N steps (observed ~1e-14)."""
import numpy as np
import pytest
from cases import oracle_mesh_of, channel_case, default_controls

FIELD_TOL = 1e-10


def build(lib, oracle, dims, weir, **kw):
    from shallowinterfoam_b200 import capi
    L = (5.0, 1.0, 2.0)
    pm = lib.polymesh_box(*dims, *L, weir=weir)
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    if kw.get("momentumPredictor"):
        # make all three velocity components non-trivial (the plain channel has Uy = 0 up to round-off: a solve of pure noise)
        C_ = pm.geometry()["C"]
        cs["U"][:, 1] = 0.2 * np.sin(2 * np.pi * C_[:, 0] / L[0]) * np.cos(np.pi * C_[:, 1] / L[1]) * cs["alpha"]
        cs["U"][:, 2] = 0.1 * np.cos(2 * np.pi * C_[:, 0] / L[0]) * np.sin(np.pi * C_[:, 2] / L[2])
    dm = pm.device_mesh()
    co = default_controls(oracle.OrcControls); cg = default_controls(capi.Region3dControls)
    for k, v in kw.items():
        setattr(co, k, v); setattr(cg, k, v)
    ro = oracle.Region(m, co, oracle.BC(m, cs["tA"], 1, cs["rvA"]), oracle.BC(m, cs["tU"], 3, cs["rvU"]), oracle.BC(m, cs["tP"], 1),
                       cs["alpha"], cs["U"], cs["pd"])
    rg = dm.region3d(cg, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    return pm, m, dm, ro, rg


def compare(ro, rg):
    fo_, fg = ro.fields(), rg.fields()
    for k in ("alpha1", "U", "pd", "phi", "p", "rho"):
        scale = max(np.abs(fo_[k]).max(), 1e-300)
        assert np.abs(fo_[k] - fg[k]).max() <= FIELD_TOL * scale, k
    po, pg = ro.perfs(), rg.perfs()
    assert [p[2] for p in po] == [p[2] for p in pg]
    for a, b in zip(po, pg):
        assert abs(a[0] - b[0]) <= 1e-8 * a[0] and abs(a[1] - b[1]) <= 1e-8 * a[1] + 1e-13 * a[0]
        assert a[3] == b[3] and a[4] == b[4]
    ho, hg = ro.res_hist(), rg.res_hist()
    assert ho.shape == hg.shape and np.allclose(ho, hg, rtol=1e-8, atol=1e-13)
    assert np.allclose(ro.alpha_stats(), rg.alpha_stats(), rtol=1e-12, atol=1e-25)
    assert np.allclose(ro.cont_errs(), rg.cont_errs(), rtol=1e-6, atol=1e-16)


@pytest.mark.parametrize("dims,weir,kw,nsteps", [
    ((20, 6, 10), (0.45, 0.55, 0.3), {}, 3),
    ((16, 4, 8), None, dict(nAlphaSubCycles=1, nCorr=2, adjustPhi=0, divUScheme=3), 2),
    ((16, 4, 8), None, dict(nAlphaSubCycles=3, nAlphaCorr=2, nNonOrthCorr=1, mulesIncludeOwn=1), 2),
    # f1: momentum predictor on (the reference's default), b200PBiCG + DILU per velocity component [REF region3d/UEqn.H:19-34]
    ((20, 6, 10), (0.45, 0.55, 0.3), dict(momentumPredictor=1, UTol=1e-7, URelTol=0.0), 3),
    ((16, 4, 8), None, dict(momentumPredictor=1, UTol=1e-6, URelTol=0.1, divUScheme=3, nCorr=2), 2),
])
def test_region_step_matches_oracle(lib, oracle, dims, weir, kw, nsteps):
    pm, m, dm, ro, rg = build(lib, oracle, dims, weir, **kw)
    compare(ro, rg)                                  # initial phi / rho / interface
    ro.correct_phi(); rg.correct_phi()               # shallowInterFoam.C:85-90
    compare(ro, rg)
    ro.clear_log(); rg.clear_log()
    for s in range(nsteps):
        ro.step(); rg.step()
        compare(ro, rg)
        ro.clear_log(); rg.clear_log()
    st = rg.alpha_stats()
    rg.release(); dm.release()


def test_step_host_round_trip_equals_resident_steps(lib, oracle):
    """the end-to-end (host-buffer) step and the resident step are the same computation"""
    pm, m, dm, ro, rg = build(lib, oracle, (16, 4, 8), None)
    pm2, m2, dm2, ro2, rg2 = build(lib, oracle, (16, 4, 8), None)
    rg.correct_phi(); rg2.correct_phi()
    f = rg2.fields()
    a, U, pd, phi, p = f["alpha1"].copy(), f["U"].copy().reshape(-1), f["pd"].copy(), f["phi"].copy(), f["p"].copy()
    for s in range(2):
        rg.step()
        rg2.step_host(a, U, pd, phi, p)
    g = rg.fields()
    assert np.array_equal(g["alpha1"], a) and np.array_equal(g["U"].reshape(-1), U) and np.array_equal(g["pd"], pd)
    assert np.array_equal(g["phi"], phi) and np.array_equal(g["p"], p)
    for r in (rg, rg2):
        r.release()
    dm.release(); dm2.release()


def test_mixed_patch_upload_and_patch_internal_gather(lib, oracle):
    """host-evaluated mixed BCs (the sif* coupling path): upload refValue/refGrad/valueFraction, read patch-internal values"""
    from shallowinterfoam_b200 import capi
    L = (5.0, 1.0, 2.0)
    pm = lib.polymesh_box(12, 4, 6, *L)
    cs = channel_case(pm, *L)
    dm = pm.device_mesh()
    tA = list(cs["tA"]); p_out = pm.patchNames.index("outlet"); tA[p_out] = 2
    cg = default_controls(capi.Region3dControls)
    rg = dm.region3d(cg, dm.bc(tA, 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    sl = pm.patch_slice("outlet"); n = sl.stop - sl.start
    rg.set_patch_mixed(0, p_out, np.full(n, 0.25), np.zeros(n), np.ones(n))
    cells = pm.owner[pm.nInternalFaces:][sl]
    assert np.array_equal(rg.patch_internal(0, p_out), cs["alpha"][cells])
    assert np.array_equal(rg.patch_internal(1, p_out).reshape(-1, 3), cs["U"][cells])
    with pytest.raises(capi.B200Error):
        rg.set_patch_mixed(0, pm.patchNames.index("inlet"), np.zeros(1), np.zeros(1), np.zeros(1))   # not a mixed patch
    rg.step()
    f = rg.fields()
    assert np.isfinite(f["alpha1"]).all() and np.isfinite(f["U"]).all()
    rg.release(); dm.release()


def test_courant_number_on_device(lib, oracle):
    """f3: mean / max Courant number and velocity magnitude reduced on the device [REF include/CourantNoMultiRegion.H:78-92]"""
    pm, m, dm, ro, rg = build(lib, oracle, (20, 6, 10), (0.45, 0.55, 0.3))
    ro.correct_phi(); rg.correct_phi()
    for s in range(2):
        co, cg = ro.courant(), rg.courant()
        assert co[1] > 0 and np.allclose(co, cg, rtol=1e-10, atol=0)   # (phi itself agrees to 1e-10)
        # independent formula from the downloaded phi (numpy)
        g = pm.geometry(); F = pm.nInternalFaces
        phi = np.abs(rg.fields()["phi"]); dt = 0.005
        assert np.isclose(cg[0], (g["deltaCoeffs"][:F] * phi[:F]).sum() / g["magSf"][:F].sum() * dt, rtol=1e-12)
        assert np.isclose(cg[1], (g["deltaCoeffs"] * phi / g["magSf"]).max() * dt, rtol=1e-14)
        ro.step(); rg.step()
    rg.release(); dm.release()


def test_set_reference_closed_domain(lib, oracle):
    """a11: fvMatrix::setReference(pdRefCell, pdRefValue) [REF region3d/pEqn.H:30]: in a closed box no pd patch fixes the value,
    pd.needReference() is true, and the reference cell pins the level of the otherwise singular pd / pcorr systems"""
    from shallowinterfoam_b200 import capi
    L = (2.0, 1.0, 1.0)
    pm = lib.polymesh_box(12, 5, 6, *L)
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    # closed tank at rest with a wavy interface: every U patch is a wall, every pd patch zeroGradient, alpha zeroGradient
    tA = [1] * pm.nPatches; tU = [0] * pm.nPatches; tP = [1] * pm.nPatches
    C_ = pm.geometry()["C"]
    U0 = np.zeros((pm.nCells, 3)); rvU = np.zeros((pm.nBoundaryFaces, 3))
    U0[:, 0] = 0.3 * np.sin(np.pi * C_[:, 0] / L[0]) * np.cos(np.pi * C_[:, 2] / L[2])       # a sloshing cell: no flow through the walls
    U0[:, 2] = -0.3 * (L[2] / L[0]) * np.cos(np.pi * C_[:, 0] / L[0]) * np.sin(np.pi * C_[:, 2] / L[2])
    dm = pm.device_mesh()
    for ref_cell, ref_val in ((7, 0.0), (40, 3.5)):
        co = default_controls(oracle.OrcControls); cg = default_controls(capi.Region3dControls)
        for c in (co, cg):
            c.pdRefCell = ref_cell; c.pdRefValue = ref_val; c.adjustPhi = 1
        ro = oracle.Region(m, co, oracle.BC(m, tA, 1), oracle.BC(m, tU, 3, rvU), oracle.BC(m, tP, 1), cs["alpha"], U0, cs["pd"])
        rg = dm.region3d(cg, dm.bc(tA, 1), dm.bc(tU, 3, rvU), dm.bc(tP, 1), cs["alpha"], U0, cs["pd"])
        ro.correct_phi(); rg.correct_phi()
        for s in range(2):
            ro.step(); rg.step()
        compare(ro, rg)
        pg = rg.perfs()
        assert all(p[3] == 1 and p[4] == 0 for p in pg)             # converged, not singular
        f = rg.fields()
        assert np.isfinite(f["pd"]).all()
        rg.release()
    # without a reference cell the same system has no unique level: the two runs above must differ by a (nearly) uniform shift of pd
    dm.release()


def test_restore_restarts_from_saved_fields(lib, oracle):
    """b200_region3d_restore: the checkpoint the bench uses so that every leg times the SAME time steps"""
    pm, m, dm, ro, rg = build(lib, oracle, (16, 4, 8), None)
    rg.correct_phi()
    rg.step(2)
    ck = rg.fields()
    rg.clear_log(); rg.step(2)
    a = rg.fields(); ia = [p[2] for p in rg.perfs()]
    rg.restore(ck["alpha1"], ck["U"], ck["pd"], ck["phi"])
    rg.clear_log(); rg.step(2)
    b = rg.fields(); ib = [p[2] for p in rg.perfs()]
    assert ia == ib
    for k in ("alpha1", "U", "pd", "phi", "p", "rho"):
        assert np.array_equal(a[k], b[k]), k
    rg.release(); dm.release()
