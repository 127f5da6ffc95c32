"""a10 / a12 `corrected` branch: non-orthogonal correction of fvm::laplacian(rUAf, pd) and of the snGrads of pEqn.H
[REF region3d/pEqn.H:16-45; readRegion3dPISOControls.H nNonOrthogonalCorrectors] [UPSTREAM surfaceInterpolation::makeCorrectionVectors,
correctedSnGrad::correction, gaussLaplacianScheme::fvmLaplacian (faceFluxCorrection, source -= V*div), fvMatrix::flux].
 * analytic: on a sheared (non-orthogonal, not skewed) mesh the corrected Laplacian of a LINEAR field vanishes to round-off in every
   cell, the uncorrected one does not (the whole point of the correction);
 * the correction field is bitwise the oracle's on sheared, jittered (configs[3] skewed) and polyhedral meshes;
 * region steps with `corrected` schemes and nNonOrthogonalCorrectors = 1, 2 keep the parity bars (iterations equal, fields 1e-10);
 * on an orthogonal mesh `corrected` is a no-op (mesh.orthogonal()), bitwise."""
import numpy as np
import pytest
from oracle import mesh_oracle as mo
from cases import oracle_mesh_of, channel_case, default_controls

L = (5.0, 1.0, 2.0)
SHEAR = (0.35, 0.2, 0.25)


def test_sheared_mesh_bit_exact_and_non_orthogonal():
    from shallowinterfoam_b200 import capi
    lib = capi.load()
    dims = (9, 5, 6)
    pm = lib.polymesh_box(*dims, *L, shear=SHEAR)
    ref = mo.shear_points(mo.block_mesh_box(*dims, *L), *SHEAR)
    assert np.array_equal(pm.points, ref["points"])
    g, gr = pm.geometry(), mo.geometry(ref)
    for k in g:
        assert np.array_equal(g[k], gr[k]), k
    F = pm.nInternalFaces
    # not skewed: the face centre lies on the line between the two cell centres (affine maps keep midpoints)
    d = g["C"][pm.neighbour] - g["C"][pm.owner[:F]]
    e = g["Cf"][:F] - 0.5 * (g["C"][pm.neighbour] + g["C"][pm.owner[:F]])
    assert np.abs(e).max() < 1e-13 and np.abs(g["weights"][:F] - 0.5).max() < 1e-13
    # non-orthogonal: Sf is not parallel to d
    cosang = np.einsum("ij,ij->i", g["Sf"][:F], d) / (g["magSf"][:F] * np.linalg.norm(d, axis=1))
    assert cosang.min() < 0.97


@pytest.mark.parametrize("kind", ["shear", "jitter", "poly"])
def test_correction_bitwise_and_linear_field_exact(lib, oracle, kind):
    dims = (10, 6, 7)
    if kind == "shear":
        pm = lib.polymesh_box(*dims, *L, shear=SHEAR)
    elif kind == "jitter":
        pm = lib.polymesh_box(*dims, *L, weir=(0.45, 0.6, 0.3), jitter=(0.3, 12345))
    else:
        base = lib.polymesh_box(*dims, *L, jitter=(0.25, 7))
        pm = base.agglomerate(mo.block_groups(*dims))
    m = oracle_mesh_of(pm)
    dm = pm.device_mesh()
    deg, orth = dm.non_orthogonality()
    dego, ortho = oracle.non_orthogonality(m)
    assert not orth and not ortho and abs(deg - dego) <= 1e-12 * dego and deg > 1.0
    g = pm.geometry(); F = pm.nInternalFaces; B = pm.nBoundaryFaces
    rng = np.random.default_rng(5)
    gam = rng.uniform(0.5, 1.5, F + B)
    bcO = oracle.BC(m, [0] * pm.nPatches, 1); bcG = dm.bc([0] * pm.nPatches, 1)
    vf = rng.uniform(-1, 1, m.N); vfb = rng.uniform(-1, 1, B)
    for gm in (None, gam):
        co = oracle.non_orth_correction(m, bcO, vf, vfb, gm)
        cg = dm.non_orth_correction(vf, vfb, gm)
        assert np.array_equal(co, cg)
        assert np.all(cg[F:] == 0.0) and np.abs(cg[:F]).max() > 0           # physical boundary faces carry no correction
    if kind != "shear":
        dm.release(); return
    # analytic check: phi = a.x is reproduced exactly by the Gauss-linear gradient on an unskewed mesh, so with ANY face diffusivity
    #   gamma|Sf| (dc (phiN - phiP) + corrVec.grad(phi)_f)  =  gamma|Sf| n.a  =  gamma Sf.a     exactly, face by face,
    # while the uncorrected flux gamma|Sf| dc (d.a) is off by the non-orthogonality (random gamma: no cancellation between faces)
    a = np.array([0.7, -1.3, 0.4])
    phi = g["C"] @ a; phib = g["Cf"][F:] @ a
    corr = dm.non_orth_correction(phi, phib, gam)
    gradg, _ = dm.grad_gauss(bcG, phi, phib)
    interior = np.ones(m.N, dtype=bool); interior[pm.owner[F:]] = False
    assert np.abs(gradg[interior] - a).max() < 1e-12                          # exact gradient away from the boundary
    flux_unc = np.zeros(F + B)
    flux_unc[:F] = g["deltaCoeffs"][:F] * (gam[:F] * g["magSf"][:F]) * (phi[pm.neighbour] - phi[pm.owner[:F]])
    exact = np.zeros(F + B); exact[:F] = gam[:F] * (g["Sf"][:F] @ a)
    deepf = interior[pm.owner[:F]] & interior[pm.neighbour]                   # faces whose two cell gradients are exact
    scale = np.abs(exact).max()
    assert deepf.sum() > 50
    assert np.abs((flux_unc + corr)[:F][deepf] - exact[:F][deepf]).max() < 1e-12 * scale     # corrected: exact to round-off
    assert np.abs(flux_unc[:F][deepf] - exact[:F][deepf]).max() > 1e-2 * scale               # uncorrected: first-order inconsistent
    # and cell by cell: the corrected Laplacian equals the exact sum over the faces
    deep = interior.copy()
    deep[pm.owner[:F][~interior[pm.neighbour]]] = False; deep[pm.neighbour[~interior[pm.owner[:F]]]] = False
    lap_cor = dm.surface_integrate(flux_unc + corr) * g["V"]; lap_ex = dm.surface_integrate(exact) * g["V"]
    lap_unc = dm.surface_integrate(flux_unc) * g["V"]
    assert deep.sum() > 10 and np.abs(lap_cor[deep] - lap_ex[deep]).max() < 1e-12 * scale
    assert np.abs(lap_unc[deep] - lap_ex[deep]).max() > 1e-3 * scale
    dm.release()


@pytest.mark.parametrize("kind,nNonOrth", [("shear", 1), ("jitter", 2), ("jitter", 0)])
def test_region_steps_with_corrected_schemes(lib, oracle, kind, nNonOrth):
    from shallowinterfoam_b200 import capi
    dims = (14, 5, 8)
    pm = lib.polymesh_box(*dims, *L, shear=SHEAR) if kind == "shear" else lib.polymesh_box(*dims, *L, weir=(0.45, 0.6, 0.3), jitter=(0.3, 12345))
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    dm = pm.device_mesh()
    co = default_controls(oracle.OrcControls); cg = default_controls(capi.Region3dControls)
    for c in (co, cg):
        c.nonOrthCorrected = 1; c.nNonOrthCorr = nNonOrth
    ro = oracle.Region(m, co, oracle.BC(m, cs["tA"], 1, cs["rvA"]), oracle.BC(m, cs["tU"], 3, cs["rvU"]), oracle.BC(m, cs["tP"], 1),
                       cs["alpha"], cs["U"], cs["pd"])
    rg = dm.region3d(cg, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    # the same case with `uncorrected` schemes must give DIFFERENT fields (the correction is not dead code) ...
    cu = default_controls(oracle.OrcControls); cu.nNonOrthCorr = nNonOrth
    ru = oracle.Region(m, cu, oracle.BC(m, cs["tA"], 1, cs["rvA"]), oracle.BC(m, cs["tU"], 3, cs["rvU"]), oracle.BC(m, cs["tP"], 1),
                       cs["alpha"], cs["U"], cs["pd"])
    ro.correct_phi(); rg.correct_phi(); ru.correct_phi()
    for s in range(2):
        ro.step(); rg.step(); ru.step()
    fo_, fg, fu = ro.fields(), rg.fields(), ru.fields()
    for k in ("alpha1", "U", "pd", "phi", "p", "rho"):
        assert np.abs(fo_[k] - fg[k]).max() <= 1e-10 * max(np.abs(fo_[k]).max(), 1e-300), k
    assert np.abs(fo_["pd"] - fu["pd"]).max() > 1e-6 * np.abs(fu["pd"]).max()
    po, pg = ro.perfs(), rg.perfs()
    assert len(pg) == (1 + 2 * 3) * (nNonOrth + 1)                            # correctPhi + 2 steps x 3 correctors, each (nNonOrth+1) solves
    assert [p[2] for p in po] == [p[2] for p in pg]
    ho, hg = ro.res_hist(), rg.res_hist()
    assert ho.shape == hg.shape and np.allclose(ho, hg, rtol=1e-8, atol=1e-13)
    rg.release(); dm.release()


def test_corrected_is_a_no_op_on_an_orthogonal_mesh(lib, oracle):
    from shallowinterfoam_b200 import capi
    dims = (12, 4, 6)
    pm = lib.polymesh_box(*dims, *L)
    cs = channel_case(pm, *L)
    dm = pm.device_mesh()
    assert dm.non_orthogonality()[1] is True
    out = []
    for flag in (0, 1):
        cg = default_controls(capi.Region3dControls); cg.nonOrthCorrected = flag
        rg = dm.region3d(cg, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
        rg.correct_phi(); rg.step(2)
        out.append(rg.fields()); rg.release()
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k]), k
    dm.release()
