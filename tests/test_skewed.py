"""BASELINE configs[3] (skewed 3D box): a hex block whose interior vertices are jittered by up to 0.3 of a cell size (LCG seed 12345,
SURVEY.md section 8(d)).  Faces become non-planar and non-orthogonal: interpolation weights leave 0.5, Sf is no longer parallel to d.
Points and geometry are bit-exact against the numpy oracle; the LDU kernels and the region step keep their parity bars on it."""
import numpy as np
import pytest
from oracle import mesh_oracle as mo
from cases import oracle_mesh_of, channel_case, default_controls, random_spd

L = (5.0, 1.0, 2.0)


@pytest.mark.parametrize("dims,weir", [((10, 6, 8), None), ((12, 5, 7), (0.45, 0.6, 0.3))])
def test_skewed_mesh_points_and_geometry_bit_exact(dims, weir):
    from shallowinterfoam_b200 import capi
    lib = capi.load()   # host tools need no device
    nx, ny, nz = dims
    pm = lib.polymesh_box(*dims, *L, weir=weir, jitter=(0.3, 12345))
    ref = mo.block_mesh_box(*dims, *L)
    if weir:
        ref = mo.subset_mesh(ref, mo.weir_keep_mask(*dims, *weir))
    plain = ref["points"]
    ref = mo.jitter_points(ref, 0.3, L[0] / nx, L[1] / ny, L[2] / nz, 12345)
    assert np.array_equal(pm.points, ref["points"])
    moved = np.any(pm.points != plain, axis=1)
    assert moved.sum() > 0 and np.abs(pm.points - plain).max() <= 0.3 * max(L[0] / nx, L[1] / ny, L[2] / nz)
    F = pm.nInternalFaces
    on_b = np.zeros(len(plain), dtype=bool); on_b[np.unique(pm.faces[F:])] = True
    assert not np.any(moved & on_b)                                  # the boundary keeps its shape
    g, gr = pm.geometry(), mo.geometry(ref)
    for k in g:
        assert np.array_equal(g[k], gr[k]), k
    assert np.all(g["V"] > 0) and abs(g["V"].sum() - mo.geometry(dict(ref, points=plain))["V"].sum()) < 1e-12 * g["V"].sum()
    assert np.abs(g["weights"][:F] - 0.5).max() > 0.02               # genuinely skewed
    # closed cells: sum of outward face-area vectors vanishes
    s = np.zeros((pm.nCells, 3)); np.add.at(s, pm.owner, g["Sf"]); np.add.at(s, pm.neighbour, -g["Sf"][:F])
    assert np.abs(s).max() < 1e-14


def test_skewed_ldu_and_region_parity(lib, oracle):
    from shallowinterfoam_b200 import capi
    dims = (14, 5, 8)
    pm = lib.polymesh_box(*dims, *L, weir=(0.45, 0.6, 0.3), jitter=(0.3, 12345))
    m = oracle_mesh_of(pm)
    cs = channel_case(pm, *L)
    dm = pm.device_mesh()
    # fv operators on skewed geometry: bitwise
    rng = np.random.default_rng(3)
    vf = rng.uniform(0, 1, m.N)
    bcO = oracle.BC(m, cs["tA"], 1, cs["rvA"]); bcG = dm.bc(cs["tA"], 1, cs["rvA"])
    vfb = oracle.bc_evaluate(m, bcO, vf)
    assert np.array_equal(oracle.interpolate(m, vf, vfb), dm.interpolate(vf, vfb))
    assert np.array_equal(oracle.sngrad(m, bcO, vf, vfb), dm.sngrad(bcG, vf, vfb))
    # region steps: same bars as tests/test_region.py
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
    po, pg = ro.perfs(), rg.perfs()
    assert [p[2] for p in po] == [p[2] for p in pg]
    ho, hg = ro.res_hist(), rg.res_hist()
    assert ho.shape == hg.shape and np.allclose(ho, hg, rtol=1e-8, atol=1e-13)
    rg.release(); dm.release()
