"""Host mesh tools of the product (C++) vs the numpy oracle: integers bit-exact (north_star: "mesh addressing,
renumbering and decomposition must be bit-exact"); geometry bit-exact too (same operation order)."""
import numpy as np
import pytest
from oracle import mesh_oracle as mo
from shallowinterfoam_b200 import capi


@pytest.fixture(scope="module")
def hostlib():
    return capi.load()   # host tools need no device


def test_library_exports_every_declared_symbol(hostlib):
    assert hostlib.missing_exports() == []
    # every function declared in include/b200foam.h is in EXPORTS
    import os, re
    hdr = open(os.path.join(os.path.dirname(capi.__file__), "..", "include", "b200foam.h")).read()
    declared = set(re.findall(r"\b(b200_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTS)


@pytest.mark.parametrize("dims,weir", [((50, 10, 20), None), ((12, 5, 7), (0.45, 0.6, 0.3)), ((40, 8, 16), (0.45, 0.5, 0.3)), ((1, 1, 1), None), ((3, 1, 2), None)])
def test_box_and_weir_mesh_bit_exact(hostlib, dims, weir):
    nx, ny, nz = dims
    pm = hostlib.polymesh_box(nx, ny, nz, 5.0, 1.0, 2.0, weir=weir)
    ref = mo.block_mesh_box(nx, ny, nz, 5.0, 1.0, 2.0)
    if weir:
        ref = mo.subset_mesh(ref, mo.weir_keep_mask(nx, ny, nz, *weir))
    assert pm.nCells == ref["nCells"]
    assert np.array_equal(pm.points, ref["points"])
    assert np.array_equal(pm.faces, ref["faces"])
    assert np.array_equal(pm.owner, ref["owner"])
    assert np.array_equal(pm.neighbour, ref["neighbour"])
    assert pm.patches == ref["patches"]
    F = pm.nInternalFaces
    if F:
        assert mo.check_upper_triangular(pm.owner[:F], pm.neighbour)
    g, gr = pm.geometry(), mo.geometry(ref)
    for k in g:
        assert np.array_equal(g[k], gr[k].reshape(g[k].shape)), k
    # sanity: closed cells, volume sum
    assert abs(g["V"].sum() - 10.0 * pm.nCells / (nx * ny * nz)) < 1e-12
    closed = np.zeros((pm.nCells, 3))
    np.add.at(closed, pm.owner, g["Sf"]); np.add.at(closed, pm.neighbour, -g["Sf"][:F])
    assert np.abs(closed).max() < 1e-14


def test_config_sizes_match_survey(hostlib):
    pm = hostlib.polymesh_box(50, 10, 20, 5.0, 1.0, 2.0)
    assert (pm.nCells, pm.nInternalFaces, pm.nBoundaryFaces) == (10000, 28300, 3400)


@pytest.mark.parametrize("n", [(2, 1, 1), (2, 2, 1), (4, 1, 2), (3, 1, 1)])
def test_simple_decomposition_and_submeshes_bit_exact(hostlib, n):
    pm = hostlib.polymesh_box(20, 6, 8, 5.0, 1.0, 2.0, weir=(0.45, 0.6, 0.3))
    ref = mo.subset_mesh(mo.block_mesh_box(20, 6, 8, 5.0, 1.0, 2.0), mo.weir_keep_mask(20, 6, 8, 0.45, 0.6, 0.3))
    proc = pm.decompose_simple(n)
    proc_ref = mo.simple_decomp(mo.geometry(ref)["C"], n)
    assert np.array_equal(proc, proc_ref)
    nr = n[0] * n[1] * n[2]
    counts = np.bincount(proc, minlength=nr)
    assert counts.min() > 0
    subs = mo.decompose(ref, proc_ref, nr)
    tot_cells = 0
    for r in range(nr):
        sub = pm.submesh(proc, nr, r)
        s = subs[r]
        assert np.array_equal(sub.cellMap, s["cells"])
        assert np.array_equal(sub.faceMap, s["faceIds"])
        assert np.array_equal(sub.faceFlip.astype(bool), s["flip"])
        assert np.array_equal(sub.owner, s["owner"]) and np.array_equal(sub.neighbour, s["neighbour"])
        assert [(p[0], p[1], p[2]) for p in sub.patches] == [(p[0], p[1], p[2]) for p in s["patches"]]
        assert list(sub.patchNeighbRank) == [p[3] for p in s["patches"]]
        if sub.nInternalFaces:
            assert mo.check_upper_triangular(sub.owner[:sub.nInternalFaces], sub.neighbour)
        tot_cells += sub.nCells
    assert tot_cells == pm.nCells
    # both sides of a processor patch list the same global faces in the same order
    for r in range(nr):
        a = pm.submesh(proc, nr, r)
        for p in range(a.nPatches):
            q = int(a.patchNeighbRank[p])
            if q < 0:
                continue
            b = pm.submesh(proc, nr, q)
            pb = list(b.patchNeighbRank).index(r)
            fa = a.faceMap[a.patchStart[p]:a.patchStart[p] + a.patchSize[p]]
            fb = b.faceMap[b.patchStart[pb]:b.patchStart[pb] + b.patchSize[pb]]
            assert np.array_equal(fa, fb)
