"""CPU ORACLE (test infrastructure only) -- mesh, LDU addressing, geometry,
`simple` decomposition and level renumbering, restated in numpy.

PARITY UNPINNED: foam-extend-3.1 is not vendored in /root/reference (it is linked
through Make/options:22-31) and is not installed here, so every function below is a
restatement of the published foam-extend-3.1 algorithm named in its docstring, not a
run of the real binary.  Nothing in the product path may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

All integer outputs here are compared BIT-EXACTLY with the product's independent C++
implementation (shallowinterfoam_b200/csrc/host_mesh.cpp).
"""
import numpy as np

VSMALL = 1.0e-300


# ----------------------------------------------------------------------------
# blockMesh single hex block  [UPSTREAM foam-extend-3.1 src/mesh/blockMesh
#   (blockCreate.C cell/point order; createPatches / polyMeshFromShapeMesh.C face order)]
# SURVEY.md Appendix A.1.
# ----------------------------------------------------------------------------
def _vtx(i, j, k, nx, ny):
    return i + (nx + 1) * (j + (ny + 1) * k)


def block_mesh_box(nx, ny, nz, lx, ly, lz, patch_spec=None):
    """Hex box in blockMesh ordering.

    cells  : c = i + nx*(j + ny*k)
    points : v = i + (nx+1)*(j + (ny+1)*k), coordinate = L*(i/n)
    internal faces: upper-triangular order -- for each cell ascending, its faces to the
        higher-numbered neighbours c+1, c+nx, c+nx*ny in that (ascending) order.
    boundary faces: patch by patch in `patch_spec` order; a patch is a list of block
        sides; side loops are  x: for k, for j   y: for i, for k   z: for i, for j.
    Face vertex order is chosen so the area vector points out of the owner.
    """
    if patch_spec is None:
        patch_spec = [("inlet", ["xmin"]), ("outlet", ["xmax"]),
                      ("walls", ["ymin", "ymax", "zmin"]), ("atmosphere", ["zmax"])]
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    # point order: i fastest
    pts = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
    vid = _vtx(I, J, K, nx, ny).ravel()
    pts[vid, 0] = lx * (I.ravel() / nx)
    pts[vid, 1] = ly * (J.ravel() / ny)
    pts[vid, 2] = lz * (K.ravel() / nz)

    N = nx * ny * nz
    # ---- internal faces, cell-by-cell, neighbour ascending
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cid = (ci + nx * (cj + ny * ck))
    has = np.stack([ci < nx - 1, cj < ny - 1, ck < nz - 1], axis=-1)  # x+,y+,z+
    # order faces by (cell, dir)
    cell_flat = cid.ravel()
    order = np.argsort(cell_flat, kind="stable")
    ci_s, cj_s, ck_s = ci.ravel()[order], cj.ravel()[order], ck.ravel()[order]
    has_s = has.reshape(-1, 3)[order]
    owner_l, nei_l, faces_l = [], [], []
    # build a (N,3) table then select in row-major order => (cell asc, dir asc)
    own_tab = np.repeat(np.arange(N)[:, None], 3, axis=1)
    nei_tab = np.stack([np.arange(N) + 1, np.arange(N) + nx, np.arange(N) + nx * ny], axis=1)
    fv = np.empty((N, 3, 4), dtype=np.int64)
    i, j, k = ci_s, cj_s, ck_s
    V = lambda a, b, c: _vtx(a, b, c, nx, ny)
    # x+ face at i+1: (j,k),(j+1,k),(j+1,k+1),(j,k+1)
    fv[:, 0, 0] = V(i + 1, j, k); fv[:, 0, 1] = V(i + 1, j + 1, k)
    fv[:, 0, 2] = V(i + 1, j + 1, k + 1); fv[:, 0, 3] = V(i + 1, j, k + 1)
    # y+ face at j+1: (i,k),(i,k+1),(i+1,k+1),(i+1,k)
    fv[:, 1, 0] = V(i, j + 1, k); fv[:, 1, 1] = V(i, j + 1, k + 1)
    fv[:, 1, 2] = V(i + 1, j + 1, k + 1); fv[:, 1, 3] = V(i + 1, j + 1, k)
    # z+ face at k+1: (i,j),(i+1,j),(i+1,j+1),(i,j+1)
    fv[:, 2, 0] = V(i, j, k + 1); fv[:, 2, 1] = V(i + 1, j, k + 1)
    fv[:, 2, 2] = V(i + 1, j + 1, k + 1); fv[:, 2, 3] = V(i, j + 1, k + 1)
    sel = has_s.ravel()
    owner_int = own_tab.ravel()[sel]
    nei_int = nei_tab.ravel()[sel]
    faces_int = fv.reshape(-1, 4)[sel]
    F = owner_int.size

    # ---- boundary faces
    def side(name):
        if name in ("xmin", "xmax"):
            kk, jj = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")  # k outer, j inner
            kk, jj = kk.ravel(), jj.ravel()
            if name == "xmin":
                c = 0 + nx * (jj + ny * kk)
                f = np.stack([V(0, jj, kk), V(0, jj, kk + 1), V(0, jj + 1, kk + 1), V(0, jj + 1, kk)], 1)
            else:
                c = (nx - 1) + nx * (jj + ny * kk)
                f = np.stack([V(nx, jj, kk), V(nx, jj + 1, kk), V(nx, jj + 1, kk + 1), V(nx, jj, kk + 1)], 1)
        elif name in ("ymin", "ymax"):
            ii, kk = np.meshgrid(np.arange(nx), np.arange(nz), indexing="ij")  # i outer, k inner
            ii, kk = ii.ravel(), kk.ravel()
            if name == "ymin":
                c = ii + nx * (0 + ny * kk)
                f = np.stack([V(ii, 0, kk), V(ii + 1, 0, kk), V(ii + 1, 0, kk + 1), V(ii, 0, kk + 1)], 1)
            else:
                c = ii + nx * ((ny - 1) + ny * kk)
                f = np.stack([V(ii, ny, kk), V(ii, ny, kk + 1), V(ii + 1, ny, kk + 1), V(ii + 1, ny, kk)], 1)
        else:
            ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")  # i outer, j inner
            ii, jj = ii.ravel(), jj.ravel()
            if name == "zmin":
                c = ii + nx * (jj + ny * 0)
                f = np.stack([V(ii, jj, 0), V(ii, jj + 1, 0), V(ii + 1, jj + 1, 0), V(ii + 1, jj, 0)], 1)
            else:
                c = ii + nx * (jj + ny * (nz - 1))
                f = np.stack([V(ii, jj, nz), V(ii + 1, jj, nz), V(ii + 1, jj + 1, nz), V(ii, jj + 1, nz)], 1)
        return c, f

    own_b, faces_b, patches = [], [], []
    start = F
    for pname, sides in patch_spec:
        n0 = start
        for s in sides:
            c, f = side(s)
            own_b.append(c); faces_b.append(f); start += c.size
        patches.append((pname, n0, start - n0))
    owner = np.concatenate([owner_int] + own_b).astype(np.int32)
    faces = np.concatenate([faces_int] + faces_b).astype(np.int32)
    return dict(points=pts, faces=faces, owner=owner, neighbour=nei_int.astype(np.int32),
                patches=patches, nCells=N)


def subset_mesh(mesh, keep, exposed_name="weir"):
    """Remove cells (subsetMesh semantics [UPSTREAM src/dynamicMesh/fvMeshSubset]):
    kept cells keep their relative order; kept internal faces keep relative order (stays
    upper-triangular); each original patch keeps its surviving faces in order; internal
    faces that lose exactly one cell become boundary faces of a new last patch
    `exposed_name`, ordered by original face index, flipped when the surviving cell was the
    neighbour.  Points are kept as they are (unused points are harmless)."""
    own, nei = mesh["owner"], mesh["neighbour"]
    F = nei.size
    keep = np.asarray(keep, dtype=bool)
    newid = np.full(mesh["nCells"], -1, dtype=np.int64)
    newid[keep] = np.arange(keep.sum())
    ko, kn = keep[own[:F]], keep[nei]
    both = ko & kn
    one = ko ^ kn
    faces = mesh["faces"]
    owner_i = newid[own[:F][both]]
    nei_i = newid[nei[both]]
    faces_i = faces[:F][both]
    own_b, faces_b, patches = [], [], []
    start = owner_i.size
    for pname, s, n in mesh["patches"]:
        m = keep[own[s:s + n]]
        own_b.append(newid[own[s:s + n][m]]); faces_b.append(faces[s:s + n][m])
        patches.append((pname, start, int(m.sum()))); start += int(m.sum())
    idx = np.nonzero(one)[0]
    surv_is_owner = ko[idx]
    c = np.where(surv_is_owner, own[:F][idx], nei[idx])
    fexp = faces[:F][idx].copy()
    flip = ~surv_is_owner
    fexp[flip] = fexp[flip][:, ::-1]
    own_b.append(newid[c]); faces_b.append(fexp)
    patches.append((exposed_name, start, idx.size))
    return dict(points=mesh["points"], faces=np.concatenate([faces_i] + faces_b).astype(np.int32),
                owner=np.concatenate([owner_i] + own_b).astype(np.int32),
                neighbour=nei_i.astype(np.int32), patches=patches, nCells=int(keep.sum()))


def agglomerate(mesh, group):
    """Polyhedral mesh by agglomeration (BASELINE configs[3]: irregular LDU addressing): cells with the same `group` key merge into
    one cell.  New cell labels follow the first appearance of a group among the old cells; internal faces between two merged cells
    vanish; the others keep their points, are flipped when the new owner would exceed the new neighbour, and are sorted by
    (owner, neighbour, original face) = upper-triangular order (two cells may share several faces: lduAddressing does not mind);
    boundary faces and patches keep their order.  [UPSTREAM polyMesh ordering rules; restated by b200_polymesh_agglomerate]"""
    group = np.asarray(group)
    uniq, first = np.unique(group, return_index=True)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    newid = rank[np.searchsorted(uniq, group)]
    own, nei = mesh["owner"], mesh["neighbour"]
    F = nei.size
    o2, n2 = newid[own[:F]], newid[nei]
    idx = np.nonzero(o2 != n2)[0]
    o3, n3 = o2[idx], n2[idx]
    flip = o3 > n3
    lo, hi = np.where(flip, n3, o3), np.where(flip, o3, n3)
    fv = mesh["faces"][:F][idx].copy()
    fv[flip] = fv[flip][:, ::-1]
    perm = np.lexsort((idx, hi, lo))
    shift = F - idx.size
    patches = [(n, s_ - shift, z) for (n, s_, z) in mesh["patches"]]
    return dict(points=mesh["points"], faces=np.concatenate([fv[perm], mesh["faces"][F:]]).astype(np.int32),
                owner=np.concatenate([lo[perm], newid[own[F:]]]).astype(np.int32), neighbour=hi[perm].astype(np.int32),
                patches=patches, nCells=len(uniq))


def block_groups(nx, ny, nz):
    """group keys that merge the 2x2x2 blocks with (I+J+K) % 3 != 0 of an nx x ny x nz hex box (only complete blocks): new cells are
    single hexes (6 faces) or 2x2x2 super-cells (24 faces); neighbouring super-cells share 4 faces"""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = i // 2, j // 2, k // 2
    full = (2 * I + 1 < nx) & (2 * J + 1 < ny) & (2 * K + 1 < nz)
    merged = full & ((I + J + K) % 3 != 0)
    cell = (i + nx * (j + ny * k))
    key = np.where(merged, nx * ny * nz + (I + (nx // 2 + 1) * (J + (ny // 2 + 1) * K)), cell)
    out = np.empty(nx * ny * nz, dtype=np.int64)
    out[cell.ravel()] = key.ravel()
    return out


def jitter_points(mesh, frac, dx, dy, dz, seed=12345):
    """skewed mesh (BASELINE configs[3]): interior vertices (not on any boundary face) move by (2u-1)*frac*(dx,dy,dz),
    u from the 64-bit LCG documented in include/b200foam.h (b200_polymesh_jitter); pure-Python integers, so no overflow semantics."""
    pts = mesh["points"].copy()
    nI = mesh["neighbour"].size
    on_b = np.zeros(len(pts), dtype=bool)
    on_b[np.unique(mesh["faces"][nI:])] = True
    x = int(seed); M = (1 << 64) - 1
    d = (dx, dy, dz)
    for p in np.nonzero(~on_b)[0]:
        for k in range(3):
            x = (x * 6364136223846793005 + 1442695040888963407) & M
            u = float(x >> 11) * (1.0 / 9007199254740992.0)
            pts[p, k] = pts[p, k] + (2.0 * u - 1.0) * frac * d[k]
    out = dict(mesh); out["points"] = pts
    return out


def shear_points(mesh, sxy, sxz, syz):
    """affine shear (include/b200foam.h b200_polymesh_shear): x += sxy*y + sxz*z ; y += syz*z -- non-orthogonal, not skewed"""
    pts = mesh["points"].copy()
    y, z = pts[:, 1].copy(), pts[:, 2].copy()
    pts[:, 0] = pts[:, 0] + (sxy * y + sxz * z)
    pts[:, 1] = y + syz * z
    out = dict(mesh); out["points"] = pts
    return out


def weir_keep_mask(nx, ny, nz, x0=0.45, x1=0.5, zfrac=0.3):
    """cells removed: cell-centre x in [x0,x1)*L and z < zfrac*H  (SURVEY.md section 8(d) cfg2)."""
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cid = (ci + nx * (cj + ny * ck)).ravel()
    xc = (ci.ravel() + 0.5) / nx
    zc = (ck.ravel() + 0.5) / nz
    rem = (xc >= x0) & (xc < x1) & (zc < zfrac)
    keep = np.ones(nx * ny * nz, dtype=bool)
    keep[cid[rem]] = False
    return keep


# ----------------------------------------------------------------------------
# lduAddressing  [UPSTREAM src/foam/matrices/lduMatrix/lduAddressing/lduAddressing.C
#   calcLosort / calcOwnerStart / calcLosortStart]
# ----------------------------------------------------------------------------
def ldu_addressing(lower, upper, N):
    F = lower.size
    owner_start = np.zeros(N + 1, dtype=np.int32)
    np.add.at(owner_start, lower.astype(np.int64) + 1, 1)
    owner_start = np.cumsum(owner_start).astype(np.int32)
    losort = np.argsort(upper, kind="stable").astype(np.int32)   # faces sorted by upper, ties by face index
    losort_start = np.zeros(N + 1, dtype=np.int32)
    np.add.at(losort_start, upper.astype(np.int64) + 1, 1)
    losort_start = np.cumsum(losort_start).astype(np.int32)
    return owner_start, losort, losort_start


def check_upper_triangular(lower, upper):
    ok = np.all(lower < upper)
    key = lower.astype(np.int64) * (upper.max() + 1 if upper.size else 1) + upper
    return bool(ok and np.all(np.diff(key) > 0))


# ----------------------------------------------------------------------------
# Geometry [UPSTREAM src/foam/meshes/primitiveMesh/primitiveMeshFaceCentresAndAreas.C,
#   primitiveMeshCellCentresAndVols.C; src/finiteVolume/interpolation/surfaceInterpolation/
#   surfaceInterpolation/surfaceInterpolation.C makeWeights/makeDeltaCoeffs; fvPatch::delta = Cf-Cn]
# Quad faces only (all generated meshes are hex / hex-derived).
# ----------------------------------------------------------------------------
def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                     a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def _dot(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]


def geometry(mesh):
    pts, faces, own, nei = mesh["points"], mesh["faces"], mesh["owner"], mesh["neighbour"]
    N, F, nF = mesh["nCells"], nei.size, faces.shape[0]
    p = [pts[faces[:, q]] for q in range(4)]
    fC = p[0].copy()
    for q in range(1, 4):
        fC = fC + p[q]
    fC = fC / 4.0
    sumN = np.zeros((nF, 3)); sumA = np.zeros(nF); sumAc = np.zeros((nF, 3))
    for q in range(4):
        a, b = p[q], p[(q + 1) % 4]
        c = a + b + fC
        n = _cross(b - a, fC - a)
        an = np.sqrt(_dot(n, n))
        sumN = sumN + n
        sumA = sumA + an
        sumAc = sumAc + an[:, None] * c
    Cf = (1.0 / 3.0) * sumAc / (sumA + VSMALL)[:, None]
    Sf = 0.5 * sumN
    magSf = np.sqrt(_dot(Sf, Sf))
    # cell centres & volumes
    cEst = np.zeros((N, 3)); nCF = np.zeros(N)
    np.add.at(cEst, own, Cf); np.add.at(nCF, own, 1.0)
    np.add.at(cEst, nei, Cf[:F]); np.add.at(nCF, nei, 1.0)
    cEst = cEst / nCF[:, None]
    C = np.zeros((N, 3)); Vol = np.zeros(N)
    pv = np.maximum(_dot(Sf, Cf - cEst[own]), VSMALL)
    pc = (3.0 / 4.0) * Cf + (1.0 / 4.0) * cEst[own]
    np.add.at(C, own, pv[:, None] * pc); np.add.at(Vol, own, pv)
    pv = np.maximum(_dot(Sf[:F], cEst[nei] - Cf[:F]), VSMALL)
    pc = (3.0 / 4.0) * Cf[:F] + (1.0 / 4.0) * cEst[nei]
    np.add.at(C, nei, pv[:, None] * pc); np.add.at(Vol, nei, pv)
    C = C / Vol[:, None]
    Vol = Vol * (1.0 / 3.0)
    # weights / deltaCoeffs
    w = np.ones(nF); dc = np.empty(nF)
    SfdOwn = np.abs(_dot(Sf[:F], Cf[:F] - C[own[:F]]))
    SfdNei = np.abs(_dot(Sf[:F], C[nei] - Cf[:F]))
    w[:F] = SfdNei / (SfdOwn + SfdNei)
    d = C[nei] - C[own[:F]]
    dc[:F] = 1.0 / np.sqrt(_dot(d, d))
    d = Cf[F:] - C[own[F:]]
    dc[F:] = 1.0 / np.sqrt(_dot(d, d))
    return dict(Sf=Sf, magSf=magSf, Cf=Cf, C=C, V=Vol, weights=w, deltaCoeffs=dc)


# ----------------------------------------------------------------------------
# simpleGeomDecomp [UPSTREAM src/decompositionMethods/decompositionMethods/simpleGeomDecomp,
#   geomDecomp rotDelta]  SURVEY.md Appendix A.8
# ----------------------------------------------------------------------------
def _assign_to_processor_group(n_items, n_groups):
    jump = n_items // n_groups
    jump_b = n_items - jump * n_groups
    grp = np.empty(n_items, dtype=np.int32)
    ind = 0
    j = 0
    for j in range(jump_b):
        grp[ind:ind + jump + 1] = j
        ind += jump + 1
    for j in range(jump_b, n_groups):
        grp[ind:ind + jump] = j
        ind += jump
    return grp


def simple_decomp(C, n, delta=0.001):
    nx, ny, nz = n
    d = 1 - 0.5 * delta * delta
    d2 = d * d
    a = delta
    a2 = a * a
    R = np.array([[d2, -a * d, a],
                  [a * d - a2 * d, a * a2 + d2, -2 * a * d],
                  [a * d2 + a2, a * d - a2 * d, d2 - a2]])
    rp = np.stack([R[r, 0] * C[:, 0] + R[r, 1] * C[:, 1] + R[r, 2] * C[:, 2] for r in range(3)], axis=1)
    N = C.shape[0]
    proc = np.zeros(N, dtype=np.int32)
    mult = [1, nx, nx * ny]
    for dirn, ng in enumerate((nx, ny, nz)):
        order = np.argsort(rp[:, dirn], kind="stable")
        g = _assign_to_processor_group(N, ng)
        grp = np.empty(N, dtype=np.int32)
        grp[order] = g
        proc += grp * mult[dirn]
    return proc


# ----------------------------------------------------------------------------
# DIC level schedule + device renumbering (this repo's own design; SURVEY.md A.4 "Level schedule")
#   fwd level(c) = 0 if c has no lower neighbour else 1 + max(level(l)) over faces with u == c
#   device row order = cells sorted by (level, original index)
# ----------------------------------------------------------------------------
def forward_levels(lower, upper, N):
    lev = np.zeros(N, dtype=np.int32)
    # faces are sorted by lower, so a single ascending pass is a valid topological sweep
    # (vectorised per owner-run is awkward; do the plain loop in chunks via python ints)
    lo = lower.tolist(); up = upper.tolist(); L = lev.tolist()
    for f in range(len(lo)):
        v = L[lo[f]] + 1
        if v > L[up[f]]:
            L[up[f]] = v
    return np.asarray(L, dtype=np.int32)


def level_renumbering(lower, upper, N):
    lev = forward_levels(lower, upper, N)
    perm = np.argsort(lev, kind="stable").astype(np.int32)   # device row -> original cell
    inv = np.empty(N, dtype=np.int32)
    inv[perm] = np.arange(N, dtype=np.int32)
    return lev, perm, inv


def sweep_smem_bytes(n_rows, n_slots, n_halo, d):
    return 8 * ((n_rows + n_halo + 4) // 2 * 2) + 8 * n_slots * (d + (d + 3) // 4) + 64


def sweep_plan(lower, upper, N, band_levels=16, max_rows=4096, smem_budget=225000, lanes=256):
    """DIC sweep plan of the device rows (this repo's own design; restates csrc/host_mesh.cpp buildSweepPlan).
    Three potentials that never decrease along a dependency l -> u: level, potJ / potK = most faces of stride class 1 / 2 on any
    path (stride classes = the three dominant magnitudes of upper - lower).  Tile class = (level // band, potK // binK, potJ // binJ);
    oversized classes are cut into consecutive pieces of their (level, index) order; tiles are ordered by their longest-path level
    in the tile graph; device row order = (tile, level, original index).  Where the stride classes shred the mesh (more than three
    times the ideal number of tiles) the lateral potential becomes the cell index itself (potK = index // indexBins, potJ = 0).
    CHUNKED plans (a row with more than 3 lower or upper neighbours: the rows are applied as chains of <= 3-source chunks whose list
    schedule decides the budget cuts) are not restated: the result then only carries level / chunked, and the tests check the library's
    plan through its properties (permutation, topological tile order) and the arithmetic."""
    lower = np.asarray(lower, dtype=np.int64); upper = np.asarray(upper, dtype=np.int64)
    F = lower.size
    band_levels = max(1, band_levels); max_rows = max(8, min(max_rows, 32768)); lanes = max(32, min(lanes, 1024)) // 32 * 32
    lev = forward_levels(lower.astype(np.int32), upper.astype(np.int32), N).astype(np.int64)
    out = dict(level=lev.astype(np.int32), bandLevels=band_levels, maxRows=max_rows, chunked=False, indexBins=0)
    if F and (np.bincount(lower, minlength=N).max() > 3 or np.bincount(upper, minlength=N).max() > 3):
        out["chunked"] = True
        return out
    if N == 0:
        out.update(rowCell=np.zeros(0, np.int32), cellRow=np.zeros(0, np.int32), tileStart=np.zeros(1, np.int32),
                   tileLevel=np.zeros(0, np.int32), potJ=np.zeros(0, np.int32), potK=np.zeros(0, np.int32), binJ=1, binK=1)
        return out
    n_levels = int(lev.max()) + 1
    # 1. stride classes
    d = upper - lower
    vals, cnts = np.unique(d, return_counts=True)
    removed = np.zeros(vals.size, dtype=bool)
    cen = []
    for _ in range(3):
        if removed.all():
            break
        c_ = np.where(removed, -1, cnts)
        best = int(np.argmax(c_))                      # ties: smallest value first
        c = int(vals[best]); cen.append(c)
        removed |= (2 * vals > c) & (vals < 2 * c)
    cen.sort()
    cls = np.zeros(F, dtype=np.int64)
    if len(cen) > 1:
        cls[d * d >= cen[0] * cen[1]] = 1
    if len(cen) > 2:
        cls[d * d >= cen[1] * cen[2]] = 2
    # 2. potentials (sequential ascending face pass)
    pj = [0] * N; pk = [0] * N
    lo = lower.tolist(); up = upper.tolist(); cl = cls.tolist()
    for f in range(F):
        l, u, c = lo[f], up[f], cl[f]
        a = pj[l] + (1 if c == 1 else 0); b = pk[l] + (1 if c == 2 else 0)
        if a > pj[u]: pj[u] = a
        if b > pk[u]: pk[u] = b
    potJ = np.asarray(pj, dtype=np.int64); potK = np.asarray(pk, dtype=np.int64)
    extJ, extK = int(potJ.max()) + 1, int(potK.max()) + 1
    # 3. bins
    n_triples = np.unique((lev * extJ + potJ) * extK + potK).size
    A = max(1, max_rows * n_triples // (band_levels * N))
    best = None
    for bj in range(1, min(extJ, A) + 1):
        bk = min(extK, max(1, A // bj))
        cost = -(-extJ // bj) + -(-extK // bk); prod = bj * bk
        if best is None or cost < best[0] or (cost == best[0] and prod > best[1]):
            best = (cost, prod, bj, bk)
    binJ, binK = best[2], best[3]

    # 4. classes split into connected components (root = smallest cell), sort, pieces
    def make_pieces(potJ, potK, binJ, binK, extJ, extK):
        nJb, nKb = -(-extJ // binJ), -(-extK // binK)
        cls_key = ((lev // band_levels) * nKb + potK // binK) * nJb + potJ // binJ
        same = cls_key[lower] == cls_key[upper]
        if same.any():
            from scipy.sparse import coo_matrix
            from scipy.sparse.csgraph import connected_components
            g = coo_matrix((np.ones(int(same.sum()), dtype=np.int8), (lower[same], upper[same])), shape=(N, N))
            _, lab = connected_components(g, directed=False)
        else:
            lab = np.arange(N)
        root = np.full(lab.max() + 1, N, dtype=np.int64)
        np.minimum.at(root, lab, np.arange(N))
        root = root[lab]
        order = np.lexsort((np.arange(N), lev, root, cls_key))
        ck = cls_key[order] * (N + 1) + root[order]
        bounds = np.concatenate([[0], np.nonzero(np.diff(ck))[0] + 1, [N]])
        pieces = []
        for b0, e0 in zip(bounds[:-1], bounds[1:]):
            n = int(e0 - b0); npc = -(-n // max_rows); base, rem = divmod(n, npc); at = int(b0)
            for q in range(npc):
                sz = base + (1 if q < rem else 0); pieces.append((at, at + sz)); at += sz
        return order, pieces

    order, pieces = make_pieces(potJ, potK, binJ, binK, extJ, extK)
    # 4b. fall back to the cell index as the lateral potential where the stride classes shred the mesh
    n_bands = -(-n_levels // band_levels)
    ideal = -(-N // max_rows) + n_bands
    index_bins = 0
    if len(pieces) > 3 * ideal:
        rows_per_band = max(1, N // n_bands)
        Q = max(1, -(-rows_per_band // max_rows))
        ib = max(1, -(-N // Q))
        potK2 = np.arange(N, dtype=np.int64) // ib
        order2, pieces2 = make_pieces(np.zeros(N, dtype=np.int64), potK2, 1, 1, 1, -(-N // ib))
        if len(pieces2) < len(pieces):
            order, pieces, index_bins = order2, pieces2, ib
            potJ, potK, binJ, binK = np.zeros(N, dtype=np.int64), potK2, 1, 1
    owner_start, losort, losort_start = ldu_addressing(lower.astype(np.int32), upper.astype(np.int32), N)
    # 5. budget
    while True:
        piece_of = np.empty(N, dtype=np.int64)
        for p, (b, e) in enumerate(pieces):
            piece_of[order[b:e]] = p
        nxt = []; changed = False
        for p, (b, e) in enumerate(pieces):
            cells = order[b:e]; n = e - b
            dF = int((losort_start[cells + 1] - losort_start[cells]).max()); dB = int((owner_start[cells + 1] - owner_start[cells]).max())
            lows = np.concatenate([lower[losort[losort_start[c]:losort_start[c + 1]]] for c in cells]) if dF else np.zeros(0, np.int64)
            ups = np.concatenate([upper[owner_start[c]:owner_start[c + 1]] for c in cells]) if dB else np.zeros(0, np.int64)
            hF = np.unique(lows[piece_of[lows] != p]).size; hB = np.unique(ups[piece_of[ups] != p]).size
            runs = np.diff(np.concatenate([[0], np.nonzero(np.diff(lev[cells]))[0] + 1, [n]]))   # rows per level of the piece
            n_steps = int((-(-runs // lanes)).sum()); S = n_steps * lanes                        # steps of <= lanes rows
            need = max(sweep_smem_bytes(n, S, hF, dF), sweep_smem_bytes(n, S, hB, dB))
            if (need > smem_budget or n + max(hF, hB) + 2 > 8191 or n_steps > 256) and n > 1:
                mid = b + n // 2; nxt += [(b, mid), (mid, e)]; changed = True
            else:
                nxt.append((b, e))
        pieces = nxt
        if not changed:
            break
    nT = len(pieces)
    # 6. tile graph levels (longest path) -- Kahn in piece order
    a, b = piece_of[lower], piece_of[upper]
    m = a != b
    edges = np.unique(np.stack([a[m], b[m]], axis=1), axis=0) if m.any() else np.zeros((0, 2), np.int64)
    tl = [0] * nT; indeg = [0] * nT; adj = [[] for _ in range(nT)]
    for x, y in edges.tolist():
        adj[x].append(y); indeg[y] += 1
    queue = [t for t in range(nT) if indeg[t] == 0]
    h = 0
    while h < len(queue):
        t = queue[h]; h += 1
        for v in adj[t]:
            tl[v] = max(tl[v], tl[t] + 1); indeg[v] -= 1
            if indeg[v] == 0:
                queue.append(v)
    assert len(queue) == nT, "tile graph is cyclic"
    ticket = np.argsort(np.asarray(tl), kind="stable")
    row_cell = np.concatenate([order[pieces[t][0]:pieces[t][1]] for t in ticket]).astype(np.int32)
    cell_row = np.empty(N, dtype=np.int32); cell_row[row_cell] = np.arange(N, dtype=np.int32)
    tile_start = np.concatenate([[0], np.cumsum([pieces[t][1] - pieces[t][0] for t in ticket])]).astype(np.int32)
    out.update(rowCell=row_cell, cellRow=cell_row, tileStart=tile_start, tileLevel=np.asarray(tl, np.int32)[ticket],
               potJ=potJ.astype(np.int32), potK=potK.astype(np.int32), binJ=binJ, binK=binK, indexBins=index_bins)
    return out


# ----------------------------------------------------------------------------
# decomposePar-style sub-mesh construction for `simple` decomposition
# [UPSTREAM applications/utilities/parallelProcessing/decomposePar/domainDecomposition*.C]
#   local cells: ascending global id; local internal faces: global internal faces with both
#   cells on this rank, ascending global face id; physical patches keep order; processor
#   patches appended after physical patches, ascending neighbour rank, faces in ascending
#   global face id (so both sides enumerate identically); the side holding the global
#   neighbour cell sees the face flipped.
# ----------------------------------------------------------------------------
def decompose(mesh, proc, n_ranks):
    own, nei = mesh["owner"], mesh["neighbour"]
    F = nei.size
    subs = []
    po, pn = proc[own[:F]], proc[nei]
    for r in range(n_ranks):
        cells = np.nonzero(proc == r)[0].astype(np.int32)
        g2l = np.full(mesh["nCells"], -1, dtype=np.int64)
        g2l[cells] = np.arange(cells.size)
        fint = np.nonzero((po == r) & (pn == r))[0]
        face_ids = [fint]
        owner_l = [g2l[own[fint]]]
        flip = [np.zeros(fint.size, dtype=bool)]
        patches = []
        start = fint.size
        for pname, s, n in mesh["patches"]:
            ids = s + np.nonzero(proc[own[s:s + n]] == r)[0]
            face_ids.append(ids); owner_l.append(g2l[own[ids]]); flip.append(np.zeros(ids.size, dtype=bool))
            patches.append((pname, start, ids.size, -1)); start += ids.size
        for q in range(n_ranks):
            if q == r:
                continue
            m = ((po == r) & (pn == q)) | ((po == q) & (pn == r))
            ids = np.nonzero(m)[0]
            if ids.size == 0:
                continue
            mine_is_owner = po[ids] == r
            lc = np.where(mine_is_owner, g2l[own[ids]], g2l[nei[ids]])
            face_ids.append(ids); owner_l.append(lc); flip.append(~mine_is_owner)
            patches.append(("procBoundary%dto%d" % (r, q), start, ids.size, q)); start += ids.size
        subs.append(dict(cells=cells,
                         faceIds=np.concatenate(face_ids).astype(np.int32),
                         flip=np.concatenate(flip),
                         owner=np.concatenate(owner_l).astype(np.int32),
                         neighbour=g2l[nei[fint]].astype(np.int32),
                         patches=patches, nCells=int(cells.size)))
    return subs
