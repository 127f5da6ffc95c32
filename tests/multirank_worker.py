"""Worker of tests/test_multirank.py: one process per rank (gloo for the plumbing).  Loads the SIMT-emulated test build of the
library and routes its all-reduces / halo swaps through torch.distributed -- or, with B200_WORKER_LIB=gpu, the real library with
one GPU per rank (NCCL + NVLink peer memory) -- and checks the decomposed path against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import channel_case, default_controls, oracle_mesh_of, random_spd  # noqa: E402
from oracle import foam_oracle as fo, multirank_oracle as mro  # noqa: E402
from shallowinterfoam_b200 import capi  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    on_gpu = os.environ.get("B200_WORKER_LIB") == "gpu"     # the real library: one GPU per rank, NCCL + NVLink peer memory
    if on_gpu:
        lib = capi.load()
        for o in os.environ.get("B200_WORKER_OPTS", "").split():
            k, v = o.split("="); lib.set_option(k, float(v))
        ids = [lib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        lib.init(rank, rank, world, ids[0])
    else:
        lib = capi.B200Lib(os.path.join(ROOT, "tests", "simt_emu", "libb200foam_emu.so"))
    dp = C.POINTER(C.c_double)

    @C.CFUNCTYPE(None, dp, C.c_int, C.c_int)
    def allreduce(p, count, op):
        a = np.ctypeslib.as_array(p, shape=(count,))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=[dist.ReduceOp.SUM, dist.ReduceOp.MAX, dist.ReduceOp.MIN][op])

    @C.CFUNCTYPE(None, dp, dp, C.c_int, C.c_int)
    def sendrecv(send, recv, count, peer):
        s = torch.from_numpy(np.ctypeslib.as_array(send, shape=(count,)).copy())
        r = torch.from_numpy(np.ctypeslib.as_array(recv, shape=(count,)))
        req = dist.isend(s, peer)
        dist.recv(r, peer)
        req.wait()

    if not on_gpu:
        lib.c.b200_emu_set_comm(allreduce, sendrecv)
        lib.init(0, rank, world)

    res = decomposed_checks(lib, rank, world)
    dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK world=%d %s" % (world, res))


def decomposed_checks(lib, rank, world, dims=(16, 4, 8)):
    """The decomposed hot path of `lib` (already initialised for `world` ranks) against the N-rank / serial oracle on a small
    weir channel: Amul bitwise, two consecutive b200PCG solves (rank-local DIC; iteration counts, residual history 1e-8), two
    region steps vs the SERIAL oracle (fields 1e-7 with tight solver tolerances), and correctPhi + 2 region steps vs the N-RANK oracle
    at the default tolerances (equal iteration counts, histories 1e-8, fields 1e-10).  Collective: every rank calls it.  Returns the measured errors;
    raises AssertionError on a mismatch.  Also used by bench.py's untimed pre-flight on the real GPUs (N > 1)."""
    res = {}
    L = (5.0, 1.0, 2.0)
    gpm = lib.polymesh_box(*dims, *L, weir=(0.45, 0.6, 0.3))
    n = (world, 1, 1) if world != 4 else (2, 1, 2)
    proc = gpm.decompose_simple(n)
    subs = [gpm.submesh(proc, world, r) for r in range(world)]
    pm = subs[rank]
    gm = oracle_mesh_of(gpm)

    # ---------------- a1/a4: decomposed Amul and PCG (block-Jacobi DIC) vs the N-rank oracle
    diag, upper = random_spd(gm, seed=4, shift=(1e-3, 1e-2))
    rng = np.random.default_rng(5)
    psi = rng.uniform(0, 1, gm.N); b = rng.uniform(-1, 1, gm.N)
    osubs = []
    for s in subs:
        om = oracle_mesh_of(s)
        osubs.append(dict(mesh=om, patches=[(s.patchNames[p], int(s.patchStart[p]), int(s.patchSize[p]), int(s.patchNeighbRank[p])) for p in range(s.nPatches)]))
    def restrict(s):
        F = s.nInternalFaces
        d = diag[s.cellMap]; u = upper[s.faceMap[:F]]
        bou = np.zeros(s.nBoundaryFaces)
        for p in range(s.nPatches):
            if s.patchNeighbRank[p] >= 0:
                sl = slice(s.patchStart[p] - F, s.patchStart[p] - F + s.patchSize[p])
                bou[sl] = -upper[s.faceMap[s.patchStart[p]:s.patchStart[p] + s.patchSize[p]]]
        return d, u, bou
    mats = [restrict(s) for s in subs]
    ldu = pm.ldu()
    d, u, bou = mats[rank]
    A_loc = ldu.amul(d, u, psi[pm.cellMap], bou=bou)                      # halo over the (emulated) comm layer
    A_ref = mro.amul_multi(osubs, [m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats], [psi[s.cellMap] for s in subs])[rank]
    assert np.array_equal(A_loc, A_ref), "decomposed Amul differs from the N-rank oracle"
    res["amul_bitwise"] = True
    A_ser = fo.amul(gm, diag, upper, upper, psi)[pm.cellMap]
    assert np.allclose(A_loc, A_ser, rtol=1e-13, atol=1e-13), "decomposed Amul differs from the serial Amul"
    xg, pg, hg = ldu.pcg_solve(d, u, np.zeros(pm.nCells), b[pm.cellMap], tol=1e-9, maxIter=300, bou=bou)
    xo, po, ho = mro.pcg_solve_multi(osubs, [m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats],
                                     [np.zeros(s.nCells) for s in subs], [b[s.cellMap] for s in subs], tol=1e-9, maxIter=300)
    assert pg.nIterations == po["nIterations"], (pg.nIterations, po["nIterations"])
    assert np.allclose(hg, ho, rtol=1e-8, atol=1e-13 * ho[0])
    assert np.abs(xg - xo[rank]).max() <= 1e-10 * np.abs(xo[rank]).max()
    res["pcg_iterations"] = [int(pg.nIterations)]
    res["pcg_hist_rel"] = float(np.max(np.abs(hg - ho) / np.maximum(ho, 1e-13 * ho[0])))
    # a SECOND solve on the same handle (PISO correctors 2, 3 ...): the first one stopped after a backward sweep and left the
    # ready-flag vector z populated; stale halo values must not be taken for published ones
    b2 = rng.uniform(-1, 1, gm.N)
    xg2, pg2, hg2 = ldu.pcg_solve(d, u, xo[rank], b2[pm.cellMap], tol=1e-9, maxIter=300, bou=bou)
    xo2, po2, ho2 = mro.pcg_solve_multi(osubs, [m[0] for m in mats], [m[1] for m in mats], [m[2] for m in mats],
                                        [v.copy() for v in xo], [b2[s.cellMap] for s in subs], tol=1e-9, maxIter=300)
    assert pg2.nIterations == po2["nIterations"], (pg2.nIterations, po2["nIterations"])
    assert np.allclose(hg2, ho2, rtol=1e-8, atol=1e-13 * ho2[0])
    assert np.abs(xg2 - xo2[rank]).max() <= 1e-10 * np.abs(xo2[rank]).max()
    res["pcg_iterations"].append(int(pg2.nIterations))
    res["pcg_hist_rel"] = max(res["pcg_hist_rel"], float(np.max(np.abs(hg2 - ho2) / np.maximum(ho2, 1e-13 * ho2[0]))))
    xs, ps, hs = fo.pcg_solve(gm, diag, upper, np.zeros(gm.N), b, tol=1e-9, maxIter=300)
    assert np.abs(xg - xs[pm.cellMap]).max() <= 1e-6 * np.abs(xs).max()      # same solution, different (rank-local) preconditioner

    # ---------------- the decomposed region step vs the serial oracle step (tight solver tolerances)
    cs = channel_case(gpm, *L)
    def tight(c):
        c.pdTol = 1e-13; c.pdRelTol = 0.0; c.pdFinalTol = 1e-13; c.pcorrTol = 1e-13; c.maxIter = 2000
        return c
    ro = fo.Region(gm, tight(default_controls(fo.OrcControls)), fo.BC(gm, cs["tA"], 1, cs["rvA"]), fo.BC(gm, cs["tU"], 3, cs["rvU"]),
                   fo.BC(gm, cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    F = pm.nInternalFaces
    bmap = pm.faceMap[F:]                                   # global face of every local boundary face
    PR = 3
    def types(t):
        return [PR if pm.patchNeighbRank[p] >= 0 else t[gpm.patchNames.index(pm.patchNames[p])] for p in range(pm.nPatches)]
    phys = np.array([pm.patchNeighbRank[p] < 0 for p in range(pm.nPatches) for _ in range(pm.patchSize[p])])
    rvA = np.zeros(pm.nBoundaryFaces); rvU = np.zeros((pm.nBoundaryFaces, 3))
    gB0 = gpm.nInternalFaces
    rvA[phys] = cs["rvA"][bmap[phys] - gB0]; rvU[phys] = cs["rvU"][bmap[phys] - gB0]
    dm = pm.device_mesh()                                    # collective: halo swap of cell centres
    rg = dm.region3d(tight(default_controls(capi.Region3dControls)), dm.bc(types(cs["tA"]), 1, rvA), dm.bc(types(cs["tU"]), 3, rvU),
                     dm.bc(types(cs["tP"]), 1), cs["alpha"][pm.cellMap], cs["U"][pm.cellMap], cs["pd"][pm.cellMap])
    res["fields_linf"] = 0.0
    def check(tag, tol):
        fo_, fg = ro.fields(), rg.fields()
        for k in ("alpha1", "U", "pd", "p", "rho"):
            ref = fo_[k][pm.cellMap]
            scale = max(np.abs(fo_[k]).max(), 1e-300)
            err = np.abs(ref - fg[k]).max() / scale
            assert err <= tol, (tag, k, err)
            res["fields_linf"] = max(res["fields_linf"], float(err))
        sign = np.where(pm.faceFlip != 0, -1.0, 1.0)
        err = np.abs(fo_["phi"][pm.faceMap] * sign - fg["phi"]).max() / np.abs(fo_["phi"]).max()
        assert err <= tol, (tag, "phi", err)
        res["fields_linf"] = max(res["fields_linf"], float(err))
    check("init", 1e-13)
    ro.correct_phi(); rg.correct_phi()
    check("correctPhi", 1e-9)
    for s in range(2):
        ro.step(); rg.step()
        check("step%d" % s, 1e-7)
    so, sg = ro.alpha_stats()[-1], rg.alpha_stats()[-1]
    assert np.allclose(so, sg, rtol=1e-8, atol=1e-12)
    rg.release()

    # ---------------- the decomposed region step vs the N-RANK oracle (one thread per sub-domain, rank-local DIC) at the DEFAULT
    # solver tolerances: identical PCG iteration counts, residual histories to 1e-8, fields to 1e-10
    infos = [dict(owner=sb.owner, nInternalFaces=sb.nInternalFaces, geom=sb.geometry(),
                  patches=[(sb.patchNames[p], int(sb.patchStart[p]), int(sb.patchSize[p]), int(sb.patchNeighbRank[p])) for p in range(sb.nPatches)])
             for sb in subs]
    coupled = fo.couple_submeshes(infos)
    # momentum predictor ON here (the reference's default [REF region3d/readRegion3dPISOControls.H:9-10]): b200PBiCG + DILU with processor
    # interfaces; all three velocity components made non-trivial so that none of the solves works on round-off noise
    gC = gpm.geometry()["C"]
    Uswirl = cs["U"].copy()
    Uswirl[:, 1] = 0.2 * np.sin(2 * np.pi * gC[:, 0] / L[0]) * np.cos(np.pi * gC[:, 1] / L[1]) * cs["alpha"]
    Uswirl[:, 2] = 0.1 * np.cos(2 * np.pi * gC[:, 0] / L[0]) * np.sin(np.pi * gC[:, 2] / L[2])
    def mom(c):
        c.momentumPredictor = 1; c.UTol = 1e-7; c.URelTol = 0.0
        return c
    ranks = []
    for q, sb in enumerate(subs):
        nbrRank, nbrCell, Cn, sg = coupled[q]
        kinds = [1 if r_ >= 0 else 0 for r_ in sb.patchNeighbRank]
        om = fo.Mesh(sb.owner, sb.neighbour, sb.patches, sg, sb.nCells, kinds, Cn)
        Fq = sb.nInternalFaces; bm = sb.faceMap[Fq:]
        ph = np.array([sb.patchNeighbRank[p] < 0 for p in range(sb.nPatches) for _ in range(sb.patchSize[p])])
        ra = np.zeros(sb.nBoundaryFaces); ru = np.zeros((sb.nBoundaryFaces, 3))
        ra[ph] = cs["rvA"][bm[ph] - gB0]; ru[ph] = cs["rvU"][bm[ph] - gB0]
        def ty(t, sb=sb):
            return [PR if sb.patchNeighbRank[p] >= 0 else t[gpm.patchNames.index(sb.patchNames[p])] for p in range(sb.nPatches)]
        ranks.append(dict(mesh=om, ctl=mom(default_controls(fo.OrcControls)), bcAlpha=fo.BC(om, ty(cs["tA"]), 1, ra), bcU=fo.BC(om, ty(cs["tU"]), 3, ru),
                          bcPd=fo.BC(om, ty(cs["tP"]), 1), alpha=cs["alpha"][sb.cellMap], U=Uswirl[sb.cellMap], pd=cs["pd"][sb.cellMap],
                          nbrRank=nbrRank, nbrCell=nbrCell))
    wo = fo.World(ranks)
    rg = dm.region3d(mom(default_controls(capi.Region3dControls)), dm.bc(types(cs["tA"]), 1, rvA), dm.bc(types(cs["tU"]), 3, rvU),
                     dm.bc(types(cs["tP"]), 1), cs["alpha"][pm.cellMap], Uswirl[pm.cellMap], cs["pd"][pm.cellMap])
    res["nrank_fields_linf"] = 0.0
    def check_n(tag, tol=1e-10):
        fo_, fg = wo.fields(rank), rg.fields()
        for k in ("alpha1", "U", "pd", "phi", "p", "rho"):
            err = np.abs(fo_[k] - fg[k]).max() / max(np.abs(fo_[k]).max(), 1e-300)
            assert err <= tol, (tag, k, err)
            res["nrank_fields_linf"] = max(res["nrank_fields_linf"], float(err))
        po, pg_ = wo.perfs(), rg.perfs()
        assert [p[2] for p in po] == [p[2] for p in pg_], (tag, [p[2] for p in po], [p[2] for p in pg_])
        ho_, hg_ = wo.res_hist(), rg.res_hist()
        assert ho_.shape == hg_.shape and np.allclose(ho_, hg_, rtol=1e-8, atol=1e-13), tag
        assert np.allclose(wo.alpha_stats(), rg.alpha_stats(), rtol=1e-12, atol=1e-25), tag
    check_n("init")
    wo.correct_phi(); rg.correct_phi()
    check_n("correctPhi")
    for s in range(2):
        wo.step(); rg.step()
        check_n("step%d" % s)
    assert np.allclose(wo.courant(), rg.courant(), rtol=1e-10)
    res["nrank_pcg_iterations"] = [int(p[2]) for p in rg.perfs()]
    rg.release(); dm.release(); ldu.release()
    return res


if __name__ == "__main__":
    main()
