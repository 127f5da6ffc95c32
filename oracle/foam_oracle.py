"""ctypes front-end of the CPU oracle (oracle/foam_oracle.cpp).  TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_MIXED, BC_PROCESSOR, BC_CALCULATED = 0, 1, 2, 3, 4

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


class OrcMesh(C.Structure):
    _fields_ = [("N", C.c_int), ("F", C.c_int), ("B", C.c_int), ("nPatches", C.c_int),
                ("own", ip), ("nei", ip), ("pStart", ip), ("pSize", ip), ("pKind", ip),
                ("Sf", dp), ("magSf", dp), ("Cf", dp), ("C", dp), ("V", dp), ("w", dp), ("dc", dp), ("Cn", dp)]


class OrcBC(C.Structure):
    _fields_ = [("type", ip), ("refValue", dp), ("refGrad", dp), ("valueFraction", dp)]


class OrcPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("nIterations", C.c_int), ("converged", C.c_int), ("singular", C.c_int)]


class OrcControls(C.Structure):
    _fields_ = [("deltaT", C.c_double),
                ("nCorr", C.c_int), ("nNonOrthCorr", C.c_int), ("nAlphaCorr", C.c_int), ("nAlphaSubCycles", C.c_int),
                ("cAlpha", C.c_double),
                ("rho1", C.c_double), ("rho2", C.c_double), ("nu1", C.c_double), ("nu2", C.c_double), ("sigma", C.c_double),
                ("g", C.c_double * 3),
                ("pdTol", C.c_double), ("pdRelTol", C.c_double), ("pdFinalTol", C.c_double), ("pdFinalRelTol", C.c_double),
                ("pcorrTol", C.c_double), ("pcorrRelTol", C.c_double),
                ("maxIter", C.c_int), ("minIter", C.c_int),
                ("adjustPhi", C.c_int), ("mulesIncludeOwn", C.c_int), ("divUScheme", C.c_int),
                ("momentumPredictor", C.c_int), ("nonOrthCorrected", C.c_int), ("pdRefCell", C.c_int), ("reserved0", C.c_int),
                ("pdRefValue", C.c_double), ("UTol", C.c_double), ("URelTol", C.c_double)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        if "pdRefCell" not in k:
            self.pdRefCell = -1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "foam_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_region_create.restype = C.c_void_p
        for n in ("orc_region_num_solves", "orc_region_res_hist", "orc_region_alpha_stats"):
            getattr(_LIB, n).restype = C.c_int
    return _LIB


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Mesh:
    """Holds the arrays alive and exposes the OrcMesh struct."""

    def __init__(self, owner, neighbour, patches, geom, nCells, patch_kinds=None, Cn=None):
        self.owner = i32(owner); self.neighbour = i32(neighbour)
        self.N = int(nCells); self.F = int(self.neighbour.size); self.B = int(self.owner.size - self.F)
        self.patches = list(patches)
        self.pStart = i32([p[1] for p in patches]); self.pSize = i32([p[2] for p in patches])
        self.pKind = i32(patch_kinds if patch_kinds is not None else [0] * len(patches))
        self.Sf = f64(geom["Sf"]); self.magSf = f64(geom["magSf"]); self.Cf = f64(geom["Cf"])
        self.C = f64(geom["C"]); self.V = f64(geom["V"]); self.w = f64(geom["weights"]); self.dc = f64(geom["deltaCoeffs"])
        self.Cn = None if Cn is None else f64(Cn)     # [B,3] neighbour-rank cell centres of processor faces (decomposed runs)
        self.s = OrcMesh(self.N, self.F, self.B, len(patches), _i(self.owner), _i(self.neighbour), _i(self.pStart),
                         _i(self.pSize), _i(self.pKind), _d(self.Sf), _d(self.magSf), _d(self.Cf), _d(self.C),
                         _d(self.V), _d(self.w), _d(self.dc), None if self.Cn is None else _d(self.Cn))

    @property
    def ref(self):
        return C.byref(self.s)

    def patch_slice(self, name):
        for p in self.patches:
            if p[0] == name:
                return slice(p[1] - self.F, p[1] - self.F + p[2])
        raise KeyError(name)


class BC:
    def __init__(self, mesh, types, nc=1, refValue=None, refGrad=None, valueFraction=None):
        B = mesh.B
        self.types = i32(types)
        self.refValue = f64(np.zeros(nc * B) if refValue is None else refValue).reshape(-1)
        self.refGrad = f64(np.zeros(nc * B) if refGrad is None else refGrad).reshape(-1)
        self.valueFraction = f64(np.zeros(B) if valueFraction is None else valueFraction)
        self.s = OrcBC(_i(self.types), _d(self.refValue), _d(self.refGrad), _d(self.valueFraction))

    @property
    def ref(self):
        return C.byref(self.s)


# ---- thin wrappers (numpy in, numpy out) -----------------------------------------------------
def amul(m, diag, upper, lower, psi, bou=None, psiNbr=None):
    out = np.empty(m.N)
    lib().orc_amul(m.ref, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(f64(psi)), _d(out),
                   _d(f64(bou)) if bou is not None else None, _d(f64(psiNbr)) if psiNbr is not None else None)
    return out


def sumA(m, diag, upper, lower, bou=None):
    out = np.empty(m.N)
    lib().orc_sumA(m.ref, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(out), _d(f64(bou)) if bou is not None else None)
    return out


def dic_calc_rd(m, diag, upper):
    out = np.empty(m.N)
    lib().orc_dic_calc_rd(m.ref, _d(f64(diag)), _d(f64(upper)), _d(out))
    return out


def dic_precondition(m, rD, upper, rA):
    out = np.empty(m.N)
    lib().orc_dic_precondition(m.ref, _d(f64(rD)), _d(f64(upper)), _d(f64(rA)), _d(out))
    return out


def pcg_solve(m, diag, upper, x0, b, tol=1e-6, relTol=0.0, minIter=0, maxIter=1000, bou=None):
    x = f64(x0).copy()
    perf = OrcPerf()
    hist = np.full(maxIter + 2, -1.0)
    lib().orc_pcg_solve(m.ref, _d(f64(diag)), _d(f64(upper)), _d(f64(bou)) if bou is not None else None, _d(x),
                        _d(f64(b)), C.c_double(tol), C.c_double(relTol), C.c_int(minIter), C.c_int(maxIter),
                        C.byref(perf), _d(hist))
    return x, perf, hist[:perf.nIterations + 1].copy()


def dilu_calc_rd(m, diag, upper, lower):
    out = np.empty(m.N)
    lib().orc_dilu_calc_rd(m.ref, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(out))
    return out


def dilu_precondition(m, rD, upper, lower, rA, transpose=False):
    out = np.empty(m.N)
    lib().orc_dilu_precondition(m.ref, _d(f64(rD)), _d(f64(upper)), _d(f64(lower)), _d(f64(rA)), _d(out), C.c_int(1 if transpose else 0))
    return out


def pbicg_solve(m, diag, upper, lower, x0, b, tol=1e-6, relTol=0.0, minIter=0, maxIter=1000):
    x = f64(x0).copy()
    perf = OrcPerf()
    hist = np.full(maxIter + 2, -1.0)
    lib().orc_pbicg_solve(m.ref, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(x), _d(f64(b)), C.c_double(tol), C.c_double(relTol),
                          C.c_int(minIter), C.c_int(maxIter), C.byref(perf), _d(hist))
    return x, perf, hist[:perf.nIterations + 1].copy()


def bc_evaluate(m, bc, internal, nc=1, bval=None):
    out = np.zeros(nc * m.B) if bval is None else f64(bval).copy().reshape(-1)
    lib().orc_bc_evaluate(m.ref, bc.ref, _d(f64(internal).reshape(-1)), _d(out), C.c_int(nc))
    return out


def interpolate(m, vf, vfb, nc=1):
    out = np.empty(nc * (m.F + m.B))
    lib().orc_interpolate(m.ref, _d(f64(vf).reshape(-1)), _d(f64(vfb).reshape(-1)), _d(out), C.c_int(nc))
    return out


def sngrad(m, bc, vf, vfb):
    out = np.empty(m.F + m.B)
    lib().orc_sngrad(m.ref, bc.ref, _d(f64(vf)), _d(f64(vfb)), _d(out))
    return out


def surface_integrate(m, ssf):
    out = np.empty(m.N)
    lib().orc_surface_integrate(m.ref, _d(f64(ssf)), _d(out))
    return out


def grad_gauss(m, bc, vf, vfb):
    g = np.empty(3 * m.N); gb = np.empty(3 * m.B)
    lib().orc_grad_gauss(m.ref, bc.ref, _d(f64(vf)), _d(f64(vfb)), _d(g), _d(gb))
    return g.reshape(-1, 3), gb.reshape(-1, 3)


def reconstruct_tensor(m):
    out = np.empty(6 * m.N)
    lib().orc_reconstruct_tensor(m.ref, _d(out))
    return out


def reconstruct(m, invT, ssf):
    out = np.empty(3 * m.N)
    lib().orc_reconstruct(m.ref, _d(f64(invT)), _d(f64(ssf)), _d(out))
    return out.reshape(-1, 3)


def flux_limited(m, scheme, faceFlux, vf, vfb, grad=None):
    out = np.empty(m.F + m.B)
    lib().orc_flux_limited(m.ref, C.c_int(scheme), _d(f64(faceFlux)), _d(f64(vf)), _d(f64(vfb)),
                           _d(f64(grad).reshape(-1)) if grad is not None else None, _d(out))
    return out


def mules_limiter(m, bc, psi, psib, psi0, phiBD, phiCorr, deltaT, psiMax=1.0, psiMin=0.0, nIter=3, includeOwn=0):
    lam = np.empty(m.F + m.B)
    lib().orc_mules_limiter(m.ref, _d(lam), _d(f64(psi)), _d(f64(psib)), _d(f64(psi0)), _d(f64(phiBD)), _d(f64(phiCorr)),
                            C.c_double(deltaT), C.c_double(psiMax), C.c_double(psiMin), C.c_int(nIter),
                            C.c_int(includeOwn), None, bc.ref, None)
    return lam


def mules_explicit_solve(m, bc, psi, psib, psi0, phi, phiPsi, deltaT, psiMax=1.0, psiMin=0.0, includeOwn=0):
    psi = f64(psi).copy(); phiPsi = f64(phiPsi).copy(); lam = np.empty(m.F + m.B)
    lib().orc_mules_explicit_solve(m.ref, _d(psi), _d(f64(psib)), _d(f64(psi0)), _d(f64(phi)), _d(phiPsi),
                                   C.c_double(deltaT), C.c_double(psiMax), C.c_double(psiMin), C.c_int(includeOwn),
                                   _d(lam), bc.ref, None)
    return psi, phiPsi, lam


def laplacian_assemble(m, bc, gammaf, pb, phi=None):
    diag = np.empty(m.N); upper = np.empty(m.F); source = np.empty(m.N); ic = np.empty(m.B); bcv = np.empty(m.B)
    lib().orc_laplacian_assemble(m.ref, bc.ref, _d(f64(gammaf)), _d(f64(pb)), _d(f64(phi)) if phi is not None else None,
                                 _d(diag), _d(upper), _d(source), _d(ic), _d(bcv))
    return diag, upper, source, ic, bcv


def add_boundary(m, ic, bcv, diag, source):
    d = f64(diag).copy(); s = f64(source).copy()
    lib().orc_add_boundary(m.ref, _d(f64(ic)), _d(f64(bcv)), _d(d), _d(s))
    return d, s


def matrix_flux(m, upper, lower, ic, bcv, psi):
    out = np.empty(m.F + m.B)
    lib().orc_matrix_flux(m.ref, _d(f64(upper)), _d(f64(lower)), _d(f64(ic)), _d(f64(bcv)), _d(f64(psi)), _d(out))
    return out


def non_orthogonality(m):
    """(degrees, orthogonal) of mesh.orthogonal() [UPSTREAM surfaceInterpolation::makeCorrectionVectors] on one rank"""
    s = np.empty(2)
    lib().orc_non_orth_sums(m.ref, _d(s))
    deg = float(np.degrees(np.arcsin(min(s[0] / s[1], 1.0)))) if s[1] > 0 else 0.0
    return deg, deg < 0.1


def non_orth_correction(m, bc, vf, vfb, gammaf=None):
    """(gammaf*magSf) * correctedSnGrad::correction(vf) on every face (single rank)"""
    g, _ = grad_gauss(m, bc, vf, vfb)
    out = np.empty(m.F + m.B)
    lib().orc_non_orth_correction(m.ref, _d(f64(gammaf)) if gammaf is not None else None, _d(f64(g).reshape(-1)), None, _d(out))
    return out


def alpha_stats(m, a, ab):
    out = np.empty(3)
    lib().orc_alpha_stats(m.ref, _d(f64(a)), _d(f64(ab)), _d(out))
    return out


class Region:
    """Oracle replay of region3d/solve3d.H for a laminar 3D-only region."""

    def __init__(self, mesh, controls, bcAlpha, bcU, bcPd, alpha, U, pd):
        self.mesh = mesh; self.ctl = controls
        self._keep = (bcAlpha, bcU, bcPd)
        self.h = C.c_void_p(lib().orc_region_create(mesh.ref, C.byref(controls), bcAlpha.ref, bcU.ref, bcPd.ref,
                                                    _d(f64(alpha)), _d(f64(U).reshape(-1)), _d(f64(pd))))
        lib().orc_region_init_interface(self.h)

    def correct_phi(self):
        lib().orc_region_correct_phi(self.h)

    def step(self, n=1):
        for _ in range(n):
            lib().orc_region_step(self.h)

    def fields(self):
        m = self.mesh
        a = np.empty(m.N); U = np.empty(3 * m.N); pd = np.empty(m.N); phi = np.empty(m.F + m.B); p = np.empty(m.N); rho = np.empty(m.N)
        lib().orc_region_get(self.h, _d(a), _d(U), _d(pd), _d(phi), _d(p), _d(rho))
        return dict(alpha1=a, U=U.reshape(-1, 3), pd=pd, phi=phi, p=p, rho=rho)

    def perfs(self):
        n = lib().orc_region_num_solves(self.h)
        arr = (OrcPerf * n)()
        lib().orc_region_get_perfs(self.h, arr)
        return [(p.initialResidual, p.finalResidual, p.nIterations, p.converged, p.singular) for p in arr]

    def res_hist(self):
        n = lib().orc_region_res_hist(self.h, None, 0)
        out = np.empty(n)
        lib().orc_region_res_hist(self.h, _d(out), n)
        return out

    def alpha_stats(self):
        n = lib().orc_region_alpha_stats(self.h, None, 0)
        out = np.empty(n)
        lib().orc_region_alpha_stats(self.h, _d(out), n)
        return out.reshape(-1, 3)

    def cont_errs(self):
        out = np.empty(3)
        lib().orc_region_cont_errs(self.h, _d(out))
        return out

    def timers(self):
        out = np.empty(3)
        lib().orc_region_timers(self.h, _d(out))
        return out

    def courant(self):
        out = np.empty(3)
        lib().orc_region_courant(self.h, _d(out))
        return out

    def clear_log(self):
        lib().orc_region_clear_log(self.h)

    def __del__(self):
        try:
            lib().orc_region_destroy(self.h)
        except Exception:
            pass


# ---- N-rank emulation (one thread per sub-domain): orc_world_* -------------------------------------------------------------------
def couple_submeshes(subs):
    """subs[r] = dict(owner [F+B] local cells, nInternalFaces, patches [(name, start, size, neighbRank)], geom dict of this sub-mesh).
    Returns per rank (nbrRank [B], nbrCell [B], Cn [B,3], geom') where geom' carries, on processor faces, the weights / deltaCoeffs of
    the internal face they were, seen from this side [UPSTREAM surfaceInterpolation.C makeWeights / makeDeltaCoeffs; processorFvPatch]:
    w = |Sf.(Cn-Cf)| / (|Sf.(Cf-C)| + |Sf.(Cn-Cf)|), dc = 1/|Cn - C|.  Both sides of a processor patch list the same faces in the
    same order (decomposePar), so face i of patch r->q meets face i of patch q->r."""
    out = []
    for r, s in enumerate(subs):
        F = s["nInternalFaces"]; B = s["owner"].size - F
        nbrRank = np.full(B, -1, dtype=np.int32); nbrCell = np.zeros(B, dtype=np.int32)
        Cn = np.zeros((B, 3))
        g = {k: np.array(v, dtype=np.float64, copy=True) for k, v in s["geom"].items()}
        for (name, start, size, q) in s["patches"]:
            if q < 0 or size == 0:
                continue
            t = subs[q]
            pq = [p for p in t["patches"] if p[3] == r][0]
            assert pq[2] == size
            cells_q = t["owner"][pq[1]:pq[1] + size]
            sl = slice(start - F, start - F + size)
            nbrRank[sl] = q; nbrCell[sl] = cells_q
            cn = np.asarray(t["geom"]["C"]).reshape(-1, 3)[cells_q]
            Cn[sl] = cn
            fsl = slice(start, start + size)
            Sf = g["Sf"].reshape(-1, 3)[fsl]; Cf = g["Cf"].reshape(-1, 3)[fsl]
            co = g["C"].reshape(-1, 3)[s["owner"][fsl]]
            dO, dN = Cf - co, cn - Cf
            sO = np.abs((Sf[:, 0] * dO[:, 0] + Sf[:, 1] * dO[:, 1]) + Sf[:, 2] * dO[:, 2])
            sN = np.abs((Sf[:, 0] * dN[:, 0] + Sf[:, 1] * dN[:, 1]) + Sf[:, 2] * dN[:, 2])
            g["weights"][fsl] = sN / (sO + sN)
            d = cn - co
            g["deltaCoeffs"][fsl] = 1.0 / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        out.append((nbrRank, nbrCell, Cn, g))
    return out


class World:
    """N regions on the sub-meshes of one decomposition, stepped by one thread each (halo swaps + rank-ordered reductions in shared
    memory).  ranks[r] = dict(mesh=Mesh (pKind 1 + Cn on processor patches), ctl, bcAlpha, bcU, bcPd, alpha, U, pd, nbrRank, nbrCell)."""

    def __init__(self, ranks):
        L = lib()
        L.orc_world_create.restype = C.c_void_p
        self.n_ranks = len(ranks)
        self.regions = []
        for rk in ranks:
            reg = Region.__new__(Region)
            reg.mesh = rk["mesh"]; reg.ctl = rk["ctl"]; reg._keep = (rk["bcAlpha"], rk["bcU"], rk["bcPd"])
            reg.h = C.c_void_p(L.orc_region_create(rk["mesh"].ref, C.byref(rk["ctl"]), rk["bcAlpha"].ref, rk["bcU"].ref, rk["bcPd"].ref,
                                                   _d(f64(rk["alpha"])), _d(f64(rk["U"]).reshape(-1)), _d(f64(rk["pd"]))))
            self.regions.append(reg)
        self._nr = [i32(rk["nbrRank"]) for rk in ranks]; self._nc = [i32(rk["nbrCell"]) for rk in ranks]
        R = self.n_ranks
        regs = (C.c_void_p * R)(*[r.h for r in self.regions])
        pr = (ip * R)(*[_i(a) for a in self._nr]); pc = (ip * R)(*[_i(a) for a in self._nc])
        self.h = C.c_void_p(L.orc_world_create(C.c_int(R), regs, pr, pc))

    def correct_phi(self):
        lib().orc_world_correct_phi(self.h)

    def step(self, n=1):
        lib().orc_world_step(self.h, C.c_int(n))

    def perfs(self):
        return self.regions[0].perfs()        # global quantities: identical on every rank

    def res_hist(self):
        return self.regions[0].res_hist()

    def alpha_stats(self):
        return self.regions[0].alpha_stats()

    def cont_errs(self):
        return self.regions[0].cont_errs()

    def courant(self):
        out = np.empty(3)
        lib().orc_world_courant(self.h, _d(out))
        return out

    def clear_log(self):
        for r in self.regions:
            r.clear_log()

    def fields(self, rank):
        return self.regions[rank].fields()

    @staticmethod
    def best_split(dims, n):
        """n = nx*ny*nz sub-domains of a dims box with the least processor-patch area (what one writes into decomposeParDict)"""
        best = None
        for a in range(1, n + 1):
            if n % a:
                continue
            for b in range(1, n // a + 1):
                if (n // a) % b:
                    continue
                c = n // a // b
                if a > dims[0] or b > dims[1] or c > dims[2]:
                    continue
                area = (a - 1) * dims[1] * dims[2] + (b - 1) * dims[0] * dims[2] + (c - 1) * dims[0] * dims[1]
                if best is None or area < best[0]:
                    best = (area, (a, b, c))
        return best[1]

    @staticmethod
    def build(dims, lengths, weir, channel_fields, set_controls, n_proc):
        """the synthetic channel of bench.py decomposed n_proc ways (`simple`, mesh_oracle.decompose) -> (total cells, World)"""
        from . import mesh_oracle as mo
        mesh = mo.block_mesh_box(*dims, *lengths)
        if weir:
            mesh = mo.subset_mesh(mesh, mo.weir_keep_mask(*dims, *weir))
        g = mo.geometry(mesh)
        split = World.best_split(dims, n_proc)
        proc = mo.simple_decomp(g["C"], split)
        subs = mo.decompose(mesh, proc, n_proc)
        F = mesh["neighbour"].size
        infos = []
        for s in subs:
            ids = s["faceIds"]; flip = s["flip"]
            sg = dict(Sf=np.where(flip[:, None], -g["Sf"][ids], g["Sf"][ids]), magSf=g["magSf"][ids], Cf=g["Cf"][ids], C=g["C"][s["cells"]],
                      V=g["V"][s["cells"]], weights=g["weights"][ids].copy(), deltaCoeffs=g["deltaCoeffs"][ids].copy())
            nIF = s["neighbour"].size
            sg["weights"][nIF:] = 1.0
            infos.append(dict(owner=s["owner"], nInternalFaces=nIF, patches=s["patches"], geom=sg))
        coupled = couple_submeshes(infos)
        ranks = []
        for r, s in enumerate(subs):
            nbrRank, nbrCell, Cn, sg = coupled[r]
            nIF = s["neighbour"].size
            patches3 = [(p[0], p[1], p[2]) for p in s["patches"]]
            kinds = [1 if p[3] >= 0 else 0 for p in s["patches"]]
            m = Mesh(s["owner"], s["neighbour"], patches3, sg, s["nCells"], kinds, Cn)
            names = [p[0] for p in s["patches"]]
            slices = {n: m.patch_slice(n) for n in names}
            cs = channel_fields(names, sg["C"], sg["Cf"][nIF:], slices, lengths)
            ctl = set_controls(OrcControls())
            ranks.append(dict(mesh=m, ctl=ctl, bcAlpha=BC(m, cs["tA"], 1, cs["rvA"]), bcU=BC(m, cs["tU"], 3, cs["rvU"]), bcPd=BC(m, cs["tP"], 1),
                              alpha=cs["alpha"], U=cs["U"], pd=cs["pd"], nbrRank=nbrRank, nbrCell=nbrCell))
        return int(mesh["nCells"]), World(ranks)
