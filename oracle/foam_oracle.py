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
                ("Sf", dp), ("magSf", dp), ("Cf", dp), ("C", dp), ("V", dp), ("w", dp), ("dc", dp)]


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

    def __init__(self, owner, neighbour, patches, geom, nCells, patch_kinds=None):
        self.owner = i32(owner); self.neighbour = i32(neighbour)
        self.N = int(nCells); self.F = int(self.neighbour.size); self.B = int(self.owner.size - self.F)
        self.patches = list(patches)
        self.pStart = i32([p[1] for p in patches]); self.pSize = i32([p[2] for p in patches])
        self.pKind = i32(patch_kinds if patch_kinds is not None else [0] * len(patches))
        self.Sf = f64(geom["Sf"]); self.magSf = f64(geom["magSf"]); self.Cf = f64(geom["Cf"])
        self.C = f64(geom["C"]); self.V = f64(geom["V"]); self.w = f64(geom["weights"]); self.dc = f64(geom["deltaCoeffs"])
        self.s = OrcMesh(self.N, self.F, self.B, len(patches), _i(self.owner), _i(self.neighbour), _i(self.pStart),
                         _i(self.pSize), _i(self.pKind), _d(self.Sf), _d(self.magSf), _d(self.Cf), _d(self.C),
                         _d(self.V), _d(self.w), _d(self.dc))

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
                            C.c_int(includeOwn), None, bc.ref)
    return lam


def mules_explicit_solve(m, bc, psi, psib, psi0, phi, phiPsi, deltaT, psiMax=1.0, psiMin=0.0, includeOwn=0):
    psi = f64(psi).copy(); phiPsi = f64(phiPsi).copy(); lam = np.empty(m.F + m.B)
    lib().orc_mules_explicit_solve(m.ref, _d(psi), _d(f64(psib)), _d(f64(psi0)), _d(f64(phi)), _d(phiPsi),
                                   C.c_double(deltaT), C.c_double(psiMax), C.c_double(psiMin), C.c_int(includeOwn),
                                   _d(lam), bc.ref)
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
