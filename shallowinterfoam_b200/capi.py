"""ctypes binding of include/b200foam.h (the C-ABI of libb200foam.so).

The product library is the nvcc-built, sm_100a `libb200foam.so` next to this file.  There is no CPU fallback:
`load()` raises if the library is missing, and every compute call raises `B200Error` without a CUDA device.
`B200Lib(path)` can also wrap another build of the same ABI (tests use the SIMT-emulated test build).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200foam.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)

BC_FIXED_VALUE, BC_ZERO_GRADIENT, BC_MIXED, BC_PROCESSOR, BC_CALCULATED = 0, 1, 2, 3, 4
SCHEME_VANLEER, SCHEME_INTERFACE_COMPRESSION, SCHEME_UPWIND, SCHEME_LINEAR = 0, 1, 2, 3

EXPORTS = [
    "b200_init", "b200_nccl_unique_id", "b200_finalize", "b200_last_error", "b200_device_count",
    "b200_kernel_launch_count", "b200_set_option", "b200_timers_reset", "b200_timers_get", "b200_timer_name",
    "b200_ldu_get_or_create", "b200_ldu_release", "b200_ldu_get_renumbering", "b200_ldu_get_tiling", "b200_ldu_get_addressing", "b200_debug_get_sweep_trace",
    "b200_amul", "b200_sumA", "b200_dic_calc_reciprocal_d", "b200_dic_precondition", "b200_pcg_solve",
    "b200_mesh_create", "b200_mesh_release", "b200_mesh_ldu",
    "b200_interpolate", "b200_sngrad", "b200_surface_integrate", "b200_grad_gauss", "b200_reconstruct",
    "b200_flux_limited", "b200_mules_explicit_solve", "b200_mules_limiter",
    "b200_laplacian_assemble", "b200_matrix_flux", "b200_alpha_stats",
    "b200_region3d_create", "b200_region3d_release", "b200_region3d_correct_phi", "b200_region3d_step",
    "b200_region3d_step_host", "b200_region3d_get_fields", "b200_region3d_set_fields",
    "b200_region3d_set_patch_mixed", "b200_region3d_get_patch_internal", "b200_region3d_num_solves",
    "b200_region3d_get_perfs", "b200_region3d_get_res_hist", "b200_region3d_get_alpha_stats",
    "b200_region3d_get_cont_errs", "b200_region3d_clear_log", "b200_region3d_get_courant", "b200_region3d_restore", "b200_ldu_invalidate",
    "b200_pbicg_solve", "b200_dilu_calc_reciprocal_d", "b200_dilu_precondition", "b200_polymesh_shear", "b200_mesh_non_orthogonality", "b200_non_orth_correction",
    "b200_region3d_capture", "b200_region3d_num_captured", "b200_region3d_get_captured_solve", "b200_region3d_get_captured_mules",
    "b200_polymesh_box", "b200_polymesh_box_weir", "b200_polymesh_jitter", "b200_polymesh_agglomerate", "b200_polymesh_release", "b200_polymesh_sizes",
    "b200_polymesh_get", "b200_polymesh_patch_name", "b200_polymesh_geometry", "b200_decompose_simple",
    "b200_polymesh_submesh", "b200_polymesh_submesh_maps", "b200_mesh_create_from_polymesh",
]


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200foam error %d: %s" % (code, msg))
        self.code = code


class PatchDesc(C.Structure):
    _fields_ = [("nFaces", C.c_int32), ("faceCells", ip), ("neighbRank", C.c_int32), ("tag", C.c_int32)]


class SolverControls(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("minIter", C.c_int32), ("maxIter", C.c_int32)]


class SolverPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double), ("nIterations", C.c_int32),
                ("converged", C.c_int32), ("singular", C.c_int32)]


class BCStruct(C.Structure):
    _fields_ = [("type", ip), ("refValue", dp), ("refGrad", dp), ("valueFraction", dp)]


class MeshDesc(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nInternalFaces", C.c_int32), ("nBoundaryFaces", C.c_int32), ("nPatches", C.c_int32),
                ("owner", ip), ("neighbour", ip), ("patchStart", ip), ("patchSize", ip), ("patchNeighbRank", ip),
                ("Sf", dp), ("magSf", dp), ("Cf", dp), ("C", dp), ("V", dp), ("weights", dp), ("deltaCoeffs", dp)]


class Region3dControls(C.Structure):
    _fields_ = [("deltaT", C.c_double),
                ("nCorr", C.c_int32), ("nNonOrthCorr", C.c_int32), ("nAlphaCorr", C.c_int32), ("nAlphaSubCycles", C.c_int32),
                ("cAlpha", C.c_double),
                ("rho1", C.c_double), ("rho2", C.c_double), ("nu1", C.c_double), ("nu2", C.c_double), ("sigma", C.c_double),
                ("g", C.c_double * 3),
                ("pdTol", C.c_double), ("pdRelTol", C.c_double), ("pdFinalTol", C.c_double), ("pdFinalRelTol", C.c_double),
                ("pcorrTol", C.c_double), ("pcorrRelTol", C.c_double),
                ("maxIter", C.c_int32), ("minIter", C.c_int32),
                ("adjustPhi", C.c_int32), ("mulesIncludeOwn", C.c_int32), ("divUScheme", C.c_int32),
                ("momentumPredictor", C.c_int32), ("nonOrthCorrected", C.c_int32), ("pdRefCell", C.c_int32), ("reserved0", C.c_int32),
                ("pdRefValue", C.c_double), ("UTol", C.c_double), ("URelTol", C.c_double)]

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        if "pdRefCell" not in k:
            self.pdRefCell = -1


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return None if a is None else a.ctypes.data_as(dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(ip)


class BC:
    """Boundary description of one field (types per patch + per-face mixed data)."""

    def __init__(self, nB, types, nc=1, refValue=None, refGrad=None, valueFraction=None):
        self.types = i32(types)
        self.nc = nc
        self.refValue = f64(np.zeros(nc * nB) if refValue is None else refValue).reshape(-1).copy()
        self.refGrad = f64(np.zeros(nc * nB) if refGrad is None else refGrad).reshape(-1).copy()
        self.valueFraction = f64(np.zeros(nB) if valueFraction is None else valueFraction).copy()
        self.s = BCStruct(_i(self.types), _d(self.refValue), _d(self.refGrad), _d(self.valueFraction))

    @property
    def ref(self):
        return C.byref(self.s)


class B200Lib:
    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise ImportError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback." % path)
        self.path = path
        self.c = C.CDLL(path)
        c = self.c
        c.b200_last_error.restype = C.c_char_p
        c.b200_timer_name.restype = C.c_char_p
        c.b200_polymesh_patch_name.restype = C.c_char_p
        c.b200_kernel_launch_count.restype = C.c_longlong
        c.b200_mesh_ldu.restype = C.c_void_p
        c.b200_mesh_ldu.argtypes = [C.c_void_p]
        self._initialised = False

    # ------------------------------------------------------------------ plumbing
    def check(self, rc):
        if rc != 0:
            raise B200Error(rc, self.c.b200_last_error().decode())

    def missing_exports(self):
        return [n for n in EXPORTS if not hasattr(self.c, n)]

    def device_count(self):
        return int(self.c.b200_device_count())

    def init(self, device=0, rank=0, n_ranks=1, nccl_id=None):
        buf = None if nccl_id is None else (C.c_char * len(nccl_id)).from_buffer_copy(nccl_id)
        self.check(self.c.b200_init(C.c_int(device), C.c_int(rank), C.c_int(n_ranks), buf,
                                    C.c_size_t(0 if nccl_id is None else len(nccl_id))))
        self._initialised = True

    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        self.check(self.c.b200_nccl_unique_id(buf))
        return bytes(buf)

    def set_option(self, name, value):
        self.check(self.c.b200_set_option(name.encode(), C.c_double(value)))

    def launches(self):
        return int(self.c.b200_kernel_launch_count())

    def timers_reset(self):
        self.check(self.c.b200_timers_reset())

    def timers(self):
        n = 16
        ms = np.zeros(n); calls = np.zeros(n, dtype=np.int64)
        self.check(self.c.b200_timers_get(_d(ms), calls.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(n)))
        out = {}
        for i in range(n):
            name = self.c.b200_timer_name(C.c_int(i)).decode()
            if name:
                out[name] = (float(ms[i]), int(calls[i]))
        return out

    # ------------------------------------------------------------------ host mesh tools
    def polymesh_box(self, nx, ny, nz, lx, ly, lz, weir=None, jitter=None, shear=None):
        """jitter = (frac, seed): skew the mesh by moving interior vertices by up to frac of a cell size (b200_polymesh_jitter);
        shear = (sxy, sxz, syz): affine shear, non-orthogonal but not skewed (b200_polymesh_shear)"""
        h = C.c_void_p()
        if weir is None:
            self.check(self.c.b200_polymesh_box(C.c_int32(nx), C.c_int32(ny), C.c_int32(nz), C.c_double(lx), C.c_double(ly),
                                                C.c_double(lz), C.byref(h)))
        else:
            x0, x1, zf = weir
            self.check(self.c.b200_polymesh_box_weir(C.c_int32(nx), C.c_int32(ny), C.c_int32(nz), C.c_double(lx), C.c_double(ly),
                                                     C.c_double(lz), C.c_double(x0), C.c_double(x1), C.c_double(zf), C.byref(h)))
        if jitter is not None:
            self.c.b200_polymesh_jitter.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64]
            self.check(self.c.b200_polymesh_jitter(h, C.c_double(jitter[0]), C.c_double(lx / nx), C.c_double(ly / ny), C.c_double(lz / nz),
                                                   C.c_uint64(jitter[1])))
        if shear is not None:
            self.check(self.c.b200_polymesh_shear(h, C.c_double(shear[0]), C.c_double(shear[1]), C.c_double(shear[2])))
        return PolyMesh(self, h)

    # ------------------------------------------------------------------ lduAddressing-level API
    def ldu_create(self, nCells, lower, upper, patches=(), key=None):
        """patches: list of (faceCells int32 array, neighbRank)"""
        lower = i32(lower); upper = i32(upper)
        keep = [i32(p[0]) for p in patches]
        arr = (PatchDesc * max(1, len(patches)))()
        for k, p in enumerate(patches):
            arr[k] = PatchDesc(keep[k].size, _i(keep[k]), int(p[1]), 0)
        h = C.c_void_p()
        self.check(self.c.b200_ldu_get_or_create(C.c_void_p(key), C.c_int32(nCells), C.c_int32(lower.size), _i(lower), _i(upper),
                                                 C.c_int32(len(patches)), arr, C.byref(h)))
        return Ldu(self, h, nCells, lower.size, sum(k.size for k in keep))


class Ldu:
    def __init__(self, lib, h, N, F, B):
        self.lib, self.h, self.N, self.F, self.B = lib, h, N, F, B

    def release(self):
        if self.h:
            self.lib.check(self.lib.c.b200_ldu_release(self.h))
            self.h = None

    def renumbering(self):
        rowCell = np.empty(self.N, dtype=np.int32); level = np.empty(self.N, dtype=np.int32); nl = C.c_int32()
        self.lib.check(self.lib.c.b200_ldu_get_renumbering(self.h, _i(rowCell), _i(level), C.byref(nl)))
        return rowCell, level, nl.value

    def tiling(self):
        bl, tr, nt = C.c_int32(), C.c_int32(), C.c_int32()
        self.lib.check(self.lib.c.b200_ldu_get_tiling(self.h, C.byref(bl), C.byref(tr), C.byref(nt), None))
        ts = np.empty(nt.value + 1, dtype=np.int32)
        self.lib.check(self.lib.c.b200_ldu_get_tiling(self.h, None, None, None, _i(ts)))
        return bl.value, tr.value, ts

    def sweep_trace(self):
        _, _, ts = self.tiling()
        nt = ts.size - 1
        out = np.zeros(8 * nt, dtype=np.uint64); tl = np.zeros(nt, dtype=np.int32)
        self.lib.check(self.lib.c.b200_debug_get_sweep_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_ulonglong)), C.c_int32(out.size), _i(tl)))
        return out.reshape(nt, 8), tl, ts

    def addressing(self):
        os_ = np.empty(self.N + 1, dtype=np.int32); lo = np.empty(self.F, dtype=np.int32); ls = np.empty(self.N + 1, dtype=np.int32)
        self.lib.check(self.lib.c.b200_ldu_get_addressing(self.h, _i(os_), _i(lo), _i(ls)))
        return os_, lo, ls

    def amul(self, diag, upper, psi, lower=None, bou=None, psiNbr=None):
        out = np.empty(self.N)
        upper = f64(upper); lower_ = None if lower is None else f64(lower)
        bou_ = None if bou is None else f64(bou); nbr_ = None if psiNbr is None else f64(psiNbr)
        self.lib.check(self.lib.c.b200_amul(self.h, _d(f64(diag)), _d(upper), _d(lower_), _d(f64(psi)), _d(out), _d(bou_), _d(nbr_)))
        return out

    def sumA(self, diag, upper, lower=None, bou=None):
        out = np.empty(self.N)
        lower_ = None if lower is None else f64(lower); bou_ = None if bou is None else f64(bou)
        self.lib.check(self.lib.c.b200_sumA(self.h, _d(f64(diag)), _d(f64(upper)), _d(lower_), _d(out), _d(bou_)))
        return out

    def dic_calc_rd(self, diag, upper):
        out = np.empty(self.N)
        self.lib.check(self.lib.c.b200_dic_calc_reciprocal_d(self.h, _d(f64(diag)), _d(f64(upper)), _d(out)))
        return out

    def dic_precondition(self, rD, upper, rA):
        out = np.empty(self.N)
        self.lib.check(self.lib.c.b200_dic_precondition(self.h, _d(f64(rD)), _d(f64(upper)), _d(f64(rA)), _d(out)))
        return out

    def pcg_solve(self, diag, upper, x0, b, tol=1e-6, relTol=0.0, minIter=0, maxIter=1000, bou=None, lower=None):
        x = f64(x0).copy()
        ctl = SolverControls(tol, relTol, minIter, maxIter)
        perf = SolverPerf()
        hist = np.full(maxIter + 2, -1.0)
        bou_ = None if bou is None else f64(bou); lower_ = None if lower is None else f64(lower)
        self.lib.check(self.lib.c.b200_pcg_solve(self.h, _d(f64(diag)), _d(f64(upper)), _d(lower_), _d(bou_), _d(x), _d(f64(b)),
                                                 C.byref(ctl), C.byref(perf), _d(hist), C.c_int32(hist.size)))
        return x, perf, hist[:perf.nIterations + 1].copy()


    # ---- f1: b200PBiCG + DILU
    def dilu_calc_rd(self, diag, upper, lower):
        out = np.empty(self.N)
        self.lib.check(self.lib.c.b200_dilu_calc_reciprocal_d(self.h, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(out)))
        return out

    def dilu_precondition(self, rD, upper, lower, rA, transpose=False):
        out = np.empty(self.N)
        self.lib.check(self.lib.c.b200_dilu_precondition(self.h, _d(f64(rD)), _d(f64(upper)), _d(f64(lower)), _d(f64(rA)), _d(out),
                                                         C.c_int32(1 if transpose else 0)))
        return out

    def pbicg_solve(self, diag, upper, lower, x0, b, tol=1e-6, relTol=0.0, minIter=0, maxIter=1000, bou=None, intc=None):
        x = f64(x0).copy()
        ctl = SolverControls(tol, relTol, minIter, maxIter)
        perf = SolverPerf()
        hist = np.full(maxIter + 2, -1.0)
        bou_ = None if bou is None else f64(bou); int_ = None if intc is None else f64(intc)
        self.lib.check(self.lib.c.b200_pbicg_solve(self.h, _d(f64(diag)), _d(f64(upper)), _d(f64(lower)), _d(bou_), _d(int_), _d(x), _d(f64(b)),
                                                   C.byref(ctl), C.byref(perf), _d(hist), C.c_int32(hist.size)))
        return x, perf, hist[:perf.nIterations + 1].copy()


class PolyMesh:
    """Host polyMesh produced by the in-library generators (b200_polymesh_*)."""

    def __init__(self, lib, h):
        self.lib, self.h = lib, h
        n = [C.c_int32() for _ in range(5)]
        lib.check(lib.c.b200_polymesh_sizes(h, *[C.byref(x) for x in n]))
        self.nCells, self.nInternalFaces, self.nBoundaryFaces, self.nPatches, self.nPoints = [x.value for x in n]
        nF = self.nInternalFaces + self.nBoundaryFaces
        self.points = np.empty((self.nPoints, 3)); self.faces = np.empty((nF, 4), dtype=np.int32)
        self.owner = np.empty(nF, dtype=np.int32); self.neighbour = np.empty(self.nInternalFaces, dtype=np.int32)
        self.patchStart = np.empty(self.nPatches, dtype=np.int32); self.patchSize = np.empty(self.nPatches, dtype=np.int32)
        self.patchNeighbRank = np.empty(self.nPatches, dtype=np.int32)
        lib.check(lib.c.b200_polymesh_get(h, _d(self.points), _i(self.faces), _i(self.owner), _i(self.neighbour),
                                          _i(self.patchStart), _i(self.patchSize), _i(self.patchNeighbRank)))
        self.patchNames = [lib.c.b200_polymesh_patch_name(h, C.c_int32(p)).decode() for p in range(self.nPatches)]
        self._geom = None

    @property
    def patches(self):
        return [(self.patchNames[p], int(self.patchStart[p]), int(self.patchSize[p])) for p in range(self.nPatches)]

    def patch_slice(self, name):
        p = self.patchNames.index(name)
        s = int(self.patchStart[p]) - self.nInternalFaces
        return slice(s, s + int(self.patchSize[p]))

    def geometry(self):
        if self._geom is None:
            nF = self.nInternalFaces + self.nBoundaryFaces
            g = dict(Sf=np.empty((nF, 3)), magSf=np.empty(nF), Cf=np.empty((nF, 3)), C=np.empty((self.nCells, 3)),
                     V=np.empty(self.nCells), weights=np.empty(nF), deltaCoeffs=np.empty(nF))
            self.lib.check(self.lib.c.b200_polymesh_geometry(self.h, _d(g["Sf"]), _d(g["magSf"]), _d(g["Cf"]), _d(g["C"]),
                                                             _d(g["V"]), _d(g["weights"]), _d(g["deltaCoeffs"])))
            self._geom = g
        return self._geom

    def agglomerate(self, group):
        """polyhedral mesh: cells with equal group keys merge (b200_polymesh_agglomerate)"""
        h = C.c_void_p()
        g = np.ascontiguousarray(group, dtype=np.int64)
        self.lib.check(self.lib.c.b200_polymesh_agglomerate(self.h, g.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(h)))
        return PolyMesh(self.lib, h)

    def decompose_simple(self, n):
        proc = np.empty(self.nCells, dtype=np.int32)
        self.lib.check(self.lib.c.b200_decompose_simple(self.h, C.c_int32(n[0]), C.c_int32(n[1]), C.c_int32(n[2]), _i(proc)))
        return proc

    def submesh(self, proc, n_ranks, rank):
        h = C.c_void_p()
        proc = i32(proc)
        self.lib.check(self.lib.c.b200_polymesh_submesh(self.h, _i(proc), C.c_int32(n_ranks), C.c_int32(rank), C.byref(h)))
        sub = PolyMesh(self.lib, h)
        nF = sub.nInternalFaces + sub.nBoundaryFaces
        sub.cellMap = np.empty(sub.nCells, dtype=np.int32); sub.faceMap = np.empty(nF, dtype=np.int32)
        sub.faceFlip = np.empty(nF, dtype=np.int32)
        self.lib.check(self.lib.c.b200_polymesh_submesh_maps(h, _i(sub.cellMap), _i(sub.faceMap), _i(sub.faceFlip)))
        return sub

    def ldu(self, key=None):
        """device lduAddressing cache of this mesh (faceCells per patch taken from owner)."""
        F = self.nInternalFaces
        patches = [(self.owner[self.patchStart[p]:self.patchStart[p] + self.patchSize[p]], int(self.patchNeighbRank[p]))
                   for p in range(self.nPatches)]
        return self.lib.ldu_create(self.nCells, self.owner[:F], self.neighbour, patches, key)

    def device_mesh(self):
        h = C.c_void_p()
        self.lib.check(self.lib.c.b200_mesh_create_from_polymesh(self.h, C.byref(h)))
        return DeviceMesh(self.lib, h, self)

    def release(self):
        if self.h:
            self.lib.c.b200_polymesh_release(self.h)
            self.h = None


class DeviceMesh:
    """b200_mesh_t: lduAddressing + fvMesh geometry resident on the device."""

    def __init__(self, lib, h, pm):
        self.lib, self.h, self.pm = lib, h, pm
        self.N, self.F, self.B = pm.nCells, pm.nInternalFaces, pm.nBoundaryFaces
        self.ldu = Ldu(lib, C.c_void_p(lib.c.b200_mesh_ldu(h)), self.N, self.F, self.B)

    def release(self):
        if self.h:
            self.lib.check(self.lib.c.b200_mesh_release(self.h))
            self.h = None

    def bc(self, types, nc=1, refValue=None, refGrad=None, valueFraction=None):
        return BC(self.B, types, nc, refValue, refGrad, valueFraction)

    # ---- a13
    def interpolate(self, vf, vfb, nc=1):
        out = np.empty(nc * (self.F + self.B))
        self.lib.check(self.lib.c.b200_interpolate(self.h, _d(f64(vf).reshape(-1)), _d(f64(vfb).reshape(-1)), C.c_int32(nc), _d(out)))
        return out

    def sngrad(self, bc, vf, vfb):
        out = np.empty(self.F + self.B)
        self.lib.check(self.lib.c.b200_sngrad(self.h, bc.ref, _d(f64(vf)), _d(f64(vfb)), _d(out)))
        return out

    def surface_integrate(self, ssf):
        out = np.empty(self.N)
        self.lib.check(self.lib.c.b200_surface_integrate(self.h, _d(f64(ssf)), _d(out)))
        return out

    def grad_gauss(self, bc, vf, vfb):
        g = np.empty(3 * self.N); gb = np.empty(3 * self.B)
        self.lib.check(self.lib.c.b200_grad_gauss(self.h, bc.ref, _d(f64(vf)), _d(f64(vfb)), _d(g), _d(gb)))
        return g.reshape(-1, 3), gb.reshape(-1, 3)

    def reconstruct(self, ssf):
        out = np.empty(3 * self.N)
        self.lib.check(self.lib.c.b200_reconstruct(self.h, _d(f64(ssf)), _d(out)))
        return out.reshape(-1, 3)

    def non_orthogonality(self):
        """(degrees over this rank's internal faces, mesh.orthogonal() over all ranks)"""
        d = C.c_double(); o = C.c_int32()
        self.lib.check(self.lib.c.b200_mesh_non_orthogonality(self.h, C.byref(d), C.byref(o)))
        return d.value, bool(o.value)

    def non_orth_correction(self, vf, vfb, gammaf=None):
        out = np.empty(self.F + self.B)
        g = None if gammaf is None else f64(gammaf)
        self.lib.check(self.lib.c.b200_non_orth_correction(self.h, _d(g), _d(f64(vf)), _d(f64(vfb)), _d(out)))
        return out

    # ---- a7/a8
    def flux_limited(self, scheme, bc, faceFlux, vf, vfb):
        out = np.empty(self.F + self.B)
        self.lib.check(self.lib.c.b200_flux_limited(self.h, C.c_int32(scheme), bc.ref, _d(f64(faceFlux)), _d(f64(vf)), _d(f64(vfb)), _d(out)))
        return out

    # ---- a5/a6
    def mules_explicit_solve(self, bc, psi, psib, psi0, phi, phiPsi, deltaT, psiMax=1.0, psiMin=0.0):
        psi = f64(psi).copy(); phiPsi = f64(phiPsi).copy(); lam = np.empty(self.F + self.B)
        self.lib.check(self.lib.c.b200_mules_explicit_solve(self.h, bc.ref, _d(psi), _d(f64(psib)), _d(f64(psi0)), _d(f64(phi)),
                                                            _d(phiPsi), C.c_double(deltaT), C.c_double(psiMax), C.c_double(psiMin), _d(lam)))
        return psi, phiPsi, lam

    def mules_limiter(self, bc, psi, psib, psi0, phiBD, phiCorr, deltaT, psiMax=1.0, psiMin=0.0, nIter=3):
        lam = np.empty(self.F + self.B)
        self.lib.check(self.lib.c.b200_mules_limiter(self.h, bc.ref, _d(f64(psi)), _d(f64(psib)), _d(f64(psi0)), _d(f64(phiBD)),
                                                     _d(f64(phiCorr)), C.c_double(deltaT), C.c_double(psiMax), C.c_double(psiMin),
                                                     C.c_int32(nIter), _d(lam)))
        return lam

    # ---- a10/a11/a12
    def laplacian_assemble(self, bc, gammaf, pb, phi=None):
        diag = np.empty(self.N); upper = np.empty(self.F); source = np.empty(self.N); ic = np.empty(self.B); bcv = np.empty(self.B)
        phi_ = None if phi is None else f64(phi)
        self.lib.check(self.lib.c.b200_laplacian_assemble(self.h, bc.ref, _d(f64(gammaf)), _d(f64(pb)), _d(phi_), _d(diag), _d(upper),
                                                          _d(source), _d(ic), _d(bcv)))
        return diag, upper, source, ic, bcv

    def matrix_flux(self, upper, lower, ic, bcv, psi):
        out = np.empty(self.F + self.B)
        self.lib.check(self.lib.c.b200_matrix_flux(self.h, _d(f64(upper)), _d(f64(lower)), _d(f64(ic)), _d(f64(bcv)), _d(f64(psi)), _d(out)))
        return out

    def alpha_stats(self, a, ab):
        out = np.empty(3)
        self.lib.check(self.lib.c.b200_alpha_stats(self.h, _d(f64(a)), _d(f64(ab)), _d(out)))
        return out

    def region3d(self, controls, bcAlpha, bcU, bcPd, alpha1, U, pd):
        h = C.c_void_p()
        self.lib.check(self.lib.c.b200_region3d_create(self.h, C.byref(controls), bcAlpha.ref, bcU.ref, bcPd.ref, _d(f64(alpha1)),
                                                       _d(f64(U).reshape(-1)), _d(f64(pd)), C.byref(h)))
        return Region3D(self.lib, h, self, (bcAlpha, bcU, bcPd))


class Region3D:
    """Device-resident 3D region: replays region3d/solve3d.H on the GPU."""

    def __init__(self, lib, h, mesh, keep):
        self.lib, self.h, self.mesh, self._keep = lib, h, mesh, keep

    def release(self):
        if self.h:
            self.lib.check(self.lib.c.b200_region3d_release(self.h))
            self.h = None

    def correct_phi(self):
        self.lib.check(self.lib.c.b200_region3d_correct_phi(self.h))

    def step(self, n=1):
        self.lib.check(self.lib.c.b200_region3d_step(self.h, C.c_int32(n)))

    def step_host(self, alpha1, U, pd, phi, p):
        """end-to-end step on HOST buffers (in place)."""
        self.lib.check(self.lib.c.b200_region3d_step_host(self.h, _d(alpha1), _d(U), _d(pd), _d(phi), _d(p)))

    def fields(self):
        m = self.mesh
        a = np.empty(m.N); U = np.empty(3 * m.N); pd = np.empty(m.N); phi = np.empty(m.F + m.B); p = np.empty(m.N); rho = np.empty(m.N)
        self.lib.check(self.lib.c.b200_region3d_get_fields(self.h, _d(a), _d(U), _d(pd), _d(phi), _d(p), _d(rho)))
        return dict(alpha1=a, U=U.reshape(-1, 3), pd=pd, phi=phi, p=p, rho=rho)

    def set_fields(self, alpha1=None, U=None, pd=None, phi=None):
        a = None if alpha1 is None else f64(alpha1); u = None if U is None else f64(U).reshape(-1)
        q = None if pd is None else f64(pd); f = None if phi is None else f64(phi)
        self.lib.check(self.lib.c.b200_region3d_set_fields(self.h, _d(a), _d(u), _d(q), _d(f)))

    def set_patch_mixed(self, field, patch, refValue, refGrad, valueFraction):
        self.lib.check(self.lib.c.b200_region3d_set_patch_mixed(self.h, C.c_int32(field), C.c_int32(patch), _d(f64(refValue).reshape(-1)),
                                                                _d(f64(refGrad).reshape(-1)), _d(f64(valueFraction))))

    def patch_internal(self, field, patch):
        n = int(self.mesh.pm.patchSize[patch]) * (3 if field == 1 else 1)
        out = np.empty(n)
        self.lib.check(self.lib.c.b200_region3d_get_patch_internal(self.h, C.c_int32(field), C.c_int32(patch), _d(out)))
        return out

    def perfs(self):
        n = self.lib.c.b200_region3d_num_solves(self.h)
        arr = (SolverPerf * max(n, 1))()
        self.lib.check(self.lib.c.b200_region3d_get_perfs(self.h, arr, C.c_int32(n)))
        return [(arr[i].initialResidual, arr[i].finalResidual, arr[i].nIterations, arr[i].converged, arr[i].singular) for i in range(n)]

    def res_hist(self):
        n = self.lib.c.b200_region3d_get_res_hist(self.h, None, C.c_int32(0))
        out = np.empty(max(n, 0))
        if n > 0:
            self.lib.c.b200_region3d_get_res_hist(self.h, _d(out), C.c_int32(n))
        return out

    def alpha_stats(self):
        n = self.lib.c.b200_region3d_get_alpha_stats(self.h, None, C.c_int32(0))
        out = np.empty(max(n, 0))
        if n > 0:
            self.lib.c.b200_region3d_get_alpha_stats(self.h, _d(out), C.c_int32(n))
        return out.reshape(-1, 3)

    def cont_errs(self):
        out = np.empty(3)
        self.lib.check(self.lib.c.b200_region3d_get_cont_errs(self.h, _d(out)))
        return out

    def clear_log(self):
        self.lib.check(self.lib.c.b200_region3d_clear_log(self.h))

    def capture(self, on=True):
        self.lib.check(self.lib.c.b200_region3d_capture(self.h, C.c_int32(1 if on else 0)))

    def captured_solves(self):
        m = self.mesh; out = []
        for i in range(self.lib.c.b200_region3d_num_captured(self.h, C.c_int32(0))):
            d = dict(diag=np.empty(m.N), upper=np.empty(m.F), source=np.empty(m.N), x0=np.empty(m.N), bou=np.empty(m.B)); tr = np.empty(2)
            self.lib.check(self.lib.c.b200_region3d_get_captured_solve(self.h, C.c_int32(i), _d(d["diag"]), _d(d["upper"]), _d(d["source"]),
                                                                       _d(d["x0"]), _d(d["bou"]), _d(tr)))
            d["tol"], d["relTol"] = float(tr[0]), float(tr[1])
            out.append(d)
        return out

    def captured_mules(self):
        m = self.mesh; out = []
        for i in range(self.lib.c.b200_region3d_num_captured(self.h, C.c_int32(1))):
            d = dict(psi=np.empty(m.N), psib=np.empty(m.B), psi0=np.empty(m.N), phi=np.empty(m.F + m.B), phiPsi=np.empty(m.F + m.B))
            dt = C.c_double()
            self.lib.check(self.lib.c.b200_region3d_get_captured_mules(self.h, C.c_int32(i), _d(d["psi"]), _d(d["psib"]), _d(d["psi0"]),
                                                                       _d(d["phi"]), _d(d["phiPsi"]), C.byref(dt)))
            d["deltaT"] = dt.value
            out.append(d)
        return out

    def courant(self):
        """(meanCoNum, CoNum, velMag) of the current phi, reduced on the device"""
        out = np.empty(3)
        self.lib.check(self.lib.c.b200_region3d_get_courant(self.h, _d(out)))
        return out

    def restore(self, alpha1, U, pd, phi):
        """restart from saved fields: set_fields + rho + interface.correct()"""
        self.lib.check(self.lib.c.b200_region3d_restore(self.h, _d(f64(alpha1)), _d(f64(U).reshape(-1)), _d(f64(pd)), _d(f64(phi))))


_LIB = None


def checkerboard_groups(nx, ny, nz):
    """group keys for PolyMesh.agglomerate: the complete 2x2x2 blocks with (I + J + K) % 3 != 0 of an nx x ny x nz hex box merge into
    24-face super-cells, the other cells stay hexes -- the synthetic polyhedral mesh of BASELINE configs[3] (restated, and the resulting
    mesh compared bit by bit, by oracle/mesh_oracle.py block_groups / agglomerate)"""
    i = np.arange(nx, dtype=np.int64)[:, None, None]; j = np.arange(ny, dtype=np.int64)[None, :, None]; k = np.arange(nz, dtype=np.int64)[None, None, :]
    I, J, K = i // 2, j // 2, k // 2
    merged = (2 * I + 1 < nx) & (2 * J + 1 < ny) & (2 * K + 1 < nz) & ((I + J + K) % 3 != 0)
    cell = i + nx * (j + ny * k)
    key = np.where(merged, nx * ny * nz + (I + (nx // 2 + 1) * (J + (ny // 2 + 1) * K)), cell)
    out = np.empty(nx * ny * nz, dtype=np.int64)
    out[cell.ravel()] = key.ravel()
    return out


def load():
    """The product library (nvcc, sm_100a).  Raises when it is missing -- no fallback of any kind."""
    global _LIB
    if _LIB is None:
        _LIB = B200Lib(LIB_PATH)
    return _LIB
