// api.cu -- the extern "C" boundary declared in include/b200foam.h.  Host pointers in, host pointers out; no C++
// exception leaves this file.
#include <map>
#include <memory>
#include "comm.cuh"
#include "device_mesh.cuh"
#include "pcg.cuh"
#include "region3d.cuh"

using namespace b200;

static thread_local std::string g_lastError;
static std::map<const void*, DeviceMesh*> g_registry;

// FNV-1a over the addressing a cache entry was built from: a hit whose arrays hash differently (mesh motion / topology change rebuilt
// the lduAddressing at the same address) is rebuilt instead of being trusted
static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint64_t addressingHash(int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper, int32_t nPatches,
                               const b200_patch_desc* patches) {
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, &nCells, 4); h = fnv1a(h, &nFaces, 4); h = fnv1a(h, &nPatches, 4);
    if (nFaces) { h = fnv1a(h, lower, 4 * (size_t)nFaces); h = fnv1a(h, upper, 4 * (size_t)nFaces); }
    for (int p = 0; p < nPatches; p++) {
        h = fnv1a(h, &patches[p].nFaces, 4); h = fnv1a(h, &patches[p].neighbRank, 4);
        if (patches[p].nFaces) h = fnv1a(h, patches[p].faceCells, 4 * (size_t)patches[p].nFaces);
    }
    return h;
}

#define API_BEGIN try {
#define API_END                                                      \
    return B200_OK;                                                  \
    }                                                                \
    catch (const b200::Error& e) { g_lastError = e.what(); return e.code; } \
    catch (const std::exception& e) { g_lastError = e.what(); return B200_ERR_STATE; }

struct b200_polymesh { PolyMesh pm; };


static DeviceMesh* asMesh(b200_ldu_t* l) { return reinterpret_cast<DeviceMesh*>(l); }
static DeviceMesh* asMesh(b200_mesh_t* l) { return reinterpret_cast<DeviceMesh*>(l); }
static const DeviceMesh* asMesh(const b200_ldu_t* l) { return reinterpret_cast<const DeviceMesh*>(l); }

static void requireSymmetric(const double* upper, const double* lower) {
    B200_REQUIRE(lower == nullptr || lower == upper, B200_ERR_UNSUPPORTED,
                 "b200PCG: asymmetric matrix (PCG is registered in the symMatrix table only)");
}

extern "C" {

const char* b200_last_error(void) { return g_lastError.c_str(); }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int b200_init(int device, int rank, int nRanks, const void* ncclId, size_t ncclIdBytes) {
    API_BEGIN
    Runtime& r = rt();
    B200_REQUIRE(!r.initialised || (r.device == device && r.rank == rank && r.nRanks == nRanks), B200_ERR_STATE,
                 "b200_init called twice with different arguments");
    if (!r.initialised) {
        r.device = device; r.rank = rank; r.nRanks = nRanks;
        requireDevice();
        commInit(rank, nRanks, ncclId, ncclIdBytes);
    }
    API_END
}
int b200_nccl_unique_id(void* out128) {
    API_BEGIN
    commUniqueId(out128);
    API_END
}
int b200_finalize(void) {
    API_BEGIN
    for (auto& kv : g_registry) delete kv.second;
    g_registry.clear();
    commFinalize();
    API_END
}
long long b200_kernel_launch_count(void) { return rt().launches; }
int b200_set_option(const char* name, double value) {
    API_BEGIN
    B200_REQUIRE(name, B200_ERR_ARG, "b200_set_option: null name");
    rt().opts[name] = value;
    API_END
}
int b200_timers_reset(void) {
    API_BEGIN
    Timers& t = timers();
    t.collect();
    t.enabled = true;
    for (int i = 0; i < T_COUNT; i++) { t.ms[i] = 0; t.calls[i] = 0; }
    API_END
}
int b200_timers_get(double* outMs, long long* outCalls, int cap) {
    API_BEGIN
    Timers& t = timers();
    t.collect();
    for (int i = 0; i < T_COUNT && i < cap; i++) { outMs[i] = t.ms[i]; if (outCalls) outCalls[i] = t.calls[i]; }
    API_END
}
const char* b200_timer_name(int i) { return timerName(i); }

// ---------------------------------------------------------------------------------------------- ldu cache
int b200_ldu_get_or_create(const void* key, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                           int32_t nPatches, const b200_patch_desc* patches, b200_ldu_t** out) {
    API_BEGIN
    requireDevice();
    B200_REQUIRE(out && nCells >= 0 && nFaces >= 0 && (nFaces == 0 || (lower && upper)), B200_ERR_ARG, "b200_ldu_get_or_create: bad arguments");
    const uint64_t hash = addressingHash(nCells, nFaces, lower, upper, nPatches, patches);
    if (key) {
        auto it = g_registry.find(key);
        if (it != g_registry.end()) {
            if (it->second->addrHash == hash) {
                *out = reinterpret_cast<b200_ldu_t*>(it->second);
                return B200_OK;
            }
            delete it->second;          // same address, different addressing: the mesh changed under the key
            g_registry.erase(it);
        }
    }
    std::vector<Patch> pt; std::vector<int32_t> fc;
    int32_t start = nFaces;
    for (int p = 0; p < nPatches; p++) {
        B200_REQUIRE(patches[p].nFaces == 0 || patches[p].faceCells, B200_ERR_ARG, "patch without faceCells");
        Patch q; q.name = "patch" + std::to_string(p); q.start = start; q.size = patches[p].nFaces; q.neighbRank = patches[p].neighbRank;
        fc.insert(fc.end(), patches[p].faceCells, patches[p].faceCells + patches[p].nFaces);
        start += q.size; pt.push_back(q);
    }
    std::unique_ptr<DeviceMesh> m(new DeviceMesh());
    m->build(nCells, nFaces, lower, upper, pt, fc);
    m->key = key; m->addrHash = hash;
    DeviceMesh* raw = m.release();
    if (key) g_registry[key] = raw;
    *out = reinterpret_cast<b200_ldu_t*>(raw);
    API_END
}
int b200_ldu_release(b200_ldu_t* ldu) {
    API_BEGIN
    DeviceMesh* m = asMesh(ldu);
    if (m) { if (m->key) g_registry.erase(m->key); delete m; }
    API_END
}
int b200_ldu_invalidate(const void* key) {
    API_BEGIN
    auto it = g_registry.find(key);
    if (it != g_registry.end()) { delete it->second; g_registry.erase(it); }
    API_END
}
int b200_ldu_get_renumbering(const b200_ldu_t* ldu, int32_t* rowCell, int32_t* cellLevel, int32_t* nLevels) {
    API_BEGIN
    const DeviceMesh* m = asMesh(ldu);
    if (rowCell) memcpy(rowCell, m->rowCell.data(), 4 * (size_t)m->N);
    if (cellLevel) memcpy(cellLevel, m->level.data(), 4 * (size_t)m->N);
    if (nLevels) *nLevels = m->nLevels;
    API_END
}
int b200_ldu_get_tiling(const b200_ldu_t* ldu, int32_t* bandLevels, int32_t* tileRows, int32_t* nTiles, int32_t* tileStart) {
    API_BEGIN
    const DeviceMesh* m = asMesh(ldu);
    if (bandLevels) *bandLevels = m->plan.bandLevels;
    if (tileRows) *tileRows = m->plan.maxRows;
    if (nTiles) *nTiles = m->nTiles;
    if (tileStart) memcpy(tileStart, m->plan.tileStart.data(), 4 * m->plan.tileStart.size());
    API_END
}
int b200_debug_get_sweep_trace(const b200_ldu_t* ldu, unsigned long long* out, int32_t cap, int32_t* tileLevel) {
    API_BEGIN
    const DeviceMesh* m = asMesh(ldu);
    size_t n = std::min<size_t>((size_t)cap, 8 * (size_t)m->nTiles);
    syncStream();
    B200_CHECK_CUDA(cudaMemcpy(out, m->dTrace.p, 8 * n, cudaMemcpyDeviceToHost));
    if (tileLevel) memcpy(tileLevel, m->plan.tileLevel.data(), 4 * (size_t)m->nTiles);
    API_END
}
int b200_ldu_get_addressing(const b200_ldu_t* ldu, int32_t* ownerStart, int32_t* losort, int32_t* losortStart) {
    API_BEGIN
    const DeviceMesh* m = asMesh(ldu);
    if (ownerStart) memcpy(ownerStart, m->ownerStart.data(), 4 * (size_t)(m->N + 1));
    if (losort) memcpy(losort, m->losort.data(), 4 * (size_t)m->F);
    if (losortStart) memcpy(losortStart, m->losortStart.data(), 4 * (size_t)(m->N + 1));
    API_END
}

// ---------------------------------------------------------------------------------------------- a1..a4
static void stageMatrix(DeviceMesh& m, PcgWorkspace& w, const double* diag, const double* upper, const double* bou) {
    w.diag.ensure(m.N); w.upS.ensure(m.Fs + 1);
    m.uploadCells(diag, w.diag);
    m.uploadFaces(upper, w.upS, false);
    if (bou && m.B) { w.bou.ensure(m.B); m.uploadBoundary(bou, w.bou); }
}

int b200_amul(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, const double* psi,
              double* Apsi, const double* bouCoeffs, const double* psiNbr) {
    API_BEGIN
    requireDevice();
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, bouCoeffs);
    DevBuf<double> loS;
    const double* lo = w.upS;
    if (lower && lower != upper) { loS.alloc(m.Fs + 1); m.uploadFaces(lower, loS, false); lo = loS; }
    w.x.ensure(m.N); w.b.ensure(m.N);
    m.uploadCells(psi, w.x);
    if (bouCoeffs && psiNbr && m.B) {  // caller-supplied neighbour values (single process, explicit halo)
        { ScopedTimer t(T_AMUL);
          if (m.N) B200_LAUNCH_AMUL(0, false, m.view(), (const double*)w.diag.p, (const double*)w.upS.p, lo,
                               (const double*)w.x.p, w.b.p, (PcgScalars*)nullptr, (double*)nullptr, (unsigned*)nullptr); }
        m.uploadBoundary(psiNbr, w.recvBuf);
        B200_LAUNCH(k_interface_update, gridFor(m.nBRows), BLOCK, rt().stream, m.view(), (const double*)w.bou.p, (const double*)w.recvBuf.p, w.b.p,
                    m.nBRows, (const int*)w.bRowOfSlot.p);
    } else {
        amulDevice(m, w.diag, w.upS, lo, w.x, w.b, bouCoeffs ? w.bou.p : nullptr);
    }
    m.downloadCells(w.b, Apsi);
    API_END
}

int b200_sumA(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, double* sumA,
              const double* bouCoeffs) {
    API_BEGIN
    requireDevice();
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, bouCoeffs);
    DevBuf<double> loS;
    const double* lo = w.upS;
    if (lower && lower != upper) { loS.alloc(m.Fs + 1); m.uploadFaces(lower, loS, false); lo = loS; }
    w.b.ensure(m.N);
    if (m.N) B200_LAUNCH(k_sumA, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)w.diag.p, (const double*)w.upS.p, lo,
                         bouCoeffs ? (const double*)w.bou.p : (const double*)nullptr, w.b.p);
    m.downloadCells(w.b, sumA);
    API_END
}

int b200_dic_calc_reciprocal_d(b200_ldu_t* ldu, const double* diag, const double* upper, double* rD) {
    API_BEGIN
    requireDevice();
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, nullptr);
    dicCalcReciprocalD(m, w.diag, w.upS, w.rD);
    m.downloadCells(w.rD, rD);
    API_END
}

int b200_dic_precondition(b200_ldu_t* ldu, const double* rD, const double* upper, const double* rA, double* wA) {
    API_BEGIN
    requireDevice();
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    w.upS.ensure(m.Fs + 1);
    m.uploadFaces(upper, w.upS, false);
    m.uploadCells(rD, w.rD);
    m.uploadCells(rA, w.rA);
    dicPrecondition(m, w.rD, w.upS, w.rA, w.wA);
    m.downloadCells(w.wA, wA);
    API_END
}

int b200_pcg_solve(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower,
                   const double* bouCoeffs, double* x, const double* b, const b200_solver_controls* ctl,
                   b200_solver_perf* perf, double* resHist, int32_t histCap) {
    API_BEGIN
    requireDevice();
    requireSymmetric(upper, lower);
    B200_REQUIRE(ctl && perf && x && b, B200_ERR_ARG, "b200_pcg_solve: null argument");
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, bouCoeffs);
    w.x.ensure(m.N); w.b.ensure(m.N);
    m.uploadCells(x, w.x);
    m.uploadCells(b, w.b);
    std::vector<double> hist;
    *perf = pcgSolveDevice(m, w.diag, w.upS, bouCoeffs ? w.bou.p : nullptr, w.x, w.b, *ctl, resHist ? &hist : nullptr);
    m.downloadCells(w.x, x);
    if (resHist) for (int i = 0; i < histCap && i < (int)hist.size(); i++) resHist[i] = hist[i];
    API_END
}

// ---------------------------------------------------------------------------------------------- f1: b200PBiCG + DILU
static const double* stageLower(DeviceMesh& m, PcgWorkspace& w, const double* lower) {
    w.loS.ensure(m.Fs + 1);
    m.uploadFaces(lower, w.loS, false);
    return w.loS.p;
}
int b200_dilu_calc_reciprocal_d(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, double* rD) {
    API_BEGIN
    requireDevice();
    B200_REQUIRE(ldu && diag && upper && lower && rD, B200_ERR_ARG, "b200_dilu_calc_reciprocal_d: null argument");
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, nullptr);
    const double* lo = stageLower(m, w, lower);
    diluCalcReciprocalD(m, w.diag, w.upS, lo, w.rD);
    m.downloadCells(w.rD, rD);
    API_END
}
int b200_dilu_precondition(b200_ldu_t* ldu, const double* rD, const double* upper, const double* lower, const double* rA, double* wA,
                           int32_t transpose) {
    API_BEGIN
    requireDevice();
    B200_REQUIRE(ldu && rD && upper && lower && rA && wA, B200_ERR_ARG, "b200_dilu_precondition: null argument");
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    w.upS.ensure(m.Fs + 1);
    m.uploadFaces(upper, w.upS, false);
    const double* lo = stageLower(m, w, lower);
    m.uploadCells(rD, w.rD);
    m.uploadCells(rA, w.rA);
    diluPrecondition(m, w.rD, w.upS, lo, w.rA, w.wA, transpose != 0);
    m.downloadCells(w.wA, wA);
    API_END
}
int b200_pbicg_solve(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, const double* bouCoeffs,
                     const double* intCoeffs, double* x, const double* b, const b200_solver_controls* ctl, b200_solver_perf* perf,
                     double* resHist, int32_t histCap) {
    API_BEGIN
    requireDevice();
    B200_REQUIRE(ldu && diag && upper && ctl && perf && x && b, B200_ERR_ARG, "b200_pbicg_solve: null argument");
    DeviceMesh& m = *asMesh(ldu);
    PcgWorkspace& w = pcgWorkspace(m);
    stageMatrix(m, w, diag, upper, bouCoeffs);
    const double* lo = (lower && lower != upper) ? stageLower(m, w, lower) : (const double*)w.upS.p;
    if (intCoeffs && m.B) { w.intc.ensure(m.B); m.uploadBoundary(intCoeffs, w.intc); }
    w.x.ensure(m.N); w.b.ensure(m.N);
    m.uploadCells(x, w.x);
    m.uploadCells(b, w.b);
    std::vector<double> hist;
    *perf = pbicgSolveDevice(m, w.diag, w.upS, lo, bouCoeffs ? w.bou.p : nullptr, intCoeffs ? w.intc.p : nullptr, w.x, w.b, *ctl, resHist ? &hist : nullptr);
    m.downloadCells(w.x, x);
    if (resHist) for (int i = 0; i < histCap && i < (int)hist.size(); i++) resHist[i] = hist[i];
    API_END
}

// ---------------------------------------------------------------------------------------------- mesh
int b200_mesh_create(const b200_mesh_desc* d, b200_mesh_t** out) {
    API_BEGIN
    requireDevice();
    B200_REQUIRE(d && out, B200_ERR_ARG, "b200_mesh_create: null argument");
    std::vector<Patch> pt; std::vector<int32_t> fc(d->owner + d->nInternalFaces, d->owner + d->nInternalFaces + d->nBoundaryFaces);
    for (int p = 0; p < d->nPatches; p++) {
        Patch q; q.name = "patch" + std::to_string(p); q.start = d->patchStart[p]; q.size = d->patchSize[p];
        q.neighbRank = d->patchNeighbRank ? d->patchNeighbRank[p] : -1;
        pt.push_back(q);
    }
    std::unique_ptr<DeviceMesh> m(new DeviceMesh());
    m->build(d->nCells, d->nInternalFaces, d->owner, d->neighbour, pt, fc);
    m->setGeometry(d->Sf, d->magSf, d->Cf, d->C, d->V, d->weights, d->deltaCoeffs);
    m->coupleGeometry();   // collective when the mesh has processor patches
    *out = reinterpret_cast<b200_mesh_t*>(m.release());
    API_END
}
int b200_mesh_release(b200_mesh_t* mesh) {
    API_BEGIN
    delete asMesh(mesh);
    API_END
}
b200_ldu_t* b200_mesh_ldu(b200_mesh_t* mesh) { return reinterpret_cast<b200_ldu_t*>(mesh); }

// ---------------------------------------------------------------------------------------------- host mesh tools
int b200_polymesh_box(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz, b200_polymesh_t** out) {
    API_BEGIN
    B200_REQUIRE(nx > 0 && ny > 0 && nz > 0 && out, B200_ERR_ARG, "b200_polymesh_box: bad arguments");
    *out = new b200_polymesh{makeBox(nx, ny, nz, lx, ly, lz)};
    API_END
}
int b200_polymesh_box_weir(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz, double x0, double x1,
                           double zfrac, b200_polymesh_t** out) {
    API_BEGIN
    B200_REQUIRE(nx > 0 && ny > 0 && nz > 0 && out, B200_ERR_ARG, "b200_polymesh_box_weir: bad arguments");
    PolyMesh box = makeBox(nx, ny, nz, lx, ly, lz);
    *out = new b200_polymesh{subsetMesh(box, weirKeepMask(nx, ny, nz, x0, x1, zfrac), "weir")};
    API_END
}
int b200_polymesh_jitter(b200_polymesh_t* pm, double frac, double dx, double dy, double dz, uint64_t seed) {
    API_BEGIN
    B200_REQUIRE(pm && frac >= 0.0 && frac < 0.5, B200_ERR_ARG, "b200_polymesh_jitter: frac must be in [0, 0.5)");
    jitterPoints(pm->pm, frac, dx, dy, dz, seed);
    API_END
}
int b200_polymesh_shear(b200_polymesh_t* pm, double sxy, double sxz, double syz) {
    API_BEGIN
    B200_REQUIRE(pm, B200_ERR_ARG, "b200_polymesh_shear: null mesh");
    shearPoints(pm->pm, sxy, sxz, syz);
    API_END
}
int b200_polymesh_agglomerate(const b200_polymesh_t* pm, const int64_t* group, b200_polymesh_t** out) {
    API_BEGIN
    B200_REQUIRE(pm && group && out, B200_ERR_ARG, "b200_polymesh_agglomerate: null argument");
    *out = new b200_polymesh{agglomerate(pm->pm, group)};
    API_END
}
int b200_polymesh_release(b200_polymesh_t* pm) { delete pm; return B200_OK; }
int b200_polymesh_sizes(const b200_polymesh_t* pm, int32_t* nCells, int32_t* nIF, int32_t* nBF, int32_t* nPatches, int32_t* nPoints) {
    API_BEGIN
    if (nCells) *nCells = pm->pm.nCells;
    if (nIF) *nIF = pm->pm.nInternalFaces();
    if (nBF) *nBF = pm->pm.nBoundaryFaces();
    if (nPatches) *nPatches = (int32_t)pm->pm.patches.size();
    if (nPoints) *nPoints = (int32_t)(pm->pm.points.size() / 3);
    API_END
}
int b200_polymesh_get(const b200_polymesh_t* pm, double* points, int32_t* faces4, int32_t* owner, int32_t* neighbour,
                      int32_t* patchStart, int32_t* patchSize, int32_t* patchNeighbRank) {
    API_BEGIN
    const PolyMesh& m = pm->pm;
    if (points) memcpy(points, m.points.data(), 8 * m.points.size());
    if (faces4) memcpy(faces4, m.faces.data(), 16 * m.faces.size());
    if (owner) memcpy(owner, m.owner.data(), 4 * m.owner.size());
    if (neighbour) memcpy(neighbour, m.neighbour.data(), 4 * m.neighbour.size());
    for (size_t p = 0; p < m.patches.size(); p++) {
        if (patchStart) patchStart[p] = m.patches[p].start;
        if (patchSize) patchSize[p] = m.patches[p].size;
        if (patchNeighbRank) patchNeighbRank[p] = m.patches[p].neighbRank;
    }
    API_END
}
const char* b200_polymesh_patch_name(const b200_polymesh_t* pm, int32_t patch) {
    if (!pm || patch < 0 || patch >= (int)pm->pm.patches.size()) return "";
    return pm->pm.patches[patch].name.c_str();
}
int b200_polymesh_geometry(const b200_polymesh_t* pm, double* Sf, double* magSf, double* Cf, double* C, double* V,
                           double* weights, double* deltaCoeffs) {
    API_BEGIN
    Geometry g = computeGeometry(pm->pm);
    if (Sf) memcpy(Sf, g.Sf.data(), 8 * g.Sf.size());
    if (magSf) memcpy(magSf, g.magSf.data(), 8 * g.magSf.size());
    if (Cf) memcpy(Cf, g.Cf.data(), 8 * g.Cf.size());
    if (C) memcpy(C, g.C.data(), 8 * g.C.size());
    if (V) memcpy(V, g.V.data(), 8 * g.V.size());
    if (weights) memcpy(weights, g.weights.data(), 8 * g.weights.size());
    if (deltaCoeffs) memcpy(deltaCoeffs, g.deltaCoeffs.data(), 8 * g.deltaCoeffs.size());
    API_END
}
int b200_decompose_simple(const b200_polymesh_t* pm, int32_t npx, int32_t npy, int32_t npz, int32_t* procOfCell) {
    API_BEGIN
    B200_REQUIRE(npx > 0 && npy > 0 && npz > 0 && procOfCell, B200_ERR_ARG, "b200_decompose_simple: bad arguments");
    Geometry g = computeGeometry(pm->pm);
    std::vector<int32_t> proc = decomposeSimple(g.C, pm->pm.nCells, npx, npy, npz);
    memcpy(procOfCell, proc.data(), 4 * proc.size());
    API_END
}
int b200_polymesh_submesh(const b200_polymesh_t* pm, const int32_t* procOfCell, int32_t nRanks, int32_t rank,
                          b200_polymesh_t** out) {
    API_BEGIN
    std::vector<int32_t> proc(procOfCell, procOfCell + pm->pm.nCells);
    *out = new b200_polymesh{subMesh(pm->pm, proc, nRanks, rank)};
    API_END
}
int b200_polymesh_submesh_maps(const b200_polymesh_t* sub, int32_t* cellMap, int32_t* faceMap, int32_t* faceFlip) {
    API_BEGIN
    const PolyMesh& m = sub->pm;
    if (cellMap) memcpy(cellMap, m.cellMap.data(), 4 * m.cellMap.size());
    if (faceMap) memcpy(faceMap, m.faceMap.data(), 4 * m.faceMap.size());
    if (faceFlip) for (size_t i = 0; i < m.faceFlip.size(); i++) faceFlip[i] = m.faceFlip[i];
    API_END
}
int b200_mesh_create_from_polymesh(const b200_polymesh_t* pm, b200_mesh_t** out) {
    API_BEGIN
    requireDevice();
    const PolyMesh& m = pm->pm;
    Geometry g = computeGeometry(m);
    std::vector<int32_t> ps, pz, pr;
    for (const Patch& p : m.patches) { ps.push_back(p.start); pz.push_back(p.size); pr.push_back(p.neighbRank); }
    b200_mesh_desc d;
    d.nCells = m.nCells; d.nInternalFaces = m.nInternalFaces(); d.nBoundaryFaces = m.nBoundaryFaces(); d.nPatches = (int32_t)m.patches.size();
    d.owner = m.owner.data(); d.neighbour = m.neighbour.data(); d.patchStart = ps.data(); d.patchSize = pz.data(); d.patchNeighbRank = pr.data();
    d.Sf = g.Sf.data(); d.magSf = g.magSf.data(); d.Cf = g.Cf.data(); d.C = g.C.data(); d.V = g.V.data();
    d.weights = g.weights.data(); d.deltaCoeffs = g.deltaCoeffs.data();
    int rc = b200_mesh_create(&d, out);
    if (rc != B200_OK) return rc;
    API_END
}


// ---------------------------------------------------------------------------------------------- fv building blocks
namespace {
struct Stage {   // host-pointer API staging on the mesh's cached buffers
    DeviceMesh& m; FvWorkspace& w; int next = 0;
    explicit Stage(DeviceMesh& m_) : m(m_), w(fvWorkspace(m_)) {}
    double* buf(size_t n) { B200_REQUIRE(next < 16, B200_ERR_STATE, "staging exhausted"); DevBuf<double>& b = w.stg[next++]; b.ensure(n ? n : 1); return b.p; }
    double* cells(const double* h) { double* d = buf(m.N); m.uploadCells(h, d); return d; }
    double* faces(const double* h) { double* d = buf(m.nFaceField()); m.uploadFaces(h, d, true); return d; }
    double* facesInt(const double* h) { double* d = buf(m.Fs + 1); m.uploadFaces(h, d, false); return d; }
    double* bnd(const double* h, int nc = 1) { double* d = buf((size_t)nc * m.B); m.uploadBoundary(h, d, nc); return d; }
};
DeviceMesh& geoMesh(b200_mesh_t* mesh) {
    requireDevice();
    B200_REQUIRE(mesh, B200_ERR_ARG, "null mesh");
    DeviceMesh& m = *asMesh(mesh);
    B200_REQUIRE(m.hasGeometry, B200_ERR_ARG, "mesh has no geometry");
    return m;
}
int nmaxOf(const DeviceMesh& m) { return std::max(std::max(m.N, m.B), 1); }
}  // namespace

int b200_mesh_non_orthogonality(b200_mesh_t* mesh, double* degrees, int32_t* orthogonal) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    if (degrees) *degrees = nonOrthogonalityDegrees(m.nonOrthSum[0], m.nonOrthSum[1]);   // this rank's faces
    if (orthogonal) *orthogonal = m.orthogonal ? 1 : 0;                                    // global
    API_END
}
int b200_non_orth_correction(b200_mesh_t* mesh, const double* gammaf, const double* vf, const double* vfBoundary, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    B200_REQUIRE(vf && vfBoundary && out, B200_ERR_ARG, "b200_non_orth_correction: null argument");
    Stage s(m);
    FvWorkspace& w = s.w;
    double* d = s.cells(vf); double* db = s.bnd(vfBoundary); double* g = gammaf ? s.faces(gammaf) : nullptr; double* o = s.buf(m.nFaceField());
    if (m.N) B200_LAUNCH(k_grad_scalar, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)d, (const double*)db, w.gx.p, w.gy.p, w.gz.p);
    if (rt().nRanks > 1 && meshHasProcessorPatches(m)) {
        const double* src[3] = {w.gx.p, w.gy.p, w.gz.p};
        for (int k = 0; k < 3; k++) { haloPack(m, src[k], w.haloSend.p); haloExchange(m, w.haloSend.p, w.gb.p + (size_t)k * m.B); }
    }
    B200_LAUNCH(k_nonorth_ffc, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), (const double*)g, (const double*)w.gx.p, (const double*)w.gy.p,
                (const double*)w.gz.p, (const double*)w.gb.p, o);
    m.downloadFaces(o, out, true);
    API_END
}
int b200_interpolate(b200_mesh_t* mesh, const double* vf, const double* vfBoundary, int32_t nCmpt, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    B200_REQUIRE(nCmpt == 1 || nCmpt == 3, B200_ERR_ARG, "b200_interpolate: nCmpt must be 1 or 3");
    size_t nF = (size_t)m.F + m.B;
    std::vector<double> hc(m.N), hb(m.B), ho(nF);
    for (int k = 0; k < nCmpt; k++) {
        Stage s(m);
        for (int i = 0; i < m.N; i++) hc[i] = vf[(size_t)nCmpt * i + k];
        for (int i = 0; i < m.B; i++) hb[i] = vfBoundary[(size_t)nCmpt * i + k];
        double* d = s.cells(hc.data()); double* db = s.bnd(hb.data()); double* o = s.buf(m.nFaceField());
        B200_LAUNCH(k_interpolate, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), (const double*)d, (const double*)db, o);
        m.downloadFaces(o, ho.data(), true);
        for (size_t i = 0; i < nF; i++) out[(size_t)nCmpt * i + k] = ho[i];
    }
    API_END
}
int b200_sngrad(b200_mesh_t* mesh, const b200_bc* bc, const double* vf, const double* vfBoundary, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    DevBC dbc; dbc.upload(m, bc, 1);
    Stage s(m);
    double* d = s.cells(vf); double* db = s.bnd(vfBoundary); double* o = s.buf(m.nFaceField());
    B200_LAUNCH(k_sngrad, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), dbc.view(), (const double*)d, (const double*)db, o);
    m.downloadFaces(o, out, true);
    API_END
}
int b200_surface_integrate(b200_mesh_t* mesh, const double* ssf, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    Stage s(m);
    double* f = s.faces(ssf); double* o = s.buf(m.N);
    if (m.N) B200_LAUNCH(k_surface_integrate, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)f, o, 1.0);
    m.downloadCells(o, out);
    API_END
}
int b200_grad_gauss(b200_mesh_t* mesh, const b200_bc* bc, const double* vf, const double* vfBoundary, double* grad, double* gradBoundary) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    DevBC dbc; dbc.upload(m, bc, 1);
    Stage s(m);
    FvWorkspace& w = s.w;
    double* d = s.cells(vf); double* db = s.bnd(vfBoundary);
    if (m.N) B200_LAUNCH(k_grad_scalar, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)d, (const double*)db, w.gx.p, w.gy.p, w.gz.p);
    m.downloadCellVec(w.gx, w.gy, w.gz, grad);
    if (gradBoundary && m.B) {
        B200_LAUNCH(k_grad_boundary, gridFor(m.B), BLOCK, rt().stream, m.view(), dbc.view(), (const double*)d, (const double*)db, (const double*)w.gx.p,
                    (const double*)w.gy.p, (const double*)w.gz.p, w.gb.p);
        double* aos = s.buf(3 * (size_t)m.B);
        B200_LAUNCH(k_soa_to_aos, gridFor(m.B), BLOCK, rt().stream, (const double*)w.gb.p, aos, m.B, 3);
        m.downloadBoundary(aos, gradBoundary, 3);
    }
    API_END
}
int b200_reconstruct(b200_mesh_t* mesh, const double* ssf, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    Stage s(m);
    double* f = s.faces(ssf); double* ox = s.buf(m.N); double* oy = s.buf(m.N); double* oz = s.buf(m.N);
    const double* invT = reconTensor(m);
    if (m.N) B200_LAUNCH(k_reconstruct<0>, gridFor(m.N), BLOCK, rt().stream, m.view(), invT, (const double*)f, (const double*)nullptr,
                         (const double*)nullptr, (const double*)nullptr, ox, oy, oz);
    m.downloadCellVec(ox, oy, oz, out);
    API_END
}
int b200_flux_limited(b200_mesh_t* mesh, int32_t scheme, const b200_bc* bc, const double* faceFlux, const double* vf,
                      const double* vfBoundary, double* out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    B200_REQUIRE(scheme >= 0 && scheme <= 3, B200_ERR_ARG, "b200_flux_limited: unknown scheme");
    (void)bc;
    Stage s(m);
    FvWorkspace& w = s.w;
    double* ff = s.faces(faceFlux); double* d = s.cells(vf); double* db = s.bnd(vfBoundary); double* o = s.buf(m.nFaceField());
    ScopedTimer t(T_FLUX);
    if (scheme == B200_SCHEME_VANLEER && m.N)
        B200_LAUNCH(k_grad_scalar, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)d, (const double*)db, w.gx.p, w.gy.p, w.gz.p);
    B200_LAUNCH(k_flux_limited, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), (int)scheme, (const double*)ff, (const double*)d, (const double*)db,
                (const double*)w.gx.p, (const double*)w.gy.p, (const double*)w.gz.p, (const double*)w.gb.p, o);
    m.downloadFaces(o, out, true);
    API_END
}
int b200_mules_explicit_solve(b200_mesh_t* mesh, const b200_bc* bc, double* psi, const double* psiBoundary, const double* psi0,
                              const double* phi, double* phiPsi, double deltaT, double psiMax, double psiMin, double* lambdaOut) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    DevBC dbc; dbc.upload(m, bc, 1);
    Stage s(m);
    double* dpsi = s.cells(psi); double* dpsib = s.bnd(psiBoundary); double* dpsi0 = s.cells(psi0);
    double* dphi = s.faces(phi); double* dphiPsi = s.faces(phiPsi);
    double* dlam = lambdaOut ? s.buf(m.nFaceField()) : nullptr;
    mulesExplicitDevice(m, dbc.view(), dpsi, dpsib, dpsi0, dphi, dphiPsi, deltaT, psiMax, psiMin, (int)rt().opt("mules_include_own", 0), 3, dlam);
    m.downloadCells(dpsi, psi);
    m.downloadFaces(dphiPsi, phiPsi, true);
    if (lambdaOut) m.downloadFaces(dlam, lambdaOut, true);
    API_END
}
int b200_mules_limiter(b200_mesh_t* mesh, const b200_bc* bc, const double* psi, const double* psiBoundary, const double* psi0,
                       const double* phiBD, const double* phiCorr, double deltaT, double psiMax, double psiMin,
                       int32_t nLimiterIter, double* lambda) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    B200_REQUIRE(nLimiterIter >= 0 && lambda, B200_ERR_ARG, "b200_mules_limiter: bad arguments");
    DevBC dbc; dbc.upload(m, bc, 1);
    Stage s(m);
    double* dpsi = s.cells(psi); double* dpsib = s.bnd(psiBoundary); double* dpsi0 = s.cells(psi0);
    double* dbd = s.faces(phiBD); double* dcorr = s.faces(phiCorr); double* dlam = s.buf(m.nFaceField());
    mulesLimiterDevice(m, dbc.view(), dpsi, dpsib, dpsi0, dbd, dcorr, deltaT, psiMax, psiMin, (int)rt().opt("mules_include_own", 0), nLimiterIter, dlam);
    m.downloadFaces(dlam, lambda, true);
    API_END
}
int b200_laplacian_assemble(b200_mesh_t* mesh, const b200_bc* bc, const double* gammaf, const double* psiBoundary, const double* phi,
                            double* diag, double* upper, double* source, double* internalCoeffs, double* boundaryCoeffs) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    DevBC dbc; dbc.upload(m, bc, 1);
    Stage s(m);
    double* g = s.faces(gammaf); double* pb = s.bnd(psiBoundary); double* dphi = phi ? s.faces(phi) : nullptr;
    double* dd = s.buf(m.N); double* du = s.buf(m.nFaceField()); double* ds = s.buf(m.N); double* dic = s.buf(m.B); double* dbc2 = s.buf(m.B);
    ScopedTimer t(T_ASSEMBLY);
    if (m.N) B200_LAUNCH(k_lap_row, gridFor(m.N), BLOCK, rt().stream, m.view(), (const double*)g, (const double*)dphi, dd, du, ds, (const double*)nullptr);
    if (m.B) B200_LAUNCH(k_lap_boundary, gridFor(m.B), BLOCK, rt().stream, m.view(), dbc.view(), (const double*)g, (const double*)pb, dic, dbc2);
    m.downloadCells(dd, diag); m.downloadFaces(du, upper, false); m.downloadCells(ds, source);
    m.downloadBoundary(dic, internalCoeffs); m.downloadBoundary(dbc2, boundaryCoeffs);
    API_END
}
int b200_matrix_flux(b200_mesh_t* mesh, const double* upper, const double* lower, const double* internalCoeffs,
                     const double* boundaryCoeffs, const double* psi, double* flux) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    Stage s(m);
    double* du = s.facesInt(upper); double* dl = (lower && lower != upper) ? s.facesInt(lower) : du;
    double* dic = s.bnd(internalCoeffs); double* dbc2 = s.bnd(boundaryCoeffs); double* dpsi = s.cells(psi); double* o = s.buf(m.nFaceField());
    B200_LAUNCH(k_matrix_flux<false>, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), (const double*)du, (const double*)dl, (const double*)dic,
                (const double*)dbc2, (const double*)dpsi, (const double*)nullptr, o, (const double*)nullptr);
    m.downloadFaces(o, flux, true);
    API_END
}
int b200_alpha_stats(b200_mesh_t* mesh, const double* alpha, const double* alphaBoundary, double* out3) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    Stage s(m);
    FvWorkspace& w = s.w;
    double* a = s.cells(alpha); double* ab = s.bnd(alphaBoundary);
    B200_LAUNCH(k_alpha_stats, gridFor(nmaxOf(m)), BLOCK, rt().stream, m.view(), (const double*)a, (const double*)ab, w.scal.p, w.partials.p, w.counter.p);
    if (rt().nRanks > 1) { allReduce(w.scal.p, 2, RED_SUM); allReduce(w.scal.p + 2, 1, RED_MIN); allReduce(w.scal.p + 3, 1, RED_MAX); }
    double h[4];
    w.scal.download(h, 4); syncStream();
    out3[0] = h[0] / h[1]; out3[1] = h[2]; out3[2] = h[3];
    API_END
}

// ---------------------------------------------------------------------------------------------- region3d
static Region3D* asRegion(b200_region3d_t* r) { return reinterpret_cast<Region3D*>(r); }

int b200_region3d_create(b200_mesh_t* mesh, const b200_region3d_controls* ctl, const b200_bc* bcAlpha, const b200_bc* bcU,
                         const b200_bc* bcPd, const double* alpha1, const double* U, const double* pd, b200_region3d_t** out) {
    API_BEGIN
    DeviceMesh& m = geoMesh(mesh);
    B200_REQUIRE(ctl && bcAlpha && bcU && bcPd && alpha1 && U && pd && out, B200_ERR_ARG, "b200_region3d_create: null argument");
    std::unique_ptr<Region3D> r(new Region3D());
    r->create(&m, *ctl, bcAlpha, bcU, bcPd, alpha1, U, pd);
    *out = reinterpret_cast<b200_region3d_t*>(r.release());
    API_END
}
int b200_region3d_release(b200_region3d_t* r) {
    API_BEGIN
    delete asRegion(r);
    API_END
}
int b200_region3d_correct_phi(b200_region3d_t* r) {
    API_BEGIN
    asRegion(r)->correctPhi();
    API_END
}
int b200_region3d_step(b200_region3d_t* r, int32_t nSteps) {
    API_BEGIN
    for (int i = 0; i < nSteps; i++) asRegion(r)->step();
    API_END
}
int b200_region3d_set_fields(b200_region3d_t* rr, const double* alpha1, const double* U, const double* pd, const double* phi) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    DeviceMesh& m = *r.m;
    if (alpha1) { m.uploadCells(alpha1, r.alpha); r.evalAlphaBC(); }
    if (U) { m.uploadCellVec(U, r.Ux, r.Uy, r.Uz); r.evalUBC(); }
    if (pd) {
        m.uploadCells(pd, r.pd);
        r.evalPdBC(r.bcPd, r.pd, r.pdb);
    }
    if (phi) m.uploadFaces(phi, r.phi, true);
    syncStream();
    API_END
}
int b200_region3d_get_fields(b200_region3d_t* rr, double* alpha1, double* U, double* pd, double* phi, double* p, double* rho) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    DeviceMesh& m = *r.m;
    if (alpha1) m.downloadCells(r.alpha, alpha1);
    if (U) m.downloadCellVec(r.Ux, r.Uy, r.Uz, U);
    if (pd) m.downloadCells(r.pd, pd);
    if (phi) m.downloadFaces(r.phi, phi, true);
    if (p) m.downloadCells(r.p, p);
    if (rho) m.downloadCells(r.rho, rho);
    API_END
}
int b200_region3d_step_host(b200_region3d_t* rr, double* alpha1, double* U, double* pd, double* phi, double* p) {
    API_BEGIN
    B200_REQUIRE(alpha1 && U && pd && phi && p, B200_ERR_ARG, "b200_region3d_step_host: null argument");
    int rc = b200_region3d_set_fields(rr, alpha1, U, pd, phi);
    if (rc != B200_OK) return rc;
    asRegion(rr)->step();
    rc = b200_region3d_get_fields(rr, alpha1, U, pd, phi, p, nullptr);
    if (rc != B200_OK) return rc;
    API_END
}
int b200_region3d_set_patch_mixed(b200_region3d_t* rr, int32_t field, int32_t patch, const double* refValue, const double* refGrad,
                                  const double* valueFraction) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    DeviceMesh& m = *r.m;
    B200_REQUIRE(field >= 0 && field <= 2 && patch >= 0 && patch < (int)m.patches.size(), B200_ERR_ARG, "b200_region3d_set_patch_mixed: bad field/patch");
    DevBC& bc = field == 0 ? r.bcAlpha : field == 1 ? r.bcU : r.bcPd;
    B200_REQUIRE(bc.types[patch] == B200_BC_MIXED, B200_ERR_ARG, "patch is not a mixed patch");
    int off = m.patches[patch].start - m.F, n = m.patches[patch].size, nc = bc.nc;
    std::vector<double> tmp(n);
    for (int k = 0; k < nc; k++) {
        for (int i = 0; i < n; i++) tmp[i] = refValue[(size_t)nc * i + k];
        B200_CHECK_CUDA(cudaMemcpyAsync(bc.rv.p + (size_t)k * m.B + off, tmp.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, rt().stream));
        syncStream();
        for (int i = 0; i < n; i++) tmp[i] = refGrad[(size_t)nc * i + k];
        B200_CHECK_CUDA(cudaMemcpyAsync(bc.rg.p + (size_t)k * m.B + off, tmp.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, rt().stream));
        syncStream();
    }
    B200_CHECK_CUDA(cudaMemcpyAsync(bc.vf.p + off, valueFraction, 8 * (size_t)n, cudaMemcpyHostToDevice, rt().stream));
    syncStream();
    API_END
}
int b200_region3d_get_patch_internal(b200_region3d_t* rr, int32_t field, int32_t patch, double* out) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    DeviceMesh& m = *r.m;
    B200_REQUIRE(field >= 0 && field <= 2 && patch >= 0 && patch < (int)m.patches.size() && out, B200_ERR_ARG, "b200_region3d_get_patch_internal: bad arguments");
    int off = m.patches[patch].start - m.F, n = m.patches[patch].size;
    if (n == 0) return B200_OK;
    Stage s(m);
    const double* src[3] = {field == 0 ? r.alpha.p : field == 1 ? r.Ux.p : r.pd.p, r.Uy.p, r.Uz.p};
    int nc = field == 1 ? 3 : 1;
    std::vector<double> tmp(n);
    double* d = s.buf(n);
    for (int k = 0; k < nc; k++) {
        B200_LAUNCH(k_gather_patch, gridFor(n), BLOCK, rt().stream, (const int*)(m.dBFaceRow.p + off), n, src[k], d);
        B200_CHECK_CUDA(cudaMemcpyAsync(tmp.data(), d, 8 * (size_t)n, cudaMemcpyDeviceToHost, rt().stream));
        syncStream();
        for (int i = 0; i < n; i++) out[(size_t)nc * i + k] = tmp[i];
    }
    API_END
}
int b200_region3d_num_solves(b200_region3d_t* r) { return (int)asRegion(r)->perfs.size(); }
int b200_region3d_get_perfs(b200_region3d_t* r, b200_solver_perf* out, int32_t cap) {
    API_BEGIN
    auto& v = asRegion(r)->perfs;
    for (int i = 0; i < cap && i < (int)v.size(); i++) out[i] = v[i];
    API_END
}
int b200_region3d_get_res_hist(b200_region3d_t* r, double* out, int32_t cap) {
    auto& v = asRegion(r)->resHist;
    for (int i = 0; out && i < cap && i < (int)v.size(); i++) out[i] = v[i];
    return (int)v.size();
}
int b200_region3d_get_alpha_stats(b200_region3d_t* r, double* out, int32_t cap) {
    auto& v = asRegion(r)->alphaStats;
    for (int i = 0; out && i < cap && i < (int)v.size(); i++) out[i] = v[i];
    return (int)v.size();
}
int b200_region3d_get_cont_errs(b200_region3d_t* rr, double* out3) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    out3[0] = r.sumLocalContErr; out3[1] = r.globalContErr; out3[2] = r.cumulativeContErr;
    API_END
}
int b200_region3d_get_courant(b200_region3d_t* rr, double* out3) {
    API_BEGIN
    B200_REQUIRE(rr && out3, B200_ERR_ARG, "b200_region3d_get_courant: null argument");
    asRegion(rr)->courant(out3);
    API_END
}
int b200_region3d_restore(b200_region3d_t* rr, const double* alpha1, const double* U, const double* pd, const double* phi) {
    API_BEGIN
    B200_REQUIRE(rr && alpha1 && U && pd && phi, B200_ERR_ARG, "b200_region3d_restore: null argument");
    int rc = b200_region3d_set_fields(rr, alpha1, U, pd, phi);
    if (rc != B200_OK) return rc;
    asRegion(rr)->refreshDerived();
    API_END
}
int b200_region3d_capture(b200_region3d_t* rr, int32_t on) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    r.capture = on != 0;
    r.capSolves.clear(); r.capMules.clear();
    API_END
}
int b200_region3d_num_captured(b200_region3d_t* rr, int32_t kind) {
    Region3D& r = *asRegion(rr);
    return kind == 0 ? (int)r.capSolves.size() : (int)r.capMules.size();
}
int b200_region3d_get_captured_solve(b200_region3d_t* rr, int32_t i, double* diag, double* upper, double* source, double* x0, double* bou,
                                     double* tolRelTol) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    B200_REQUIRE(i >= 0 && i < (int)r.capSolves.size(), B200_ERR_ARG, "b200_region3d_get_captured_solve: index out of range");
    const Region3D::CapturedSolve& c = r.capSolves[i];
    if (diag) memcpy(diag, c.diag.data(), 8 * c.diag.size());
    if (upper) memcpy(upper, c.upper.data(), 8 * c.upper.size());
    if (source) memcpy(source, c.source.data(), 8 * c.source.size());
    if (x0) memcpy(x0, c.x0.data(), 8 * c.x0.size());
    if (bou) memcpy(bou, c.bou.data(), 8 * c.bou.size());
    if (tolRelTol) { tolRelTol[0] = c.tol; tolRelTol[1] = c.relTol; }
    API_END
}
int b200_region3d_get_captured_mules(b200_region3d_t* rr, int32_t i, double* psi, double* psib, double* psi0, double* phi, double* phiPsi,
                                     double* deltaT) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    B200_REQUIRE(i >= 0 && i < (int)r.capMules.size(), B200_ERR_ARG, "b200_region3d_get_captured_mules: index out of range");
    const Region3D::CapturedMules& c = r.capMules[i];
    if (psi) memcpy(psi, c.psi.data(), 8 * c.psi.size());
    if (psib) memcpy(psib, c.psib.data(), 8 * c.psib.size());
    if (psi0) memcpy(psi0, c.psi0.data(), 8 * c.psi0.size());
    if (phi) memcpy(phi, c.phi.data(), 8 * c.phi.size());
    if (phiPsi) memcpy(phiPsi, c.phiPsi.data(), 8 * c.phiPsi.size());
    if (deltaT) *deltaT = c.deltaT;
    API_END
}
int b200_region3d_clear_log(b200_region3d_t* rr) {
    API_BEGIN
    Region3D& r = *asRegion(rr);
    r.perfs.clear(); r.resHist.clear(); r.alphaStats.clear(); r.pendingStats.clear();
    API_END
}

}  // extern "C"
