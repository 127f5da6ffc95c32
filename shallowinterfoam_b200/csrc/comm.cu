// comm.cu -- NCCL-backed all-reduce and processor-patch halo exchange (one process per GPU).
#include "comm.cuh"
#include "ldu_kernels.cuh"
#ifdef B200_WITH_NCCL
#include <nccl.h>
#define B200_CHECK_NCCL(expr)                                                                              \
    do {                                                                                                   \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess) throw b200::Error(B200_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_r)); \
    } while (0)
#endif

namespace b200 {

__global__ void k_halo_pack(MeshView mv, const double* __restrict__ cf, double* __restrict__ send) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < mv.B && mv.bCoupled[b]) send[b] = cf[mv.bFaceRow[b]];
}

// processor faces get the geometry of the internal face they were before decomposition (seen from this side):
// weights = |Sf.(Cn-Cf)| / (|Sf.(Cf-C)| + |Sf.(Cn-Cf)|), deltaCoeffs = 1/|Cn - C|   [UPSTREAM surfaceInterpolation.C]
__global__ void k_couple_geometry(MeshView mv, const double* __restrict__ Cfx, const double* __restrict__ Cfy,
                                  const double* __restrict__ Cfz, double* __restrict__ w, double* __restrict__ dc,
                                  double* __restrict__ cvx, double* __restrict__ cvy, double* __restrict__ cvz) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B || !mv.bCoupled[b]) return;
    int p = mv.Fs + b, c = mv.bFaceRow[b];
    double ox = mv.Cx[c], oy = mv.Cy[c], oz = mv.Cz[c], nx = mv.Cnx[b], ny = mv.Cny[b], nz = mv.Cnz[b];
    double fx = Cfx[p], fy = Cfy[p], fz = Cfz[p], sx = mv.Sfx[p], sy = mv.Sfy[p], sz = mv.Sfz[p];
    double SfdOwn = fabs((sx * (fx - ox) + sy * (fy - oy)) + sz * (fz - oz));
    double SfdNei = fabs((sx * (nx - fx) + sy * (ny - fy)) + sz * (nz - fz));
    w[p] = SfdNei / (SfdOwn + SfdNei);
    double dx = nx - ox, dy = ny - oy, dz = nz - oz;
    dc[p] = 1.0 / sqrt((dx * dx + dy * dy) + dz * dz);
    const double ms = mv.magSf[p];
    cvx[p] = sx / ms - dx * dc[p]; cvy[p] = sy / ms - dy * dc[p]; cvz[p] = sz / ms - dz * dc[p];
}

bool meshHasProcessorPatches(const DeviceMesh& m) {
    for (const Patch& p : m.patches) if (p.neighbRank >= 0 && p.size > 0) return true;
    return false;
}
void DeviceMesh::coupleGeometry() {
    if (hasGeometry && rt().nRanks > 1) {   // mesh.orthogonal() is a global property: gSum of both sums (collective: every rank's mesh)
        DevBuf<double> s2(2);
        B200_CHECK_CUDA(cudaMemcpyAsync(s2.p, nonOrthSum, 16, cudaMemcpyHostToDevice, rt().stream));
        allReduce(s2.p, 2, RED_SUM);
        double h[2];
        B200_CHECK_CUDA(cudaMemcpyAsync(h, s2.p, 16, cudaMemcpyDeviceToHost, rt().stream));
        syncStream();
        orthogonal = nonOrthogonalityDegrees(h[0], h[1]) < 0.1;
    }
    if (!hasGeometry || !meshHasProcessorPatches(*this)) return;
    DevBuf<double> send(B);
    const double* src[3] = {dCx.p, dCy.p, dCz.p};
    double* dst[3] = {dCnx.p, dCny.p, dCnz.p};
    for (int k = 0; k < 3; k++) { haloPack(*this, src[k], send.p); haloExchange(*this, send.p, dst[k]); }
    B200_LAUNCH(k_couple_geometry, gridFor(B), BLOCK, rt().stream, view(), (const double*)dCfx.p, (const double*)dCfy.p,
                (const double*)dCfz.p, dW.p, dDc.p, dCvx.p, dCvy.p, dCvz.p);
    syncStream();
}
void haloPack(const DeviceMesh& m, const double* cellField, double* send) {
    if (m.B) B200_LAUNCH(k_halo_pack, gridFor(m.B), BLOCK, rt().stream, m.view(), cellField, send);
}

#ifdef B200_WITH_NCCL
// ---- scalar all-reduce over NVLink peer memory (gSum / gMax / gMin of PCG and MULES: 1-4 doubles, pure latency).
// Every rank owns a mailbox; an all-reduce = each rank stores its values + an epoch tag into its slot of EVERY rank's mailbox
// (peer stores over NVLink), spins on its own mailbox until all tags of this epoch are in, and combines the values in rank order
// (same order on every rank => identical result everywhere).  One 32-thread kernel, ~5 us, instead of an NCCL launch (~20 us).
// Slots are 4-deep: a rank can be at most one all-reduce ahead of its slowest peer.
struct P2PState { bool ok = false; P2PView v; unsigned long long epoch = 0; void* opened[P2P_MAX_RANKS] = {nullptr}; double* own = nullptr; };
static P2PState g_p2p;

// post != 0: the thread that holds the result also runs k_pcg_check's body (closes the previous PCG iteration, adopts the new wArA):
// one launch less per iteration
__global__ void k_p2p_allreduce(double* data, int count, int op, P2PView pv, unsigned long long epoch, PcgScalars* post, double* hist) {
    pdlTrigger();   // (launched without the PDL attribute: everything before it is complete) lets a PDL successor -- the next forward
                    // sweep -- run its prologue while this warp waits for the peers
    p2pAllreduceWarp(data, count, op, pv, epoch);
    if (post) {
        if (threadIdx.x == 0 && !post->stop) {
            if (post->pendingCheck) { post->pendingCheck = 0; pcgEndOfIteration(post, hist); }
            if (!post->stop) post->wArA = post->wArAnew;
        }
    }
}

static void p2pInit(int rank, int nRanks) {
    g_p2p = P2PState();
    if (nRanks > P2P_MAX_RANKS) return;
    cudaStream_t st = rt().stream;
    const size_t bytes = sizeof(double) * P2P_SLOTS * nRanks * P2P_ENTRY;
    if (cudaMalloc((void**)&g_p2p.own, bytes) != cudaSuccess) { cudaGetLastError(); return; }
    cudaMemsetAsync(g_p2p.own, 0, bytes, st);
    cudaIpcMemHandle_t mine;
    bool good = cudaIpcGetMemHandle(&mine, g_p2p.own) == cudaSuccess;
    if (!good) cudaGetLastError();
    // exchange the handles (and whether every rank got one) through the NCCL communicator
    char* dAll = nullptr; char* dMine = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<char> hMine(hb, 0), hAll(hb * nRanks, 0);
    memcpy(hMine.data(), &mine, sizeof(mine)); hMine[sizeof(mine)] = good ? 1 : 0;
    B200_CHECK_CUDA(cudaMalloc((void**)&dAll, hb * nRanks)); B200_CHECK_CUDA(cudaMalloc((void**)&dMine, hb));
    B200_CHECK_CUDA(cudaMemcpyAsync(dMine, hMine.data(), hb, cudaMemcpyHostToDevice, st));
    B200_CHECK_NCCL(ncclAllGather(dMine, dAll, hb, ncclChar, (ncclComm_t)rt().nccl, st));
    B200_CHECK_CUDA(cudaMemcpyAsync(hAll.data(), dAll, hb * nRanks, cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(dAll); cudaFree(dMine);
    for (int r = 0; r < nRanks; r++) good = good && hAll[hb * r + sizeof(mine)] == 1;
    g_p2p.v.rank = rank; g_p2p.v.nRanks = nRanks; g_p2p.v.mbox = g_p2p.own;
    for (int r = 0; r < nRanks && good; r++) {
        if (r == rank) { g_p2p.v.peer[r] = g_p2p.own; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, &hAll[hb * r], sizeof(h));
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
        g_p2p.opened[r] = ptr; g_p2p.v.peer[r] = (double*)ptr;
    }
    // every rank must agree (a rank that could not map a peer sends everybody back to NCCL)
    double flag = good ? 1.0 : 0.0, *dFlag = nullptr;
    B200_CHECK_CUDA(cudaMalloc((void**)&dFlag, 8));
    B200_CHECK_CUDA(cudaMemcpyAsync(dFlag, &flag, 8, cudaMemcpyHostToDevice, st));
    B200_CHECK_NCCL(ncclAllReduce(dFlag, dFlag, 1, ncclDouble, ncclMin, (ncclComm_t)rt().nccl, st));
    B200_CHECK_CUDA(cudaMemcpyAsync(&flag, dFlag, 8, cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(dFlag);
    g_p2p.ok = flag > 0.5;
}
static void p2pFinalize() {
    for (auto& p : g_p2p.opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    if (g_p2p.own) { cudaFree(g_p2p.own); g_p2p.own = nullptr; }
    g_p2p.ok = false;
}

void commInit(int rank, int nRanks, const void* id, size_t idBytes) {
    if (nRanks <= 1) return;
    B200_REQUIRE(id && idBytes >= sizeof(ncclUniqueId), B200_ERR_ARG, "b200_init: ncclUniqueId missing");
    ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    B200_CHECK_NCCL(ncclCommInitRank(&comm, nRanks, uid, rank));
    rt().nccl = comm;
    p2pInit(rank, nRanks);
}
void commFinalize() {
    p2pFinalize();
    if (rt().nccl) { ncclCommDestroy((ncclComm_t)rt().nccl); rt().nccl = nullptr; }
}
bool commUsesPeerMemory() { return g_p2p.ok && rt().opt("p2p_allreduce", 1) > 0; }
int commUniqueId(void* out128) {
    ncclUniqueId uid;
    B200_CHECK_NCCL(ncclGetUniqueId(&uid));
    memcpy(out128, &uid, sizeof(uid));
    return (int)sizeof(uid);
}
// (the option is read per call so that a checker can exercise both transports in one process; every rank must set it alike)
static bool p2pAllreduceOn() { return g_p2p.ok && rt().opt("p2p_allreduce", 1) > 0; }
void allReduce(double* p, int count, ReduceOp op) {
    if (rt().nRanks <= 1) return;
    if (p2pAllreduceOn() && count <= P2P_ENTRY - 1) {
        g_p2p.epoch++;
        B200_LAUNCH(k_p2p_allreduce, 1, 32, rt().stream, p, count, (int)op, g_p2p.v, g_p2p.epoch, (PcgScalars*)nullptr, (double*)nullptr);
        return;
    }
    ncclRedOp_t o = op == RED_SUM ? ncclSum : op == RED_MAX ? ncclMax : ncclMin;
    B200_CHECK_NCCL(ncclAllReduce(p, p, count, ncclDouble, o, (ncclComm_t)rt().nccl, rt().stream));
}
// ---- peer-memory halo (comm.cuh).  Own allocation: [2][PEER_MAX_RANKS] epoch flags, then [2][B] receive values (parity = epoch & 1).
struct PeerHaloPushView {
    int nNbr, rank;
    int myOff[PEER_MAX_RANKS], size[PEER_MAX_RANKS];      // this rank's processor-face range [myOff, myOff+size) per neighbour slot
    double* peerData[PEER_MAX_RANKS];                     // neighbour's receive buffer (parity 0) at the offset of its patch towards this rank
    long long peerStride[PEER_MAX_RANKS];                 // neighbour's B (distance between its two parity buffers)
    double* peerFlags[PEER_MAX_RANKS];                    // neighbour's flag array
};
struct PeerHalo {
    PeerHaloPushView v;
    double* own = nullptr; int B = 0;
    void* opened[PEER_MAX_RANKS] = {nullptr};
    unsigned long long epoch = 0;
    unsigned* counter = nullptr;
    int maxSize = 0;
};
__global__ void k_halo_push(MeshView mv, const double* __restrict__ cf, PeerHaloPushView hv, unsigned long long epoch, unsigned* counter) {
    pdlWait();
    pdlTrigger();
    const int slot = blockIdx.y, parity = (int)(epoch & 1ull);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hv.size[slot]) {
        const int b = hv.myOff[slot] + i;
        hv.peerData[slot][parity * hv.peerStride[slot] + i] = cf[mv.bFaceRow[b]];     // store over NVLink into the neighbour's HBM
    }
    __shared__ int isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();     // cumulative: the block's stores (ordered before this by the barrier) precede the flag
        const unsigned total = gridDim.x * gridDim.y;
        isLast = atomicInc(counter, total - 1) == total - 1;
    }
    __syncthreads();
    if (isLast && (int)threadIdx.x < hv.nNbr) {     // every block's values are out: raise this rank's flag at each neighbour
        __threadfence_system();
        volatile double* f = hv.peerFlags[threadIdx.x] + parity * PEER_MAX_RANKS + hv.rank;
        *f = (double)epoch;
    }
}
PeerHalo* peerHaloGet(DeviceMesh& m) {
    const int nRanks = rt().nRanks, rank = rt().rank;
    if (nRanks <= 1) return nullptr;
    cudaStream_t st = rt().stream;
    ncclComm_t comm = (ncclComm_t)rt().nccl;
    PeerHalo* h = new PeerHalo();
    memset(&h->v, 0, sizeof(h->v));
    h->B = m.B; h->v.rank = rank;
    bool good = nRanks <= PEER_MAX_RANKS && rt().opt("p2p_halo", 1) > 0;
    // this rank's record: IPC handle, ok flag, B, offset of the patch towards every rank (-1: none)
    struct Rec { cudaIpcMemHandle_t handle; int ok, B; int offTo[PEER_MAX_RANKS], sizeTo[PEER_MAX_RANKS]; };
    Rec mine; memset(&mine, 0, sizeof(mine));
    for (int q = 0; q < PEER_MAX_RANKS; q++) { mine.offTo[q] = -1; mine.sizeTo[q] = 0; }
    for (const Patch& p : m.patches) {
        if (p.neighbRank < 0 || p.size == 0) continue;
        if (p.neighbRank >= PEER_MAX_RANKS || mine.offTo[p.neighbRank] >= 0) { good = false; continue; }   // one patch per neighbour (decomposePar)
        mine.offTo[p.neighbRank] = p.start - m.F; mine.sizeTo[p.neighbRank] = p.size;
    }
    const size_t bytes = sizeof(double) * (2 * PEER_MAX_RANKS + 2 * (size_t)std::max(1, m.B));
    if (good && cudaMalloc((void**)&h->own, bytes) != cudaSuccess) { cudaGetLastError(); good = false; h->own = nullptr; }
    if (h->own) {
        cudaMemsetAsync(h->own, 0, bytes, st);
        if (cudaIpcGetMemHandle(&mine.handle, h->own) != cudaSuccess) { cudaGetLastError(); good = false; }
    }
    mine.ok = good ? 1 : 0; mine.B = m.B;
    std::vector<Rec> all(nRanks);
    char *dMine = nullptr, *dAll = nullptr;
    B200_CHECK_CUDA(cudaMalloc((void**)&dMine, sizeof(Rec))); B200_CHECK_CUDA(cudaMalloc((void**)&dAll, sizeof(Rec) * nRanks));
    B200_CHECK_CUDA(cudaMemcpyAsync(dMine, &mine, sizeof(Rec), cudaMemcpyHostToDevice, st));
    B200_CHECK_NCCL(ncclAllGather(dMine, dAll, sizeof(Rec), ncclChar, comm, st));
    B200_CHECK_CUDA(cudaMemcpyAsync(all.data(), dAll, sizeof(Rec) * nRanks, cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(dMine); cudaFree(dAll);
    for (int r = 0; r < nRanks; r++) good = good && all[r].ok == 1;
    int n = 0;
    for (int q = 0; q < nRanks && good; q++) {
        if (q == rank || mine.offTo[q] < 0) continue;
        if (all[q].offTo[rank] < 0 || all[q].sizeTo[rank] != mine.sizeTo[q]) { good = false; break; }   // both sides list the same faces
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); good = false; break; }
        h->opened[q] = ptr;
        double* base = (double*)ptr;
        h->v.myOff[n] = mine.offTo[q]; h->v.size[n] = mine.sizeTo[q];
        h->v.peerFlags[n] = base;
        h->v.peerData[n] = base + 2 * PEER_MAX_RANKS + all[q].offTo[rank];
        h->v.peerStride[n] = all[q].B;
        h->maxSize = std::max(h->maxSize, mine.sizeTo[q]);
        h->v.nNbr = ++n;     // slots are in rank order: peerHaloWait lists the opened ranks in the same order
    }
    // every rank must agree
    double flag = good ? 1.0 : 0.0, *dFlag = nullptr;
    B200_CHECK_CUDA(cudaMalloc((void**)&dFlag, 8));
    B200_CHECK_CUDA(cudaMemcpyAsync(dFlag, &flag, 8, cudaMemcpyHostToDevice, st));
    B200_CHECK_NCCL(ncclAllReduce(dFlag, dFlag, 1, ncclDouble, ncclMin, comm, st));
    B200_CHECK_CUDA(cudaMemcpyAsync(&flag, dFlag, 8, cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    cudaFree(dFlag);
    if (flag < 0.5) { peerHaloRelease(h); return nullptr; }
    B200_CHECK_CUDA(cudaMalloc((void**)&h->counter, sizeof(unsigned)));
    B200_CHECK_CUDA(cudaMemsetAsync(h->counter, 0, sizeof(unsigned), st));
    B200_CHECK_CUDA(cudaStreamSynchronize(st));
    return h;
}
void peerHaloRelease(PeerHalo* h) {
    if (!h) return;
    for (auto& p : h->opened) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
    if (h->own) cudaFree(h->own);
    if (h->counter) cudaFree(h->counter);
    delete h;
}
void peerHaloPush(PeerHalo* h, const DeviceMesh& m, const double* cellField, bool pdl, cudaStream_t stream) {
    h->epoch++;
    if (h->v.nNbr == 0) return;
    dim3 grid(std::max(1, gridFor(h->maxSize)), h->v.nNbr);
    B200_LAUNCH_PDL(pdl && !stream, k_halo_push, grid, BLOCK, 0, stream ? stream : rt().stream, m.view(), cellField, h->v, h->epoch, h->counter);
}
const double* peerHaloRecv(const PeerHalo* h) { return h->own + 2 * PEER_MAX_RANKS + (size_t)(h->epoch & 1ull) * h->B; }
PeerHaloWait peerHaloWait(const PeerHalo* h) {
    PeerHaloWait w; memset(&w, 0, sizeof(w));
    w.nNbr = h->v.nNbr;
    int n = 0;
    for (int q = 0; q < PEER_MAX_RANKS; q++) if (h->opened[q]) w.nbrRank[n++] = q;
    w.flags = h->own + (size_t)(h->epoch & 1ull) * PEER_MAX_RANKS;
    w.tag = (double)h->epoch;
    return w;
}

P2PFold p2pFoldNext() {
    P2PFold f; memset(&f, 0, sizeof(f));
    if (!p2pAllreduceOn() || rt().opt("p2p_fold", 1) <= 0) return f;
    f.on = 1; f.v = g_p2p.v; f.epoch = ++g_p2p.epoch;
    return f;
}
bool allReduceSumThenPcgCheck(double* p, int count, PcgScalars* sc, double* hist) {
    if (!(p2pAllreduceOn() && count <= P2P_ENTRY - 1)) return false;
    g_p2p.epoch++;
    B200_LAUNCH(k_p2p_allreduce, 1, 32, rt().stream, p, count, (int)RED_SUM, g_p2p.v, g_p2p.epoch, sc, hist);
    return true;
}
void haloExchange(const DeviceMesh& m, const double* send, double* recv) {
    if (rt().nRanks <= 1) return;
    B200_CHECK_NCCL(ncclGroupStart());
    for (const Patch& p : m.patches) {
        if (p.neighbRank < 0 || p.size == 0) continue;
        int off = p.start - m.F;
        B200_CHECK_NCCL(ncclSend(send + off, p.size, ncclDouble, p.neighbRank, (ncclComm_t)rt().nccl, rt().stream));
        B200_CHECK_NCCL(ncclRecv(recv + off, p.size, ncclDouble, p.neighbRank, (ncclComm_t)rt().nccl, rt().stream));
    }
    B200_CHECK_NCCL(ncclGroupEnd());
}
#elif defined(B200_EMU)
// TEST BUILD ONLY: the SIMT-emulated library has no NCCL; the multi-process CPU tests (gloo, world_size 2) register
// host callbacks that move the (host-resident) "device" buffers through torch.distributed.
typedef void (*emu_allreduce_fn)(double* p, int count, int op);
typedef void (*emu_sendrecv_fn)(const double* send, double* recv, int count, int peer);
static emu_allreduce_fn g_emuAllreduce = nullptr;
static emu_sendrecv_fn g_emuSendrecv = nullptr;
extern "C" void b200_emu_set_comm(emu_allreduce_fn a, emu_sendrecv_fn s) { g_emuAllreduce = a; g_emuSendrecv = s; }
void commInit(int, int nRanks, const void*, size_t) {
    B200_REQUIRE(nRanks <= 1 || (g_emuAllreduce && g_emuSendrecv), B200_ERR_STATE, "emulated build: register b200_emu_set_comm first");
}
void commFinalize() {}
bool commUsesPeerMemory() { return false; }
int commUniqueId(void* out128) { memset(out128, 0, 128); return 128; }
void allReduce(double* p, int count, ReduceOp op) { if (rt().nRanks > 1) g_emuAllreduce(p, count, (int)op); }
bool allReduceSumThenPcgCheck(double*, int, PcgScalars*, double*) { return false; }
P2PFold p2pFoldNext() { P2PFold f; memset(&f, 0, sizeof(f)); return f; }
struct PeerHalo {};
PeerHalo* peerHaloGet(DeviceMesh&) { return nullptr; }
void peerHaloRelease(PeerHalo*) {}
void peerHaloPush(PeerHalo*, const DeviceMesh&, const double*, bool, cudaStream_t) {}
const double* peerHaloRecv(const PeerHalo*) { return nullptr; }
PeerHaloWait peerHaloWait(const PeerHalo*) { PeerHaloWait w; memset(&w, 0, sizeof(w)); return w; }
void haloExchange(const DeviceMesh& m, const double* send, double* recv) {
    if (rt().nRanks <= 1) return;
    for (const Patch& p : m.patches) {
        if (p.neighbRank < 0 || p.size == 0) continue;
        int off = p.start - m.F;
        g_emuSendrecv(send + off, recv + off, p.size, p.neighbRank);
    }
}
#else
void commInit(int, int nRanks, const void*, size_t) {
    B200_REQUIRE(nRanks <= 1, B200_ERR_UNSUPPORTED, "this build has no NCCL: nRanks must be 1");
}
void commFinalize() {}
bool commUsesPeerMemory() { return false; }
int commUniqueId(void* out128) { memset(out128, 0, 128); return 128; }
void allReduce(double*, int, ReduceOp) {}
void haloExchange(const DeviceMesh&, const double*, double*) {}
bool allReduceSumThenPcgCheck(double*, int, PcgScalars*, double*) { return false; }
P2PFold p2pFoldNext() { P2PFold f; memset(&f, 0, sizeof(f)); return f; }
struct PeerHalo {};
PeerHalo* peerHaloGet(DeviceMesh&) { return nullptr; }
void peerHaloRelease(PeerHalo*) {}
void peerHaloPush(PeerHalo*, const DeviceMesh&, const double*, bool, cudaStream_t) {}
const double* peerHaloRecv(const PeerHalo*) { return nullptr; }
PeerHaloWait peerHaloWait(const PeerHalo*) { PeerHaloWait w; memset(&w, 0, sizeof(w)); return w; }
#endif

}  // namespace b200
