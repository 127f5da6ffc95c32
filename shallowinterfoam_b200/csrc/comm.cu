// comm.cu -- NCCL-backed all-reduce and processor-patch halo exchange (one process per GPU).
#include "comm.cuh"
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
                                  const double* __restrict__ Cfz, double* __restrict__ w, double* __restrict__ dc) {
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
}

bool meshHasProcessorPatches(const DeviceMesh& m) {
    for (const Patch& p : m.patches) if (p.neighbRank >= 0 && p.size > 0) return true;
    return false;
}
void DeviceMesh::coupleGeometry() {
    if (!hasGeometry || !meshHasProcessorPatches(*this)) return;
    DevBuf<double> send(B);
    const double* src[3] = {dCx.p, dCy.p, dCz.p};
    double* dst[3] = {dCnx.p, dCny.p, dCnz.p};
    for (int k = 0; k < 3; k++) { haloPack(*this, src[k], send.p); haloExchange(*this, send.p, dst[k]); }
    B200_LAUNCH(k_couple_geometry, gridFor(B), BLOCK, rt().stream, view(), (const double*)dCfx.p, (const double*)dCfy.p,
                (const double*)dCfz.p, dW.p, dDc.p);
    syncStream();
}
void haloPack(const DeviceMesh& m, const double* cellField, double* send) {
    if (m.B) B200_LAUNCH(k_halo_pack, gridFor(m.B), BLOCK, rt().stream, m.view(), cellField, send);
}

#ifdef B200_WITH_NCCL
void commInit(int rank, int nRanks, const void* id, size_t idBytes) {
    if (nRanks <= 1) return;
    B200_REQUIRE(id && idBytes >= sizeof(ncclUniqueId), B200_ERR_ARG, "b200_init: ncclUniqueId missing");
    ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    B200_CHECK_NCCL(ncclCommInitRank(&comm, nRanks, uid, rank));
    rt().nccl = comm;
}
void commFinalize() {
    if (rt().nccl) { ncclCommDestroy((ncclComm_t)rt().nccl); rt().nccl = nullptr; }
}
int commUniqueId(void* out128) {
    ncclUniqueId uid;
    B200_CHECK_NCCL(ncclGetUniqueId(&uid));
    memcpy(out128, &uid, sizeof(uid));
    return (int)sizeof(uid);
}
void allReduce(double* p, int count, ReduceOp op) {
    if (rt().nRanks <= 1) return;
    ncclRedOp_t o = op == RED_SUM ? ncclSum : op == RED_MAX ? ncclMax : ncclMin;
    B200_CHECK_NCCL(ncclAllReduce(p, p, count, ncclDouble, o, (ncclComm_t)rt().nccl, rt().stream));
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
int commUniqueId(void* out128) { memset(out128, 0, 128); return 128; }
void allReduce(double* p, int count, ReduceOp op) { if (rt().nRanks > 1) g_emuAllreduce(p, count, (int)op); }
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
int commUniqueId(void* out128) { memset(out128, 0, 128); return 128; }
void allReduce(double*, int, ReduceOp) {}
void haloExchange(const DeviceMesh&, const double*, double*) {}
#endif

}  // namespace b200
