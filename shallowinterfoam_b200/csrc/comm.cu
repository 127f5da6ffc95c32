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

bool meshHasProcessorPatches(const DeviceMesh& m) {
    for (const Patch& p : m.patches) if (p.neighbRank >= 0 && p.size > 0) return true;
    return false;
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
