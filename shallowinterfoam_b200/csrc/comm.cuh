// comm.cuh -- rank-to-rank plumbing: scalar all-reduces (gSum/gMax/gMin) and processor-patch halo swaps over NCCL.
// [UPSTREAM Pstream reduce; processorFvPatchField initInterfaceMatrixUpdate/updateInterfaceMatrix, patchNeighbourField]
#pragma once
#include "device_mesh.cuh"

namespace b200 {
void commInit(int rank, int nRanks, const void* id, size_t idBytes);
void commFinalize();
int commUniqueId(void* out128);
enum ReduceOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
void allReduce(double* devPtr, int count, ReduceOp op);  // in place on rt().stream; no-op when nRanks == 1
struct PcgScalars;
// sum all-reduce over peer memory whose kernel also runs the PCG convergence check (k_pcg_check's body); false: not available, the
// caller does allReduce + k_pcg_check
bool allReduceSumThenPcgCheck(double* devPtr, int count, PcgScalars* sc, double* hist);
bool commUsesPeerMemory();   // scalar all-reduces run over NVLink peer memory (mailboxes) instead of NCCL
// exchange per-boundary-face values on processor patches: recv[b] = neighbour rank's send[b'] (same face)
void haloExchange(const DeviceMesh& m, const double* send, double* recv);
bool meshHasProcessorPatches(const DeviceMesh& m);
// send[b] = cellField[faceCell row of b] on processor faces
void haloPack(const DeviceMesh& m, const double* cellField, double* send);

// ---- halo swap over NVLink peer memory, fused into the producer / consumer kernels of the PCG iteration (no NCCL launch):
// the producer kernel (k_halo_push) gathers the processor-face values and stores them straight into the NEIGHBOUR rank's receive
// buffer (cudaIpc-mapped), then raises an epoch flag there; the consumer kernel (k_iface_rows) spins on its own flags before it reads.
// Amul runs between the two, so the NVLink transfer overlaps it.  [UPSTREAM processorFvPatchField::initInterfaceMatrixUpdate /
// updateInterfaceMatrix: the same non-blocking send ... compute ... receive split]
constexpr int PEER_MAX_RANKS = 16;
struct PeerHaloWait {      // POD for the consumer kernel; nNbr = 0: nothing to wait for
    int nNbr;
    int nbrRank[PEER_MAX_RANKS];
    const double* flags;   // this rank's flag row of the current parity: flags[q] = epoch tag written by rank q
    double tag;
};
struct PeerHalo;           // per mesh (comm.cu)
// collective over all ranks (first multi-rank solve on a mesh); null when peer mapping is unavailable or switched off (p2p_halo = 0)
PeerHalo* peerHaloGet(DeviceMesh& m);
void peerHaloRelease(PeerHalo* h);
// launches the push of cellField's processor-face values for a new epoch; afterwards peerHaloRecv / peerHaloWait describe that epoch
void peerHaloPush(PeerHalo* h, const DeviceMesh& m, const double* cellField, bool pdl = false);
const double* peerHaloRecv(const PeerHalo* h);   // [B] receive buffer of the current epoch (read with ld.global.cg after the wait)
PeerHaloWait peerHaloWait(const PeerHalo* h);
__device__ __forceinline__ void peerHaloWaitDevice(const PeerHaloWait& w) {
    if (w.nNbr == 0) return;
    if ((int)threadIdx.x < w.nNbr) {
        const volatile double* f = w.flags + w.nbrRank[threadIdx.x];
        while (*f != w.tag) {}
        __threadfence_system();
    }
    __syncthreads();
}
}  // namespace b200
