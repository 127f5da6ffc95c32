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

// ---- scalar all-reduce over NVLink peer memory (gSum / gMax / gMin of PCG and MULES: 1-7 doubles, pure latency).
// Every rank owns a mailbox; an all-reduce = each rank stores its values + an epoch tag into its slot of EVERY rank's mailbox
// (peer stores over NVLink), spins on its own mailbox until all tags of this epoch are in, and combines the values in rank order
// (same order on every rank => identical result everywhere).  Slots are 4-deep: a rank can be at most one all-reduce ahead of its
// slowest peer.  The exchange is ONE WARP's work, so besides the stand-alone kernel (k_p2p_allreduce) it can ride in the tail of the
// kernel that produces the values: the last CTA of the backward sweep and of the interface update (P2PFold) -- no launch of its own.
constexpr int P2P_MAX_RANKS = 16, P2P_SLOTS = 4, P2P_ENTRY = 8;   // entry = 7 values + tag
struct P2PView { double* mbox; double* peer[P2P_MAX_RANKS]; int rank, nRanks; };
struct P2PFold { int on; P2PView v; unsigned long long epoch; };   // on = 0: the caller runs the all-reduce as a separate step
// a new epoch for an all-reduce folded into another kernel; on = 0 when peer memory is unavailable or switched off (p2p_allreduce = 0)
P2PFold p2pFoldNext();
#ifndef B200_EMU
// executed by the 32 threads of ONE warp (lane = threadIdx.x & 31); data[0..count) in place; op: 0 sum, 1 max, 2 min
__device__ __forceinline__ void p2pAllreduceWarp(double* data, int count, int op, const P2PView& pv, unsigned long long epoch) {
    const int lane = threadIdx.x & 31, slot = (int)(epoch & (P2P_SLOTS - 1));
    const double tag = (double)epoch;
    if (lane < pv.nRanks) {
        volatile double* dst = pv.peer[lane] + ((size_t)slot * pv.nRanks + pv.rank) * P2P_ENTRY;
        for (int i = 0; i < count; i++) dst[i] = ((volatile double*)data)[i];
        __threadfence_system();
        dst[P2P_ENTRY - 1] = tag;
        volatile double* src = pv.mbox + ((size_t)slot * pv.nRanks + lane) * P2P_ENTRY;
        while (src[P2P_ENTRY - 1] != tag) {}
        __threadfence_system();
    }
    __syncwarp();
    if (lane < count) {
        volatile double* base = pv.mbox + (size_t)slot * pv.nRanks * P2P_ENTRY;
        double acc = base[lane];
        for (int r = 1; r < pv.nRanks; r++) {
            double v = base[(size_t)r * P2P_ENTRY + lane];
            acc = op == 0 ? acc + v : op == 1 ? (acc > v ? acc : v) : (acc < v ? acc : v);
        }
        data[lane] = acc;
    }
    __syncwarp();
}
#else
__device__ __forceinline__ void p2pAllreduceWarp(double*, int, int, const P2PView&, unsigned long long) {}
#endif
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
// stream != null: launch there instead of the library stream (the caller orders it against the producer of cellField and against its reuse)
void peerHaloPush(PeerHalo* h, const DeviceMesh& m, const double* cellField, bool pdl = false, cudaStream_t stream = nullptr);
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
