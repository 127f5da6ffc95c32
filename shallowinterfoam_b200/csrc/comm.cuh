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
bool commUsesPeerMemory();   // scalar all-reduces run over NVLink peer memory (mailboxes) instead of NCCL
// exchange per-boundary-face values on processor patches: recv[b] = neighbour rank's send[b'] (same face)
void haloExchange(const DeviceMesh& m, const double* send, double* recv);
bool meshHasProcessorPatches(const DeviceMesh& m);
// send[b] = cellField[faceCell row of b] on processor faces
void haloPack(const DeviceMesh& m, const double* cellField, double* send);
}  // namespace b200
