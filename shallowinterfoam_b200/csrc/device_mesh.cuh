// device_mesh.cuh -- the device-resident mesh: renumbered rows + "owner-slot SELL-32" face layout.
//
// Layout:
//  * rows   : cells renumbered by forward DIC level (level-major, original index within a level).  All cell
//             fields live in row order; vectors are SoA (x[N], y[N], z[N]).
//  * slices : 32 consecutive rows = one warp.
//  * faces  : every internal face is stored ONCE, with its ORIGINAL owner's row, in a sliced-ELL slot:
//             pos = sliceOff[s] + k*32 + lane   (k = index of the face among the owner's faces, ascending
//             original face id).  Face fields (upper, phi, weights, Sf ...) are [Fs + B]: slots, then the B
//             boundary faces in original order.  upIdx[pos] = row of the face's neighbour cell (-1 = padding).
//  * nbr side: each row also lists the faces where it is the original NEIGHBOUR, ascending original face id,
//             as packed entries (ownerRow << KB | k) in a second sliced-ELL (nbrOff); -1 = padding.
//             => a row gathers  [nbr-side faces asc] + [owner faces asc] + [boundary faces asc], which is
//             exactly the per-cell summation order of upstream's sequential face loops, with no atomics.
//  * boundary: per-slice bit mask of rows owning boundary faces + CSR of their boundary faces.
#pragma once
#include <memory>
#include "common.cuh"
#include "host_mesh.h"

namespace b200 {

struct MeshView {  // POD handed to kernels by value
    int N, F, B, nSlices, KB, Fs;
    int uniOw, uniNw;      // > 0: every slice has this many owner / neighbour-side slots per row (positions are arithmetic)
    const int* sliceOff;   // [nSlices+1]
    const int* nbrOff;     // [nSlices+1]
    const int* upIdx;      // [Fs]
    const int* loPk;       // [Ls]
    const unsigned* bMask; // [nSlices]
    const int* bBase;      // [nSlices]
    const int* bRowStart;  // [nBRows+1]
    const int* bRowFaces;  // [B]
    const int* bFaceRow;   // [B] row of the face cell
    const int* bPatch;     // [B] patch index
    const int* bCoupled;   // [B] 1 = processor face
    // geometry (null when created through b200_ldu_get_or_create)
    const double* V;                 // [N]
    const double *Cx, *Cy, *Cz;      // [N]
    const double *Sfx, *Sfy, *Sfz;   // [Fs+B]
    const double *magSf, *w, *dc;    // [Fs+B]  (processor faces: internal-face weights / deltaCoeffs, see coupleGeometry)
    const double *Cnx, *Cny, *Cnz;   // [B] neighbour cell centre of processor faces
    // non-orthogonal correction vectors  corrVec = Sf/magSf - d*deltaCoeffs  (d = C[N] - C[P]; 0 on physical boundary faces; processor
    // faces as the internal faces they were) [UPSTREAM surfaceInterpolation.C makeCorrectionVectors]
    const double *cvx, *cvy, *cvz;   // [Fs+B]
};

// slot of the face a packed neighbour-side reference names; uniform slice widths make it arithmetic (no look-up of the owner's slice)
__device__ __forceinline__ int facePosOf(const MeshView& mv, int pk, int& l) {
    l = pk >> mv.KB;
    int k = pk & ((1 << mv.KB) - 1);
    if (mv.uniOw > 0) return (((l >> 5) * mv.uniOw + k) << 5) + (l & 31);
    return mv.sliceOff[l >> 5] + (k << 5) + (l & 31);
}

constexpr int SWEEP_THREADS = 256;    // most threads (= lanes) of a sweep CTA
// DIC sweep tiles (host_mesh.h SweepPlan, uploaded).  One descriptor per tile and sweep direction.
struct SweepTile {
    int row0, nRows;       // device rows [row0, row0+nRows)
    int slot0, nSteps;     // first slot of the tile in fwdSlotRow / bwdSlotRow; steps of this direction: S = nSteps*lanes slots
    int blk, d;            // operand block offset (doubles) and ELL width (0..3); block = d coefficient columns [S] followed by
                           // one column of 4 packed 16-bit BYTE offsets into the tile's value array: 3 sources + the entry's own row
    int halo, nHalo;       // halo list offset and length
    int virt;              // chunked plan (host_mesh.h): an entry may be one CHUNK of a row; its own-row offset carries OWN_PARTIAL / OWN_CARRY
    int bulkLo, bulkHi;    // hybrid publish: local rows of the first SWEEP_BULK_STEPS processing steps (bulkHi == 0: not eligible)
    unsigned stepWarps[2]; // 4 bits per step (tiles of <= 16 steps): warps that hold entries in that step; the others skip their operand loads
    int pad[3];
};
// flag bits in an entry's own-row byte offset (rows are 8 bytes apart, so the low three bits are free)
constexpr unsigned OWN_PARTIAL = 1u;   // not the last chunk of its row: the result stays in the tile
constexpr unsigned OWN_CARRY = 2u;     // continues the chunk this thread computed in the previous step: start value = that result
struct SweepView {
    int nTiles;
    const SweepTile *fwd, *bwd;              // [nTiles] in ticket order
    const int *fwdSlotRow, *bwdSlotRow;      // device row of each slot's entry (-1 = inert lane)
    int lanes;                               // rows per step = threads of a sweep CTA
    const int *fwdPos, *bwdPos;              // face slot of each coefficient entry of the operand blocks (-1 = padding)
    const int *fwdHaloRow, *bwdHaloRow;      // device rows a tile polls
    int pollNs;                              // back-off between polls of a not-yet-published halo value
    int regOps;                              // 1 (default): tiles with <= 3 sources per row and <= 16 steps keep their operands in registers (dic_kernels.cuh)
    int bulkTail;                            // publishMode 2: the last `bulkTail` steps of a tile store their rows in the loop, the earlier ones leave by one TMA bulk store
    int publishMode;                         // 2: hybrid (see bulkTail); 0: one TMA bulk store per tile after its last step; 1 (default): rows are stored in the step that computes them
    unsigned long long* trace;               // debug: [8*nTiles] globaltimer stamps per tile (ticket, staged, halo ready, steps done, published) or null
};

struct PcgWorkspace;  // pcg.cu

struct DeviceMesh {
    int N = 0, F = 0, B = 0, nSlices = 0, KB = 0, Fs = 0, Ls = 0, nLevels = 0, nBRows = 0, uniOw = 0, uniNw = 0;
    bool hasGeometry = false;
    std::vector<Patch> patches;            // start = boundary-face offset + F (absolute), as given
    std::vector<int32_t> lower, upper;     // original LDU addressing (host)
    std::vector<int32_t> faceCells;        // [B]
    std::vector<int32_t> level, rowCell, cellRow;
    std::vector<int32_t> ownerStart, losort, losortStart;
    std::vector<int32_t> facePosHost;      // [F] slot of original face
    std::vector<int32_t> levelStart;       // [nLevels+1] number of cells below each level (diagnostic)
    // DIC sweep plan (host_mesh.h buildSweepPlan): rows are ordered (tile in ticket order, level, original index)
    SweepPlan plan;
    int nTiles = 0;
    size_t sweepSmem = 0;                  // dynamic shared memory of the sweep kernels (bytes)
    DevBuf<SweepTile> dTilesF, dTilesB;
    DevBuf<int> dFwdSlotRow, dBwdSlotRow, dFwdPos, dBwdPos, dFwdHaloRow, dBwdHaloRow;
    DevBuf<double> dFwdBlk0, dBwdBlk0;     // operand-block templates: coefficient columns 0, source-offset columns filled
    size_t fwdBlkSize = 0, bwdBlkSize = 0; // doubles
    DevBuf<unsigned long long> dTrace;     // debug trace of the last sweep (option sweep_trace)
    DevBuf<int> dSliceOff, dNbrOff, dUpIdx, dLoPk, dBBase, dBRowStart, dBRowFaces, dBFaceRow, dBPatch, dBCoupled;
    DevBuf<unsigned> dBMask;
    DevBuf<int> dSlotFace, dFacePos, dRowCell, dCellRow;
    DevBuf<double> dCnx, dCny, dCnz;
    DevBuf<double> dCvx, dCvy, dCvz;
    // mesh.orthogonal() [UPSTREAM surfaceInterpolation::makeCorrectionVectors]: asin(min(sum(magSf*|corrVec|)/sum(magSf), 1)) < 0.1 degrees
    // over the internal faces of ALL ranks; `corrected` snGrad / laplacian schemes are no-ops on an orthogonal mesh
    double nonOrthSum[2] = {0.0, 0.0};    // this rank's sum(magSf*|corrVec|), sum(magSf)
    bool orthogonal = true;
    DevBuf<double> dV, dCx, dCy, dCz, dSfx, dSfy, dSfz, dMagSf, dW, dDc, dCfx, dCfy, dCfz;
    // staging buffers for the host-pointer API
    DevBuf<double> stageA, stageB;
    PcgWorkspace* pcg = nullptr;
    std::shared_ptr<void> fvHolder;        // FvWorkspace (region3d.cu)
    std::vector<double> hostV;             // cell volumes, original order (deltaN of interfaceProperties)
    const void* key = nullptr;
    uint64_t addrHash = 0;                 // hash of the addressing this cache was built from (api.cu registry validation)

    ~DeviceMesh();
    void build(int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
               const std::vector<Patch>& patches, const std::vector<int32_t>& faceCells);
    void setGeometry(const double* Sf, const double* magSf, const double* Cf, const double* C, const double* V,
                     const double* w, const double* dc);
    MeshView view() const;
    SweepView sweep() const {
        return SweepView{nTiles, dTilesF.p, dTilesB.p, dFwdSlotRow.p, dBwdSlotRow.p, plan.lanes, dFwdPos.p, dBwdPos.p, dFwdHaloRow.p, dBwdHaloRow.p, (int)rt().opt("sweep_poll_ns", 0), (int)rt().opt("sweep_reg_ops", 1), (int)rt().opt("sweep_bulk_tail", 4), publishModeOf(),
                         rt().opt("sweep_trace", 0) > 0 ? dTrace.p : nullptr};
    }
    // how a tile's rows reach global memory: option sweep_publish (0 bulk store per tile, 1 in the step loop, 2 hybrid); default: hybrid
    // once the sweep is throughput-bound (more than 8 tiles per SM: 7.88 M cells 592 vs 630 us per precondition), in-loop below that
    // (985k cells: 162.6 vs 167.5 us) -- profiles/r2_ncu_summary.md
    int publishModeOf() const {
        const int o = (int)rt().opt("sweep_publish", -1);
        return o >= 0 ? o : (nTiles > 8 * rt().numSMs ? 2 : 1);
    }
    void coupleGeometry();   // collective: processor-face weights/deltaCoeffs/neighbour centres from a halo swap of C
    size_t nFaceField() const { return (size_t)Fs + B; }

    // ---- host <-> device layout conversion (all on rt().stream) -------------------------------------------
    void uploadCells(const double* host, double* dev);                 // scalar [N] original order -> rows
    void downloadCells(const double* dev, double* host);
    void uploadCellVec(const double* hostXYZ, double* dx, double* dy, double* dz);  // AoS [3N] -> SoA rows
    void downloadCellVec(const double* dx, const double* dy, const double* dz, double* hostXYZ);
    void uploadFaces(const double* host, double* dev, bool withBoundary);  // [F(+B)] -> [Fs(+B)]
    void downloadFaces(const double* dev, double* host, bool withBoundary);
    void uploadFaceVec(const double* hostXYZ, double* dx, double* dy, double* dz);  // [3(F+B)] AoS -> SoA slots+B
    void uploadBoundary(const double* host, double* dev, int nc = 1);  // [nc*B] plain copy
    void downloadBoundary(const double* dev, double* host, int nc = 1);
    double* stage(DevBuf<double>& buf, size_t n) { buf.ensure(n); return buf.p; }
};

double nonOrthogonalityDegrees(double sumMagSfCorr, double sumMagSf);
void fillSentinel(double* p, size_t n);
void fillValue(double* p, size_t n, double v);

}  // namespace b200
