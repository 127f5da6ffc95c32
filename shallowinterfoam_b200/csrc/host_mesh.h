// host_mesh.h -- host-side polyMesh tools of the product (C++): blockMesh-ordered hex box, subsetMesh-style weir,
// primitiveMesh geometry, surfaceInterpolation weights/deltaCoeffs, lduAddressing derivations, simpleGeomDecomp,
// decomposePar sub-meshes, DIC level renumbering.  Integer results are bit-exact against oracle/mesh_oracle.py.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

struct Patch {
    std::string name;
    int32_t start = 0, size = 0;
    int32_t neighbRank = -1;  // >= 0: processor patch
};

struct PolyMesh {
    int32_t nCells = 0;
    std::vector<double> points;                 // [3*nPoints]
    std::vector<std::array<int32_t, 4>> faces;  // quads
    std::vector<int32_t> owner;                 // [nFaces]
    std::vector<int32_t> neighbour;             // [nInternalFaces]
    std::vector<Patch> patches;
    // provenance of a decomposePar sub-mesh (empty for a global mesh)
    std::vector<int32_t> cellMap, faceMap;
    std::vector<uint8_t> faceFlip;
    int32_t nInternalFaces() const { return (int32_t)neighbour.size(); }
    int32_t nFaces() const { return (int32_t)owner.size(); }
    int32_t nBoundaryFaces() const { return nFaces() - nInternalFaces(); }
};

struct Geometry {
    std::vector<double> Sf, magSf, Cf, C, V, weights, deltaCoeffs;
};

// [UPSTREAM blockMesh single-block ordering; SURVEY.md Appendix A.1]
PolyMesh makeBox(int nx, int ny, int nz, double lx, double ly, double lz);
// [UPSTREAM subsetMesh semantics] keep[c] != 0 cells survive; exposed faces -> new last patch
PolyMesh subsetMesh(const PolyMesh& m, const std::vector<uint8_t>& keep, const std::string& exposedName);
PolyMesh agglomerate(const PolyMesh& m, const int64_t* group);
void jitterPoints(PolyMesh& m, double frac, double dx, double dy, double dz, uint64_t seed);
// affine shear of every point: x += sxy*y + sxz*z ; y += syz*z.  Faces stay planar and face centres stay on the line between the
// cell centres (no skewness), but Sf is no longer parallel to d: a purely NON-ORTHOGONAL mesh
void shearPoints(PolyMesh& m, double sxy, double sxz, double syz);
std::vector<uint8_t> weirKeepMask(int nx, int ny, int nz, double x0, double x1, double zfrac);

// [UPSTREAM primitiveMeshFaceCentresAndAreas.C, primitiveMeshCellCentresAndVols.C, surfaceInterpolation.C]
Geometry computeGeometry(const PolyMesh& m);

// [UPSTREAM lduAddressing.C calcOwnerStart / calcLosort / calcLosortStart]
void lduDerived(int32_t nCells, const std::vector<int32_t>& lower, const std::vector<int32_t>& upper,
                std::vector<int32_t>& ownerStart, std::vector<int32_t>& losort, std::vector<int32_t>& losortStart);

// [UPSTREAM simpleGeomDecomp::decompose, geomDecomp rotDelta; SURVEY.md Appendix A.8]
std::vector<int32_t> decomposeSimple(const std::vector<double>& C, int32_t nCells, int npx, int npy, int npz,
                                     double delta = 0.001);
// [UPSTREAM decomposePar domainDecomposition: processor-patch construction]
PolyMesh subMesh(const PolyMesh& m, const std::vector<int32_t>& proc, int nRanks, int rank);

// forward DIC level of every cell (0 = no lower neighbour) and the device row order (level, original index)
void levelRenumbering(int32_t nCells, const int32_t* lower, const int32_t* upper, int32_t nFaces,
                      std::vector<int32_t>& level, std::vector<int32_t>& rowCell, std::vector<int32_t>& cellRow,
                      int32_t& nLevels);

// DIC sweep plan: CLOSED tiles on a tile DAG.
//  * three integer potentials per cell, each non-decreasing along every dependency l -> u (l < u):
//      level  = longest path from a source (the classic wavefront number),
//      potJ/K = largest number of "stride class 1 / class 2" faces on any path to the cell.  The stride classes come from
//               the histogram of (upper - lower): its three dominant magnitudes (1, nx, nx*ny on a blockMesh-ordered hex block,
//               also after subsetMesh) => potJ = j, potK = k on such meshes; on any other mesh they are still monotone.
//  * tile class = (level / bandLevels, potJ / binJ, potK / binK).  Because every component is monotone, the tile graph is acyclic
//    for ANY mesh; classes larger than the row / shared-memory budget are cut into consecutive pieces of their (level, index) order,
//    which keeps it acyclic.  A tile's external dependencies all lie in tiles that precede it in the ticket order (tiles sorted by
//    their longest-path level in the tile graph), so a CTA waits once, then sweeps its local levels out of shared memory.
//  * device row order = (tile in ticket order, level, original index).
struct SweepPlan {
    // parameters used
    int bandLevels = 16, binJ = 1, binK = 1, maxRows = 2048, smemBudget = 110 * 1024, lanes = 128;
    int stride[3] = {0, 0, 0};                      // detected stride-class centres (0 = absent)
    std::vector<int32_t> level, potJ, potK;         // per cell
    int32_t nLevels = 0;
    // CHUNKED plans (some row has more than 3 lower or more than 3 upper neighbours: polyhedral cells).  Such a row is applied as a chain
    // of CHUNKS of at most 3 consecutive sources, each chunk one (step, lane) entry that starts from the previous chunk's partial result --
    // the arithmetic order inside the row is untouched, every entry has at most 3 sources (the register-resident fast path of the
    // kernels), and a heavy row's early sources are consumed while the later ones are still being computed.  xlevel = the row's level in
    // that expanded graph (== level when nothing is chunked); the bands are cut on xlevel.
    bool chunked = false;
    int indexBins = 0;                              // > 0: the stride-class potentials fragmented; potK = cell index / indexBins, potJ = 0
    std::vector<int32_t> xlevel;
    int32_t nXLevels = 0;
    std::vector<int32_t> rowCell, cellRow;          // device row <-> original cell
    int nTiles = 0, nTileLevels = 0;
    std::vector<int32_t> tileStart;                 // [nTiles+1] first row
    std::vector<int32_t> tileLevel;                 // [nTiles] longest-path level of the tile in the tile graph
    // a tile is executed, per sweep direction, as STEPS of at most `lanes` entries (one per thread, a barrier after every step); step k
    // of tile t owns the operand slots [k*lanes, (k+1)*lanes) of the tile's operand block (unused lanes are inert).  Unchunked plans:
    // step = up to `lanes` rows of one level, the same steps in both directions (walked downwards by the backward sweep).
    std::vector<int32_t> fwdStepStart, bwdStepStart;   // [nTiles+1] steps of the tiles before t (S = nSteps*lanes slots)
    // own: per slot, the tile-local row the entry belongs to (SWEEP_OWN_NONE = inert lane) plus the flags
    //   SWEEP_OWN_PARTIAL : not the row's last chunk (the result must not be published)
    //   SWEEP_OWN_CARRY   : continues the chunk this lane computed in the previous step (start value = that result, still in a register)
    std::vector<uint16_t> fwdOwn, bwdOwn;           // [fwdStepStart.back()*lanes], [bwdStepStart.back()*lanes]
    // per-tile ELL of the two sweeps in SLOT layout: column j of tile t = entries [ell[t] + j*S, +S), S = nSteps*lanes.
    // idx: source in the tile's value array (0..nRows-1 = row of this tile, nRows+h = halo entry h, nRows+nHalo = padding);
    // face: ORIGINAL face index of the coefficient (-1 = padding).  Forward columns are in ascending face order, backward
    // columns in descending face order = the order upstream's sequential loops apply them to the row.
    std::vector<int32_t> fwdEll, fwdD, fwdHalo;     // [nTiles+1] offsets, [nTiles] widths, [nTiles+1] halo offsets
    std::vector<int32_t> bwdEll, bwdD, bwdHalo;
    std::vector<uint16_t> fwdIdx, bwdIdx;           // [fwdEll.back()], [bwdEll.back()]
    std::vector<int32_t> fwdFace, bwdFace;
    std::vector<int32_t> fwdHaloRow, bwdHaloRow;    // device rows polled by the tile
    size_t smemFwd = 0, smemBwd = 0;                // largest per-tile shared-memory need (bytes)
};
constexpr uint16_t SWEEP_OWN_NONE = 0xffff, SWEEP_OWN_PARTIAL = 0x8000, SWEEP_OWN_CARRY = 0x4000, SWEEP_OWN_ROW = 0x1fff;
// shared-memory bytes of one tile (the formula the kernels use): the value array (rows, start values overwritten in place by the
// results; halo; padding; trash; +1 alignment shift) and the operand block: d coefficient columns and ceil(d/4) columns of
// packed 16-bit offsets
inline size_t sweepSmemBytes(int nRows, int nSlots, int nHalo, int d) {
    return 8 * ((size_t)((nRows + nHalo + 3 + 1) / 2 * 2)) + 8 * (size_t)nSlots * (size_t)(d + (d + 3) / 4) + 64;
}
constexpr int SWEEP_MAX_SLOTS = 8191;   // nRows + nHalo + 2 (sources are stored as 16-bit BYTE offsets into the value array)
constexpr int SWEEP_MAX_STEPS = 256;    // steps per tile (staged in shared memory)
void buildSweepPlan(int32_t nCells, const int32_t* lower, const int32_t* upper, int32_t nFaces, int bandLevels, int maxRows,
                    int smemBudget, int lanes, SweepPlan& plan);

}  // namespace b200
