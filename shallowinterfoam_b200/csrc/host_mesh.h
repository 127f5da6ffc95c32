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

// DIC tiling: cells are grouped into bands of `bandLevels` consecutive forward levels; inside a band the
// cells, in ascending original index, are cut into chunks of `tileRows`; tile = (band, chunk), numbered band-major.  Every
// dependency l -> u (l < u, level(l) < level(u)) goes to the same or a LATER tile, so tiles processed in order never deadlock.
// Device row order = (tile, level, original index).  tileSegStart/segRow: CSR of the per-tile level segments (segments longer
// than `maxSeg` rows are split); segRow has one extra entry (= nCells).
void tileRenumbering(int32_t nCells, const int32_t* lower, const int32_t* upper, int32_t nFaces, int bandLevels, int tileRows,
                     int maxSeg, std::vector<int32_t>& level, std::vector<int32_t>& rowCell, std::vector<int32_t>& cellRow,
                     int32_t& nLevels, std::vector<int32_t>& tileStart, std::vector<int32_t>& tileSegStart,
                     std::vector<int32_t>& segRow);

}  // namespace b200
