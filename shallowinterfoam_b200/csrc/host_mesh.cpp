// host_mesh.cpp -- see host_mesh.h.  Plain host C++ (no CUDA).  Compiled with -ffp-contract=off.
#include "host_mesh.h"
#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace b200 {

static const double VSMALL = 1.0e-300;

// ------------------------------------------------------------------------------------------------ box
// cells c = i + nx*(j + ny*k); points v = i + (nx+1)*(j + (ny+1)*k); internal faces in upper-triangular
// order (cell ascending, then neighbour c+1, c+nx, c+nx*ny); patches inlet(x-), outlet(x+),
// walls(y-,y+,z-), atmosphere(z+); side loops  x: for k for j,  y: for i for k,  z: for i for j.
PolyMesh makeBox(int nx, int ny, int nz, double lx, double ly, double lz) {
    PolyMesh m;
    m.nCells = nx * ny * nz;
    auto V = [=](int i, int j, int k) { return (int32_t)(i + (nx + 1) * (j + (ny + 1) * k)); };
    m.points.resize(3 * (size_t)(nx + 1) * (ny + 1) * (nz + 1));
    for (int k = 0; k <= nz; k++)
        for (int j = 0; j <= ny; j++)
            for (int i = 0; i <= nx; i++) {
                size_t v = V(i, j, k);
                m.points[3 * v + 0] = lx * ((double)i / (double)nx);
                m.points[3 * v + 1] = ly * ((double)j / (double)ny);
                m.points[3 * v + 2] = lz * ((double)k / (double)nz);
            }
    size_t nInt = (size_t)(nx - 1) * ny * nz + (size_t)nx * (ny - 1) * nz + (size_t)nx * ny * (nz - 1);
    m.faces.reserve(nInt + 2 * ((size_t)ny * nz + (size_t)nx * nz + (size_t)nx * ny));
    m.owner.reserve(m.faces.capacity());
    m.neighbour.reserve(nInt);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                int32_t c = i + nx * (j + ny * k);
                if (i < nx - 1) {
                    m.faces.push_back({V(i + 1, j, k), V(i + 1, j + 1, k), V(i + 1, j + 1, k + 1), V(i + 1, j, k + 1)});
                    m.owner.push_back(c); m.neighbour.push_back(c + 1);
                }
                if (j < ny - 1) {
                    m.faces.push_back({V(i, j + 1, k), V(i, j + 1, k + 1), V(i + 1, j + 1, k + 1), V(i + 1, j + 1, k)});
                    m.owner.push_back(c); m.neighbour.push_back(c + nx);
                }
                if (k < nz - 1) {
                    m.faces.push_back({V(i, j, k + 1), V(i + 1, j, k + 1), V(i + 1, j + 1, k + 1), V(i, j + 1, k + 1)});
                    m.owner.push_back(c); m.neighbour.push_back(c + nx * ny);
                }
            }
    auto addSide = [&](int side) {
        switch (side) {
            case 0:  // xmin
                for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) {
                    m.faces.push_back({V(0, j, k), V(0, j, k + 1), V(0, j + 1, k + 1), V(0, j + 1, k)});
                    m.owner.push_back(0 + nx * (j + ny * k)); }
                break;
            case 1:  // xmax
                for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) {
                    m.faces.push_back({V(nx, j, k), V(nx, j + 1, k), V(nx, j + 1, k + 1), V(nx, j, k + 1)});
                    m.owner.push_back((nx - 1) + nx * (j + ny * k)); }
                break;
            case 2:  // ymin
                for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++) {
                    m.faces.push_back({V(i, 0, k), V(i + 1, 0, k), V(i + 1, 0, k + 1), V(i, 0, k + 1)});
                    m.owner.push_back(i + nx * (0 + ny * k)); }
                break;
            case 3:  // ymax
                for (int i = 0; i < nx; i++) for (int k = 0; k < nz; k++) {
                    m.faces.push_back({V(i, ny, k), V(i, ny, k + 1), V(i + 1, ny, k + 1), V(i + 1, ny, k)});
                    m.owner.push_back(i + nx * ((ny - 1) + ny * k)); }
                break;
            case 4:  // zmin
                for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
                    m.faces.push_back({V(i, j, 0), V(i, j + 1, 0), V(i + 1, j + 1, 0), V(i + 1, j, 0)});
                    m.owner.push_back(i + nx * (j + ny * 0)); }
                break;
            default:  // zmax
                for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
                    m.faces.push_back({V(i, j, nz), V(i + 1, j, nz), V(i + 1, j + 1, nz), V(i, j + 1, nz)});
                    m.owner.push_back(i + nx * (j + ny * (nz - 1))); }
        }
    };
    struct Spec { const char* name; std::vector<int> sides; };
    std::vector<Spec> spec = {{"inlet", {0}}, {"outlet", {1}}, {"walls", {2, 3, 4}}, {"atmosphere", {5}}};
    for (auto& s : spec) {
        Patch p; p.name = s.name; p.start = (int32_t)m.faces.size();
        for (int side : s.sides) addSide(side);
        p.size = (int32_t)m.faces.size() - p.start;
        m.patches.push_back(p);
    }
    return m;
}

std::vector<uint8_t> weirKeepMask(int nx, int ny, int nz, double x0, double x1, double zfrac) {
    std::vector<uint8_t> keep((size_t)nx * ny * nz, 1);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                double xc = ((double)i + 0.5) / (double)nx, zc = ((double)k + 0.5) / (double)nz;
                if (xc >= x0 && xc < x1 && zc < zfrac) keep[i + nx * (j + ny * k)] = 0;
            }
    return keep;
}

PolyMesh subsetMesh(const PolyMesh& m, const std::vector<uint8_t>& keep, const std::string& exposedName) {
    PolyMesh s;
    int32_t F = m.nInternalFaces();
    std::vector<int32_t> newId(m.nCells, -1);
    int32_t n = 0;
    for (int32_t c = 0; c < m.nCells; c++) if (keep[c]) newId[c] = n++;
    s.nCells = n;
    s.points = m.points;
    for (int32_t f = 0; f < F; f++)
        if (keep[m.owner[f]] && keep[m.neighbour[f]]) {
            s.faces.push_back(m.faces[f]); s.owner.push_back(newId[m.owner[f]]); s.neighbour.push_back(newId[m.neighbour[f]]);
        }
    for (const Patch& p : m.patches) {
        Patch q = p; q.start = (int32_t)s.faces.size();
        for (int32_t f = p.start; f < p.start + p.size; f++)
            if (keep[m.owner[f]]) { s.faces.push_back(m.faces[f]); s.owner.push_back(newId[m.owner[f]]); }
        q.size = (int32_t)s.faces.size() - q.start;
        s.patches.push_back(q);
    }
    Patch e; e.name = exposedName; e.start = (int32_t)s.faces.size();
    for (int32_t f = 0; f < F; f++) {
        bool ko = keep[m.owner[f]], kn = keep[m.neighbour[f]];
        if (ko == kn) continue;
        auto fv = m.faces[f];
        if (ko) { s.faces.push_back(fv); s.owner.push_back(newId[m.owner[f]]); }
        else { s.faces.push_back({fv[3], fv[2], fv[1], fv[0]}); s.owner.push_back(newId[m.neighbour[f]]); }
    }
    e.size = (int32_t)s.faces.size() - e.start;
    s.patches.push_back(e);
    return s;
}

// ------------------------------------------------------------------------------------------- geometry
Geometry computeGeometry(const PolyMesh& m) {
    Geometry g;
    int32_t N = m.nCells, F = m.nInternalFaces(), nF = m.nFaces();
    g.Sf.assign(3 * (size_t)nF, 0); g.magSf.assign(nF, 0); g.Cf.assign(3 * (size_t)nF, 0);
    g.C.assign(3 * (size_t)N, 0); g.V.assign(N, 0); g.weights.assign(nF, 1.0); g.deltaCoeffs.assign(nF, 0);
    const double* P = m.points.data();
    for (int32_t f = 0; f < nF; f++) {
        const double* p[4];
        for (int q = 0; q < 4; q++) p[q] = P + 3 * (size_t)m.faces[f][q];
        double fC[3];
        for (int k = 0; k < 3; k++) fC[k] = (((p[0][k] + p[1][k]) + p[2][k]) + p[3][k]) / 4.0;
        double sumN[3] = {0, 0, 0}, sumA = 0, sumAc[3] = {0, 0, 0};
        for (int q = 0; q < 4; q++) {
            const double* a = p[q]; const double* b = p[(q + 1) % 4];
            double e1[3], e2[3], c[3], n[3];
            for (int k = 0; k < 3; k++) { c[k] = (a[k] + b[k]) + fC[k]; e1[k] = b[k] - a[k]; e2[k] = fC[k] - a[k]; }
            n[0] = e1[1] * e2[2] - e1[2] * e2[1];
            n[1] = e1[2] * e2[0] - e1[0] * e2[2];
            n[2] = e1[0] * e2[1] - e1[1] * e2[0];
            double an = std::sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
            for (int k = 0; k < 3; k++) { sumN[k] = sumN[k] + n[k]; sumAc[k] = sumAc[k] + an * c[k]; }
            sumA = sumA + an;
        }
        for (int k = 0; k < 3; k++) {
            g.Cf[3 * (size_t)f + k] = ((1.0 / 3.0) * sumAc[k]) / (sumA + VSMALL);
            g.Sf[3 * (size_t)f + k] = 0.5 * sumN[k];
        }
        const double* S = &g.Sf[3 * (size_t)f];
        g.magSf[f] = std::sqrt((S[0] * S[0] + S[1] * S[1]) + S[2] * S[2]);
    }
    std::vector<double> cEst(3 * (size_t)N, 0.0), nCF(N, 0.0);
    for (int32_t f = 0; f < nF; f++) { int32_t c = m.owner[f]; for (int k = 0; k < 3; k++) cEst[3 * (size_t)c + k] += g.Cf[3 * (size_t)f + k]; nCF[c] += 1.0; }
    for (int32_t f = 0; f < F; f++) { int32_t c = m.neighbour[f]; for (int k = 0; k < 3; k++) cEst[3 * (size_t)c + k] += g.Cf[3 * (size_t)f + k]; nCF[c] += 1.0; }
    for (int32_t c = 0; c < N; c++) for (int k = 0; k < 3; k++) cEst[3 * (size_t)c + k] /= nCF[c];
    auto dot3 = [](const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; };
    for (int32_t f = 0; f < nF; f++) {
        int32_t c = m.owner[f];
        double d[3], pc[3];
        for (int k = 0; k < 3; k++) { d[k] = g.Cf[3 * (size_t)f + k] - cEst[3 * (size_t)c + k]; pc[k] = (3.0 / 4.0) * g.Cf[3 * (size_t)f + k] + (1.0 / 4.0) * cEst[3 * (size_t)c + k]; }
        double pv = std::max(dot3(&g.Sf[3 * (size_t)f], d), VSMALL);
        for (int k = 0; k < 3; k++) g.C[3 * (size_t)c + k] += pv * pc[k];
        g.V[c] += pv;
    }
    for (int32_t f = 0; f < F; f++) {
        int32_t c = m.neighbour[f];
        double d[3], pc[3];
        for (int k = 0; k < 3; k++) { d[k] = cEst[3 * (size_t)c + k] - g.Cf[3 * (size_t)f + k]; pc[k] = (3.0 / 4.0) * g.Cf[3 * (size_t)f + k] + (1.0 / 4.0) * cEst[3 * (size_t)c + k]; }
        double pv = std::max(dot3(&g.Sf[3 * (size_t)f], d), VSMALL);
        for (int k = 0; k < 3; k++) g.C[3 * (size_t)c + k] += pv * pc[k];
        g.V[c] += pv;
    }
    for (int32_t c = 0; c < N; c++) { for (int k = 0; k < 3; k++) g.C[3 * (size_t)c + k] /= g.V[c]; g.V[c] = g.V[c] * (1.0 / 3.0); }
    for (int32_t f = 0; f < F; f++) {
        int32_t o = m.owner[f], n = m.neighbour[f];
        double dO[3], dN[3], d[3];
        for (int k = 0; k < 3; k++) { dO[k] = g.Cf[3 * (size_t)f + k] - g.C[3 * (size_t)o + k]; dN[k] = g.C[3 * (size_t)n + k] - g.Cf[3 * (size_t)f + k]; d[k] = g.C[3 * (size_t)n + k] - g.C[3 * (size_t)o + k]; }
        double SfdOwn = std::fabs(dot3(&g.Sf[3 * (size_t)f], dO)), SfdNei = std::fabs(dot3(&g.Sf[3 * (size_t)f], dN));
        g.weights[f] = SfdNei / (SfdOwn + SfdNei);
        g.deltaCoeffs[f] = 1.0 / std::sqrt(dot3(d, d));
    }
    for (int32_t f = F; f < nF; f++) {
        int32_t o = m.owner[f];
        double d[3];
        for (int k = 0; k < 3; k++) d[k] = g.Cf[3 * (size_t)f + k] - g.C[3 * (size_t)o + k];
        g.deltaCoeffs[f] = 1.0 / std::sqrt(dot3(d, d));
    }
    return g;
}

// ----------------------------------------------------------------------------------------- addressing
void lduDerived(int32_t N, const std::vector<int32_t>& lower, const std::vector<int32_t>& upper,
                std::vector<int32_t>& ownerStart, std::vector<int32_t>& losort, std::vector<int32_t>& losortStart) {
    size_t F = lower.size();
    ownerStart.assign(N + 1, 0); losortStart.assign(N + 1, 0); losort.resize(F);
    for (size_t f = 0; f < F; f++) { ownerStart[lower[f] + 1]++; losortStart[upper[f] + 1]++; }
    for (int32_t c = 0; c < N; c++) { ownerStart[c + 1] += ownerStart[c]; losortStart[c + 1] += losortStart[c]; }
    std::vector<int32_t> cur(losortStart.begin(), losortStart.end() - 1);
    for (size_t f = 0; f < F; f++) losort[cur[upper[f]]++] = (int32_t)f;  // stable: ascending face index per cell
}

// -------------------------------------------------------------------------------------- decomposition
static std::vector<int32_t> assignToProcessorGroup(int32_t nItems, int nGroups) {
    std::vector<int32_t> grp(nItems);
    int32_t jump = nItems / nGroups, jumpb = nItems - jump * nGroups, ind = 0;
    for (int j = 0; j < jumpb; j++) for (int32_t k = 0; k < jump + 1; k++) grp[ind++] = j;
    for (int j = jumpb; j < nGroups; j++) for (int32_t k = 0; k < jump; k++) grp[ind++] = j;
    return grp;
}

std::vector<int32_t> decomposeSimple(const std::vector<double>& C, int32_t N, int npx, int npy, int npz, double delta) {
    double d = 1 - 0.5 * delta * delta, d2 = d * d, a = delta, a2 = a * a;
    double R[3][3] = {{d2, -a * d, a}, {a * d - a2 * d, a * a2 + d2, -2 * a * d}, {a * d2 + a2, a * d - a2 * d, d2 - a2}};
    std::vector<int32_t> proc(N, 0);
    int np[3] = {npx, npy, npz}, mult[3] = {1, npx, npx * npy};
    std::vector<double> key(N);
    std::vector<int32_t> order(N);
    for (int dir = 0; dir < 3; dir++) {
        for (int32_t c = 0; c < N; c++) key[c] = (R[dir][0] * C[3 * (size_t)c] + R[dir][1] * C[3 * (size_t)c + 1]) + R[dir][2] * C[3 * (size_t)c + 2];
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return key[x] < key[y]; });
        std::vector<int32_t> g = assignToProcessorGroup(N, np[dir]);
        for (int32_t i = 0; i < N; i++) proc[order[i]] += g[i] * mult[dir];
    }
    return proc;
}

PolyMesh subMesh(const PolyMesh& m, const std::vector<int32_t>& proc, int nRanks, int rank) {
    PolyMesh s;
    int32_t F = m.nInternalFaces();
    std::vector<int32_t> g2l(m.nCells, -1);
    for (int32_t c = 0; c < m.nCells; c++) if (proc[c] == rank) { g2l[c] = s.nCells++; s.cellMap.push_back(c); }
    s.points = m.points;
    auto push = [&](int32_t f, int32_t ownLocal, bool flip) {
        auto fv = m.faces[f];
        if (flip) s.faces.push_back({fv[3], fv[2], fv[1], fv[0]}); else s.faces.push_back(fv);
        s.owner.push_back(ownLocal); s.faceMap.push_back(f); s.faceFlip.push_back(flip ? 1 : 0);
    };
    for (int32_t f = 0; f < F; f++)
        if (proc[m.owner[f]] == rank && proc[m.neighbour[f]] == rank) { push(f, g2l[m.owner[f]], false); s.neighbour.push_back(g2l[m.neighbour[f]]); }
    for (const Patch& p : m.patches) {
        Patch q = p; q.start = (int32_t)s.faces.size();
        for (int32_t f = p.start; f < p.start + p.size; f++) if (proc[m.owner[f]] == rank) push(f, g2l[m.owner[f]], false);
        q.size = (int32_t)s.faces.size() - q.start;
        s.patches.push_back(q);
    }
    for (int q = 0; q < nRanks; q++) {
        if (q == rank) continue;
        Patch pp; pp.name = "procBoundary" + std::to_string(rank) + "to" + std::to_string(q); pp.neighbRank = q; pp.start = (int32_t)s.faces.size();
        for (int32_t f = 0; f < F; f++) {
            int32_t po = proc[m.owner[f]], pn = proc[m.neighbour[f]];
            if (po == rank && pn == q) push(f, g2l[m.owner[f]], false);
            else if (po == q && pn == rank) push(f, g2l[m.neighbour[f]], true);
        }
        pp.size = (int32_t)s.faces.size() - pp.start;
        if (pp.size > 0) s.patches.push_back(pp);
    }
    return s;
}

// ------------------------------------------------------------------------------------ level renumbering
void levelRenumbering(int32_t N, const int32_t* lower, const int32_t* upper, int32_t F, std::vector<int32_t>& level,
                      std::vector<int32_t>& rowCell, std::vector<int32_t>& cellRow, int32_t& nLevels) {
    level.assign(N, 0);
    for (int32_t f = 0; f < F; f++) {  // faces sorted by lower: one ascending pass is a topological sweep
        if (!(lower[f] < upper[f])) throw std::runtime_error("lduAddressing is not upper-triangular (lower >= upper)");
        if (f > 0 && lower[f] < lower[f - 1]) throw std::runtime_error("lduAddressing faces are not sorted by lower");
        int32_t v = level[lower[f]] + 1;
        if (v > level[upper[f]]) level[upper[f]] = v;
    }
    nLevels = 0;
    for (int32_t c = 0; c < N; c++) nLevels = std::max(nLevels, level[c] + 1);
    if (N == 0) nLevels = 0;
    std::vector<int32_t> start(nLevels + 1, 0);
    for (int32_t c = 0; c < N; c++) start[level[c] + 1]++;
    for (int32_t l = 0; l < nLevels; l++) start[l + 1] += start[l];
    rowCell.resize(N); cellRow.resize(N);
    for (int32_t c = 0; c < N; c++) { int32_t r = start[level[c]]++; rowCell[r] = c; cellRow[c] = r; }
}

void tileRenumbering(int32_t N, const int32_t* lower, const int32_t* upper, int32_t F, int bandLevels, int tileRows, int maxSeg,
                     std::vector<int32_t>& level, std::vector<int32_t>& rowCell, std::vector<int32_t>& cellRow, int32_t& nLevels,
                     std::vector<int32_t>& tileStart, std::vector<int32_t>& tileSegStart, std::vector<int32_t>& segRow) {
    std::vector<int32_t> lvRow, lvCell;
    levelRenumbering(N, lower, upper, F, level, lvRow, lvCell, nLevels);   // level[] + validity checks
    int nBands = (nLevels + bandLevels - 1) / bandLevels;
    // cells of each band in ascending original index (counting sort by band keeps index order)
    std::vector<int32_t> bandStart(nBands + 1, 0);
    for (int32_t c = 0; c < N; c++) bandStart[level[c] / bandLevels + 1]++;
    for (int b = 0; b < nBands; b++) bandStart[b + 1] += bandStart[b];
    std::vector<int32_t> byBand(N), cur(bandStart.begin(), bandStart.end() - 1);
    for (int32_t c = 0; c < N; c++) byBand[cur[level[c] / bandLevels]++] = c;
    rowCell.resize(N); cellRow.resize(N);
    tileStart.assign(1, 0); tileSegStart.assign(1, 0); segRow.clear();
    int32_t row = 0;
    std::vector<int32_t> cnt(bandLevels + 1);
    for (int b = 0; b < nBands; b++)
        for (int32_t c0 = bandStart[b]; c0 < bandStart[b + 1]; c0 += tileRows) {
            int32_t c1 = std::min<int32_t>(c0 + tileRows, bandStart[b + 1]);
            // inside the tile: stable counting sort by level (keeps ascending index inside a level)
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int32_t i = c0; i < c1; i++) cnt[level[byBand[i]] - b * bandLevels + 1]++;
            for (int t = 0; t < bandLevels; t++) cnt[t + 1] += cnt[t];
            std::vector<int32_t> pos(cnt.begin(), cnt.end() - 1);
            for (int32_t i = c0; i < c1; i++) { int32_t c = byBand[i]; int32_t r = row + pos[level[c] - b * bandLevels]++; rowCell[r] = c; cellRow[c] = r; }
            for (int t = 0; t < bandLevels; t++)
                for (int32_t s0 = cnt[t]; s0 < cnt[t + 1]; s0 += maxSeg) segRow.push_back(row + s0);
            row += c1 - c0;
            tileStart.push_back(row);
            tileSegStart.push_back((int32_t)segRow.size());
        }
    segRow.push_back(N);
}

}  // namespace b200
