// host_mesh.cpp -- see host_mesh.h.  Plain host C++ (no CUDA).  Compiled with -ffp-contract=off.
#include "host_mesh.h"
#include <map>
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

// polyhedral mesh by agglomeration (include/b200foam.h b200_polymesh_agglomerate)
PolyMesh agglomerate(const PolyMesh& m, const int64_t* group) {
    PolyMesh s;
    const int32_t F = m.nInternalFaces();
    std::map<int64_t, int32_t> idOf;
    std::vector<int32_t> newId(m.nCells);
    for (int32_t c = 0; c < m.nCells; c++) {
        auto it = idOf.find(group[c]);
        if (it == idOf.end()) it = idOf.emplace(group[c], (int32_t)idOf.size()).first;
        newId[c] = it->second;
    }
    s.nCells = (int32_t)idOf.size();
    s.points = m.points;
    struct Rec { int32_t lo, hi, f; bool flip; };
    std::vector<Rec> recs;
    for (int32_t f = 0; f < F; f++) {
        int32_t o = newId[m.owner[f]], n = newId[m.neighbour[f]];
        if (o == n) continue;
        recs.push_back(o < n ? Rec{o, n, f, false} : Rec{n, o, f, true});
    }
    std::sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) {
        return a.lo != b.lo ? a.lo < b.lo : a.hi != b.hi ? a.hi < b.hi : a.f < b.f;
    });
    for (const Rec& r : recs) {
        auto fv = m.faces[r.f];
        if (r.flip) s.faces.push_back({fv[3], fv[2], fv[1], fv[0]}); else s.faces.push_back(fv);
        s.owner.push_back(r.lo); s.neighbour.push_back(r.hi);
    }
    const int32_t shift = F - (int32_t)recs.size();
    for (int32_t f = F; f < m.nFaces(); f++) { s.faces.push_back(m.faces[f]); s.owner.push_back(newId[m.owner[f]]); }
    for (const Patch& p : m.patches) { Patch q = p; q.start = p.start - shift; s.patches.push_back(q); }
    return s;
}

// ------------------------------------------------------------------------------------------- geometry
// vertex jitter of the interior points (include/b200foam.h b200_polymesh_jitter)
void jitterPoints(PolyMesh& m, double frac, double dx, double dy, double dz, uint64_t seed) {
    const size_t nP = m.points.size() / 3;
    std::vector<uint8_t> onBoundary(nP, 0);
    for (int32_t f = m.nInternalFaces(); f < m.nFaces(); f++)
        for (int q = 0; q < 4; q++) onBoundary[m.faces[f][q]] = 1;
    uint64_t x = seed;
    const double d[3] = {dx, dy, dz};
    for (size_t p = 0; p < nP; p++) {
        if (onBoundary[p]) continue;
        for (int k = 0; k < 3; k++) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
            m.points[3 * p + k] = m.points[3 * p + k] + (2.0 * u - 1.0) * frac * d[k];
        }
    }
}

void shearPoints(PolyMesh& m, double sxy, double sxz, double syz) {
    const size_t nP = m.points.size() / 3;
    for (size_t p = 0; p < nP; p++) {
        const double y = m.points[3 * p + 1], z = m.points[3 * p + 2];
        m.points[3 * p + 0] = m.points[3 * p + 0] + (sxy * y + sxz * z);
        m.points[3 * p + 1] = y + syz * z;
    }
}

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

// --------------------------------------------------------------------------------------------- sweep plan
namespace {
struct Piece { int32_t b, e; };   // range in the (class, level, index)-sorted cell order

// chunk boundaries of one row: lv = the (estimated) levels of its sources in application order, -1 = no constraint (halo).  Chunks of at
// most 3 consecutive sources; minimises the level of the last chunk, then the number of chunks.  Returns that level; cuts (optional)
// receives the chunk ends (exclusive source positions, ascending).
int chunkRow(const int* lv, int m, std::vector<int>* cuts, std::vector<int>& f, std::vector<int>& g, std::vector<int>& from) {
    if (cuts) cuts->clear();
    if (m == 0) return 0;
    f.resize(m + 1); g.resize(m + 1); from.resize(m + 1);
    f[0] = -1; g[0] = 0;
    for (int i = 1; i <= m; i++) {
        int bestF = 1 << 30, bestG = 1 << 30, bestK = 1, mx = -1;
        for (int k = 1; k <= 3 && k <= i; k++) {
            mx = std::max(mx, lv[i - k]);
            int v = std::max(f[i - k], mx) + 1, c = g[i - k] + 1;
            if (v < bestF || (v == bestF && c < bestG)) { bestF = v; bestG = c; bestK = k; }
        }
        f[i] = bestF; g[i] = bestG; from[i] = bestK;
    }
    if (cuts) {
        for (int i = m; i > 0; i -= from[i]) cuts->push_back(i);
        std::reverse(cuts->begin(), cuts->end());
    }
    return f[m];
}

// one (step, lane) entry of a tile's sweep
struct Entry { int row; int j0, j1; bool partial, carry; };   // tile-local row, source range [j0, j1) in application order
struct TileSchedule { int nSteps = 0; std::vector<Entry> slot; };   // slot[k*W + lane]; row < 0 = inert
}  // namespace

void buildSweepPlan(int32_t N, const int32_t* lower, const int32_t* upper, int32_t F, int bandLevels, int maxRows, int smemBudget,
                    int lanes, SweepPlan& P) {
    P = SweepPlan();
    P.lanes = std::max(32, std::min(lanes, 1024)) / 32 * 32;
    const int W = P.lanes;
    P.bandLevels = std::max(1, bandLevels);
    P.maxRows = std::max(8, std::min(maxRows, 32768));
    P.smemBudget = smemBudget;
    std::vector<int32_t> lvRow, lvCell;
    levelRenumbering(N, lower, upper, F, P.level, lvRow, lvCell, P.nLevels);   // level[] + validity checks
    P.tileStart.assign(1, 0); P.fwdStepStart.assign(1, 0); P.bwdStepStart.assign(1, 0);
    P.fwdEll.assign(1, 0); P.bwdEll.assign(1, 0); P.fwdHalo.assign(1, 0); P.bwdHalo.assign(1, 0);
    P.potJ.assign(N, 0); P.potK.assign(N, 0); P.rowCell.resize(N); P.cellRow.resize(N);
    P.xlevel = P.level; P.nXLevels = P.nLevels;
    if (N == 0) return;
    std::vector<int32_t> ownerStart, losort, losortStart;
    {
        std::vector<int32_t> lo(lower, lower + F), up(upper, upper + F);
        lduDerived(N, lo, up, ownerStart, losort, losortStart);
    }
    // sources of a cell in the order the sweep applies them: forward = faces with upper == c ascending, backward = faces with lower == c descending
    auto nSrc = [&](int32_t c, int dir) { return dir == 0 ? losortStart[c + 1] - losortStart[c] : ownerStart[c + 1] - ownerStart[c]; };
    auto srcFace = [&](int32_t c, int dir, int j) { return dir == 0 ? losort[losortStart[c] + j] : ownerStart[c + 1] - 1 - j; };
    auto srcCell = [&](int32_t c, int dir, int j) { int32_t f = srcFace(c, dir, j); return dir == 0 ? lower[f] : upper[f]; };
    std::vector<int> dpF, dpG, dpFrom, lvBuf;
    // ---- 0. chunked?  expanded forward levels (cells in ascending index = a topological order)
    long long nChunks[2] = {0, 0};
    for (int32_t c = 0; c < N; c++) {
        if (nSrc(c, 0) > 3 || nSrc(c, 1) > 3) P.chunked = true;
        for (int dir = 0; dir < 2; dir++) nChunks[dir] += std::max(1, (nSrc(c, dir) + 2) / 3);
    }
    if (P.chunked) {
        P.nXLevels = 0;
        for (int32_t c = 0; c < N; c++) {
            const int m = nSrc(c, 0);
            lvBuf.resize(m);
            for (int j = 0; j < m; j++) lvBuf[j] = P.xlevel[srcCell(c, 0, j)];
            P.xlevel[c] = chunkRow(lvBuf.data(), m, nullptr, dpF, dpG, dpFrom);
            P.nXLevels = std::max(P.nXLevels, P.xlevel[c] + 1);
        }
    }
    // chunked plans: half the band height -- the chunk chains roughly double the steps a band of levels needs (most of all in the backward
    // sweep, whose rows have the most sources), and tiles of at most SWEEP_REG_STEPS steps keep the register-resident operands
    if (P.chunked) P.bandLevels = std::max(1, P.bandLevels / 2);
    const int B = P.bandLevels;
    const std::vector<int32_t>& LV = P.xlevel;     // the level the tiling works with
    const int32_t nLV = P.nXLevels;
    // rows of one tile: unchunked = the classic budget; chunked = what fills `lanes` entries per level of a band
    int rowsCap = P.maxRows;
    if (P.chunked) {
        double perRow = (double)std::max(nChunks[0], nChunks[1]) / N;
        rowsCap = std::max(8, std::min(P.maxRows, (int)((double)W * B / perRow)));
    }

    // ---- 1. stride classes of (upper - lower): up to three dominant magnitudes
    int nc = 0; long long cen[3] = {0, 0, 0};
    {
        std::vector<int32_t> d(F);
        for (int32_t f = 0; f < F; f++) d[f] = upper[f] - lower[f];
        std::sort(d.begin(), d.end());
        std::vector<std::pair<int32_t, int32_t>> hist;
        for (int32_t f = 0; f < F; f++) { if (hist.empty() || hist.back().first != d[f]) hist.push_back({d[f], 0}); hist.back().second++; }
        std::vector<char> removed(hist.size(), 0);
        for (int it = 0; it < 3; it++) {
            int best = -1;
            for (size_t i = 0; i < hist.size(); i++)
                if (!removed[i] && (best < 0 || hist[i].second > hist[best].second)) best = (int)i;
            if (best < 0) break;
            long long c = hist[best].first;
            cen[nc++] = c;
            for (size_t i = 0; i < hist.size(); i++) { long long v = hist[i].first; if (!removed[i] && 2 * v > c && v < 2 * c) removed[i] = 1; }
        }
        for (int a = 0; a < nc; a++) for (int b = a + 1; b < nc; b++) if (cen[b] < cen[a]) std::swap(cen[a], cen[b]);
        for (int k = 0; k < nc; k++) P.stride[k] = (int)cen[k];
    }
    auto classOf = [&](long long d) { int c = 0; if (nc > 1 && d * d >= cen[0] * cen[1]) c = 1; if (nc > 2 && d * d >= cen[1] * cen[2]) c = 2; return c; };

    std::vector<int32_t> order(N), root(N);
    std::vector<long long> ckey(N);
    std::vector<Piece> pieces;
    // pieces of the tile classes (band, potK / binK, potJ / binJ): split into connected components (a weir makes two disjoint sets of
    // cells share a class: before and behind it; together they would need two steps per level); cells sorted by (class, component,
    // level, index); components -> pieces of at most rowsCap cells.  Components of one class have no path between them (any connecting
    // tile would have the same key), so the tile graph stays acyclic.
    // lateralCuts: oversized classes (and, in the budget loop, oversized pieces) are cut into consecutive CELL-INDEX ranges instead of level
    // ranges: index ranges of one class are ordered like the index itself (l < u), so the tile graph stays acyclic, and the pieces of a
    // class run side by side (pipelined along the index) instead of one after the other.  Used with the index potential / chunked plans;
    // the stride-class plans of structured meshes keep their level cuts.
    bool lateralCuts = P.chunked;
    auto makePieces = [&](long long extJ, long long extK) {
        const long long nJb = (extJ + P.binJ - 1) / P.binJ, nKb = (extK + P.binK - 1) / P.binK;
        for (int32_t c = 0; c < N; c++)
            ckey[c] = ((((long long)(LV[c] / B) * nKb + P.potK[c] / P.binK) * nJb + P.potJ[c] / P.binJ) * nLV) + LV[c];
        std::iota(root.begin(), root.end(), 0);
        auto find = [&](int32_t x) { while (root[x] != x) { root[x] = root[root[x]]; x = root[x]; } return x; };
        // (chunked plans are list-scheduled, a class of several components costs them nothing: no split, fewer and fuller tiles)
        for (int32_t f = 0; f < F && !P.chunked; f++)
            if (ckey[lower[f]] / nLV == ckey[upper[f]] / nLV) {
                int32_t a = find(lower[f]), b = find(upper[f]);
                if (a != b) { if (a < b) root[b] = a; else root[a] = b; }   // the root is the smallest cell of the component
            }
        for (int32_t c = 0; c < N; c++) root[c] = P.chunked ? 0 : find(c);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
            long long ca = ckey[a] / nLV, cb = ckey[b] / nLV;
            if (ca != cb) return ca < cb;
            if (root[a] != root[b]) return root[a] < root[b];
            return LV[a] < LV[b];
        });
        pieces.clear();
        for (int32_t b0 = 0; b0 < N;) {
            long long cls = ckey[order[b0]] / nLV;
            int32_t rt0 = root[order[b0]];
            int32_t e0 = b0;
            while (e0 < N && ckey[order[e0]] / nLV == cls && root[order[e0]] == rt0) e0++;
            int32_t n = e0 - b0, np = (n + rowsCap - 1) / rowsCap, base = n / np, rem = n % np, at = b0;
            if (lateralCuts && np > 1) {   // cut by cell index (lateral), each piece in (level, index) order
                std::sort(order.begin() + b0, order.begin() + e0);
                for (int q = 0; q < np; q++) {
                    int32_t sz = base + (q < rem ? 1 : 0);
                    std::stable_sort(order.begin() + at, order.begin() + at + sz, [&](int32_t a, int32_t b) { return LV[a] < LV[b]; });
                    pieces.push_back({at, at + sz}); at += sz;
                }
            } else
                for (int q = 0; q < np; q++) { int32_t sz = base + (q < rem ? 1 : 0); pieces.push_back({at, at + sz}); at += sz; }
            b0 = e0;
        }
    };
    // ---- 2. potentials: most class-1 / class-2 faces on any path (faces are sorted by lower => one ascending pass)
    for (int32_t f = 0; f < F; f++) {
        int32_t l = lower[f], u = upper[f];
        int c = classOf((long long)u - l);
        int32_t a = P.potJ[l] + (c == 1 ? 1 : 0), b = P.potK[l] + (c == 2 ? 1 : 0);
        if (a > P.potJ[u]) P.potJ[u] = a;
        if (b > P.potK[u]) P.potK[u] = b;
    }
    auto binsAndPieces = [&]() {
        long long extJ = 0, extK = 0;
        for (int32_t c = 0; c < N; c++) { extJ = std::max<long long>(extJ, P.potJ[c] + 1); extK = std::max<long long>(extK, P.potK[c] + 1); }
        // ---- 3. bin sizes: A = target cells of one (potJ, potK) bin column per level; minimise the number of bins crossed
        long long nTriples = 0;
        {
            std::vector<long long> key(N);
            for (int32_t c = 0; c < N; c++) key[c] = ((long long)LV[c] * extJ + P.potJ[c]) * extK + P.potK[c];
            std::sort(key.begin(), key.end());
            for (int32_t c = 0; c < N; c++) if (c == 0 || key[c] != key[c - 1]) nTriples++;
        }
        long long A = std::max<long long>(1, (long long)rowsCap * nTriples / ((long long)B * N));
        long long bestBj = 1, bestBk = 1, bestCost = -1, bestProd = 0;
        for (long long bj = 1; bj <= std::min(extJ, A); bj++) {
            long long bk = std::min(extK, std::max<long long>(1, A / bj));
            long long cost = (extJ + bj - 1) / bj + (extK + bk - 1) / bk, prod = bj * bk;
            if (bestCost < 0 || cost < bestCost || (cost == bestCost && prod > bestProd)) { bestCost = cost; bestProd = prod; bestBj = bj; bestBk = bk; }
        }
        P.binJ = (int)bestBj; P.binK = (int)bestBk;
        // ---- 4. classes -> connected components -> pieces
        makePieces(extJ, extK);
    };
    binsAndPieces();
    // ---- 4b. the stride-class potentials only describe (deformed) structured numberings; where they shred the mesh (polyhedral cells
    //          renumbered in between), fall back to the one lateral potential every LDU addressing has: the cell index itself (l < u on
    //          every face).  Class = (band, index / indexBins): aligned index ranges, so a tile depends only on tiles with a smaller or equal
    //          band and range, and the tile graph's depth is bands + ranges.
    {
        const long long nBands = (nLV + B - 1) / B;
        const long long ideal = (N + rowsCap - 1) / rowsCap + nBands;
        if ((long long)pieces.size() > 3 * ideal) {
            const size_t before = pieces.size();
            const std::vector<int32_t> keepJ = P.potJ, keepK = P.potK;
            const int keepBj = P.binJ, keepBk = P.binK;
            // ALIGNED ranges (the same cuts in every band: a tile then depends only on tiles with a smaller or equal band AND range, depth =
            // bands + ranges; cuts that differ from band to band add a diagonal dependency per band and double the depth), fine enough that
            // the fullest (band, range) class fits one tile -- the sparse classes become small tiles, which costs SMs, not critical path
            const long long rowsPerBand = std::max<long long>(1, N / nBands);
            long long Q = std::max<long long>(1, (rowsPerBand + rowsCap - 1) / rowsCap);
            std::vector<int32_t> pop;
            for (;;) {
                const long long ib = std::max<long long>(1, (N + Q - 1) / Q);
                pop.assign((size_t)(nBands * ((N + ib - 1) / ib)), 0);
                int32_t mx = 0;
                for (int32_t c = 0; c < N; c++) mx = std::max(mx, ++pop[(size_t)(LV[c] / B) * ((N + ib - 1) / ib) + c / ib]);
                if (mx <= rowsCap || ib <= 64) break;
                Q += std::max<long long>(1, Q / 4);
            }
            P.indexBins = (int)std::max<long long>(1, (N + Q - 1) / Q);
            for (int32_t c = 0; c < N; c++) { P.potJ[c] = 0; P.potK[c] = c / P.indexBins; }
            P.binJ = 1; P.binK = 1;
            const bool keepLat = lateralCuts;
            lateralCuts = true;
            makePieces(1, (N + P.indexBins - 1) / P.indexBins);
            if (pieces.size() >= before) {   // no better: keep the stride classes
                lateralCuts = keepLat;
                P.potJ = keepJ; P.potK = keepK; P.binJ = keepBj; P.binK = keepBk; P.indexBins = 0;
                long long extJ = 0, extK = 0;
                for (int32_t c = 0; c < N; c++) { extJ = std::max<long long>(extJ, P.potJ[c] + 1); extK = std::max<long long>(extK, P.potK[c] + 1); }
                makePieces(extJ, extK);
            }
        }
    }

    // ---- 5. schedule of one piece and one direction -> steps of <= W entries
    std::vector<int32_t> pieceOf(N), localOf(N, -1);
    // unchunked: step = up to W rows of one level (the rows of a piece are sorted by level), the same steps for both directions
    auto scheduleLevels = [&](const Piece& pc, TileSchedule& ts) {
        ts.slot.clear(); ts.nSteps = 0;
        int used = W;   // entries in the current step
        for (int32_t i = pc.b; i < pc.e; i++) {
            bool newLevel = i == pc.b || LV[order[i]] != LV[order[i - 1]];
            if (newLevel || used == W) { ts.slot.resize((size_t)(ts.nSteps + 1) * W, Entry{-1, 0, 0, false, false}); ts.nSteps++; used = 0; }
            ts.slot[(size_t)(ts.nSteps - 1) * W + used++] = Entry{i - pc.b, 0, 0, false, false};
        }
        for (Entry& e : ts.slot) if (e.row >= 0) e.j1 = 0;   // (source ranges are filled by the ELL builder: all sources)
    };
    // chunked: list scheduling of the chunks on W lanes, longest remaining chain first
    struct Node { int row, j0, j1, prev, height, waiting, step, lane; };
    std::vector<Node> nodes; std::vector<int> lastNode, est, cuts, succStart, succ, ready, nextReady;
    auto scheduleChunks = [&](const Piece& pc, int p, int dir, TileSchedule& ts) {
        const int n = pc.e - pc.b;
        nodes.clear(); lastNode.assign(n, -1); est.assign(n, 0);
        // chunk boundaries from in-tile level estimates (topological order: forward = piece order, backward = reversed)
        for (int q = 0; q < n; q++) {
            const int r = dir == 0 ? q : n - 1 - q;
            const int32_t c = order[pc.b + r];
            const int m = nSrc(c, dir);
            lvBuf.resize(m);
            for (int j = 0; j < m; j++) { int32_t o = srcCell(c, dir, j); lvBuf[j] = pieceOf[o] == p ? est[localOf[o]] : -1; }
            est[r] = chunkRow(lvBuf.data(), m, &cuts, dpF, dpG, dpFrom);
            if (m == 0) { nodes.push_back(Node{r, 0, 0, -1, 0, 0, -1, -1}); lastNode[r] = (int)nodes.size() - 1; continue; }
            int j0 = 0, prev = -1;
            for (int e : cuts) { nodes.push_back(Node{r, j0, e, prev, 0, 0, -1, -1}); prev = (int)nodes.size() - 1; j0 = e; }
            lastNode[r] = prev;
        }
        // dependants (CSR): node -> nodes waiting for it (in-tile sources' last chunks, and the previous chunk of the same row)
        const int nn = (int)nodes.size();
        succStart.assign(nn + 1, 0);
        auto forDeps = [&](int v, auto&& fn) {
            const Node& nd = nodes[v];
            if (nd.prev >= 0) fn(nd.prev);
            const int32_t c = order[pc.b + nd.row];
            for (int j = nd.j0; j < nd.j1; j++) { int32_t o = srcCell(c, dir, j); if (pieceOf[o] == p) fn(lastNode[localOf[o]]); }
        };
        for (int v = 0; v < nn; v++) forDeps(v, [&](int d) { succStart[d + 1]++; nodes[v].waiting++; });
        for (int v = 0; v < nn; v++) succStart[v + 1] += succStart[v];
        succ.assign(succStart[nn], 0);
        { std::vector<int> at(succStart.begin(), succStart.end() - 1); for (int v = 0; v < nn; v++) forDeps(v, [&](int d) { succ[at[d]++] = v; }); }
        for (int v = nn - 1; v >= 0; v--)   // nodes were created in topological order
            for (int q = succStart[v]; q < succStart[v + 1]; q++) nodes[v].height = std::max(nodes[v].height, nodes[succ[q]].height + 1);
        auto prio = [&](int a, int b) { return nodes[a].height != nodes[b].height ? nodes[a].height < nodes[b].height : a > b; };   // heap: top = highest, then first created
        ready.clear();
        for (int v = 0; v < nn; v++) if (nodes[v].waiting == 0) ready.push_back(v);
        std::make_heap(ready.begin(), ready.end(), prio);
        ts.slot.clear(); ts.nSteps = 0;
        int done = 0;
        std::vector<int> picked;
        while (done < nn) {
            const int k = ts.nSteps++;
            ts.slot.resize((size_t)ts.nSteps * W, Entry{-1, 0, 0, false, false});
            picked.clear();
            while ((int)picked.size() < W && !ready.empty()) { std::pop_heap(ready.begin(), ready.end(), prio); picked.push_back(ready.back()); ready.pop_back(); }
            Entry* st = &ts.slot[(size_t)k * W];
            // continuations of the previous step keep their lane (the partial result stays in the thread's register)
            for (int v : picked) {
                Node& nd = nodes[v];
                if (nd.prev >= 0 && nodes[nd.prev].step == k - 1) { nd.lane = nodes[nd.prev].lane; st[nd.lane] = Entry{nd.row, nd.j0, nd.j1, false, true}; }
            }
            int free = 0;
            for (int v : picked) {
                Node& nd = nodes[v];
                if (nd.lane < 0) { while (st[free].row >= 0) free++; nd.lane = free; st[free] = Entry{nd.row, nd.j0, nd.j1, false, false}; }
                nd.step = k;
                st[nd.lane].partial = lastNode[nd.row] != v;
            }
            nextReady.clear();
            for (int v : picked) for (int q = succStart[v]; q < succStart[v + 1]; q++) if (--nodes[succ[q]].waiting == 0) nextReady.push_back(succ[q]);
            for (int v : nextReady) { ready.push_back(v); std::push_heap(ready.begin(), ready.end(), prio); }
            done += (int)picked.size();
            if (picked.empty()) throw std::runtime_error("DIC sweep plan: chunk schedule stalled (internal error)");
        }
        if (dir == 1)   // the backward sweep walks a tile's steps downwards: store the schedule last step first
            for (int k = 0; k < ts.nSteps / 2; k++) std::swap_ranges(ts.slot.begin() + (size_t)k * W, ts.slot.begin() + (size_t)(k + 1) * W, ts.slot.begin() + (size_t)(ts.nSteps - 1 - k) * W);
    };
    auto setPieceMaps = [&]() {
        for (size_t p = 0; p < pieces.size(); p++)
            for (int32_t i = pieces[p].b; i < pieces[p].e; i++) { pieceOf[order[i]] = (int32_t)p; localOf[order[i]] = i - pieces[p].b; }
    };
    // ---- 6. shared-memory / index-width budget: offending pieces are halved until they fit
    std::vector<int32_t> stamp(N, -1), stampB(N, -1);
    TileSchedule tsF, tsB;
    for (;;) {
        setPieceMaps();
        std::fill(stamp.begin(), stamp.end(), -1); std::fill(stampB.begin(), stampB.end(), -1);
        std::vector<Piece> next; next.reserve(pieces.size());
        bool changed = false;
        for (size_t p = 0; p < pieces.size(); p++) {
            int n = pieces[p].e - pieces[p].b, dF = 0, dB = 0, hF = 0, hB = 0;
            for (int32_t i = pieces[p].b; i < pieces[p].e; i++) {
                int32_t c = order[i];
                dF = std::max(dF, nSrc(c, 0)); dB = std::max(dB, nSrc(c, 1));
                for (int32_t q = losortStart[c]; q < losortStart[c + 1]; q++) { int32_t l = lower[losort[q]]; if (pieceOf[l] != (int32_t)p && stamp[l] != (int32_t)p) { stamp[l] = (int32_t)p; hF++; } }
                for (int32_t f = ownerStart[c]; f < ownerStart[c + 1]; f++) { int32_t u = upper[f]; if (pieceOf[u] != (int32_t)p && stampB[u] != (int32_t)p) { stampB[u] = (int32_t)p; hB++; } }
            }
            int stepsF, stepsB;
            if (P.chunked) { scheduleChunks(pieces[p], (int)p, 0, tsF); scheduleChunks(pieces[p], (int)p, 1, tsB); stepsF = tsF.nSteps; stepsB = tsB.nSteps; dF = std::min(dF, 3); dB = std::min(dB, 3); }
            else { scheduleLevels(pieces[p], tsF); stepsF = stepsB = tsF.nSteps; }
            size_t need = std::max(sweepSmemBytes(n, stepsF * W, hF, dF), sweepSmemBytes(n, stepsB * W, hB, dB));
            bool tooBig = need > (size_t)smemBudget || n + std::max(hF, hB) + 2 > SWEEP_MAX_SLOTS || std::max(stepsF, stepsB) > SWEEP_MAX_STEPS;
            if (tooBig && n > 1) {
                int32_t mid = pieces[p].b + n / 2;
                // lateral halves keep the piece's chain of steps; when the STEPS are what does not fit (operand slots, step count), cut by level
                const bool byLevel = std::max(stepsF, stepsB) > SWEEP_MAX_STEPS ||
                                     8 * (size_t)std::max(stepsF * (dF + 1), stepsB * (dB + 1)) * W > (size_t)smemBudget * 3 / 4;
                if (lateralCuts && !byLevel) {
                    std::sort(order.begin() + pieces[p].b, order.begin() + pieces[p].e);
                    std::stable_sort(order.begin() + pieces[p].b, order.begin() + mid, [&](int32_t a, int32_t b) { return LV[a] < LV[b]; });
                    std::stable_sort(order.begin() + mid, order.begin() + pieces[p].e, [&](int32_t a, int32_t b) { return LV[a] < LV[b]; });
                }
                next.push_back({pieces[p].b, mid}); next.push_back({mid, pieces[p].e});
                changed = true;
            } else {
                if (tooBig) throw std::runtime_error("DIC sweep plan: a single row exceeds the shared-memory budget (more than ~" + std::to_string(smemBudget / (10 * W)) + " lower or upper neighbours with " + std::to_string(W) + " lanes; lower the option sweep_threads)");
                next.push_back(pieces[p]);
            }
        }
        pieces.swap(next);
        if (!changed) break;
    }
    const int nT = (int)pieces.size();
    setPieceMaps();
    // ---- 7. tile graph, longest-path levels (Kahn), ticket order = (tile level, class order)
    std::vector<long long> edges;
    for (int32_t f = 0; f < F; f++) { int32_t a = pieceOf[lower[f]], b = pieceOf[upper[f]]; if (a != b) edges.push_back(((long long)a << 32) | (unsigned)b); }
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::vector<int32_t> adjStart(nT + 1, 0), indeg(nT, 0), tl(nT, 0);
    for (long long e : edges) { adjStart[(e >> 32) + 1]++; indeg[(int32_t)(e & 0xffffffff)]++; }
    for (int t = 0; t < nT; t++) adjStart[t + 1] += adjStart[t];
    std::vector<int32_t> queue; queue.reserve(nT);
    for (int t = 0; t < nT; t++) if (indeg[t] == 0) queue.push_back(t);
    for (size_t h = 0; h < queue.size(); h++) {
        int32_t t = queue[h];
        for (int32_t q = adjStart[t]; q < adjStart[t + 1]; q++) {   // edges are sorted by source, so edges[q] belongs to t
            int32_t v = (int32_t)(edges[q] & 0xffffffff);
            tl[v] = std::max(tl[v], tl[t] + 1);
            if (--indeg[v] == 0) queue.push_back(v);
        }
    }
    if ((int)queue.size() != nT) throw std::runtime_error("DIC sweep plan: tile graph is cyclic (internal error)");
    std::vector<int32_t> ticket(nT);
    std::iota(ticket.begin(), ticket.end(), 0);
    std::stable_sort(ticket.begin(), ticket.end(), [&](int32_t a, int32_t b) { return tl[a] < tl[b]; });
    // ---- 8. rows: device row order = (tile in ticket order, level, original index).  Chunked plans (whose entries name their row
    //         explicitly) put the rows with the most faces first instead: the SELL-32 slices of the face arrays are as wide as their
    //         widest row, and 24-face cells scattered among hexes would pad every slice to their width (3.5x the coefficient traffic of Amul)
    P.nTiles = nT; P.tileLevel.resize(nT);
    std::vector<int32_t> storeOf(N, -1), store;
    int32_t row = 0;
    for (int t = 0; t < nT; t++) {
        const Piece& pc = pieces[ticket[t]];
        P.tileLevel[t] = tl[ticket[t]];
        P.nTileLevels = std::max(P.nTileLevels, tl[ticket[t]] + 1);
        store.assign(order.begin() + pc.b, order.begin() + pc.e);
        if (P.chunked) std::stable_sort(store.begin(), store.end(), [&](int32_t a, int32_t b) { return nSrc(a, 0) + nSrc(a, 1) > nSrc(b, 0) + nSrc(b, 1); });
        for (size_t i = 0; i < store.size(); i++) { int32_t c = store[i]; P.rowCell[row] = c; P.cellRow[c] = row; storeOf[c] = (int32_t)i; row++; }
        P.tileStart.push_back(row);
    }
    // ---- 9. per-tile steps and ELL (slot layout) + halo lists of the two sweeps
    P.fwdD.resize(nT); P.bwdD.resize(nT);
    std::vector<int32_t> slot(N, -1);
    std::fill(stamp.begin(), stamp.end(), -1);
    for (int t = 0; t < nT; t++) {
        const int p = ticket[t];
        const Piece& pc = pieces[p];
        const int32_t n = pc.e - pc.b;
        for (int dir = 0; dir < 2; dir++) {
            TileSchedule& ts = dir == 0 ? tsF : tsB;
            if (P.chunked) scheduleChunks(pc, p, dir, ts); else if (dir == 0) scheduleLevels(pc, ts);
            const TileSchedule& sc = P.chunked ? ts : tsF;
            std::vector<uint16_t>& idx = dir == 0 ? P.fwdIdx : P.bwdIdx;
            std::vector<int32_t>& face = dir == 0 ? P.fwdFace : P.bwdFace;
            std::vector<int32_t>& haloRow = dir == 0 ? P.fwdHaloRow : P.bwdHaloRow;
            std::vector<int32_t>& ell = dir == 0 ? P.fwdEll : P.bwdEll;
            std::vector<int32_t>& halo = dir == 0 ? P.fwdHalo : P.bwdHalo;
            std::vector<uint16_t>& own = dir == 0 ? P.fwdOwn : P.bwdOwn;
            std::vector<int32_t>& stepStart = dir == 0 ? P.fwdStepStart : P.bwdStepStart;
            const int S = sc.nSteps * W;
            int D = 0;
            for (const Entry& e : sc.slot) if (e.row >= 0) D = std::max(D, P.chunked ? e.j1 - e.j0 : nSrc(order[pc.b + e.row], dir));
            (dir == 0 ? P.fwdD : P.bwdD)[t] = D;
            size_t e0 = idx.size(), o0 = own.size();
            idx.resize(e0 + (size_t)D * S, 0xffff); face.resize(e0 + (size_t)D * S, -1);
            own.resize(o0 + (size_t)S, SWEEP_OWN_NONE);
            int nHalo = 0;
            const int32_t st = 2 * t + dir;
            for (int sl = 0; sl < S; sl++) {
                const Entry& e = sc.slot[sl];
                if (e.row < 0) continue;
                const int32_t c = order[pc.b + e.row];
                own[o0 + sl] = (uint16_t)(storeOf[c] | (e.partial ? SWEEP_OWN_PARTIAL : 0) | (e.carry ? SWEEP_OWN_CARRY : 0));
                const int j0 = P.chunked ? e.j0 : 0, j1 = P.chunked ? e.j1 : nSrc(c, dir);
                for (int j = j0; j < j1; j++) {
                    const int32_t f = srcFace(c, dir, j), o = dir == 0 ? lower[f] : upper[f];
                    int32_t id;
                    if (pieceOf[o] == p) id = storeOf[o];
                    else {
                        if (stamp[o] != st) { stamp[o] = st; slot[o] = nHalo++; haloRow.push_back(P.cellRow[o]); }
                        id = n + slot[o];
                    }
                    idx[e0 + (size_t)(j - j0) * S + sl] = (uint16_t)id; face[e0 + (size_t)(j - j0) * S + sl] = f;
                }
            }
            for (size_t e = e0; e < idx.size(); e++) if (face[e] < 0) idx[e] = (uint16_t)(n + nHalo);
            ell.push_back((int32_t)idx.size()); halo.push_back((int32_t)haloRow.size());
            stepStart.push_back(stepStart.back() + sc.nSteps);
            size_t need = sweepSmemBytes(n, S, nHalo, D);
            if (dir == 0) P.smemFwd = std::max(P.smemFwd, need); else P.smemBwd = std::max(P.smemBwd, need);
        }
    }
    if (P.fwdIdx.size() >= (size_t)1 << 31 || P.bwdIdx.size() >= (size_t)1 << 31) throw std::runtime_error("DIC sweep plan: mesh too large for 32-bit ELL offsets");
}

}  // namespace b200
