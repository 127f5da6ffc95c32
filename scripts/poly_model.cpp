// poly_model.cpp -- offline what-if for the polyhedral sweep plan (no GPU): rows with more than 3 sources are split into chained
// "virtual rows" of at most 3 sources (same arithmetic order), tiles = (band of expanded levels) x (consecutive index chunk).
// Prints source-count histograms, real / expanded level counts, and the modelled critical path of the tile DAG.
//   g++ -O2 -std=c++17 -I shallowinterfoam_b200/csrc scripts/poly_model.cpp shallowinterfoam_b200/csrc/host_mesh.cpp -o /tmp/poly_model
// Usage: poly_model n band maxRows lanes [s_ns o_ns] [hex=0]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "host_mesh.h"
using namespace b200;

static int dpLevel(const std::vector<int>& lv) {   // lv: final levels of the sources in application order (-1 = no constraint)
    const int m = (int)lv.size();
    if (m == 0) return 0;
    std::vector<int> f(m + 1, 0);
    f[0] = -1;
    for (int i = 1; i <= m; i++) {
        int best = 1 << 30, mx = -1;
        for (int k = 1; k <= 3 && k <= i; k++) { mx = std::max(mx, lv[i - k]); best = std::min(best, std::max(f[i - k], mx) + 1); }
        f[i] = best;
    }
    return f[m];
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 96, band = argc > 2 ? atoi(argv[2]) : 16, maxRows = argc > 3 ? atoi(argv[3]) : 4096;
    int W = argc > 4 ? atoi(argv[4]) : 256;
    double s = argc > 5 ? atof(argv[5]) : 100.0, o = argc > 6 ? atof(argv[6]) : 670.0;
    bool hex = argc > 7 && atoi(argv[7]) != 0;
    PolyMesh box = makeBox(n, n, n, 1.0, 1.0, 1.0), m;
    if (hex) m = box;
    else {
        std::vector<int64_t> key((size_t)n * n * n);
        for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
            int I = i / 2, J = j / 2, K = k / 2;
            bool full = 2 * I + 1 < n && 2 * J + 1 < n && 2 * K + 1 < n;
            bool merged = full && ((I + J + K) % 3 != 0);
            int64_t cell = i + (int64_t)n * (j + (int64_t)n * k);
            key[cell] = merged ? (int64_t)n * n * n + (I + (int64_t)(n / 2 + 1) * (J + (int64_t)(n / 2 + 1) * K)) : cell;
        }
        m = agglomerate(box, key.data());
    }
    const int N = m.nCells, F = m.nInternalFaces();
    std::vector<int32_t> lo(m.owner.begin(), m.owner.begin() + F), up(m.neighbour.begin(), m.neighbour.end());
    std::vector<int32_t> ownerStart, losort, losortStart;
    lduDerived(N, lo, up, ownerStart, losort, losortStart);
    // histograms
    std::vector<long long> hF(32, 0), hB(32, 0);
    for (int c = 0; c < N; c++) { hF[std::min(31, losortStart[c + 1] - losortStart[c])]++; hB[std::min(31, ownerStart[c + 1] - ownerStart[c])]++; }
    printf("N %d F %d (%.2f faces/cell)\nfwd sources:", N, F, 2.0 * F / N);
    for (int d = 0; d < 32; d++) if (hF[d]) printf(" %d:%lld", d, hF[d]);
    printf("\nbwd sources:");
    for (int d = 0; d < 32; d++) if (hB[d]) printf(" %d:%lld", d, hB[d]);
    printf("\n");
    // real and expanded forward levels
    std::vector<int> lev(N, 0), Lf(N, 0);
    long long virtF = 0, virtB = 0;
    for (int c = 0; c < N; c++) {
        std::vector<int> lv;
        int r = 0;
        for (int q = losortStart[c]; q < losortStart[c + 1]; q++) { int l = lo[losort[q]]; lv.push_back(Lf[l]); r = std::max(r, lev[l] + 1); }
        lev[c] = r; Lf[c] = dpLevel(lv);
        virtF += std::max(1, ((int)lv.size() + 2) / 3);
        virtB += std::max(1, (ownerStart[c + 1] - ownerStart[c] + 2) / 3);
    }
    // expanded backward levels (for information)
    std::vector<int> Lb(N, 0);
    for (int c = N - 1; c >= 0; c--) {
        std::vector<int> lv;
        for (int f = ownerStart[c + 1] - 1; f >= ownerStart[c]; f--) lv.push_back(Lb[up[f]]);
        Lb[c] = dpLevel(lv);
    }
    int nLev = *std::max_element(lev.begin(), lev.end()) + 1, nLf = *std::max_element(Lf.begin(), Lf.end()) + 1, nLb = *std::max_element(Lb.begin(), Lb.end()) + 1;
    printf("levels: real %d, expanded fwd %d, expanded bwd %d ; virtual rows per row fwd %.2f bwd %.2f\n", nLev, nLf, nLb, (double)virtF / N, (double)virtB / N);
    // tiles: (band of Lf, index chunk)
    std::vector<int> order(N);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return Lf[a] / band < Lf[b] / band; });
    std::vector<int> tileOf(N), tileB, tileE;
    for (int b0 = 0; b0 < N;) {
        int e0 = b0;
        while (e0 < N && Lf[order[e0]] / band == Lf[order[b0]] / band) e0++;
        int cnt = e0 - b0, np = (cnt + maxRows - 1) / maxRows, base = cnt / np, rem = cnt % np, at = b0;
        for (int q = 0; q < np; q++) { int sz = base + (q < rem ? 1 : 0); tileB.push_back(at); tileE.push_back(at + sz); at += sz; }
        b0 = e0;
    }
    const int nT = (int)tileB.size();
    for (int t = 0; t < nT; t++) for (int i = tileB[t]; i < tileE[t]; i++) tileOf[order[i]] = t;
    // per tile: in-tile expanded steps (fwd, bwd), lanes, halo
    std::vector<int> stF(nT), stB(nT), haloF(nT), haloB(nT);
    std::vector<int> loc(N, 0), stamp(N, -1);
    long long slotsF = 0, slotsB = 0, over16F = 0, over16B = 0; int maxHalo = 0;
    for (int t = 0; t < nT; t++) {
        std::vector<int> cells(order.begin() + tileB[t], order.begin() + tileE[t]);
        std::sort(cells.begin(), cells.end());
        // forward: virtual rows placed by list scheduling on W lanes
        for (int dir = 0; dir < 2; dir++) {
            std::vector<int> used;   // lanes used per step
            int nh = 0;
            auto place = [&](int earliest) { int k = earliest; for (;; k++) { if ((int)used.size() <= k) used.resize(k + 1, 0); if (used[k] < W) { used[k]++; return k; } } };
            if (dir == 1) std::reverse(cells.begin(), cells.end());
            for (int c : cells) {
                std::vector<int> src;
                if (dir == 0) for (int q = losortStart[c]; q < losortStart[c + 1]; q++) src.push_back(lo[losort[q]]);
                else for (int f = ownerStart[c + 1] - 1; f >= ownerStart[c]; f--) src.push_back(up[f]);
                std::vector<int> lv;
                for (int x : src) { if (tileOf[x] == t) lv.push_back(loc[x]); else { lv.push_back(-1); if (stamp[x] != 2 * t + dir) { stamp[x] = 2 * t + dir; nh++; } } }
                // DP for the chunk boundaries, then place the chunks (each at >= its earliest step, after its predecessor)
                const int mS = (int)lv.size();
                if (mS == 0) { loc[c] = place(0); continue; }
                std::vector<int> f(mS + 1, 0), cut(mS + 1, 0);
                f[0] = -1;
                for (int i = 1; i <= mS; i++) {
                    int best = 1 << 30, mx = -1, bk = 1;
                    for (int k = 1; k <= 3 && k <= i; k++) { mx = std::max(mx, lv[i - k]); int v = std::max(f[i - k], mx) + 1; if (v < best) { best = v; bk = k; } }
                    f[i] = best; cut[i] = bk;
                }
                std::vector<std::pair<int, int>> chunks;   // (begin, end) in source order
                for (int i = mS; i > 0; i -= cut[i]) chunks.push_back({i - cut[i], i});
                std::reverse(chunks.begin(), chunks.end());
                int prev = -1;
                for (auto& ch : chunks) {
                    int e = prev + 1;
                    for (int q = ch.first; q < ch.second; q++) e = std::max(e, lv[q] + 1);
                    prev = place(e);
                }
                loc[c] = prev;
            }
            (dir == 0 ? stF : stB)[t] = (int)used.size();
            (dir == 0 ? haloF : haloB)[t] = nh;
            (dir == 0 ? slotsF : slotsB) += (long long)used.size() * W;
            if ((int)used.size() > 16) (dir == 0 ? over16F : over16B)++;
            maxHalo = std::max(maxHalo, nh);
        }
    }
    // tile DAG critical path (tiles are in topological order: band, then index)
    std::vector<std::vector<int>> preds(nT);
    for (int f = 0; f < F; f++) { int a = tileOf[lo[f]], b = tileOf[up[f]]; if (a != b) preds[b].push_back(a); }
    std::vector<double> doneF(nT, 0.0); std::vector<int> depth(nT, 0);
    double critF = 0, critB = 0; int maxDepth = 0;
    for (int t = 0; t < nT; t++) {
        double st = 0; int d = 0;
        for (int a : preds[t]) { if (a > t) { printf("tile order violated\n"); return 1; } st = std::max(st, doneF[a]); d = std::max(d, depth[a] + 1); }
        doneF[t] = st + stF[t] * s + o; depth[t] = d; critF = std::max(critF, doneF[t]); maxDepth = std::max(maxDepth, d + 1);
    }
    std::vector<std::vector<int>> succs(nT);
    for (int t = 0; t < nT; t++) for (int a : preds[t]) succs[a].push_back(t);
    std::vector<double> doneB(nT, 0.0);
    for (int t = nT - 1; t >= 0; t--) { double st = 0; for (int a : succs[t]) st = std::max(st, doneB[a]); doneB[t] = st + stB[t] * s + o; critB = std::max(critB, doneB[t]); }
    double meanF = 0, meanB = 0; int mxF = 0, mxB = 0;
    for (int t = 0; t < nT; t++) { meanF += stF[t]; meanB += stB[t]; mxF = std::max(mxF, stF[t]); mxB = std::max(mxB, stB[t]); }
    printf("band %d maxRows %d lanes %d: tiles %d, tile-DAG depth %d | steps/tile fwd mean %.1f max %d (>16: %lld) bwd mean %.1f max %d (>16: %lld) | lane-util fwd %.2f bwd %.2f | max halo %d\n",
           band, maxRows, W, nT, maxDepth, meanF / nT, mxF, over16F, meanB / nT, mxB, over16B, (double)virtF / slotsF, (double)virtB / slotsB, maxHalo);
    printf("model critical path: fwd %.1f us, bwd %.1f us (s = %.0f ns, o = %.0f ns); work bound: %.1f us per sweep on 148 SMs\n",
           critF / 1000, critB / 1000, s, o, (meanF * s + o * nT) / 148 / 1000);
    return 0;
}
