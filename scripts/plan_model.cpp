// plan_model.cpp -- offline cost model of the DIC sweep plan (no GPU needed): builds the channel mesh, runs buildSweepPlan with the
// given parameters and prints the modelled critical path of a sweep,  max over tile-DAG paths of  sum(nSteps*s + o),  for the measured
// per-step time s and publish->visible overhead o (profiles/r1e_sweep_trace.log).  Build:
//   g++ -O2 -std=c++17 -I shallowinterfoam_b200/csrc scripts/plan_model.cpp shallowinterfoam_b200/csrc/host_mesh.cpp -o /tmp/plan_model
// Usage: plan_model nx ny nz band maxRows smem lanes [s_ns o_ns]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "host_mesh.h"
using namespace b200;

int main(int argc, char** argv) {
    int nx = argc > 1 ? atoi(argv[1]) : 200, ny = argc > 2 ? atoi(argv[2]) : 50, nz = argc > 3 ? atoi(argv[3]) : 100;
    int band = argc > 4 ? atoi(argv[4]) : 16, maxRows = argc > 5 ? atoi(argv[5]) : 4096, smem = argc > 6 ? atoi(argv[6]) : 225000;
    int lanes = argc > 7 ? atoi(argv[7]) : 256;
    double s = argc > 8 ? atof(argv[8]) : 126.0, o = argc > 9 ? atof(argv[9]) : 450.0;
    const bool poly = argc > 10 && atoi(argv[10]) != 0;   // 11th argument != 0: the polyhedral box of configs[3] (mesh_oracle.block_groups)
    PolyMesh box = makeBox(nx, ny, nz, 20.0, 5.0, 10.0);
    PolyMesh m;
    if (!poly) m = subsetMesh(box, weirKeepMask(nx, ny, nz, 0.45, 0.5, 0.3), "weir");
    else {
        std::vector<int64_t> key((size_t)nx * ny * nz);
        for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
            int I = i / 2, J = j / 2, K = k / 2;
            bool full = 2 * I + 1 < nx && 2 * J + 1 < ny && 2 * K + 1 < nz;
            bool merged = full && ((I + J + K) % 3 != 0);
            int64_t cell = i + (int64_t)nx * (j + (int64_t)ny * k);
            key[cell] = merged ? (int64_t)nx * ny * nz + (I + (int64_t)(nx / 2 + 1) * (J + (int64_t)(ny / 2 + 1) * K)) : cell;
        }
        m = agglomerate(box, key.data());
    }
    int F = m.nInternalFaces();
    SweepPlan P;
    buildSweepPlan(m.nCells, m.owner.data(), m.neighbour.data(), F, band, maxRows, smem, lanes, P);
    int nT = P.nTiles;
    std::vector<int> tileOfRow(m.nCells);
    for (int t = 0; t < nT; t++) for (int r = P.tileStart[t]; r < P.tileStart[t + 1]; r++) tileOfRow[r] = t;
    std::vector<double> cost(nT), done(nT, 0.0);
    long long totSteps = 0, maxSteps = 0, over16 = 0; double util = 0;
    for (int t = 0; t < nT; t++) {
        int ns = P.fwdStepStart[t + 1] - P.fwdStepStart[t];
        cost[t] = ns * (ns <= 16 ? s : 1.26 * s) + o; totSteps += ns; maxSteps = std::max<long long>(maxSteps, ns); if (ns > 16) over16++;
        util += (double)(P.tileStart[t + 1] - P.tileStart[t]);
    }
    // tiles are in ticket (topological) order: one forward pass gives the longest path
    std::vector<std::vector<int>> preds(nT);
    for (int f = 0; f < F; f++) {
        int a = tileOfRow[P.cellRow[m.owner[f]]], b = tileOfRow[P.cellRow[m.neighbour[f]]];
        if (a != b) preds[b].push_back(a);
    }
    double crit = 0; int maxHalo = 0;
    // streamed lateral halo (what-if, argv[11] = look-ahead steps, 0 = closed tiles): a consumer may run `la` steps + o behind a producer of
    // the SAME band instead of after it; producers of earlier bands still have to finish (their last level feeds the consumer's first step)
    const int la = argc > 11 ? atoi(argv[11]) : 0;
    std::vector<double> startT(nT, 0.0);
    std::vector<int> bandOf(nT);
    for (int t = 0; t < nT; t++) bandOf[t] = P.xlevel[P.rowCell[P.tileStart[t]]] / band;
    for (int t = 0; t < nT; t++) {
        double start = 0;
        for (int a : preds[t]) {
            if (la > 0 && bandOf[a] == bandOf[t]) start = std::max(start, std::max(startT[a] + la * s + o, done[a] - cost[t] + la * s + o));
            else start = std::max(start, done[a]);
        }
        startT[t] = start;
        done[t] = start + cost[t];
        crit = std::max(crit, done[t]);
        maxHalo = std::max(maxHalo, P.fwdHalo[t + 1] - P.fwdHalo[t]);
    }
    int maxD = 0; for (int t = 0; t < nT; t++) maxD = std::max(maxD, std::max(P.fwdD[t], P.bwdD[t]));
    // backward sweep: its own step counts (chunked plans), the tile graph reversed
    double critB = 0; long long totB = 0, maxB = 0;
    {
        std::vector<std::vector<int>> succs(nT);
        for (int t = 0; t < nT; t++) for (int a : preds[t]) succs[a].push_back(t);
        std::vector<double> doneB(nT, 0.0);
        for (int t = nT - 1; t >= 0; t--) {
            int nb = P.bwdStepStart[t + 1] - P.bwdStepStart[t]; totB += nb; maxB = std::max<long long>(maxB, nb);
            double st = 0; for (int a : succs[t]) st = std::max(st, doneB[a]);
            doneB[t] = st + nb * (nb <= 16 ? s : 1.26 * s) + o; critB = std::max(critB, doneB[t]);
        }
    }
    printf("chunked %d indexBins %d xlevels %d | bwd steps/tile mean %.1f max %lld | model crit path bwd %.1f us\n", (int)P.chunked, P.indexBins, P.nXLevels, (double)totB / nT, maxB, critB / 1000.0);
    printf("faces/cell %.2f  max sources per row %d  strides %d %d %d\n", 2.0 * F / m.nCells, maxD, P.stride[0], P.stride[1], P.stride[2]);
    printf("N %d levels %d | band %d maxRows %d lanes %d binJ %d binK %d | tiles %d tileLevels %d | steps/tile mean %.1f max %lld (>16: %lld) lane-util %.2f | smem %zu maxHalo %d | model crit path %.1f us (s=%.0f ns, o=%.0f ns)\n",
           m.nCells, P.nLevels, band, maxRows, lanes, P.binJ, P.binK, nT, P.nTileLevels, (double)totSteps / nT, maxSteps, over16,
           util / ((double)totSteps * lanes), std::max(P.smemFwd, P.smemBwd), maxHalo, crit / 1000.0, s, o);
    return 0;
}
