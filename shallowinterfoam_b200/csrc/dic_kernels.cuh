// dic_kernels.cuh -- a2/a3: DICPreconditioner as SM-local tiled data-flow sweeps.
//
// [UPSTREAM DICPreconditioner.C]  calcReciprocalD: rD[u] -= upper^2/rD[l] (faces ascending), rD = 1/rD
//                                 precondition   : wA = rD*rA; forward faces asc wA[u] -= rD[u]*upper*wA[l];
//                                                  backward faces desc wA[l] -= rD[l]*upper*wA[u]
// A row needs its lower (forward) / upper (backward) neighbours: a sparse triangular solve whose critical path is the number
// of levels (348 at 1 M cells).  Going through L2 for every level costs ~1 us per level, so rows are grouped into TILES
// (host_mesh.h tileRenumbering: band of 16 levels x chunk of 1024 cells in original index order).  One CTA owns a tile and walks
// its level segments with __syncthreads between them; values produced inside the tile are exchanged through shared memory
// (~0.1 us per level), values from other tiles are polled from L2 (the output array is pre-filled with a NaN sentinel and every
// row is published with one relaxed 64-bit store).  Tiles are handed out in order by an atomic ticket and every dependency
// points to the same or an earlier (forward) / later (backward) tile, so a running CTA only waits on CTAs that are already
// running: no grid barrier, no co-residency requirement.  Operands that do not depend on other rows (indices, coefficients)
// are prefetched one level ahead into registers.  Inside a row the updates keep upstream's face order => bitwise the
// sequential result.
#pragma once
#include "ldu_kernels.cuh"

namespace b200 {

struct SweepCtl { unsigned* ticket; unsigned* done; };

__device__ __forceinline__ int nextTicket(SweepCtl c) {
    __shared__ int sTicket;
    __syncthreads();
    if (threadIdx.x == 0) sTicket = (int)atomicAdd(c.ticket, 1u);
    __syncthreads();
    return sTicket;
}
__device__ __forceinline__ void sweepExit(SweepCtl c) {
    if (threadIdx.x == 0) {
        unsigned d = atomicInc(c.done, gridDim.x - 1);
        if (d == gridDim.x - 1) { *c.ticket = 0u; __threadfence(); }
    }
}

#define TILE_MAXD 6

struct RowWork {          // what a row needs that does not depend on other rows
    int row, cnt, more;   // row < 0: idle thread; more: dependencies beyond TILE_MAXD (generic tail)
    int dl[TILE_MAXD];
    double dc[TILE_MAXD];
    double w0, rd;
};

__device__ __forceinline__ double pollValue(const double* p) {
    double v = ldPoll(p);
    while (isSentinel(v)) {
#ifdef B200_EMU
        emu::yield();
#endif
        v = ldPoll(p);
    }
    return v;
}

// MODE 0: forward precondition sweep (w0 = rD*rA, c = rD*upper) ; MODE 1: calcReciprocalD (w0 = diag, c = upper)
template <int MODE>
__device__ __forceinline__ void loadFwd(RowWork& r, int row, bool valid, const MeshView& mv, const double* __restrict__ a0,
                                        const double* __restrict__ a1, const double* __restrict__ upS) {
    r.row = valid ? row : -1; r.cnt = 0; r.more = 0; r.w0 = 0.0; r.rd = 0.0;
#pragma unroll
    for (int j = 0; j < TILE_MAXD; j++) { r.dl[j] = 0; r.dc[j] = 0.0; }
    if (!valid) return;
    int s = row >> 5, lane = row & 31;
    if (MODE == 0) { r.rd = a0[row]; r.w0 = r.rd * a1[row]; } else { r.w0 = a0[row]; r.rd = 1.0; }
    int nb0 = mv.nbrOff[s], nbw = (mv.nbrOff[s + 1] - nb0) >> 5;
#pragma unroll
    for (int j = 0; j < TILE_MAXD; j++)
        if (j < nbw) {
            int pk = mv.loPk[nb0 + (j << 5) + lane];
            if (pk >= 0) { int l; int pos = facePosOf(mv, pk, l); r.dl[j] = l; r.dc[j] = (MODE == 0) ? r.rd * upS[pos] : upS[pos]; r.cnt = j + 1; }
        }
    if (nbw > TILE_MAXD && mv.loPk[nb0 + (TILE_MAXD << 5) + lane] >= 0) r.more = 1;
}

template <int MODE>
static __global__ void __launch_bounds__(SWEEP_BLOCK) k_dic_fwd_tiled(MeshView mv, TileView tv, const double* __restrict__ a0,
        const double* __restrict__ a1, const double* __restrict__ upS, double* y, double* __restrict__ rDout, SweepCtl ctl,
        const PcgScalars* sc) {
    __shared__ double ys[SWEEP_TILE_MAX];
    if (sc && sc->stop) return;
    for (;;) {
        int tile = nextTicket(ctl);
        if (tile >= tv.nTiles) break;
        const int tile0 = tv.tileStart[tile];
        const int seg0 = tv.tileSegStart[tile], seg1 = tv.tileSegStart[tile + 1];
        RowWork cur, nxt;
        { int b = tv.segRow[seg0], e = tv.segRow[seg0 + 1], row = b + (int)threadIdx.x; loadFwd<MODE>(cur, row, seg0 < seg1 && row < e, mv, a0, a1, upS); }
        for (int sg = seg0; sg < seg1; sg++) {
            if (sg + 1 < seg1) { int b = tv.segRow[sg + 1], e = tv.segRow[sg + 2], row = b + (int)threadIdx.x; loadFwd<MODE>(nxt, row, row < e, mv, a0, a1, upS); }
            if (cur.row >= 0) {
                double w = cur.w0;
#pragma unroll
                for (int j = 0; j < TILE_MAXD; j++)
                    if (j < cur.cnt) {
                        int l = cur.dl[j];
                        double v = (l >= tile0) ? ys[l - tile0] : pollValue(y + l);
                        if (MODE == 0) w = w - cur.dc[j] * v; else w = w - cur.dc[j] * cur.dc[j] / v;
                    }
                if (cur.more) {   // polyhedral rows with more than TILE_MAXD lower neighbours
                    int s = cur.row >> 5, lane = cur.row & 31, nb0 = mv.nbrOff[s], nbw = (mv.nbrOff[s + 1] - nb0) >> 5;
                    for (int j = TILE_MAXD; j < nbw; j++) {
                        int pk = mv.loPk[nb0 + (j << 5) + lane];
                        if (pk < 0) break;
                        int l; int pos = facePosOf(mv, pk, l);
                        double v = (l >= tile0) ? ys[l - tile0] : pollValue(y + l);
                        double u = upS[pos];
                        if (MODE == 0) w = w - cur.rd * u * v; else w = w - u * u / v;
                    }
                }
                ys[cur.row - tile0] = w;
                stPublish(y + cur.row, w);
                if (MODE == 1) rDout[cur.row] = 1.0 / w;
            }
            __syncthreads();
            cur = nxt;
        }
    }
    sweepExit(ctl);
}

// backward: z[row] = y[row] - sum_{owner faces DESC} (rD[row]*upper_f)*z[u] ; y[row] <- sentinel (re-armed) ;
// per-tile partial of sum(z*rA) (fixed rows per partial => deterministic) -> sc->wArA
__device__ __forceinline__ void loadBwd(RowWork& r, int row, bool valid, const MeshView& mv, const double* __restrict__ rD,
                                        const double* y, const double* __restrict__ upS) {
    r.row = valid ? row : -1; r.cnt = 0; r.more = 0; r.w0 = 0.0; r.rd = 0.0;
#pragma unroll
    for (int j = 0; j < TILE_MAXD; j++) { r.dl[j] = 0; r.dc[j] = 0.0; }
    if (!valid) return;
    int s = row >> 5, lane = row & 31;
    r.rd = rD[row]; r.w0 = y[row];
    int so0 = mv.sliceOff[s], ow = (mv.sliceOff[s + 1] - so0) >> 5;
    // application order = descending face index: count the real entries, then fill from the last one down
    int n = 0;
    for (int k = 0; k < ow; k++) { if (mv.upIdx[so0 + (k << 5) + lane] < 0) break; n++; }
    if (n > TILE_MAXD) { r.more = 1; return; }
#pragma unroll
    for (int j = 0; j < TILE_MAXD; j++)
        if (j < n) { int pos = so0 + ((n - 1 - j) << 5) + lane; r.dl[j] = mv.upIdx[pos]; r.dc[j] = r.rd * upS[pos]; }
    r.cnt = n;
}

static __global__ void __launch_bounds__(SWEEP_BLOCK) k_dic_bwd_tiled(MeshView mv, TileView tv, const double* __restrict__ rD,
        const double* __restrict__ upS, const double* __restrict__ rA, double* y, double* z, SweepCtl ctl, PcgScalars* sc,
        double* partials, unsigned* counter) {
    __shared__ double zs[SWEEP_TILE_MAX];
    if (sc && sc->stop) return;
    for (;;) {
        int t = nextTicket(ctl);
        if (t >= tv.nTiles) break;
        const int tile = tv.nTiles - 1 - t;
        const int tile0 = tv.tileStart[tile], tileEnd = tv.tileStart[tile + 1];
        const int seg0 = tv.tileSegStart[tile], seg1 = tv.tileSegStart[tile + 1];
        RowWork cur, nxt;
        double dotv = 0.0;
        { int sg = seg1 - 1; bool any = seg0 < seg1; int b = any ? tv.segRow[sg] : 0, e = any ? tv.segRow[sg + 1] : 0, row = b + (int)threadIdx.x;
          loadBwd(cur, row, any && row < e, mv, rD, y, upS); }
        for (int sg = seg1 - 1; sg >= seg0; sg--) {
            if (sg - 1 >= seg0) { int b = tv.segRow[sg - 1], e = tv.segRow[sg], row = b + (int)threadIdx.x; loadBwd(nxt, row, row < e, mv, rD, y, upS); }
            if (cur.row >= 0) {
                double w = cur.w0;
                if (!cur.more) {
#pragma unroll
                    for (int j = 0; j < TILE_MAXD; j++)
                        if (j < cur.cnt) {
                            int u = cur.dl[j];
                            double v = (u < tileEnd) ? zs[u - tile0] : pollValue(z + u);
                            w = w - cur.dc[j] * v;
                        }
                } else {
                    int s = cur.row >> 5, lane = cur.row & 31, so0 = mv.sliceOff[s], ow = (mv.sliceOff[s + 1] - so0) >> 5;
                    for (int k = ow - 1; k >= 0; k--) {
                        int pos = so0 + (k << 5) + lane;
                        int u = mv.upIdx[pos];
                        if (u < 0) continue;
                        double v = (u < tileEnd) ? zs[u - tile0] : pollValue(z + u);
                        w = w - cur.rd * upS[pos] * v;
                    }
                }
                zs[cur.row - tile0] = w;
                stPublish(z + cur.row, w);
                y[cur.row] = sentinel();
                dotv = dotv + w * rA[cur.row];
            }
            __syncthreads();
            cur = nxt;
        }
        if (partials) {
            double v[1] = {dotv};
            gridReduce<1, OpSum>(v, tile, tv.nTiles, partials, counter, nullptr);
        }
    }
    if (partials) gridFinalize<1, OpSum>(tv.nTiles, partials, counter, &sc->wArA);
    sweepExit(ctl);
}

}  // namespace b200
