// dic_kernels.cuh -- a2/a3: DICPreconditioner as CLOSED-TILE data-flow sweeps on a tile DAG.
//
// [UPSTREAM DICPreconditioner.C]  calcReciprocalD: rD[u] -= upper^2/rD[l] (faces ascending), rD = 1/rD
//                                 precondition   : wA = rD*rA; forward faces asc wA[u] -= rD[u]*upper*wA[l];
//                                                  backward faces desc wA[l] -= rD[l]*upper*wA[u]
// A row needs its lower (forward) / upper (backward) neighbours: a sparse triangular solve whose critical path is the number
// of levels (348 at 1 M cells).  One L2 round trip per level (~1 us) is what bounds a row-level schedule, so the rows are grouped
// into tiles by host_mesh.h buildSweepPlan: tile = (band of levels) x (bin of the two stride-class potentials).  All three tile
// coordinates are monotone along every dependency, so the tile graph is acyclic and a tile's external inputs ("halo") all
// come from tiles with an earlier ticket.  One CTA owns a tile:
//   1. stage everything that does not depend on other rows (coefficients, tile-local source indices, rD*rA) in shared memory,
//   2. poll the halo values from L2 ONCE (outputs are pre-filled with a NaN sentinel; every row is published with a relaxed
//      64-bit store, so the data is its own ready flag: one L2 round trip per tile hop instead of one per level),
//   3. walk the tile's level segments out of shared memory with __syncthreads between them (~50 ns per level).
// Tiles are handed out in ticket order (sorted by longest-path level in the tile graph), so a running CTA only waits on CTAs
// that are already running: no grid barrier, no co-residency requirement, deadlock-free for any grid size.  Inside a row the
// updates keep upstream's face order and (rD*upper)*w association => bitwise the sequential result.
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

#ifdef B200_EMU
#define B200_DYN_SMEM(name) unsigned char* name = emu::dynSmem()
#else
#define B200_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// ---- TMA bulk copies (cp.async.bulk, 1-D) + mbarrier: one elected thread streams a tile's operand block into shared memory
// while the other threads compute rD*rA and poll the halo.
#ifdef B200_EMU
// emulator: the copies are synchronous; the barrier counts completed stagings so that waiting threads really wait
struct TileBar { volatile unsigned completed; };
__device__ __forceinline__ void barInit(TileBar* b) { b->completed = 0; }
__device__ __forceinline__ void barExpect(TileBar*, unsigned) {}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, TileBar*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void barStaged(TileBar* b) { b->completed = b->completed + 1; }
__device__ __forceinline__ void barWait(TileBar* b, unsigned parity) { while ((b->completed & 1u) == parity) emu::yield(); }
__device__ __forceinline__ void fenceAsyncProxy() {}
__device__ __forceinline__ void bulkStore(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulkStoreCommit() {}
__device__ __forceinline__ void bulkStoreWaitRead() {}
__device__ __forceinline__ void bulkStoreWaitAll() {}
#else
struct TileBar { unsigned long long word; };
__device__ __forceinline__ void barStaged(TileBar*) {}   // (emulator hook: the hardware mbarrier completes by transaction count)
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void barInit(TileBar* b) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(b)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void barExpect(TileBar* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, TileBar* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(b)) : "memory");
}
__device__ __forceinline__ void barWait(TileBar* b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smemAddr(b)), "r"(parity) : "memory");
    } while (!ok);
}
// generic-proxy accesses of the previous tile (ordered by the CTA barrier) before the async proxy touches the buffer
__device__ __forceinline__ void fenceAsyncProxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA store shared -> global (lands in L2, where the polling loads of the dependent tiles look)
__device__ __forceinline__ void bulkStore(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smemAddr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkStoreCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulkStoreWaitRead() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulkStoreWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#endif
constexpr unsigned BULK_MAX = 32768;   // bytes per cp.async.bulk (multiple of 16)
__device__ __forceinline__ void bulkLoadChunks(void* dst, const void* src, size_t bytes, TileBar* b) {
    for (size_t o = 0; o < bytes; o += BULK_MAX)
        bulkLoad((char*)dst + o, (const char*)src + o, (unsigned)(bytes - o < BULK_MAX ? bytes - o : BULK_MAX), b);
}
#ifdef B200_EMU
__device__ __forceinline__ unsigned long long globalTimerNs() { return 0; }
#else
__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif
#define SWEEP_TRACE(slot) do { if (sv.trace && DIR > 0 && threadIdx.x == 0) sv.trace[8 * ti + (slot)] = globalTimerNs(); } while (0)

#ifdef B200_EMU
__device__ __forceinline__ void namedBarSync(int id, int n) { emu::namedBarrier(id, n); }
#else
__device__ __forceinline__ void namedBarSync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#endif

// spin on a shared-memory counter raised by another warp of the CTA (its data stores precede the counter store: barrier among
// the writers, then one thread raises the counter); a waiting helper warp sleeps between looks so that it does not steal issue slots
template <int SLEEP_NS>
__device__ __forceinline__ void waitFlag(volatile int* flag, int need) {
    while (*flag < need) {
#ifdef B200_EMU
        emu::yield();
#else
        if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
#endif
    }
#ifndef B200_EMU
    asm volatile("" ::: "memory");
#endif
}

struct TileSmem {
    double* vals;          // [nRows + 1 + nHalo + 2] rows of the tile (start values, overwritten in place by the results), one gap
                           // slot, halo values, padding slot (1.0), trash slot (0.0).  &vals[r] is 16-byte aligned iff row0+r is even.
    double* aux;           // [nRows] auxiliary row values (backward sweep: rA for the fused dot product), same alignment rule
    int* halo;             // [nHalo] halo list of the tile (device row to poll, or -(r+1) = local row r of the chain neighbour)
    double* blk;           // [(d + ceil(d/4)) * S] coefficient columns, then packed 16-bit byte offsets into vals (slot layout)
    int S;
};
__device__ __forceinline__ TileSmem carveTile(unsigned char* buf, const SweepTile& t, int lanes) {
    TileSmem s;
    const int nv = (t.nRows + t.nHalo + 6) & ~1, na = (t.nRows + 3) & ~1;
    s.S = t.nSteps * lanes;
    s.vals = (double*)buf + (t.row0 & 1);
    s.aux = (double*)buf + nv + (t.row0 & 1);
    s.halo = (int*)((double*)buf + nv + na);
    s.blk = (double*)(s.halo + ((t.nHalo + 3) & ~3));
    return s;
}
__device__ __forceinline__ double ldsAt(const double* vals, unsigned byteOff) { return *(const double*)((const char*)vals + byteOff); }
__device__ __forceinline__ void stsAt(double* vals, unsigned byteOff, double v) { *(double*)((char*)vals + byteOff) = v; }

// poll the tile's external halo entries (list in shared memory; negative entries were handed over along the chain) out of `src`
// (sentinel = not yet published) into vals[nRows + 1 ..); the loads of one pass are independent
__device__ __forceinline__ void pollHalo(const SweepTile& t, const TileSmem& s, const double* src, int np, int pollNs) {
    constexpr int U = 4;
    const int pid = threadIdx.x;
    for (int h0 = pid; h0 < t.nHalo; h0 += U * np) {
        const double* p[U]; double v[U]; bool need[U];
#pragma unroll
        for (int q = 0; q < U; q++) {
            int h = h0 + q * np;
            int hr = h < t.nHalo ? s.halo[h] : -1;
            need[q] = hr >= 0;
            p[q] = src + (need[q] ? hr : 0);
        }
        bool pending;
        do {
            pending = false;
#pragma unroll
            for (int q = 0; q < U; q++) if (need[q]) v[q] = ldPoll(p[q]);
#pragma unroll
            for (int q = 0; q < U; q++)
                if (need[q]) {
                    if (isSentinel(v[q])) pending = true;
                    else { s.vals[t.nRows + 1 + h0 + q * np] = v[q]; need[q] = false; }
                }
#ifdef B200_EMU
            if (pending) emu::yield();
#else
            if (pending && pollNs > 0) __nanosleep(pollNs);
#endif
        } while (pending);
    }
}

// The steps of one tile.  Step k = up to `lanes` rows of one level, one row per thread; its operands live in the slots
// [k*lanes, (k+1)*lanes) of the operand block (unused lanes carry zero coefficients and point at the trash slot).  A GPU thread
// retires ~1 dependent instruction per 4-6 cycles, so what bounds a step is the instruction count between two barriers: the
// slot layout makes every operand address a fixed increment (no segment look-ups, no validity tests), the next step's operands
// (D coefficients, one 64-bit word of packed offsets: D sources + the row's own place, start value) are fetched while the
// current gathers are in flight, and the critical path is
//   BAR -> D x LDS.64 (sources) -> D x DMUL -> D x DADD -> STS -> BAR.
// No global access inside the loop: bar.sync would wait for the L2 acknowledge of a gpu-scope store (~400 cycles per step).
// OP 0: w -= c*v ; OP 1: w -= c*c/v (calcReciprocalD).  DIR +1: steps upwards (forward sweep), -1: downwards (backward).
template <int OP, int DIR, int D>
__device__ __forceinline__ void sweepStepsD(const TileSmem& s, const SweepTile& t, const int W) {
    const int S = s.S;
    int slot = (DIR > 0 ? 0 : (t.nSteps - 1) * W) + (int)threadIdx.x;
    const double* C = s.blk;
    const uint2* PK = (const uint2*)(s.blk + D * S);
    double c0 = C[slot], c1 = D > 1 ? C[S + slot] : 0.0, c2 = D > 2 ? C[2 * S + slot] : 0.0;
    uint2 pk = PK[slot];
    double w = ldsAt(s.vals, pk.y >> 16);
    for (int it = 0; it < t.nSteps; it++) {
        const double v0 = ldsAt(s.vals, pk.x & 0xffffu);
        const double v1 = D > 1 ? ldsAt(s.vals, pk.x >> 16) : 0.0;
        const double v2 = D > 2 ? ldsAt(s.vals, pk.y & 0xffffu) : 0.0;
        const int nslot = it + 1 < t.nSteps ? slot + DIR * W : slot;
        const double n0 = C[nslot], n1 = D > 1 ? C[S + nslot] : 0.0, n2 = D > 2 ? C[2 * S + nslot] : 0.0;
        const uint2 npk = PK[nslot];
        const double nw = ldsAt(s.vals, npk.y >> 16);   // the next step's own rows still hold their start values
        if (OP == 0) { w = w - c0 * v0; if (D > 1) w = w - c1 * v1; if (D > 2) w = w - c2 * v2; }
        else { w = w - c0 * c0 / v0; if (D > 1) w = w - c1 * c1 / v1; if (D > 2) w = w - c2 * c2 / v2; }
        stsAt(s.vals, pk.y >> 16, w);
        namedBarSync(1, W);
        c0 = n0; c1 = n1; c2 = n2; w = nw; pk = npk; slot = nslot;
    }
}
template <int OP, int DIR>
__device__ __forceinline__ void sweepSteps(const SweepView& sv, const TileSmem& s, const SweepTile& t, const int W) {
    if (t.nSteps == 0) return;
    if (t.d == 3) { sweepStepsD<OP, DIR, 3>(s, t, W); return; }
    if (t.d == 2) { sweepStepsD<OP, DIR, 2>(s, t, W); return; }
    if (t.d == 1) { sweepStepsD<OP, DIR, 1>(s, t, W); return; }
    const int S = s.S;
    const unsigned short* OFF = (const unsigned short*)(s.blk + t.d * S);   // [ceil(d/4)][S][4]
    for (int k = DIR > 0 ? 0 : t.nSteps - 1; k >= 0 && k < t.nSteps; k += DIR) {   // polyhedral rows (d > 3), isolated rows (d = 0)
        if ((int)threadIdx.x < sv.stepLen[t.step0 + k]) {
            const int slot = k * W + threadIdx.x, r = sv.stepRow[t.step0 + k] - t.row0 + threadIdx.x;
            double w = s.vals[r];
            for (int j = 0; j < t.d; j++) {
                double c = s.blk[j * S + slot], v = ldsAt(s.vals, OFF[((j >> 2) * S + slot) * 4 + (j & 3)]);
                if (OP == 0) w = w - c * v; else w = w - c * c / v;
            }
            s.vals[r] = w;
        }
        namedBarSync(1, W);
    }
}

// rows [row0, row0+n) of a global array -> shared memory at dst[0..n), by TMA: the 16-byte aligned hull [lo, hi) of the range is
// copied (dst[-1] and dst[n] exist and are scratch); an odd very last row of the array is copied by the calling thread.
__device__ __forceinline__ unsigned rowsBytes(const SweepTile& t, int nTotal) {
    const int lo = t.row0 & ~1, end = t.row0 + t.nRows;
    int hi = end + (end & 1); if (hi > nTotal) hi = end & ~1;
    return (unsigned)((hi - lo) * 8);
}
__device__ __forceinline__ void stageRows(double* dst, const double* __restrict__ src, const SweepTile& t, int nTotal, TileBar* bar) {
    const int lo = t.row0 & ~1, end = t.row0 + t.nRows;
    int hi = end + (end & 1); if (hi > nTotal) hi = end & ~1;
    if (hi > lo) bulkLoadChunks(dst + (lo - t.row0), src + lo, (size_t)(hi - lo) * 8, bar);
    if (hi < end) dst[t.nRows - 1] = src[end - 1];
}
// stage one tile into a buffer (one thread): operand block, halo list, start values (+ auxiliary rows)
__device__ __forceinline__ void stageTile(const SweepView& sv, const SweepTile& t, unsigned char* buf, const int* __restrict__ haloRow,
                                          const double* __restrict__ blocks, const double* __restrict__ start, const double* __restrict__ auxSrc, TileBar* bar) {
    TileSmem s = carveTile(buf, t, sv.lanes);
    const size_t blkBytes = (size_t)(t.d + ((t.d + 3) >> 2)) * s.S * 8, haloBytes = (size_t)((t.nHalo + 3) & ~3) * 4;
    const unsigned rb = rowsBytes(t, sv.nRows);
    fenceAsyncProxy();   // generic-proxy accesses of the buffer's previous tenant (ordered by the CTA barrier) before the async writes
    barExpect(bar, (unsigned)(blkBytes + haloBytes + rb + (auxSrc ? rb : 0u)));
    bulkLoadChunks(s.blk, blocks + t.blk, blkBytes, bar);
    if (haloBytes) bulkLoadChunks(s.halo, haloRow + t.halo, haloBytes, bar);
    stageRows(s.vals, start, t, sv.nRows, bar);
    if (auxSrc) stageRows(s.aux, auxSrc, t, sv.nRows, bar);
    barStaged(bar);
}
// publish the tile's rows: the results sit contiguously in vals[0..nRows) = device rows [row0, row0+nRows): one TMA bulk
// store for the 16-byte aligned middle (it lands in L2, where the dependants poll), scalar gpu-scope stores for an odd first /
// last row.  Writers have fenced (fence.proxy.async) and passed a CTA barrier; called by one thread.
__device__ __forceinline__ void publishTile(const TileSmem& s, const SweepTile& t, double* out) {
    const int a = t.row0 & 1, e = t.nRows - ((t.row0 + t.nRows) & 1);   // local range [a, e) is aligned
    if (e > a) {
        for (int o = a; o < e; o += (int)(BULK_MAX / 8)) bulkStore(out + t.row0 + o, s.vals + o, (unsigned)((e - o < (int)(BULK_MAX / 8) ? e - o : (int)(BULK_MAX / 8)) * 8));
        bulkStoreCommit();
    }
    if (a && t.nRows > 0) stPublish(out + t.row0, s.vals[0]);
    if (e < t.nRows && e >= a) stPublish(out + t.row0 + t.nRows - 1, s.vals[t.nRows - 1]);
}

// One sweep.  A CTA takes a CHAIN of tiles by ticket (the tiles of one (binK, binJ) column band after band, or a single tile)
// and runs them in order, double-buffered in shared memory:
//   * one thread keeps the TMA busy: while tile i runs, tile i+1's operand block, halo list and start values stream into the
//     other buffer (cp.async.bulk + mbarrier);
//   * the values tile i+1 needs from tile i (the band direction: the true critical path of the triangular solve) are copied
//     from buffer to buffer -- no L2 round trip along the chain; only the transverse halo (from chains with an earlier ticket,
//     which run ahead) is polled from L2, once per tile, the data being its own ready flag (NaN sentinel);
//   * a tile's results leave with ONE TMA bulk store right after its last step.
// Chains are ticketed in a topological order of the chain graph, so a running CTA only waits on CTAs that are already running:
// no grid barrier, no co-residency requirement, deadlock-free for any grid size.
// MODE 0, DIR +1: forward precondition  y = w0 - sum_j cf_j*y[src_j]          (w0 = rD*rA from k_pcg_xr_w0, cf = rD[row]*upper)
// MODE 1, DIR +1: calcReciprocalD       Dt = diag - sum_j up_j^2/Dt[src_j] ; rD = 1/Dt   (operand block = plain upper)
// MODE 0, DIR -1: backward              z = y - sum_{owner faces DESC} cb_j*z[u_j] ; per-chain partial of sum(z*rA) -> sc->wArA
template <int MODE, int DIR>
static __global__ void __launch_bounds__(SWEEP_THREADS + SWEEP_HELPERS) k_dic_sweep(SweepView sv, const double* __restrict__ start,
        const double* __restrict__ auxSrc, const double* __restrict__ blocks, double* out, double* __restrict__ rDout,
        SweepCtl ctl, PcgScalars* sc, double* partials, unsigned* counter) {
    B200_DYN_SMEM(smem);
    __shared__ TileBar sBar[2];
    __shared__ SweepTile sTile[SWEEP_MAX_CHAIN];   // the chain's tile descriptors, in execution order
    __shared__ int sDone, sPubd, sHaloIn[8];       // tiles of the chain: finished by the compute warps / published / halo share of helper warp w in
    if (sc && sc->stop) return;
    const int tid = threadIdx.x, W = sv.lanes;      // threads [0, W): compute warps (named barrier 1); last warp: DMA
    const bool dma = tid >= W;
    const SweepTile* tiles = DIR > 0 ? sv.fwd : sv.bwd;
    const int* haloRow = DIR > 0 ? sv.fwdHaloRow : sv.bwdHaloRow;
    volatile int* done = &sDone;
    volatile int* haloIn = sHaloIn;
    volatile int* pubd = &sPubd;
    if (tid == 0) { barInit(&sBar[0]); barInit(&sBar[1]); }
    unsigned phase[2] = {0u, 0u};
    for (;;) {
        const int tk = nextTicket(ctl);
        if (tk >= sv.nChains) break;
        const int c = DIR > 0 ? tk : sv.nChains - 1 - tk;
        const int cs = sv.chainStart[c], m = sv.chainStart[c + 1] - cs;
        for (int i = tid; i < m; i += blockDim.x) sTile[i] = tiles[sv.chainTiles[DIR > 0 ? cs + i : cs + m - 1 - i]];
        if (tid < 8) sHaloIn[tid] = 0;
        if (tid == 0) { sDone = 0; sPubd = 0; }
        __syncthreads();
        double dotv = 0.0;
        if (dma) {
            // ---- helper warps, off the compute warps' critical path.
            //   * publish (helper warp 0): as soon as the compute warps have finished tile i (`done`), one TMA store sends its rows
            //     to L2, then tile i+2 is staged into its buffer (operand block, halo list, start values: TMA); `pubd` counts;
            //   * halo prefetch (all helper warps, entries dealt out statically): the transverse halo of the next tile is polled
            //     out of L2 and written into the tile's halo slots; `haloIn[w]` = tiles whose share of warp w is complete.
            //     One pass keeps up to 8 loads per lane in flight, so a tile's ~700 entries take one L2 round trip.
            const int hw = (tid - W) >> 5, lane = tid & 31, nhw = (blockDim.x - W) >> 5, hid = tid - W, nh = blockDim.x - W;
            if (hid == 0) {
                stageTile(sv, sTile[0], smem, haloRow, blocks, start, auxSrc, &sBar[0]);
                if (m > 1) stageTile(sv, sTile[1], smem + sv.bufBytes, haloRow, blocks, start, auxSrc, &sBar[1]);
            }
            __syncwarp();
            int pub = 0, pol = 0;            // next tile to publish (warp 0) / tile whose halo is being fetched
            unsigned pending = 0; bool armed = false;   // per-lane bit mask of the entries (hid + nh*k) still to arrive
            while (hw == 0 ? pub < m : pol < m) {
                if (hw == 0 && *done > pub) {           // (warp-uniform: every lane reads the same word)
                    if (lane == 0) {
                        unsigned char* buf = smem + (pub & 1) * sv.bufBytes;
                        const SweepTile t = sTile[pub];
                        fenceAsyncProxy();
                        publishTile(carveTile(buf, t, W), t, out);
                        if (pub + 2 < m) { bulkStoreWaitRead(); stageTile(sv, sTile[pub + 2], buf, haloRow, blocks, start, auxSrc, &sBar[pub & 1]); }
                        *pubd = pub + 1;
                    }
                    __syncwarp();
                    pub++;
                    continue;
                }
                const int pubNow = hw == 0 ? pub : *pubd;
                if (pol < m && pol <= pubNow + 1) {    // tile pol owns its buffer (tile pol-2 has been published and re-staged)
                    const SweepTile t = sTile[pol];
                    double* hv = carveTile(smem + (pol & 1) * sv.bufBytes, t, W).vals + t.nRows + 1;
                    if (!armed) {
                        pending = 0;
                        for (int k = 0; k < 32; k++) { int h = hid + nh * k; if (h < t.nHalo && haloRow[t.halo + h] >= 0) pending |= 1u << k; }
                        armed = true;
                    }
                    unsigned left = pending;
                    while (left) {              // one pass: up to 8 loads in flight per lane
                        constexpr int U = 8;
                        int k[U]; double v[U];
#pragma unroll
                        for (int q = 0; q < U; q++) { k[q] = left ? __ffs(left) - 1 : -1; if (left) left &= left - 1; }
#pragma unroll
                        for (int q = 0; q < U; q++) if (k[q] >= 0) v[q] = ldPoll(out + haloRow[t.halo + hid + nh * k[q]]);
#pragma unroll
                        for (int q = 0; q < U; q++) if (k[q] >= 0 && !isSentinel(v[q])) { hv[hid + nh * k[q]] = v[q]; pending &= ~(1u << k[q]); }
                    }
                    if (__all_sync(0xffffffffu, pending == 0)) {
                        __syncwarp();
                        if (lane == 0) haloIn[hw] = pol + 1;
                        pol++; armed = false;
                    }
#ifdef B200_EMU
                    else emu::yield();
#endif
                } else {
#ifdef B200_EMU
                    emu::yield();
#else
                    __nanosleep(20);
#endif
                }
            }
            if (hid == 0) bulkStoreWaitRead();   // both buffers are re-staged by the next chain
        } else {
            // ---- compute warps
            barWait(&sBar[0], phase[0]); phase[0] ^= 1u;
            for (int i = 0; i < m; i++) {
                const int b = i & 1;
                const SweepTile t = sTile[i];
                const int ti = sv.trace ? sv.chainTiles[DIR > 0 ? cs + i : cs + m - 1 - i] : 0;
                SWEEP_TRACE(0);
                unsigned char* buf = smem + b * sv.bufBytes;
                unsigned char* other = smem + (b ^ 1) * sv.bufBytes;
                const TileSmem s = carveTile(buf, t, W);
                const int n = t.nRows;
                // A. the transverse halo was fetched by the DMA warp (tiles with more than 1024 entries poll it themselves);
                //    the chain neighbour's values were handed over in shared memory
                if (t.nHalo > 32 * (int)(blockDim.x - W)) pollHalo(t, s, out, W, sv.pollNs);
                else for (int w = 0; w < (int)(blockDim.x - W) >> 5; w++) waitFlag<0>(haloIn + w, i + 1);
                if (tid == 0) { s.vals[n + 1 + t.nHalo] = 1.0; s.vals[n + 2 + t.nHalo] = 0.0; }
                SWEEP_TRACE(1);
                namedBarSync(1, W);
                if (sv.trace && DIR > 0 && tid == 0 && s.vals[n + 1 + t.nHalo] == 1.0) SWEEP_TRACE(2);
                // B. the steps
                sweepSteps<MODE, DIR>(sv, s, t, W);
                fenceAsyncProxy();    // this thread's results, for the DMA warp's TMA store (writers fence, barrier, one thread stores)
                SWEEP_TRACE(3);
                // C. hand the chain values over to tile i+1 (its buffer was staged while this tile ran)
                if (i + 1 < m) {
                    barWait(&sBar[b ^ 1], phase[b ^ 1]); phase[b ^ 1] ^= 1u;
                    const SweepTile t2 = sTile[i + 1];
                    const TileSmem s2 = carveTile(other, t2, W);
                    for (int h = tid; h < t2.nHalo; h += W) { int hr = s2.halo[h]; if (hr < 0) s2.vals[t2.nRows + 1 + h] = s.vals[-hr - 1]; }
                }
                // D. epilogue out of shared memory (before the barrier: the buffer is re-staged once `done` is raised)
                if (MODE == 1) for (int r = tid; r < n; r += W) rDout[t.row0 + r] = 1.0 / s.vals[r];
                if (DIR < 0 && partials) for (int r = tid; r < n; r += W) dotv = dotv + s.vals[r] * s.aux[r];
                SWEEP_TRACE(5);
                namedBarSync(1, W);
                if (tid == 0) *done = i + 1;
            }
        }
        if (DIR < 0 && partials) { double v[1] = {dotv}; gridReduce<1, OpSum>(v, c, sv.nChains, partials, counter, nullptr); }
    }
    if (tid == W) bulkStoreWaitAll();   // this CTA's TMA stores (helper thread 0) must have landed before the kernel ends
    if (DIR < 0 && partials) gridFinalize<1, OpSum>(sv.nChains, partials, counter, &sc->wArA);
    sweepExit(ctl);
}

// per solve: coefficient columns of the operand blocks: cf = rD[row]*upper (forward), cb = rD[row]*upper (backward).
// rD == null: the plain upper coefficients (what calcReciprocalD sweeps with; 1.0*upper is exact).
// grid = 2*nTiles (forward tiles, then backward tiles) or nTiles (forward only); block = lanes
static __global__ void __launch_bounds__(SWEEP_THREADS) k_dic_coeffs(SweepView sv, const double* __restrict__ rD,
        const double* __restrict__ upS, double* __restrict__ cf, double* __restrict__ cb) {
    const bool bw = (int)blockIdx.x >= sv.nTiles;
    const SweepTile t = bw ? sv.bwd[blockIdx.x - sv.nTiles] : sv.fwd[blockIdx.x];
    const int* pos = bw ? sv.bwdPos : sv.fwdPos;
    double* out = bw ? cb : cf;
    const int W = blockDim.x, S = t.nSteps * W;
    for (int k = 0; k < t.nSteps; k++) {
        const bool valid = (int)threadIdx.x < sv.stepLen[t.step0 + k];
        const double rd = !valid ? 0.0 : rD ? rD[sv.stepRow[t.step0 + k] + threadIdx.x] : 1.0;
        for (int j = 0; j < t.d; j++) {
            const int e = t.blk + j * S + k * W + threadIdx.x;
            const int p = valid ? pos[e] : -1;
            out[e] = p >= 0 ? rd * upS[p] : 0.0;
        }
    }
}

// The tail of upstream's PCG iteration + the start values of the next forward sweep, one streaming pass over the rows:
//   alpha = wArA/wApA ; x += alpha*pA ; rA -= alpha*wA ; sum|rA|  ->  finalResidual, nIter++, stop  [UPSTREAM PCG.C]
//   w0 = rD*rA (the forward sweep's start values) ; y <- sentinel (re-armed for the sweep)
// The last block finishes the (fixed-order) residual sum and runs the convergence test (single rank), or leaves the local sum for
// the all-reduce + k_pcg_check (multi-rank).  x == null: only w0 = rD*rA and the re-arming (b200_dic_precondition).
static __global__ void __launch_bounds__(BLOCK) k_pcg_xr_w0(int N, double* __restrict__ x, double* rA, const double* __restrict__ pA,
        const double* __restrict__ wA, const double* __restrict__ rD, double* __restrict__ w0, double* __restrict__ y,
        PcgScalars* sc, double* partials, unsigned* counter, double* hist, int doCheck) {
    if (sc && sc->stop) return;
    const bool fused = x != nullptr;
    const bool upd = fused && sc->nFwd > 0;          // an iteration's x / rA update is pending
    double alpha = 0.0;
    if (upd) {
        if (fabs(sc->wApA) / sc->normFactor <= 1.0e-300) {   // checkSingularity [UPSTREAM PCG.C]: leave x, rA, nIter untouched
            if (blockIdx.x == 0 && threadIdx.x == 0) { sc->singular = 1; sc->stop = 1; }
            return;
        }
        alpha = sc->wArA / sc->wApA;
    }
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (row < N) {
        double r = rA[row];
        if (upd) { x[row] = x[row] + alpha * pA[row]; r = r - alpha * wA[row]; rA[row] = r; }
        w0[row] = rD[row] * r;
        y[row] = sentinel();
        v[0] = fabs(r);
    }
    if (!fused) return;
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    const bool last = gridFinalizeLast<OpSum>(gridDim.x, partials, counter, &sc->sumMagR);
    if (last && threadIdx.x == 0) {
        if (upd) { if (doCheck) pcgEndOfIteration(sc, hist); else sc->pendingCheck = 1; }
        sc->nFwd = sc->nFwd + 1;
    }
}

}  // namespace b200
