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
struct TileBar { int dummy; };
__device__ __forceinline__ void barInit(TileBar*) {}
__device__ __forceinline__ void barExpect(TileBar*, unsigned) {}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, TileBar*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void barWait(TileBar*, unsigned) {}
__device__ __forceinline__ void fenceAsyncProxy() {}
__device__ __forceinline__ void bulkStore(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulkStoreCommit() {}
__device__ __forceinline__ void bulkStoreWaitRead() {}
__device__ __forceinline__ void bulkStoreWaitAll() {}
#else
struct TileBar { unsigned long long word; };
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
#define SWEEP_TRACE(slot) do { if (sv.trace && threadIdx.x == 0) sv.trace[8 * ti + (slot)] = globalTimerNs(); } while (0)

struct TileSmem {
    double* vals;          // [nRows + nHalo + 2] rows of the tile (start values, overwritten in place by the results), halo values,
                           // padding slot (1.0), trash slot (0.0).  &vals[r] is 16-byte aligned iff device row row0+r is even.
    double* blk;           // [(d + ceil(d/4)) * S] coefficient columns, then packed 16-bit byte offsets into vals (slot layout)
    int S;
};
__device__ __forceinline__ TileSmem carveTile(unsigned char* smem, const SweepTile& t, int lanes) {
    TileSmem s;
    s.S = t.nSteps * lanes;
    s.vals = (double*)smem + (t.row0 & 1);
    s.blk = (double*)smem + ((t.nRows + t.nHalo + 4) & ~1);
    return s;
}
__device__ __forceinline__ double ldsAt(const double* vals, unsigned byteOff) { return *(const double*)((const char*)vals + byteOff); }
__device__ __forceinline__ void stsAt(double* vals, unsigned byteOff, double v) { *(double*)((char*)vals + byteOff) = v; }

// poll the tile's halo rows out of `src` (sentinel = not yet published) into vals[nRows ..); loads of one pass are independent
__device__ __forceinline__ void pollHalo(const SweepTile& t, const TileSmem& s, const int* __restrict__ haloRow, const double* src, int pollNs) {
    constexpr int U = 4;
    for (int h0 = threadIdx.x; h0 < t.nHalo; h0 += U * blockDim.x) {
        const double* p[U]; double v[U]; bool need[U];
#pragma unroll
        for (int q = 0; q < U; q++) {
            int h = h0 + q * blockDim.x;
            need[q] = h < t.nHalo;
            p[q] = src + (need[q] ? haloRow[t.halo + h] : 0);
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
                    else { s.vals[t.nRows + h0 + q * blockDim.x] = v[q]; need[q] = false; }
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
// OP 0: w -= c*v ; OP 1: w -= c/v (calcReciprocalD: c = upper*upper for DIC, upper*lower for DILU, formed by k_dic_coeffs).  DIR +1: steps upwards (forward sweep), -1: downwards (backward).
// out != null: every row is also stored to out[row] (plain write-through store) in the step that computes it, so that when the
// last step ends only its own rows are still on their way to L2 -- the dependants' polls see the tile ~0.3 us after its last step
// instead of after a whole-tile publish pass.  Each 8-byte value is its own ready flag, so no ordering between the stores is needed,
// and a plain store (unlike a gpu-scope one) does not make the following bar.sync wait for the L2 acknowledge.
// VIRT (chunked plans, host_mesh.h): an entry may be one CHUNK (<= 3 consecutive sources) of a row with more sources.  Its own-row offset
// carries two flags: OWN_PARTIAL -- not the row's last chunk, the partial result stays in the tile (published values are ready flags);
// OWN_CARRY -- it continues the chunk this thread computed in the previous step, whose result is still in `w` (the own-row slot was
// prefetched before that result was stored).  A later chunk that runs two or more steps after its predecessor reads the partial result from
// the row's slot like a start value.
template <int OP, int DIR, int D, bool VIRT>
__device__ __forceinline__ void sweepStepsD(const TileSmem& s, const SweepTile& t, double* out) {
    const int W = blockDim.x, S = s.S;
    const unsigned rowBytes = 8u * (unsigned)t.nRows;
    char* outB = (char*)(out + t.row0);
    int slot = (DIR > 0 ? 0 : (t.nSteps - 1) * W) + (int)threadIdx.x;
    const double* C = s.blk;
    const uint2* PK = (const uint2*)(s.blk + D * S);
    double c0 = C[slot], c1 = D > 1 ? C[S + slot] : 0.0, c2 = D > 2 ? C[2 * S + slot] : 0.0;
    uint2 pk = PK[slot];
    double w = ldsAt(s.vals, VIRT ? (pk.y >> 16) & ~7u : pk.y >> 16);
    for (int it = 0; it < t.nSteps; it++) {
        const double v0 = ldsAt(s.vals, pk.x & 0xffffu);
        const double v1 = D > 1 ? ldsAt(s.vals, pk.x >> 16) : 0.0;
        const double v2 = D > 2 ? ldsAt(s.vals, pk.y & 0xffffu) : 0.0;
        const int nslot = it + 1 < t.nSteps ? slot + DIR * W : slot;
        const double n0 = C[nslot], n1 = D > 1 ? C[S + nslot] : 0.0, n2 = D > 2 ? C[2 * S + nslot] : 0.0;
        const uint2 npk = PK[nslot];
        const unsigned own = VIRT ? (pk.y >> 16) & ~7u : pk.y >> 16;
        const double nw = ldsAt(s.vals, VIRT ? (npk.y >> 16) & ~7u : npk.y >> 16);   // the next step's own rows still hold their start values
        if (OP == 0) { w = w - c0 * v0; if (D > 1) w = w - c1 * v1; if (D > 2) w = w - c2 * v2; }
        else { w = w - c0 / v0; if (D > 1) w = w - c1 / v1; if (D > 2) w = w - c2 / v2; }
        stsAt(s.vals, own, w);
        if (out && own < rowBytes && !(VIRT && ((pk.y >> 16) & OWN_PARTIAL))) stWeak((double*)(outB + own), w);   // (deferring it past the barrier is slower)
        __syncthreads();
        c0 = n0; c1 = n1; c2 = n2; pk = npk; slot = nslot;
        w = (VIRT && ((npk.y >> 16) & OWN_CARRY)) ? w : nw;
    }
}
// ---- register-resident operands (the default for tiles with 1..3 sources per row and at most SWEEP_REG_STEPS steps: every hex mesh).
// Per step the shared-memory loop above moves 72 bytes per row through the SM's 128 B/clk shared-memory pipe (3 sources, 3 coefficients,
// the packed offsets, the next start value, the result): 256 rows x 72 B = 144 cycles, more than the dependent-instruction chain of the
// step.  Coefficients and offsets are constant while the tile runs, so each thread loads those of ITS rows (one per step) straight from
// global memory into registers when it takes the tile -- before the halo wait, off the critical path -- and the step loop, fully
// unrolled, touches shared memory only for the three source values, the start value and the result (40 B per row).  Tiles with fewer
// than three sources carry zero coefficients that point at the padding slot (w - 0*1 = w exactly): one code path for d = 1..3.
constexpr int SWEEP_REG_STEPS = 16;
struct RegOps { double c[SWEEP_REG_STEPS][3]; uint2 pk[SWEEP_REG_STEPS]; };
#ifdef B200_EMU
__device__ __forceinline__ double ldgStream(const double* p) { return *p; }
__device__ __forceinline__ uint2 ldgStream(const uint2* p) { return *p; }
#else
__device__ __forceinline__ double ldgStream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ uint2 ldgStream(const uint2* p) { return __ldcs(p); }
#endif
// register index `it` = the it-th step in processing order (DIR < 0: the tile's steps downwards)
template <int DIR>
__device__ __forceinline__ void loadRegOps(RegOps& o, const double* __restrict__ blk, const SweepTile& t) {
    const int W = blockDim.x, S = t.nSteps * W;
    const uint2* PK = (const uint2*)(blk + t.d * S);
    if (DIR > 0) {
        // forward sweep: warps without entries in a step (t.stepWarps) skip that step's loads -- their operands are the inert ones (zero
        // coefficients, sources = the padding slot, own row = the trash slot); partly filled tiles otherwise stream mostly padding.
        // (Forward only: the same code in the backward kernel makes ptxas lay its step loop out 7 us slower, profiles/r2_ncu_summary.md.)
        const unsigned padOff = 8u * (unsigned)(t.nRows + t.nHalo), warp = threadIdx.x >> 5;
        const uint2 inert = make_uint2(padOff | (padOff << 16), padOff | ((padOff + 8u) << 16));
#pragma unroll
        for (int it = 0; it < SWEEP_REG_STEPS; it++) {
            const bool on = it < t.nSteps && warp < ((t.stepWarps[(it >> 3) & 1] >> (4 * (it & 7))) & 15u);
            const int slot = (on ? it * W : 0) + (int)threadIdx.x;
            o.c[it][0] = on ? ldgStream(blk + slot) : 0.0;
            o.c[it][1] = (on && t.d > 1) ? ldgStream(blk + S + slot) : 0.0;
            o.c[it][2] = (on && t.d > 2) ? ldgStream(blk + 2 * S + slot) : 0.0;
            o.pk[it] = on ? ldgStream(PK + slot) : inert;
        }
        return;
    }
#pragma unroll
    for (int it = 0; it < SWEEP_REG_STEPS; it++) {
        const bool on = it < t.nSteps;
        const int slot = (on ? (DIR > 0 ? it : t.nSteps - 1 - it) * W : 0) + (int)threadIdx.x;
        o.c[it][0] = on ? ldgStream(blk + slot) : 0.0;
        o.c[it][1] = (on && t.d > 1) ? ldgStream(blk + S + slot) : 0.0;
        o.c[it][2] = (on && t.d > 2) ? ldgStream(blk + 2 * S + slot) : 0.0;
        o.pk[it] = on ? ldgStream(PK + slot) : make_uint2(0u, 0u);
    }
}
// local rows [lo, hi) of the tile -> out by one TMA bulk store (16-byte aligned middle) + scalar gpu-scope stores for odd ends; one thread,
// after the barrier that follows the writers' fence.proxy.async
__device__ __forceinline__ void publishRange(const TileSmem& s, const SweepTile& t, double* out, int lo, int hi) {
    if (hi <= lo) return;
    const int a = lo + ((t.row0 + lo) & 1), e = hi - ((t.row0 + hi) & 1);
    if (e > a) {
        for (int q = a; q < e; q += (int)(BULK_MAX / 8)) bulkStore(out + t.row0 + q, s.vals + q, (unsigned)((e - q < (int)(BULK_MAX / 8) ? e - q : (int)(BULK_MAX / 8)) * 8));
        bulkStoreCommit();
    }
    if (a > lo) stPublish(out + t.row0 + lo, s.vals[lo]);
    if (e < hi && e >= a) stPublish(out + t.row0 + hi - 1, s.vals[hi - 1]);
}
// BULK > 0 (hybrid publish): the rows of the first BULK processing steps leave by ONE asynchronous TMA bulk store issued when
// they are complete -- it lands while the remaining steps run -- and only the last steps store their rows in the loop (their rows are the
// ones the dependants wait for).  A store in the step loop costs every step ~28 ns (the following bar.sync waits for it).
// (BULK is a compile-time constant: the step loop is so short that two run-time tests per step cost a quarter of it)
template <int DIR, int BULK, bool VIRT>
__device__ __forceinline__ void sweepStepsReg(const TileSmem& s, const SweepTile& t, const RegOps& o, double* out, int bulkLo = 0, int bulkHi = 0) {
    const unsigned rowBytes = 8u * (unsigned)t.nRows;
    char* outB = (char*)(out + t.row0);
    double w = ldsAt(s.vals, VIRT ? (o.pk[0].y >> 16) & ~7u : o.pk[0].y >> 16);
#pragma unroll
    for (int it = 0; it < SWEEP_REG_STEPS; it++) {
        if (it < t.nSteps) {   // (uniform: no divergence; the unrolled body indexes the register arrays statically)
            const uint2 pk = o.pk[it];
            const double v0 = ldsAt(s.vals, pk.x & 0xffffu);
            const double v1 = ldsAt(s.vals, pk.x >> 16);
            const double v2 = ldsAt(s.vals, pk.y & 0xffffu);
            // the next step's own rows still hold their start values
            const int nit = it + 1 < SWEEP_REG_STEPS ? it + 1 : it;   // (compile-time after unrolling)
            const unsigned nownRaw = o.pk[nit].y >> 16, own = VIRT ? (pk.y >> 16) & ~7u : pk.y >> 16;
            const double nw = (nit != it && nit < t.nSteps) ? ldsAt(s.vals, VIRT ? nownRaw & ~7u : nownRaw) : 0.0;
            w = w - o.c[it][0] * v0; w = w - o.c[it][1] * v1; w = w - o.c[it][2] * v2;
            stsAt(s.vals, own, w);
            if (it >= BULK) { if (out && own < rowBytes && !(VIRT && ((pk.y >> 16) & OWN_PARTIAL))) stWeak((double*)(outB + own), w); }
            if (it + 1 == BULK) fenceAsyncProxy();      // this thread's results of the bulk steps, for the TMA store below
            __syncthreads();
            if (it + 1 == BULK) { if (threadIdx.x == 0) publishRange(s, t, out, bulkLo, bulkHi); }
            w = (VIRT && nit != it && (nownRaw & OWN_CARRY)) ? w : nw;
        }
    }
}
// hybrid publish of a register-path tile: how many leading processing steps go by bulk store, and which local rows they cover
constexpr int SWEEP_BULK_STEPS = 12;      // full tiles (16 steps): steps 0..11 by bulk store, 12..15 in the loop
// (unchunked full tiles only; the row range was counted on the host: forward = the first rows, backward = the last rows, device_mesh.cu)
__device__ __forceinline__ bool bulkPlan(const SweepView& sv, const SweepTile& t, int& lo, int& hi) {
    lo = hi = 0;
    if (sv.publishMode != 2 || t.bulkHi == 0 || t.nSteps != SWEEP_REG_STEPS) return false;
    lo = t.bulkLo; hi = t.bulkHi;
    return true;
}
__device__ __forceinline__ bool tileUsesRegOps(const SweepView& sv, const SweepTile& t) {
    return sv.regOps && t.d >= 1 && t.d <= 3 && t.nSteps >= 1 && t.nSteps <= SWEEP_REG_STEPS;
}
// shared-memory operand path (calcReciprocalD; tiles with more than SWEEP_REG_STEPS steps).  Returns whether the rows were published in the loop.
// shared-memory operand path (calcReciprocalD; tiles with more than SWEEP_REG_STEPS steps).  Returns whether the rows were published in the loop.
// The last branch is the plain reference walk of a tile (any width, one entry per thread and step, straight from the slot table); it runs the
// tiles of isolated rows (d = 0), whose start values already are the results.
// (Keep it: without this loop in the function ptxas lays the unrolled register path of k_dic_bwd out 4 % slower -- measured A/B on one box,
// profiles/r2_ncu_summary.md.)
template <int OP, int DIR, bool VIRT>
__device__ __forceinline__ bool sweepSteps(const TileSmem& s, const SweepTile& t, double* out, const int* __restrict__ slotRow) {
    if (t.nSteps == 0) return false;
    if (t.d == 3) { sweepStepsD<OP, DIR, 3, VIRT>(s, t, out); return out != nullptr; }
    if (t.d == 2) { sweepStepsD<OP, DIR, 2, VIRT>(s, t, out); return out != nullptr; }
    if (t.d == 1) { sweepStepsD<OP, DIR, 1, VIRT>(s, t, out); return out != nullptr; }
    const int W = blockDim.x, S = s.S;
    const unsigned short* OFF = (const unsigned short*)(s.blk + t.d * S);   // [S][4]
    for (int k = DIR > 0 ? 0 : t.nSteps - 1; k >= 0 && k < t.nSteps; k += DIR) {
        const int slot = k * W + threadIdx.x, row = slotRow[t.slot0 + slot];
        if (row >= 0) {
            const int r = row - t.row0;
            double w = s.vals[r];
            for (int j = 0; j < t.d; j++) {
                double c = s.blk[j * S + slot], v = ldsAt(s.vals, OFF[slot * 4 + j]);
                if (OP == 0) w = w - c * v; else w = w - c / v;
            }
            s.vals[r] = w;
        }
        __syncthreads();
    }
    return false;
}
// publish of a whole tile by all threads (consecutive threads -> consecutive rows); call after the barrier that ends the steps.
// Only tiles of isolated rows (no sources) leave this way.  Measured on the 985k-cell channel: a whole-tile publish after the
// last step, by TMA bulk store or by these stores, is seen by the dependants ~1.0 us after the last step; step-by-step stores ~0.45 us.
__device__ __forceinline__ void publishTileDirect(const TileSmem& s, const SweepTile& t, double* out) {
    for (int r = threadIdx.x; r < t.nRows; r += blockDim.x) stWeak(out + t.row0 + r, s.vals[r]);
}

// stage the operand block of a tile: one elected thread issues the TMA bulk copies (the whole block is contiguous).
// offsetsOnly: just the packed-offset columns (calcReciprocalD gathers its coefficients itself)
__device__ __forceinline__ void stageBlock(const TileSmem& s, const SweepTile& t, const double* __restrict__ blocks, TileBar* bar, bool offsetsOnly) {
    if (threadIdx.x == 0) {
        const size_t skip = offsetsOnly ? (size_t)t.d * s.S : 0;
        const size_t bytes = ((size_t)(t.d + ((t.d + 3) >> 2)) * s.S - skip) * 8;
        fenceAsyncProxy();
        barExpect(bar, (unsigned)bytes);
        bulkLoadChunks(s.blk + skip, blocks + t.blk + skip, bytes, bar);
    }
}
// publish the tile's rows: the results sit contiguously in vals[0..nRows) = device rows [row0, row0+nRows): one TMA bulk
// store for the 16-byte aligned middle, scalar gpu-scope stores for an odd first / last row.  Called by one thread after the
// CTA barrier that ends the last step.
__device__ __forceinline__ void publishTile(const TileSmem& s, const SweepTile& t, double* out) {
    const int a = t.row0 & 1, e = t.nRows - ((t.row0 + t.nRows) & 1);   // local range [a, e) is aligned
    if (e > a) {
        fenceAsyncProxy();
        for (int o = a; o < e; o += (int)(BULK_MAX / 8)) bulkStore(out + t.row0 + o, s.vals + o, (unsigned)((e - o < (int)(BULK_MAX / 8) ? e - o : (int)(BULK_MAX / 8)) * 8));
        bulkStoreCommit();
    }
    if (a && t.nRows > 0) stPublish(out + t.row0, s.vals[0]);
    if (e < t.nRows && e >= a) stPublish(out + t.row0 + t.nRows - 1, s.vals[t.nRows - 1]);
}

// The tail of upstream's PCG iteration, fused into the staging of the next forward sweep (the sweep reads rA anyway):
//   alpha = wArA/wApA ; x += alpha*pA ; rA -= alpha*wA ; sum|rA|  ->  finalResidual, nIter++, stop  [UPSTREAM PCG.C]
// The last CTA of the sweep finishes the (fixed-order) residual sum and runs the convergence test (single rank), or leaves the
// local sum for the all-reduce + k_pcg_check (multi-rank).  When the test says stop, this sweep was the one wasted pass.
struct PcgFuse {
    double* x;            // null: plain sweep (b200_dic_precondition)
    double* rA;
    const double *pA, *wA;
    PcgScalars* sc;
    double* partials; unsigned* counter; double* hist;
    int doCheck;
};

// MODE 0: forward precondition sweep: y[row] = rD*rA - sum_j cf_j * y[src_j]              (cf = rD[row]*upper, per solve)
// MODE 1: calcReciprocalD           : Dt[row] = diag - sum_j (up_j*lo_j) / Dt[src_j] ; rD = 1/Dt (operand block = the products, lo = up for DIC)
// VIRT: the kernels of a chunked plan (every tile of a mesh is, or none: chosen at launch -- the unchunked kernels carry none of it)
template <int MODE, bool VIRT>
static __global__ void __launch_bounds__(SWEEP_THREADS) k_dic_fwd(SweepView sv, const double* __restrict__ a0,
        const double* a1, const double* __restrict__ coef, const double* __restrict__ blocks, double* y,
        double* __restrict__ rDout, SweepCtl ctl, const PcgScalars* sc, PcgFuse fz) {
    B200_DYN_SMEM(smem);
    __shared__ TileBar sBar;
    if (threadIdx.x == 0) barInit(&sBar);
    unsigned phase = 0;
    const int W = blockDim.x, tid = threadIdx.x;
    // prologue (may overlap the tail of the previous kernel, PDL): first ticket, tile descriptor, TMA staging of the operand block --
    // all static during the iterations of a solve.  Nothing the previous kernel writes is touched before pdlWait().
    int ti = nextTicket(ctl);
    SweepTile t; TileSmem s;
    RegOps ro;
    bool viaReg = false;   // this tile's operands live in registers (no TMA staging in flight)
    if (ti < sv.nTiles) {
        t = sv.fwd[ti]; s = carveTile(smem, t, W);
        viaReg = MODE == 0 && tileUsesRegOps(sv, t);
        if (viaReg) loadRegOps<1>(ro, blocks + t.blk, t);
        else stageBlock(s, t, blocks, &sBar, false);   // MODE 1: the coefficient columns hold the plain upper coefficients (k_dic_coeffs, rD = null)
    }
    pdlWait();
    pdlTrigger();
    const bool fused = MODE == 0 && fz.x != nullptr;
    bool quit = sc && sc->stop;
    bool upd = false;
    double alpha = 0.0;
    if (!quit && fused) {
        upd = fz.sc->nFwd > 0;          // an iteration's x / rA update is pending
        if (upd) {
            if (fabs(fz.sc->wApA) / fz.sc->normFactor <= 1.0e-300) {   // checkSingularity [UPSTREAM PCG.C]: leave x, rA, nIter untouched
                __syncthreads();   // every thread of the CTA has read the scalars
                if (blockIdx.x == 0 && threadIdx.x == 0) { fz.sc->singular = 1; fz.sc->stop = 1; }
                quit = true;
            } else alpha = fz.sc->wArA / fz.sc->wApA;
        }
    }
    if (quit) {   // hand the ticket back (the last CTA to leave resets the counter) once the staging copy has landed
        if (ti < sv.nTiles && !viaReg) barWait(&sBar, phase);
        sweepExit(ctl);
        return;
    }
    while (ti < sv.nTiles) {
        SWEEP_TRACE(0);
        const int n = t.nRows;
        // 1. row-independent operands: the operand block by TMA bulk copy (already on its way), start values by batched loads
        double mag = 0.0;
        constexpr int SU = MODE == 0 ? 4 : 8;   // loads in flight per thread (MODE 0 shares the register file with the tile's operands)
        for (int r0 = tid; r0 < n; r0 += SU * W) {
            double a[SU], b[SU], wa[SU], xv[SU], pv[SU];
#pragma unroll
            for (int k = 0; k < SU; k++) {
                int r = r0 + k * W;
                if (r < n) {
                    a[k] = a0[t.row0 + r]; b[k] = (MODE == 0) ? a1[t.row0 + r] : 1.0;
                    if (upd) { wa[k] = fz.wA[t.row0 + r]; xv[k] = fz.x[t.row0 + r]; pv[k] = fz.pA[t.row0 + r]; }
                }
            }
#pragma unroll
            for (int k = 0; k < SU; k++) {
                int r = r0 + k * W;
                if (r < n) {
                    if (upd) { fz.x[t.row0 + r] = xv[k] + alpha * pv[k]; b[k] = b[k] - alpha * wa[k]; fz.rA[t.row0 + r] = b[k]; }
                    if (fused) mag = mag + fabs(b[k]);
                    s.vals[r] = (MODE == 0) ? a[k] * b[k] : a[k];
                }
            }
        }
        if (fused) { double v[1] = {mag}; gridReduce<1, OpSum>(v, ti, sv.nTiles, fz.partials, fz.counter, nullptr); }
        if (tid == 0) { s.vals[n + t.nHalo] = 1.0; s.vals[n + t.nHalo + 1] = 0.0; }
        // 2. external inputs
        if (sv.trace) { __syncthreads(); SWEEP_TRACE(1); }
        pollHalo(t, s, sv.fwdHaloRow, y, sv.pollNs);
        if (!viaReg) { barWait(&sBar, phase); phase ^= 1u; }
        __syncthreads();
        SWEEP_TRACE(2);
        // 3. the steps, then publish
        bool inLoop;
        if (viaReg) {
            int blo, bhi;
            if (VIRT) sweepStepsReg<1, 0, true>(s, t, ro, sv.publishMode >= 1 ? y : nullptr);
            else if (bulkPlan(sv, t, blo, bhi)) {
                sweepStepsReg<1, SWEEP_BULK_STEPS, false>(s, t, ro, y, blo, bhi);
                if (tid == 0) bulkStoreWaitRead();          // (the value array is reused by the next tile)
            } else sweepStepsReg<1, 0, false>(s, t, ro, sv.publishMode >= 1 ? y : nullptr);
            inLoop = sv.publishMode >= 1;
        }
        else inLoop = sweepSteps<MODE, 1, VIRT>(s, t, sv.publishMode >= 1 ? y : nullptr, sv.fwdSlotRow);
        if (sv.publishMode == 0) {
            fenceAsyncProxy();    // this thread's results, for the TMA store below (writers fence, barrier, one thread stores)
            __syncthreads();
            SWEEP_TRACE(3);
            if (tid == 0) publishTile(s, t, y);
        } else {
            SWEEP_TRACE(3);
            if (!inLoop) publishTileDirect(s, t, y);
        }
        if (MODE == 1) for (int r = tid; r < n; r += W) rDout[t.row0 + r] = 1.0 / s.vals[r];
        if (sv.publishMode == 0 && tid == 0) bulkStoreWaitRead();   // the value array is reused by the next tile
        SWEEP_TRACE(4);
        ti = nextTicket(ctl);
        if (ti < sv.nTiles) {
            t = sv.fwd[ti]; s = carveTile(smem, t, W);
            viaReg = MODE == 0 && tileUsesRegOps(sv, t);
            if (viaReg) loadRegOps<1>(ro, blocks + t.blk, t); else stageBlock(s, t, blocks, &sBar, false);
        }
    }
    if (threadIdx.x == 0) bulkStoreWaitAll();   // this CTA's TMA stores must have landed before the kernel ends
    if (fused) {
        const bool last = gridFinalizeLast<OpSum>(sv.nTiles, fz.partials, fz.counter, &fz.sc->sumMagR);
        if (last && threadIdx.x == 0) {
            if (upd) { if (fz.doCheck) pcgEndOfIteration(fz.sc, fz.hist); else fz.sc->pendingCheck = 1; }
            fz.sc->nFwd = fz.sc->nFwd + 1;
        }
    }
    sweepExit(ctl);
}

// backward: z[row] = y[row] - sum_{owner faces DESC} cb_j * z[u_j]   (cb = rD[row]*upper) ; y[row] <- sentinel (re-armed) ;
// per-tile partial of sum(z*rA) (fixed rows per partial, fixed order => deterministic) -> *dotOut (sc->wArA; multi-rank: sc->wArAnew)
// fold (multi-rank): the last CTA also runs the ONE scalar all-reduce of the iteration's first half -- (sum|rA| left by the forward
// sweep, sum(wA*rA) of this sweep; adjacent in PcgScalars) -- over NVLink peer memory and then closes the previous iteration
// (k_pcg_check's body): no stand-alone 32-thread kernel between the sweeps and the p-update
template <bool VIRT>
static __global__ void __launch_bounds__(SWEEP_THREADS) k_dic_bwd(SweepView sv, const double* __restrict__ blocks,
        const double* __restrict__ rA, double* y, double* z, SweepCtl ctl, PcgScalars* sc, double* partials, unsigned* counter,
        double* dotOut, P2PFold fold, double* hist) {
    B200_DYN_SMEM(smem);
    __shared__ TileBar sBar;
    if (threadIdx.x == 0) barInit(&sBar);
    unsigned phase = 0;
    const int W = blockDim.x, tid = threadIdx.x;
    // prologue (PDL, see k_dic_fwd): first ticket, descriptor, operand block
    int tk = nextTicket(ctl), ti = sv.nTiles - 1 - tk;
    SweepTile t; TileSmem s;
    RegOps ro;
    bool viaReg = false;
    if (tk < sv.nTiles) {
        t = sv.bwd[ti]; s = carveTile(smem, t, W);
        viaReg = tileUsesRegOps(sv, t);
        if (viaReg) loadRegOps<-1>(ro, blocks + t.blk, t); else stageBlock(s, t, blocks, &sBar, false);
    }
    pdlWait();
    pdlTrigger();
    if (sc && sc->stop) {
        if (tk < sv.nTiles && !viaReg) barWait(&sBar, phase);
        sweepExit(ctl);
        return;
    }
    while (tk < sv.nTiles) {
        const int n = t.nRows;
        for (int r0 = tid; r0 < n; r0 += 8 * W) {
            double a[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { int r = r0 + k * W; if (r < n) a[k] = y[t.row0 + r]; }
#pragma unroll
            for (int k = 0; k < 8; k++) { int r = r0 + k * W; if (r < n) { s.vals[r] = a[k]; y[t.row0 + r] = sentinel(); } }
        }
        if (tid == 0) { s.vals[n + t.nHalo] = 1.0; s.vals[n + t.nHalo + 1] = 0.0; }
        pollHalo(t, s, sv.bwdHaloRow, z, sv.pollNs);
        if (!viaReg) { barWait(&sBar, phase); phase ^= 1u; }
        __syncthreads();
        bool inLoop;
        if (viaReg) {
            int blo, bhi;
            if (VIRT) sweepStepsReg<-1, 0, true>(s, t, ro, sv.publishMode >= 1 ? z : nullptr);
            else if (bulkPlan(sv, t, blo, bhi)) {
                sweepStepsReg<-1, SWEEP_BULK_STEPS, false>(s, t, ro, z, blo, bhi);
                if (tid == 0) bulkStoreWaitRead();
            } else sweepStepsReg<-1, 0, false>(s, t, ro, sv.publishMode >= 1 ? z : nullptr);
            inLoop = sv.publishMode >= 1;
        }
        else inLoop = sweepSteps<0, -1, VIRT>(s, t, sv.publishMode >= 1 ? z : nullptr, sv.bwdSlotRow);
        if (sv.publishMode == 0) {
            fenceAsyncProxy();
            __syncthreads();
            if (tid == 0) publishTile(s, t, z);
        } else if (!inLoop) publishTileDirect(s, t, z);
        if (partials) {
            double dotv = 0.0;
            for (int r = tid; r < n; r += W) dotv = dotv + s.vals[r] * rA[t.row0 + r];
            double v[1] = {dotv};
            gridReduce<1, OpSum>(v, ti, sv.nTiles, partials, counter, nullptr);
        }
        if (sv.publishMode == 0 && tid == 0) bulkStoreWaitRead();
        tk = nextTicket(ctl); ti = sv.nTiles - 1 - tk;
        if (tk < sv.nTiles) {
            t = sv.bwd[ti]; s = carveTile(smem, t, W);
            viaReg = tileUsesRegOps(sv, t);
            if (viaReg) loadRegOps<-1>(ro, blocks + t.blk, t); else stageBlock(s, t, blocks, &sBar, false);
        }
    }
    if (threadIdx.x == 0) bulkStoreWaitAll();
    if (partials) {
        const bool last = gridFinalizeLast<OpSum>(sv.nTiles, partials, counter, dotOut);
        if (last && fold.on) {
            if (threadIdx.x < 32) p2pAllreduceWarp(&sc->sumMagR, 2, 0, fold.v, fold.epoch);
            __syncthreads();
            if (threadIdx.x == 0 && !sc->stop) {
                if (sc->pendingCheck) { sc->pendingCheck = 0; pcgEndOfIteration(sc, hist); }
                if (!sc->stop) sc->wArA = sc->wArAnew;
            }
        }
    }
    sweepExit(ctl);
}

// per solve: coefficient columns of the operand blocks: cf = rD[row]*srcF (forward), cb = rD[row]*srcB (backward), where
//   DIC                          : srcF = srcB = upper            [UPSTREAM DICPreconditioner::precondition]
//   DILU precondition            : srcF = lower, srcB = upper     [UPSTREAM DILUPreconditioner::precondition]
//   DILU preconditionT           : srcF = upper, srcB = lower     [UPSTREAM DILUPreconditioner::preconditionT]
// rD == null (calcReciprocalD): cf = srcF*srcB, the face products upper*upper (DIC) / upper*lower (DILU) the sweep divides by Dt.
// grid = 2*nTiles (forward tiles, then backward tiles) or nTiles (forward only); block = lanes
static __global__ void __launch_bounds__(SWEEP_THREADS) k_dic_coeffs(SweepView sv, const double* __restrict__ rD,
        const double* __restrict__ srcF, const double* __restrict__ srcB, double* __restrict__ cf, double* __restrict__ cb, const PcgScalars* gate) {
    if (gate && gate->stop) return;
    const bool bw = (int)blockIdx.x >= sv.nTiles;
    const SweepTile t = bw ? sv.bwd[blockIdx.x - sv.nTiles] : sv.fwd[blockIdx.x];
    const int* pos = bw ? sv.bwdPos : sv.fwdPos;
    const int* slotRow = bw ? sv.bwdSlotRow : sv.fwdSlotRow;
    double* out = bw ? cb : cf;
    const int W = blockDim.x, S = t.nSteps * W;
    for (int k = 0; k < t.nSteps; k++) {
        const int row = slotRow[t.slot0 + k * W + threadIdx.x];
        const bool valid = row >= 0;
        const double rd = (valid && rD) ? rD[row] : 0.0;
        for (int j = 0; j < t.d; j++) {
            const int e = t.blk + j * S + k * W + threadIdx.x;
            const int p = valid ? pos[e] : -1;
            out[e] = p < 0 ? 0.0 : rD ? rd * (bw ? srcB : srcF)[p] : srcF[p] * srcB[p];
        }
    }
}

// DIC reuse (pcg.cuh): compares this solve's diag / upper with the kept copies bit by bit, refreshes the copies, and opens the gate
// (gate->stop = 0) when anything differs.  k_gate_close runs first.
static __global__ void k_gate_close(PcgScalars* gate) { if (threadIdx.x == 0 && blockIdx.x == 0) gate->stop = 1; }
static __global__ void __launch_bounds__(BLOCK) k_matrix_changed(int N, int Fs, const double* __restrict__ diag, const double* __restrict__ upS,
        double* __restrict__ keptDiag, double* __restrict__ keptUp, PcgScalars* gate) {
    bool diff = false;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N + Fs; i += stride) {
        const double* src = i < N ? diag + i : upS + (i - N);
        double* kept = i < N ? keptDiag + i : keptUp + (i - N);
        const long long a = __double_as_longlong(*src);
        if (a != __double_as_longlong(*kept)) { *kept = __longlong_as_double(a); diff = true; }
    }
    if (diff) gate->stop = 0;
}

}  // namespace b200
