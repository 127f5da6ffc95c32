// ldu_kernels.cuh -- sm_100a kernels of the lduMatrix solve: Amul / sumA (row gather on the owner-slot SELL
// layout), data-flow level-scheduled DIC sweeps, fused PCG vector updates and deterministic reductions.
// FP64, HBM-bound, no tensor cores.  Compiled with --fmad=false: every expression is evaluated in upstream's
// operation order with IEEE mul/add (foam-extend's x86-64 build has no FMA contraction), so Amul and DIC are
// bitwise equal to the sequential loops of lduMatrixATmul.C / DICPreconditioner.C.
#pragma once
#include "device_mesh.cuh"
#include "comm.cuh"

namespace b200 {

// ----------------------------------------------------------------------------- deterministic reductions
// block tree (shuffle + smem) -> partials[slot] ; the last block to finish sums all partials in a fixed order.
struct OpSum { __device__ static double id() { return 0.0; } __device__ static double op(double a, double b) { return a + b; } };
struct OpMax { __device__ static double id() { return -1.0e300; } __device__ static double op(double a, double b) { return a > b ? a : b; } };
struct OpMin { __device__ static double id() { return 1.0e300; } __device__ static double op(double a, double b) { return a < b ? a : b; } };

template <class Op>
__device__ __forceinline__ double warpReduce(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::op(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
// result valid in thread 0
template <class Op>
__device__ __forceinline__ double blockReduce(double v, double* sh /*[32]*/) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warpReduce<Op>(v);
    __syncthreads();  // protect sh reuse
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    v = (threadIdx.x < (unsigned)nw) ? sh[threadIdx.x] : Op::id();
    if (warp == 0) v = warpReduce<Op>(v);
    return v;
}
// NV values per block -> partials[slot*NV+i]; last block writes out[i] = reduce over nSlots slots (fixed order)
template <int NV, class Op0, class Op1 = Op0, class Op2 = Op0>
__device__ __forceinline__ void gridReduce(double (&v)[NV], int slot, int nSlots, double* partials, unsigned* counter,
                                           double* out0, double* out1 = nullptr, double* out2 = nullptr) {
    __shared__ double sh[32];
    __shared__ int isLast;
    double r0 = blockReduce<Op0>(v[0], sh), r1 = 0, r2 = 0;
    if (NV > 1) r1 = blockReduce<Op1>(v[1 % NV], sh);
    if (NV > 2) r2 = blockReduce<Op2>(v[2 % NV], sh);
    if (threadIdx.x == 0) {
        partials[(size_t)slot * NV] = r0;
        if (NV > 1) partials[(size_t)slot * NV + 1] = r1;
        if (NV > 2) partials[(size_t)slot * NV + 2] = r2;
    }
    (void)nSlots;
}
// called once per block after its last gridReduce slot was written
template <int NV, class Op0, class Op1 = Op0, class Op2 = Op0>
__device__ __forceinline__ void gridFinalize(int nSlots, const double* partials, unsigned* counter, double* out0,
                                             double* out1 = nullptr, double* out2 = nullptr) {
    __shared__ double sh[32];
    __shared__ int isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicInc(counter, gridDim.x - 1);  // wraps back to 0 for the next launch
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const volatile double* vp = partials;
    double a0 = Op0::id(), a1 = Op1::id(), a2 = Op2::id();
    for (int j = threadIdx.x; j < nSlots; j += blockDim.x) {
        a0 = Op0::op(a0, vp[(size_t)j * NV]);
        if (NV > 1) a1 = Op1::op(a1, vp[(size_t)j * NV + 1]);
        if (NV > 2) a2 = Op2::op(a2, vp[(size_t)j * NV + 2]);
    }
    a0 = blockReduce<Op0>(a0, sh);
    if (NV > 1) a1 = blockReduce<Op1>(a1, sh);
    if (NV > 2) a2 = blockReduce<Op2>(a2, sh);
    if (threadIdx.x == 0) {
        *out0 = a0;
        if (NV > 1 && out1) *out1 = a1;
        if (NV > 2 && out2) *out2 = a2;
    }
}

// same, but every thread of the block learns whether it was the last block (after the result was written)
template <class Op0>
__device__ __forceinline__ bool gridFinalizeLast(int nSlots, const double* partials, unsigned* counter, double* out0) {
    __shared__ double sh[32];
    __shared__ int isLast;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicInc(counter, gridDim.x - 1);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return false;
    __threadfence();
    const volatile double* vp = partials;
    double a0 = Op0::id();
    for (int j = threadIdx.x; j < nSlots; j += blockDim.x) a0 = Op0::op(a0, vp[j]);
    a0 = blockReduce<Op0>(a0, sh);
    if (threadIdx.x == 0) *out0 = a0;
    __syncthreads();
    return true;
}

// ----------------------------------------------------------------------------- PCG device scalars
struct PcgScalars {
    double sumX, sumMagR, wArAnew, nfSum, wArA, wArAold, wApA, normFactor, initialResidual, finalResidual, tol, relTol, nGlobal;
    // (sumMagR, wArAnew are adjacent: the multi-rank path all-reduces them in one call)
    int nIter, minIter, maxIter, stop, singular, converged, histCap, nFwd;   // nFwd: forward sweeps run in this solve
    int pendingCheck, pad;   // multi-rank: the forward sweep left a residual sum for k_pcg_check
};

__device__ __forceinline__ bool isSentinel(double v) { return (unsigned long long)__double_as_longlong(v) == SENTINEL_BITS; }
__device__ __forceinline__ double sentinel() { return __longlong_as_double((long long)SENTINEL_BITS); }
// gpu-scope relaxed accesses: served by L2 (never a stale L1 line), single-copy atomic for the aligned 8 bytes
#ifdef B200_EMU
__device__ __forceinline__ double ldPoll(const double* p) { return *(const volatile double*)p; }
__device__ __forceinline__ void stPublish(double* p, double v) { *(volatile double*)p = v; }
__device__ __forceinline__ void stWeak(double* p, double v) { *(volatile double*)p = v; }
#else
// plain (weak) write-through store that the compiler may not drop or move across the step barrier
__device__ __forceinline__ void stWeak(double* p, double v) { asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ldPoll(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stPublish(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#endif

// read-only loads that stay where the source puts them (volatile asm keeps its program order): the gather kernels issue ALL loads
// of a row before the first dependent FP64 instruction -- left to itself the compiler sinks each load next to its use to save
// registers, and the in-order issue then exposes one memory round trip per face
#ifdef B200_EMU
__device__ __forceinline__ double ldgOrdered(const double* p) { return *p; }
__device__ __forceinline__ int ldgOrdered(const int* p) { return *p; }
#else
__device__ __forceinline__ double ldgOrdered(const double* p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ldgOrdered(const int* p) { int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
#endif
// load that is served by L2 (values a peer GPU stored into this GPU's memory must not come from a stale L1 line)
#ifdef B200_EMU
__device__ __forceinline__ double ldNoL1(const double* p) { return *p; }
#else
__device__ __forceinline__ double ldNoL1(const double* p) { return __ldcg(p); }
#endif

// ----------------------------------------------------------------------------- a1: Amul (row gather)
// Apsi[row] = diag*psi + sum_{nbr-side faces asc} lower_f*psi[l] + sum_{owner faces asc} upper_f*psi[u]
// rowOffDiag: the row's off-diagonal products, accumulated onto acc in upstream's order.
// Generic form (any slice widths: polyhedral meshes).  The faces of each side are taken four at a time: four index loads, then the
// four slice offsets, then coefficients and psi values, all issued in program order before the four additions (ldgOrdered); padding
// entries load slot 0 / the row's own psi so that no load is predicated.  (Hex meshes take the pipelined loop of k_amul instead.)
template <bool WITH_SUM>
__device__ __forceinline__ double rowOffDiag(const MeshView& mv, const double* __restrict__ upS, const double* __restrict__ loS,
                                             const double* __restrict__ psi, int row, double acc, double& sA) {
    constexpr int GW = 4;
    const int s = row >> 5, lane = row & 31, kmask = (1 << mv.KB) - 1;
    const int nb0 = mv.nbrOff[s], nbw = (mv.nbrOff[s + 1] - nb0) >> 5;
    const int so0 = mv.sliceOff[s], ow = (mv.sliceOff[s + 1] - so0) >> 5;
    for (int j0 = 0; j0 < nbw; j0 += GW) {
        int pk[GW], so[GW];
        double cl[GW], pl[GW];
#pragma unroll
        for (int i = 0; i < GW; i++) pk[i] = j0 + i < nbw ? ldgOrdered(mv.loPk + nb0 + ((j0 + i) << 5) + lane) : -1;
#pragma unroll
        for (int i = 0; i < GW; i++) so[i] = pk[i] >= 0 ? ldgOrdered(mv.sliceOff + ((pk[i] >> mv.KB) >> 5)) : 0;
#pragma unroll
        for (int i = 0; i < GW; i++) {
            const int l = pk[i] >= 0 ? pk[i] >> mv.KB : row;
            const int pos = pk[i] >= 0 ? so[i] + ((pk[i] & kmask) << 5) + (l & 31) : 0;
            cl[i] = ldgOrdered(loS + pos); pl[i] = ldgOrdered(psi + l);
        }
#pragma unroll
        for (int i = 0; i < GW; i++)
            if (pk[i] >= 0) { acc = acc + cl[i] * pl[i]; if (WITH_SUM) sA = sA + cl[i]; }
    }
    for (int k0 = 0; k0 < ow; k0 += GW) {
        int u[GW];
        double cu[GW], pu[GW];
#pragma unroll
        for (int i = 0; i < GW; i++) u[i] = k0 + i < ow ? ldgOrdered(mv.upIdx + so0 + ((k0 + i) << 5) + lane) : -1;
#pragma unroll
        for (int i = 0; i < GW; i++) {
            cu[i] = ldgOrdered(upS + (k0 + i < ow ? so0 + ((k0 + i) << 5) + lane : 0));
            pu[i] = ldgOrdered(psi + (u[i] >= 0 ? u[i] : row));
        }
#pragma unroll
        for (int i = 0; i < GW; i++)
            if (u[i] >= 0) { acc = acc + cu[i] * pu[i]; if (WITH_SUM) sA = sA + cu[i]; }
    }
    return acc;
}

// MODE 0: plain; MODE 1: also accumulates sum(Apsi*psi) into sc->wApA (PCG); skips everything when sc->stop.
// UNI3: the mesh has uniform slice widths <= 3 (every hex mesh).
// Persistent: the grid is a fixed number of CTAs per SM and every CTA walks chunks of BLOCK rows with a grid stride, so there is
// no barrier, fence or atomic per chunk (a CTA that ends after each chunk idles its warps for the ~2 us of the reduction tail) and
// the warps of an SM drift apart, which keeps loads in flight all the time.  One partial per CTA, summed in a fixed order.
template <int MODE, bool UNI3>
static __global__ void __launch_bounds__(BLOCK) k_amul(MeshView mv, const double* __restrict__ diag,
                                                 const double* __restrict__ upS, const double* __restrict__ loS,
                                                 const double* __restrict__ psi, double* __restrict__ Apsi,
                                                 PcgScalars* sc, double* partials, unsigned* counter) {
    pdlWait();
    pdlTrigger();
    if (MODE == 1 && sc->stop) return;
    double dot = 0.0;
    const int stride = gridDim.x * blockDim.x;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (UNI3) {
        // uniform widths <= 3: the NEXT row's face indices are fetched while this row's coefficients and psi values are in flight,
        // so a row costs one exposed memory round trip (software pipeline over the persistent loop)
        const int nbw = mv.uniNw, ow = mv.uniOw, kmask = (1 << mv.KB) - 1;
        int pk[3], u[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            pk[i] = (row < mv.N && i < nbw) ? mv.loPk[(row >> 5) * 32 * nbw + (i << 5) + (row & 31)] : -1;
            u[i] = (row < mv.N && i < ow) ? mv.upIdx[(row >> 5) * 32 * ow + (i << 5) + (row & 31)] : -1;
        }
        for (; row < mv.N; row += stride) {
            const int so0 = (row >> 5) * 32 * ow, lane = row & 31, nrow = row + stride;
            const double p = ldgOrdered(psi + row), dg = ldgOrdered(diag + row);
            double cu[3], pu[3], cl[3], pl[3];
#pragma unroll
            for (int i = 0; i < 3; i++) cu[i] = i < ow ? ldgOrdered(upS + so0 + (i << 5) + lane) : 0.0;
#pragma unroll
            for (int i = 0; i < 3; i++) {      // padding entries load the row's own (valid, cached) data: unconditional loads batch
                const int l = pk[i] >= 0 ? pk[i] >> mv.KB : row, k = pk[i] >= 0 ? (pk[i] & kmask) : 0;
                cl[i] = ldgOrdered(loS + (((l >> 5) * ow + k) << 5) + (l & 31)); pl[i] = ldgOrdered(psi + l);
                pu[i] = ldgOrdered(psi + (u[i] >= 0 ? u[i] : row));
            }
            int npk[3], nu[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                npk[i] = (nrow < mv.N && i < nbw) ? ldgOrdered(mv.loPk + (nrow >> 5) * 32 * nbw + (i << 5) + (nrow & 31)) : -1;
                nu[i] = (nrow < mv.N && i < ow) ? ldgOrdered(mv.upIdx + (nrow >> 5) * 32 * ow + (i << 5) + (nrow & 31)) : -1;
            }
            double acc = dg * p;
#pragma unroll
            for (int i = 0; i < 3; i++) if (pk[i] >= 0) acc = acc + cl[i] * pl[i];
#pragma unroll
            for (int i = 0; i < 3; i++) if (u[i] >= 0) acc = acc + cu[i] * pu[i];
            Apsi[row] = acc;
            if (MODE == 1) dot = dot + acc * p;
#pragma unroll
            for (int i = 0; i < 3; i++) { pk[i] = npk[i]; u[i] = nu[i]; }
        }
    } else {
        for (; row < mv.N; row += stride) {
            const double p = psi[row];
            double unused = 0.0;
            const double acc = rowOffDiag<false>(mv, upS, loS, psi, row, diag[row] * p, unused);
            Apsi[row] = acc;
            if (MODE == 1) dot = dot + acc * p;
        }
    }
    if (MODE == 1) {
        double v[1] = {dot};
        gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
        gridFinalize<1, OpSum>(gridDim.x, partials, counter, &sc->wApA);
    }
}
inline bool uniform3(const MeshView& mv) { return mv.uniOw > 0 && mv.uniOw <= 3 && mv.uniNw <= 3; }
// CTAs of the persistent row kernels: as many as are resident at once (k_amul: 48 registers -> 5 CTAs of 256 threads per SM)
// CTAs of a persistent row kernel: exactly as many as are resident at once (a PDL successor only starts when every CTA of the
// grid has started); option amul_ctas_per_sm overrides the occupancy query
template <class K>
inline int persistentGrid(int nRows, K kernel) {
    int perSM = (int)rt().opt("amul_ctas_per_sm", 0);
#ifndef B200_EMU
    if (perSM <= 0) B200_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, BLOCK, 0));
#endif
    if (perSM <= 0) perSM = 4;
    return std::max(1, std::min(gridFor(nRows), rt().numSMs * perSM));
}
#define B200_LAUNCH_AMUL(MODE, pdl, mv, ...)                                                                                          \
    do {                                                                                                                              \
        if (uniform3(mv)) B200_LAUNCH_PDL(pdl, (k_amul<MODE, true>), persistentGrid((mv).N, k_amul<MODE, true>), BLOCK, 0, rt().stream, mv, __VA_ARGS__);    \
        else B200_LAUNCH_PDL(pdl, (k_amul<MODE, false>), persistentGrid((mv).N, k_amul<MODE, false>), BLOCK, 0, rt().stream, mv, __VA_ARGS__);               \
    } while (0)

// [UPSTREAM lduMatrixATmul.C sumA]
static __global__ void k_sumA(MeshView mv, const double* __restrict__ diag, const double* __restrict__ upS,
                       const double* __restrict__ loS, const double* __restrict__ bou, double* __restrict__ out) {
    int row = blockIdx.x * blockDim.x + threadIdx.x, s = row >> 5, lane = row & 31;
    if (row >= mv.N) return;
    double sA = diag[row];
    int nb0 = mv.nbrOff[s], nbw = (mv.nbrOff[s + 1] - nb0) >> 5;
    for (int j = 0; j < nbw; j++) {
        int pk = mv.loPk[nb0 + (j << 5) + lane];
        if (pk >= 0) { int l; int pos = facePosOf(mv, pk, l); sA = sA + loS[pos]; }
    }
    int so0 = mv.sliceOff[s], ow = (mv.sliceOff[s + 1] - so0) >> 5;
    for (int k = 0; k < ow; k++) { int pos = so0 + (k << 5) + lane; if (mv.upIdx[pos] >= 0) sA = sA + upS[pos]; }
    if (bou && ((mv.bMask[s] >> lane) & 1u)) {
        int i = mv.bBase[s] + __popc(mv.bMask[s] & ((1u << lane) - 1u));
        for (int q = mv.bRowStart[i]; q < mv.bRowStart[i + 1]; q++) { int bf = mv.bRowFaces[q]; if (mv.bCoupled[bf]) sA = sA - bou[bf]; }
    }
    out[row] = sA;
}

// processor-interface update  Apsi[faceCells] -= bouCoeffs*psiNbr, per boundary row in patch order
static __global__ void k_interface_update(MeshView mv, const double* __restrict__ bou, const double* __restrict__ psiNbr,
                                   double* __restrict__ Apsi, int nBRows, const int* __restrict__ bRowOfSlot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nBRows) return;
    int row = bRowOfSlot[i];
    double acc = Apsi[row];
    bool any = false;
    for (int q = mv.bRowStart[i]; q < mv.bRowStart[i + 1]; q++) {
        int b = mv.bRowFaces[q];
        if (mv.bCoupled[b]) { acc = acc - bou[b] * psiNbr[b]; any = true; }
    }
    if (any) Apsi[row] = acc;
}

// ----------------------------------------------------------------------------- PCG start-up
// wA = A x ; rA = b - wA ; sumA -> pA (scratch, as upstream normFactor) ; partial sum(x), sum|rA|
static __global__ void __launch_bounds__(BLOCK) k_pcg_init1(MeshView mv, const double* __restrict__ diag,
                                                      const double* __restrict__ upS, const double* __restrict__ x,
                                                      const double* __restrict__ b, const double* __restrict__ bou,
                                                      double* __restrict__ wA, double* __restrict__ rA,
                                                      double* __restrict__ pA, PcgScalars* sc, double* partials,
                                                      unsigned* counter) {
    int row = blockIdx.x * blockDim.x + threadIdx.x, s = row >> 5, lane = row & 31;
    double v[2] = {0.0, 0.0};
    if (row < mv.N) {
        double xr = x[row], d = diag[row];
        double sA = d;
        double acc = rowOffDiag<true>(mv, upS, upS, x, row, d * xr, sA);
        if (bou && ((mv.bMask[s] >> lane) & 1u)) {  // sumA: coupled patches subtract bouCoeffs
            int i = mv.bBase[s] + __popc(mv.bMask[s] & ((1u << lane) - 1u));
            for (int q = mv.bRowStart[i]; q < mv.bRowStart[i + 1]; q++) { int bf = mv.bRowFaces[q]; if (mv.bCoupled[bf]) sA = sA - bou[bf]; }
        }
        wA[row] = acc; pA[row] = sA;
        double r = b[row] - acc;
        rA[row] = r;
        v[0] = xr; v[1] = fabs(r);
    }
    gridReduce<2, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<2, OpSum>(gridDim.x, partials, counter, &sc->sumX, &sc->sumMagR);
}
// after the interface update of wA (multi-rank) rA must be recomputed; single-rank path skips this
static __global__ void k_pcg_resid(int N, const double* __restrict__ b, const double* __restrict__ wA, double* __restrict__ rA,
                            PcgScalars* sc, double* partials, unsigned* counter) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (row < N) { double r = b[row] - wA[row]; rA[row] = r; v[0] = fabs(r); }
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpSum>(gridDim.x, partials, counter, &sc->sumMagR);
}
// nfSum = sum(|wA - xRef*sumA| + |b - xRef*sumA|)
static __global__ void __launch_bounds__(BLOCK) k_pcg_init2(int N, const double* __restrict__ wA, const double* __restrict__ b,
                                                      const double* __restrict__ sumA, PcgScalars* sc, double* partials,
                                                      unsigned* counter) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double xRef = sc->sumX / sc->nGlobal;
    double v[1] = {0.0};
    if (row < N) { double t = sumA[row] * xRef; v[0] = fabs(wA[row] - t) + fabs(b[row] - t); }
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpSum>(gridDim.x, partials, counter, &sc->nfSum);
}
__device__ __forceinline__ bool pcgConverged(const PcgScalars* sc) {
    return sc->finalResidual < sc->tol || (sc->relTol > 0 && sc->finalResidual < sc->relTol * sc->initialResidual);
}
static __global__ void k_pcg_check0(PcgScalars* sc, double* hist) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    sc->normFactor = sc->nfSum + 1.0e-20;
    sc->initialResidual = sc->finalResidual = sc->sumMagR / sc->normFactor;
    sc->nIter = 0; sc->singular = 0;
    sc->converged = pcgConverged(sc) ? 1 : 0;
    sc->stop = (0 >= sc->minIter) && (0 >= sc->maxIter || sc->converged);
    sc->wArA = 1.0e20; sc->wArAold = 1.0e20;
    if (hist && sc->histCap > 0) hist[0] = sc->initialResidual;
}
// end of an iteration (x, rA updated, sum|rA| known): residual, convergence, iteration count
__device__ __forceinline__ void pcgEndOfIteration(PcgScalars* sc, double* hist) {
    sc->finalResidual = sc->sumMagR / sc->normFactor;
    sc->nIter = sc->nIter + 1;
    sc->wArAold = sc->wArA;
    if (hist && sc->nIter < sc->histCap) hist[sc->nIter] = sc->finalResidual;
    sc->converged = pcgConverged(sc) ? 1 : 0;
    sc->stop = (sc->nIter >= sc->minIter) && (sc->nIter >= sc->maxIter || sc->converged);
}
// multi-rank: runs after the ONE all-reduce of (sum|rA| left behind by the forward sweep, sum(wA*rA) of the backward sweep):
// closes the previous iteration (residual, convergence, wArAold = wArA), then adopts the new wArA
static __global__ void k_pcg_check(PcgScalars* sc, double* hist) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (sc->stop) return;
    if (sc->pendingCheck) { sc->pendingCheck = 0; pcgEndOfIteration(sc, hist); }
    if (!sc->stop) sc->wArA = sc->wArAnew;
}
// multi-rank tail of Amul: processor-interface update  Apsi[faceCells] -= bouCoeffs*psiNbr  over the rows that own boundary faces
// only (per row in patch order), and the matching correction of sum(Apsi*psi): k_amul<1> has left the sum over the un-updated Apsi
// in sc->wApA, this kernel adds sum((Apsi_new - Apsi_old)*psi) over the updated rows.  hw: peer-memory halo -- wait for the
// neighbours' values of this epoch first (no-op on the NCCL path).
static __global__ void __launch_bounds__(BLOCK) k_iface_rows(MeshView mv, const double* __restrict__ bou, const double* psiNbr,
                                                       double* __restrict__ Apsi, const double* __restrict__ psi, int nBRows,
                                                       const int* __restrict__ bRowOfSlot, PcgScalars* sc, double* partials,
                                                       unsigned* counter, PeerHaloWait hw, P2PFold fold) {
    // (programmatic dependent launch: the neighbours' flags do not depend on the local Amul, so the wait for them is the prologue; the next
    //  forward sweep may start its own prologue as soon as every CTA is here)
    peerHaloWaitDevice(hw);
    pdlWait();
    pdlTrigger();
    if (sc->stop) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (i < nBRows) {
        const int row = bRowOfSlot[i];
        const double old = Apsi[row];
        double acc = old;
        bool any = false;
        for (int q = mv.bRowStart[i]; q < mv.bRowStart[i + 1]; q++) {
            const int b = mv.bRowFaces[q];
            if (mv.bCoupled[b]) { acc = acc - bou[b] * ldNoL1(psiNbr + b); any = true; }
        }
        if (any) { Apsi[row] = acc; v[0] = (acc - old) * psi[row]; }
    }
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    __shared__ double corr;
    const bool last = gridFinalizeLast<OpSum>(gridDim.x, partials, counter, &corr);
    if (last && threadIdx.x == 0) sc->wApA = sc->wApA + corr;
    // folded gSumProd(wA, pA): the last CTA's first warp exchanges the rank's sum over NVLink peer memory right here (no kernel of its own)
    if (last && fold.on) {
        __syncthreads();
        if (threadIdx.x < 32) p2pAllreduceWarp(&sc->wApA, 1, 0, fold.v, fold.epoch);
    }
}
// pA = z + beta*pA (first iteration: pA = z);  z <- sentinel (re-armed for the next sweep)
// (two rows per thread through 16-byte accesses: the kernel is a 24 B/cell stream bound by requests in flight, not by arithmetic)
static __global__ void __launch_bounds__(BLOCK) k_pcg_p_update(int N, double* __restrict__ z, double* __restrict__ pA,
                                                         const PcgScalars* sc) {
    pdlWait();
    pdlTrigger();
    if (sc->stop) return;
    const int row = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (row >= N) return;
    const bool first = sc->nIter == 0;
    const double beta = first ? 0.0 : sc->wArA / sc->wArAold;
    if (row + 1 < N) {
        const double2 zv = *reinterpret_cast<const double2*>(z + row);
        double2 pv = *reinterpret_cast<const double2*>(pA + row);
        pv.x = first ? zv.x : zv.x + beta * pv.x;
        pv.y = first ? zv.y : zv.y + beta * pv.y;
        *reinterpret_cast<double2*>(pA + row) = pv;
        *reinterpret_cast<double2*>(z + row) = make_double2(sentinel(), sentinel());
    } else {
        const double zv = z[row];
        pA[row] = first ? zv : zv + beta * pA[row];
        z[row] = sentinel();
    }
}
// (x += alpha*pA ; rA -= alpha*wA ; sum|rA| of upstream's loop tail are fused into the NEXT forward sweep: dic_kernels.cuh)

}  // namespace b200

#include "dic_kernels.cuh"
