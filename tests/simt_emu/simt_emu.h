// simt_emu.h -- TEST INFRASTRUCTURE ONLY.  A single-threaded SIMT emulator (one ucontext fiber per CUDA
// thread, blocks run one after another) that lets g++ compile the SAME kernel sources as nvcc so that
// kernel logic can be checked in this GPU-less container (compute-sanitizer-style dev tool).
// The product never uses it: shallowinterfoam_b200 loads libb200foam.so (nvcc, sm_100a) only and fails loudly
// without a CUDA device; the emulated build is tests/simt_emu/libb200foam_emu.so and is loaded by tests alone.
#pragma once
#ifndef B200_EMU
#error "simt_emu.h is for the B200_EMU test build only"
#endif
#include <time.h>
#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)

struct uint3 { unsigned x, y, z; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
struct alignas(16) double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace emu {
struct Fiber { ucontext_t ctx; uint3 tid; bool done; };
struct WarpState { int count = 0, gen = 0; unsigned long long buf[32]; unsigned ballot = 0, ballotOut = 0; };
struct State {
    std::vector<Fiber> fibers; std::vector<char*> stacks; std::vector<WarpState> warps;
    ucontext_t sched; Fiber* cur = nullptr; int live = 0, barCount = 0, barGen = 0;
    uint3 blockIdx; dim3 blockDim, gridDim; std::function<void()> body;
};
inline State& S() { static State s; return s; }
inline void yield() { State& s = S(); swapcontext(&s.cur->ctx, &s.sched); }
inline void trampoline() {
    State& s = S();
    s.body();
    s.cur->done = true; s.live--;
    if (s.live > 0 && s.barCount == s.live) { s.barCount = 0; s.barGen++; }  // exited threads release the barrier
    swapcontext(&s.cur->ctx, &s.sched);
}
inline std::vector<unsigned char>& dynSmemBuf() { static std::vector<unsigned char> b; return b; }
inline unsigned char* dynSmem() { return dynSmemBuf().data(); }
inline void launch(dim3 grid, dim3 block, std::function<void()> body, size_t smemBytes = 0) {
    State& s = S();
    if (dynSmemBuf().size() < smemBytes + 16) dynSmemBuf().resize(smemBytes + 16);
    const size_t STK = 256 * 1024;
    unsigned nt = block.x;
    while (s.stacks.size() < nt) s.stacks.push_back((char*)malloc(STK));
    s.gridDim = grid; s.blockDim = block; s.body = body;
    for (unsigned b = 0; b < grid.x; b++) {
        s.blockIdx = {b, 0, 0};
        s.fibers.assign(nt, Fiber());
        s.warps.assign((nt + 31) / 32, WarpState());
        s.live = nt; s.barCount = 0;
        for (unsigned t = 0; t < nt; t++) {
            Fiber& f = s.fibers[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = s.stacks[t]; f.ctx.uc_stack.ss_size = STK; f.ctx.uc_link = nullptr;
            f.tid = {t, 0, 0}; f.done = false;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        while (s.live > 0)
            for (unsigned t = 0; t < nt; t++)
                if (!s.fibers[t].done) { s.cur = &s.fibers[t]; swapcontext(&s.sched, &s.fibers[t].ctx); }
    }
    s.cur = nullptr;
}
inline void syncthreads() {
    State& s = S(); int gen = s.barGen;
    if (++s.barCount == s.live) { s.barCount = 0; s.barGen++; }
    else while (s.barGen == gen) yield();
}
inline WarpState& warp() { State& s = S(); return s.warps[s.cur->tid.x / 32]; }
inline void warpBarrier() {
    WarpState& w = warp(); int gen = w.gen;
    if (++w.count == 32) { w.count = 0; w.gen++; }
    else while (w.gen == gen) yield();
}
template <class T> inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shfl size");
    WarpState& w = warp(); int lane = S().cur->tid.x & 31;
    unsigned long long bits = 0; memcpy(&bits, &v, sizeof(T)); w.buf[lane] = bits;
    warpBarrier();
    unsigned long long r = w.buf[src & 31];
    warpBarrier();
    T out; memcpy(&out, &r, sizeof(T)); return out;
}
inline unsigned ballot(int pred) {
    WarpState& w = warp(); int lane = S().cur->tid.x & 31;
    if (lane == 0) w.ballot = 0;   // safe: previous collective ended with a barrier
    warpBarrier();
    if (pred) w.ballot |= (1u << lane);
    warpBarrier();
    unsigned r = w.ballot;
    warpBarrier();
    return r;
}
}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)

inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warpBarrier(); }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { emu::yield(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl(v, src); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { int l = emu::S().cur->tid.x & 31; return emu::shfl(v, l + d < 32 ? l + d : l); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { int l = emu::S().cur->tid.x & 31; return emu::shfl(v, l ^ m); }
inline unsigned __ballot_sync(unsigned, int p) { return emu::ballot(p); }
inline int __all_sync(unsigned, int p) { return emu::ballot(p) == 0xffffffffu; }
inline int __any_sync(unsigned, int p) { return emu::ballot(p) != 0u; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }
template <class T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }

// ---- minimal CUDA runtime shim --------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent { double t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = nullptr; return 0; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline double emuNow() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent{0}; return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emuNow(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }
