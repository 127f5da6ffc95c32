// common.cuh -- shared plumbing of libb200foam: error handling, device buffers, launch macro, timers.
#pragma once
#ifdef B200_EMU
#include "simt_emu.h"   // tests/simt_emu: test-only SIMT emulator (never part of the product build)
#else
#include <cuda_runtime.h>
#endif
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "b200foam.h"

namespace b200 {

// ---- errors: no C++ exception crosses the C-ABI (api.cu catches and stores the message) ------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define B200_CHECK_CUDA(expr)                                                                            \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw b200::Error(B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                                                 __FILE__ + ":" + std::to_string(__LINE__));             \
    } while (0)
#define B200_REQUIRE(cond, code, msg)                                   \
    do {                                                                \
        if (!(cond)) throw b200::Error(code, std::string(msg));        \
    } while (0)

// ---- runtime singleton -----------------------------------------------------------------------------------
struct Runtime {
    bool initialised = false;
    int device = 0, rank = 0, nRanks = 1;
    cudaStream_t stream = nullptr;
    int numSMs = 148;
    long long launches = 0;
    void* nccl = nullptr;  // ncclComm_t (comm.cu)
    std::map<std::string, double> opts;  // tuning knobs (b200_set_option)
    double opt(const char* k, double dflt) const { auto it = opts.find(k); return it == opts.end() ? dflt : it->second; }
};
Runtime& rt();
void requireDevice();  // throws B200_ERR_NO_DEVICE when no CUDA device is usable

// bulk transfers between caller-owned (possibly pageable) host arrays and the device (host_copy.cu): chunked through pinned staging
// buffers by a few worker threads; hostCopyIn returns when the source is consumed, hostCopyOut when the data has arrived
void hostCopyIn(void* dev, const void* host, size_t bytes);
void hostCopyOut(void* host, const void* dev, size_t bytes);

// ---- kernel-class timers (CUDA events on the launching stream) -------------------------------------------
enum TimerId { T_AMUL = 0, T_DIC_FWD, T_DIC_BWD, T_DIC_RD, T_PCG_VEC, T_PCG_SOLVE, T_MULES, T_FLUX, T_ASSEMBLY, T_UEQN, T_MISC, T_STEP, T_ALLREDUCE, T_HALO, T_COUNT };
const char* timerName(int i);
struct Timers {
    bool enabled = false;
    double ms[T_COUNT] = {0};
    long long calls[T_COUNT] = {0};
    struct Pending { int id; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get();
    void collect();
};
Timers& timers();
struct ScopedTimer {  // records events around a launch sequence when timers are enabled
    int id; cudaEvent_t a = nullptr, b = nullptr;
    explicit ScopedTimer(int id_);
    ~ScopedTimer();
};

// ---- launch macro: the ONLY way kernels are launched (counts launches; works under the emulator) ----------
#ifdef B200_EMU
#define B200_LAUNCH(kernel, grid, block, stream, ...)                         \
    do {                                                                      \
        b200::rt().launches++;                                                \
        emu::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); }); \
    } while (0)
#define B200_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                        \
    do {                                                                                \
        b200::rt().launches++;                                                          \
        emu::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); }, (smem));   \
    } while (0)
#else
// dynamic shared memory above 48 KB needs the opt-in attribute (set on every launch: a cheap host-side call)
#define B200_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                                         \
    do {                                                                                                                 \
        b200::rt().launches++;                                                                                           \
        B200_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                                      \
        B200_CHECK_CUDA(cudaGetLastError());                                                                             \
    } while (0)
#define B200_LAUNCH(kernel, grid, block, stream, ...)          \
    do {                                                       \
        b200::rt().launches++;                                 \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); \
        B200_CHECK_CUDA(cudaGetLastError());                   \
    } while (0)
#endif

// ---- programmatic dependent launch (PDL): a kernel launched with pdl = true may start while its predecessor in the stream is
// still running; everything it does before pdlWait() must be independent of that predecessor.  Every kernel of the PCG iteration
// calls pdlWait() and then pdlTrigger(), so when a kernel's prologue runs, the predecessor of its predecessor is complete.
#ifdef B200_EMU
#define B200_LAUNCH_PDL(pdl, kernel, grid, block, smem, stream, ...)                    \
    do {                                                                                \
        b200::rt().launches++;                                                          \
        emu::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); }, (smem));   \
    } while (0)
__device__ __forceinline__ void pdlWait() {}
__device__ __forceinline__ void pdlTrigger() {}
#else
#define B200_LAUNCH_PDL(pdl_, kernel_, grid_, block_, smem_, stream_, ...)                                                    \
    do {                                                                                                                      \
        b200::rt().launches++;                                                                                                \
        if ((smem_) > 48 * 1024)                                                                                              \
            B200_CHECK_CUDA(cudaFuncSetAttribute(kernel_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_)));       \
        cudaLaunchConfig_t _cfg; memset(&_cfg, 0, sizeof(_cfg));                                                              \
        _cfg.gridDim = dim3(grid_); _cfg.blockDim = dim3(block_); _cfg.dynamicSmemBytes = (smem_); _cfg.stream = (stream_);   \
        cudaLaunchAttribute _attr[1];                                                                                         \
        _attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                     \
        _attr[0].val.programmaticStreamSerializationAllowed = (pdl_) ? 1 : 0;                                                 \
        _cfg.attrs = _attr; _cfg.numAttrs = 1;                                                                                \
        B200_CHECK_CUDA(cudaLaunchKernelEx(&_cfg, kernel_, __VA_ARGS__));                                                     \
    } while (0)
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdlTrigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---- device buffer ---------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t n_) {
        release();
        n = n_;
        B200_CHECK_CUDA(cudaMalloc((void**)&p, sizeof(T) * (n ? n : 1)));
    }
    void ensure(size_t n_) { if (n < n_) alloc(n_); }
    void zero() { B200_CHECK_CUDA(cudaMemsetAsync(p, 0, sizeof(T) * n, rt().stream)); }
    void upload(const T* h, size_t cnt) { B200_CHECK_CUDA(cudaMemcpyAsync(p, h, sizeof(T) * cnt, cudaMemcpyHostToDevice, rt().stream)); }
    void upload(const std::vector<T>& h) { alloc(h.size()); if (!h.empty()) upload(h.data(), h.size()); }
    void download(T* h, size_t cnt) const { B200_CHECK_CUDA(cudaMemcpyAsync(h, p, sizeof(T) * cnt, cudaMemcpyDeviceToHost, rt().stream)); }
    operator T*() const { return p; }
};
inline void syncStream() { B200_CHECK_CUDA(cudaStreamSynchronize(rt().stream)); }

constexpr int BLOCK = 256;  // rows per CTA for the row/face kernels (8 slices)
inline int gridFor(long long n, int block = BLOCK) { return (int)((n + block - 1) / block); }

// "not ready" marker of the data-flow DIC sweeps: a quiet NaN whose 32-bit halves are equal, so one
// 32-bit memset pattern fills it.  Arithmetic never produces this payload.
constexpr unsigned long long SENTINEL_BITS = 0x7FF8B2007FF8B200ull;

}  // namespace b200
