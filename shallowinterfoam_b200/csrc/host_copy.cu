// host_copy.cu -- bulk field transfers between CALLER-OWNED host arrays and the device, for the plug points
// (b200_pcg_solve, b200_mules_explicit_solve ...: foam-extend hands over plain `Field` storage, i.e. pageable memory).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one pinned buffer by ONE thread: ~10 GB/s, a third of what the
// link does -- and at 985k cells a pd solve moves 55 MB in and 8 MB out around 12 ms of device work (`e2e_plugin` in bench.py).
// Here a few worker threads copy 2 MB chunks into their own pinned buffers (two each) and issue the DMA of every chunk on their own
// stream as soon as it is staged, so the host-side copy of one chunk overlaps the DMA of the others.  Pinned / registered memory and small
// transfers go straight to cudaMemcpyAsync on the library stream.  Ordering: the worker streams wait for an event recorded on the
// library stream (the destination may still be read by earlier kernels), and the library stream waits for every worker's last event.
// hostCopyIn returns when the SOURCE has been consumed (the caller may reuse it); hostCopyOut returns when the data is in `host`.
#include "common.cuh"
#ifndef B200_EMU
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#endif

namespace b200 {

#ifdef B200_EMU
void hostCopyIn(void* dev, const void* host, size_t bytes) { if (bytes) B200_CHECK_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, rt().stream)); }
void hostCopyOut(void* host, const void* dev, size_t bytes) { if (bytes) B200_CHECK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, rt().stream)); B200_CHECK_CUDA(cudaStreamSynchronize(rt().stream)); }
#else
namespace {
constexpr size_t CHUNK = 2u << 20;
constexpr int MAX_LANES = 16;

struct Lane {
    cudaStream_t s = nullptr;
    char* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;
    cudaError_t err = cudaSuccess;
};

// a fixed set of workers that run one job (a function of the worker index) at a time
class Pool {
public:
    ~Pool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (std::thread& t : th_) if (t.joinable()) t.join();
    }
    void run(int n, const std::function<void(int)>& job) {
        while ((int)th_.size() < n) { int id = (int)th_.size(); th_.emplace_back([this, id] { loop(id); }); }
        { std::lock_guard<std::mutex> g(mu_); job_ = &job; active_ = n; pending_ = n; gen_++; }
        cv_.notify_all();
        std::unique_lock<std::mutex> l(mu_);
        doneCv_.wait(l, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
private:
    void loop(int id) {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return stop_ || (gen_ != seen && id < active_); });
                if (stop_) return;
                seen = gen_; job = job_;
            }
            (*job)(id);
            { std::lock_guard<std::mutex> g(mu_); pending_--; }
            doneCv_.notify_one();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_;
    std::condition_variable cv_, doneCv_;
    const std::function<void(int)>* job_ = nullptr;
    int active_ = 0, pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
};

struct Pipe {
    std::vector<Lane> lanes;
    cudaEvent_t fence = nullptr;
    Pool pool;
    void ensure(int n) {
        if (!fence) B200_CHECK_CUDA(cudaEventCreateWithFlags(&fence, cudaEventDisableTiming));
        while ((int)lanes.size() < n) {
            Lane L;
            B200_CHECK_CUDA(cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
            for (int k = 0; k < 2; k++) {
                B200_CHECK_CUDA(cudaHostAlloc((void**)&L.buf[k], CHUNK, cudaHostAllocDefault));
                B200_CHECK_CUDA(cudaEventCreateWithFlags(&L.ev[k], cudaEventDisableTiming));
            }
            B200_CHECK_CUDA(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
            lanes.push_back(L);
        }
    }
};
Pipe& pipe() { static Pipe* p = new Pipe(); return *p; }   // (leaked on purpose: the CUDA runtime may be gone when statics are destroyed)

bool isDeviceVisibleHost(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
int lanesFor(size_t bytes) {
    int n = (int)rt().opt("host_copy_threads", 4);
    if (n <= 1 || bytes < 2 * CHUNK) return 0;
    return std::min(std::min(n, MAX_LANES), (int)((bytes + CHUNK - 1) / CHUNK));
}
#define LANE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { L.err = e_; return; } } while (0)
}  // namespace

void hostCopyIn(void* dev, const void* host, size_t bytes) {
    if (!bytes) return;
    const int T = isDeviceVisibleHost(host) ? 0 : lanesFor(bytes);
    if (T == 0) { B200_CHECK_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, rt().stream)); return; }
    Pipe& P = pipe();
    P.ensure(T);
    B200_CHECK_CUDA(cudaEventRecord(P.fence, rt().stream));
    const size_t nChunks = (bytes + CHUNK - 1) / CHUNK;
    const int device = rt().device;
    P.pool.run(T, [&](int t) {
        Lane& L = P.lanes[t];
        L.err = cudaSuccess;
        LANE_TRY(cudaSetDevice(device));
        LANE_TRY(cudaStreamWaitEvent(L.s, P.fence, 0));
        int k = 0;
        for (size_t c = t; c < nChunks; c += T, k ^= 1) {
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            LANE_TRY(cudaEventSynchronize(L.ev[k]));        // the DMA that last read this staging buffer (an earlier chunk or call)
            memcpy(L.buf[k], (const char*)host + off, len);
            LANE_TRY(cudaMemcpyAsync((char*)dev + off, L.buf[k], len, cudaMemcpyHostToDevice, L.s));
            LANE_TRY(cudaEventRecord(L.ev[k], L.s));
        }
        LANE_TRY(cudaEventRecord(L.done, L.s));
    });
    for (int t = 0; t < T; t++) {
        B200_CHECK_CUDA(P.lanes[t].err);
        B200_CHECK_CUDA(cudaStreamWaitEvent(rt().stream, P.lanes[t].done, 0));
    }
}

void hostCopyOut(void* host, const void* dev, size_t bytes) {
    if (!bytes) return;
    const int T = isDeviceVisibleHost(host) ? 0 : lanesFor(bytes);
    if (T == 0) {
        B200_CHECK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, rt().stream));
        B200_CHECK_CUDA(cudaStreamSynchronize(rt().stream));
        return;
    }
    Pipe& P = pipe();
    P.ensure(T);
    B200_CHECK_CUDA(cudaEventRecord(P.fence, rt().stream));
    const size_t nChunks = (bytes + CHUNK - 1) / CHUNK;
    const int device = rt().device;
    P.pool.run(T, [&](int t) {
        Lane& L = P.lanes[t];
        L.err = cudaSuccess;
        LANE_TRY(cudaSetDevice(device));
        LANE_TRY(cudaStreamWaitEvent(L.s, P.fence, 0));
        int k = 0;
        size_t prevOff = 0, prevLen = 0;
        for (size_t c = t; c < nChunks; c += T, k ^= 1) {
            const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
            LANE_TRY(cudaMemcpyAsync(L.buf[k], (const char*)dev + off, len, cudaMemcpyDeviceToHost, L.s));
            LANE_TRY(cudaEventRecord(L.ev[k], L.s));
            if (prevLen) {      // hand the previous chunk over while this one is on the link
                LANE_TRY(cudaEventSynchronize(L.ev[k ^ 1]));
                memcpy((char*)host + prevOff, L.buf[k ^ 1], prevLen);
            }
            prevOff = off; prevLen = len;
        }
        if (prevLen) {
            LANE_TRY(cudaEventSynchronize(L.ev[k ^ 1]));
            memcpy((char*)host + prevOff, L.buf[k ^ 1], prevLen);
        }
    });
    for (int t = 0; t < T; t++) B200_CHECK_CUDA(P.lanes[t].err);
}
#endif

}  // namespace b200
