// runtime.cu -- process-level state: device/stream, launch counter, kernel-class timers.
#include "common.cuh"
#include "comm.cuh"

namespace b200 {

Runtime& rt() { static Runtime r; return r; }

void requireDevice() {
    Runtime& r = rt();
    if (r.initialised) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        throw Error(B200_ERR_NO_DEVICE, "libb200foam: no CUDA device available (there is no CPU fallback)");
    B200_CHECK_CUDA(cudaSetDevice(r.device));
#ifndef B200_EMU
    cudaDeviceProp prop;
    B200_CHECK_CUDA(cudaGetDeviceProperties(&prop, r.device));
    r.numSMs = prop.multiProcessorCount;
#endif
    B200_CHECK_CUDA(cudaStreamCreate(&r.stream));
    r.initialised = true;
}

static const char* kTimerNames[T_COUNT] = {"amul", "dic_fwd", "dic_bwd", "dic_calc_rD", "pcg_vec", "pcg_solve",
                                           "mules", "flux", "assembly", "ueqn", "misc", "step", "allreduce", "halo"};
const char* timerName(int i) { return (i >= 0 && i < T_COUNT) ? kTimerNames[i] : ""; }
Timers& timers() { static Timers t; return t; }
cudaEvent_t Timers::get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; B200_CHECK_CUDA(cudaEventCreate(&e)); return e;
}
void Timers::collect() {
    for (auto& p : pending) {
        cudaEventSynchronize(p.b);
        float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
        this->ms[p.id] += ms; calls[p.id]++;
        pool.push_back(p.a); pool.push_back(p.b);
    }
    pending.clear();
}
ScopedTimer::ScopedTimer(int id_) : id(id_) {
    Timers& t = timers();
    if (!t.enabled) return;
    a = t.get(); b = t.get();
    cudaEventRecord(a, rt().stream);
}
ScopedTimer::~ScopedTimer() {
    if (!a) return;
    cudaEventRecord(b, rt().stream);
    timers().pending.push_back({id, a, b});
}

}  // namespace b200
