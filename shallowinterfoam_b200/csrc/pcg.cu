// pcg.cu -- b200PCG driver: PCG::solve [UPSTREAM PCG.C] with the DIC preconditioner, every vector resident in HBM.
// The iteration is 5 kernels + a 1-thread convergence kernel; all scalars (alpha, beta, residuals, iteration count,
// stop flag) stay on the device.  The host enqueues iterations in batches and only polls the stop flag through a
// pinned mirror, so launch latency overlaps execution; iterations enqueued past convergence are no-ops.
#include "pcg.cuh"
#include <memory>
#include "comm.cuh"

namespace b200 {

PcgWorkspace::~PcgWorkspace() {
    peerHaloRelease(peer);
    if (hsc) cudaFreeHost(hsc);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (evP) cudaEventDestroy(evP);
    if (evPush) cudaEventDestroy(evPush);
    if (pushStream) cudaStreamDestroy(pushStream);
}
DeviceMesh::~DeviceMesh() { delete pcg; }

__global__ void k_brow_of_slot(MeshView mv, int* out) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= mv.N) return;
    int s = row >> 5, lane = row & 31;
    if ((mv.bMask[s] >> lane) & 1u) out[mv.bBase[s] + __popc(mv.bMask[s] & ((1u << lane) - 1u))] = row;
}
__global__ void k_dot(int N, const double* __restrict__ a, const double* __restrict__ b, double* out, double* partials,
                      unsigned* counter, const PcgScalars* sc) {
    if (sc && sc->stop) return;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {row < N ? a[row] * b[row] : 0.0};
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpSum>(gridDim.x, partials, counter, out);
}

PcgWorkspace& pcgWorkspace(DeviceMesh& m) {
    if (m.pcg) return *m.pcg;
    PcgWorkspace* w = new PcgWorkspace();
    m.pcg = w;
    size_t N = m.N;
    w->pA.alloc(N); w->wA.alloc(N); w->rA.alloc(N); w->y.alloc(N); w->z.alloc(N); w->rD.alloc(N); w->Dt.alloc(N);
    int nBlocks = gridFor(m.N) + 1;
    w->partials.alloc(3 * (size_t)std::max(nBlocks, m.nTiles) + 8);
    w->cf.alloc(m.fwdBlkSize); w->cb.alloc(m.bwdBlkSize);   // operand blocks: templates carry the (static) source offsets
    if (m.fwdBlkSize) B200_CHECK_CUDA(cudaMemcpyAsync(w->cf.p, m.dFwdBlk0.p, 8 * m.fwdBlkSize, cudaMemcpyDeviceToDevice, rt().stream));
    if (m.bwdBlkSize) B200_CHECK_CUDA(cudaMemcpyAsync(w->cb.p, m.dBwdBlk0.p, 8 * m.bwdBlkSize, cudaMemcpyDeviceToDevice, rt().stream));
    w->sc.alloc(1); w->counters.alloc(8); w->counters.zero();
    B200_CHECK_CUDA(cudaMallocHost((void**)&w->hsc, 2 * sizeof(PcgScalars)));
    for (auto& e : w->ev) B200_CHECK_CUDA(cudaEventCreate(&e));
    fillSentinel(w->y, N); fillSentinel(w->z, N);
    w->bRowOfSlot.alloc(m.nBRows + 1);
    if (m.N) B200_LAUNCH(k_brow_of_slot, gridFor(m.N), BLOCK, rt().stream, m.view(), w->bRowOfSlot.p);
    if (m.B) { w->sendBuf.alloc(m.B); w->recvBuf.alloc(m.B); }
    // persistent sweep grid: a few CTAs per SM; chunks are handed out by ticket so any size is deadlock-free
    w->sweepGrid = std::max(1, std::min(gridFor(m.N), rt().numSMs * 4));
    w->rowsPerLevel = m.nLevels > 0 ? (double)m.N / m.nLevels : 1.0;
    syncStream();
    return *w;
}

// separate ticket / done counters per direction: with PDL the backward sweep's CTAs take tickets while the forward sweep still runs
static SweepCtl sweepCtl(PcgWorkspace& w, bool bwd = false) { return bwd ? SweepCtl{w.counters.p + 3, w.counters.p + 4} : SweepCtl{w.counters.p + 1, w.counters.p + 2}; }
// CTAs of a tiled sweep: as many as fit (shared memory bound); tiles are handed out in ticket order, so any grid size is deadlock-free
static int sweepGridOf(const DeviceMesh& m, const PcgWorkspace&) {
    int forced = (int)rt().opt("sweep_grid", 0);
    int perSM = std::max(1, std::min(8, (int)((size_t)(227 * 1024) / (m.sweepSmem + 2048))));
    int cap = (int)rt().opt("sweep_ctas_per_sm", 0);
    if (cap > 0) perSM = std::min(perSM, cap);
    int g = forced > 0 ? forced : rt().numSMs * perSM;
    return std::max(1, std::min(g, std::max(1, m.nTiles)));
}
static int sweepThreads(const DeviceMesh& m) { return m.plan.lanes; }   // one row per thread per step
// gate (optional): device flag of the DIC reuse; when it says "unchanged" the kernels return at once and rD / cf / cb stay as they are
// loS == upS: DIC (rD[u] -= upper*upper/rD[l]); loS != upS: DILU (rD[u] -= upper*lower/rD[l]).  cf is used as scratch for the products.
static void launchCalcRD(DeviceMesh& m, PcgWorkspace& w, const double* diag, const double* upS, const double* loS, double* rD, double* cf,
                         const PcgScalars* gate = nullptr) {
    fillSentinel(w.Dt, m.N);
    // forward operand blocks <- face products (fully parallel), so that the sweep stages them with one TMA copy per tile
    B200_LAUNCH(k_dic_coeffs, m.nTiles, sweepThreads(m), rt().stream, m.sweep(), (const double*)nullptr, upS, loS, cf, (double*)nullptr, gate);
    if (m.plan.chunked)
        B200_LAUNCH_SMEM((k_dic_fwd<1, true>), sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), diag, (const double*)nullptr, upS,
                         (const double*)cf, w.Dt.p, rD, sweepCtl(w), gate, PcgFuse{});
    else
        B200_LAUNCH_SMEM((k_dic_fwd<1, false>), sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), diag, (const double*)nullptr, upS,
                         (const double*)cf, w.Dt.p, rD, sweepCtl(w), gate, PcgFuse{});
}
// forward operand blocks <- rD[row]*srcF, backward <- rD[row]*srcB (DIC: both upper; DILU: lower/upper, transposed: upper/lower)
static void launchCoeffs(DeviceMesh& m, PcgWorkspace& w, const double* rD, const double* srcF, const double* srcB, double* cf, double* cb,
                         const PcgScalars* gate = nullptr) {
    (void)w;
    B200_LAUNCH(k_dic_coeffs, 2 * m.nTiles, sweepThreads(m), rt().stream, m.sweep(), rD, srcF, srcB, cf, cb, gate);
}
// DIC reuse: closes the gate, then re-opens it if diag / upper differ from the kept copies (refreshed in the same pass)
static const PcgScalars* launchMatrixGate(DeviceMesh& m, PcgWorkspace& w, const double* diag, const double* upS) {
    if (rt().opt("dic_reuse", 1) <= 0) { w.keptValid = false; return nullptr; }
    if (!w.keptValid) {
        w.keptDiag.ensure(m.N); w.keptUp.ensure(m.Fs); if (!w.gate.p) w.gate.alloc(1);
        fillSentinel(w.keptDiag, m.N); fillSentinel(w.keptUp, m.Fs);   // differs from every coefficient
        w.keptValid = true;
    }
    B200_LAUNCH(k_gate_close, 1, 32, rt().stream, w.gate.p);
    const int n = m.N + m.Fs;
    B200_LAUNCH(k_matrix_changed, std::max(1, std::min(gridFor(n), rt().numSMs * 8)), BLOCK, rt().stream, m.N, m.Fs, diag, upS, w.keptDiag.p,
                w.keptUp.p, w.gate.p);
    return w.gate.p;
}
// pdl: the kernel may start (ticket, descriptor, operand-block staging) while its predecessor in the stream is still running
static void launchFwd(DeviceMesh& m, PcgWorkspace& w, const double* rD, const double* rA, const PcgScalars* sc, PcgFuse fz = PcgFuse{}, bool pdl = false,
                      const double* cf = nullptr) {
    if (m.plan.chunked)
        B200_LAUNCH_PDL(pdl, (k_dic_fwd<0, true>), sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), rD, rA, (const double*)nullptr,
                        cf ? cf : (const double*)w.cf.p, w.y.p, (double*)nullptr, sweepCtl(w), sc, fz);
    else
        B200_LAUNCH_PDL(pdl, (k_dic_fwd<0, false>), sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), rD, rA, (const double*)nullptr,
                        cf ? cf : (const double*)w.cf.p, w.y.p, (double*)nullptr, sweepCtl(w), sc, fz);
}
static void launchBwd(DeviceMesh& m, PcgWorkspace& w, const double* rA, PcgScalars* sc, double* part, unsigned* cnt, double* dotOut = nullptr, bool pdl = false,
                      const double* cb = nullptr, P2PFold fold = P2PFold{}, double* hist = nullptr) {
    if (m.plan.chunked)
        B200_LAUNCH_PDL(pdl, k_dic_bwd<true>, sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), cb ? cb : (const double*)w.cb.p, (const double*)rA, w.y.p, w.z.p,
                        sweepCtl(w, true), sc, part, cnt, dotOut ? dotOut : (sc ? &sc->wArA : nullptr), fold, hist);
    else
        B200_LAUNCH_PDL(pdl, k_dic_bwd<false>, sweepGridOf(m, w), sweepThreads(m), m.sweepSmem, rt().stream, m.sweep(), cb ? cb : (const double*)w.cb.p, (const double*)rA, w.y.p, w.z.p,
                        sweepCtl(w, true), sc, part, cnt, dotOut ? dotOut : (sc ? &sc->wArA : nullptr), fold, hist);
}

void amulDevice(DeviceMesh& m, const double* diag, const double* upS, const double* loS, const double* psi, double* Apsi,
                const double* bou) {
    PcgWorkspace& w = pcgWorkspace(m);
    ScopedTimer t(T_AMUL);
    bool halo = bou && rt().nRanks > 1 && meshHasProcessorPatches(m);
    if (halo) haloPack(m, psi, w.sendBuf);
    if (m.N) B200_LAUNCH_AMUL(0, false, m.view(), diag, upS, loS, psi, Apsi, (PcgScalars*)nullptr,
                         (double*)nullptr, (unsigned*)nullptr);
    if (halo) {
        haloExchange(m, w.sendBuf, w.recvBuf);
        B200_LAUNCH(k_interface_update, gridFor(m.nBRows), BLOCK, rt().stream, m.view(), bou, w.recvBuf.p, Apsi, m.nBRows, w.bRowOfSlot.p);
    }
}

void dicCalcReciprocalD(DeviceMesh& m, const double* diag, const double* upS, double* rD) {
    PcgWorkspace& w = pcgWorkspace(m);
    ScopedTimer t(T_DIC_RD);
    if (!m.N) return;
    w.keptValid = false;   // the operand blocks no longer belong to the kept matrix
    launchCalcRD(m, w, diag, upS, upS, rD, w.cf.p);
}

void dicPrecondition(DeviceMesh& m, const double* rD, const double* upS, const double* rA, double* wA) {
    PcgWorkspace& w = pcgWorkspace(m);
    if (!m.N) return;
    w.keptValid = false;
    // y / z double as ready flags (NaN sentinel = not yet published).  A PCG solve that stopped leaves whichever sweep ran last fully
    // populated (the kernels after the stop flag return at once and never re-arm it), so every entry point re-arms both first.
    fillSentinel(w.y, m.N); fillSentinel(w.z, m.N);
    { ScopedTimer t(T_DIC_RD); launchCoeffs(m, w, rD, upS, upS, w.cf.p, w.cb.p); }
    { ScopedTimer t(T_DIC_FWD); launchFwd(m, w, rD, rA, nullptr); }
    { ScopedTimer t(T_DIC_BWD); launchBwd(m, w, rA, nullptr, nullptr, nullptr); }
    B200_CHECK_CUDA(cudaMemcpyAsync(wA, w.z.p, 8 * (size_t)m.N, cudaMemcpyDeviceToDevice, rt().stream));
    fillSentinel(w.z, m.N);
}

b200_solver_perf pcgSolveDevice(DeviceMesh& m, const double* diag, const double* upS, const double* bou, double* x,
                                const double* b, const b200_solver_controls& ctl, std::vector<double>* resHist) {
    PcgWorkspace& w = pcgWorkspace(m);
    ScopedTimer tt(T_PCG_SOLVE);
    cudaStream_t st = rt().stream;
    const int N = m.N, grid = std::max(1, gridFor(N)), nChunks = gridFor(N);
    const bool multi = rt().nRanks > 1;
    const bool halo = multi && bou && meshHasProcessorPatches(m);
    MeshView mv = m.view();
    int histCap = resHist ? ctl.maxIter + 2 : 0;
    if (histCap > 0) w.hist.ensure(histCap);
    PcgScalars h; memset(&h, 0, sizeof(h));
    h.tol = ctl.tolerance; h.relTol = ctl.relTol; h.minIter = ctl.minIter; h.maxIter = ctl.maxIter; h.histCap = histCap;
    double nGlobal = (double)N;
    h.nGlobal = nGlobal;
    w.hsc[0] = h;
    B200_CHECK_CUDA(cudaMemcpyAsync(w.sc.p, &w.hsc[0], sizeof(PcgScalars), cudaMemcpyHostToDevice, st));
    PcgScalars* sc = w.sc.p;
    double* part = w.partials.p; unsigned* cnt = w.counters.p;
    double* hist = histCap > 0 ? w.hist.p : nullptr;
    if (multi) allReduce(&sc->nGlobal, 1, RED_SUM);
    // --- start-up: wA = A x, rA = b - wA, normFactor, initial residual
    const double* bouOrNull = halo ? bou : nullptr;
    if (halo) haloPack(m, x, w.sendBuf);
    B200_LAUNCH(k_pcg_init1, grid, BLOCK, st, mv, diag, upS, (const double*)x, b, bouOrNull, w.wA.p, w.rA.p, w.pA.p, sc, part, cnt);
    if (halo) {
        haloExchange(m, w.sendBuf, w.recvBuf);
        B200_LAUNCH(k_interface_update, gridFor(m.nBRows), BLOCK, st, mv, bou, w.recvBuf.p, w.wA.p, m.nBRows, w.bRowOfSlot.p);
        B200_LAUNCH(k_pcg_resid, grid, BLOCK, st, N, b, w.wA.p, w.rA.p, sc, part, cnt);
    }
    if (multi) { allReduce(&sc->sumX, 1, RED_SUM); allReduce(&sc->sumMagR, 1, RED_SUM); }
    B200_LAUNCH(k_pcg_init2, grid, BLOCK, st, N, w.wA.p, b, w.pA.p, sc, part, cnt);
    if (multi) allReduce(&sc->nfSum, 1, RED_SUM);
    B200_LAUNCH(k_pcg_check0, 1, 32, st, sc, hist);
    B200_CHECK_CUDA(cudaMemcpyAsync(&w.hsc[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    B200_CHECK_CUDA(cudaEventRecord(w.ev[0], st));
    // calcReciprocalD is enqueued before the host learns whether it is needed (one sweep; keeps the GPU busy)
    if (N) {
        ScopedTimer t(T_DIC_RD);
        const PcgScalars* gate = launchMatrixGate(m, w, diag, upS);
        launchCalcRD(m, w, diag, upS, upS, w.rD.p, w.cf.p, gate); launchCoeffs(m, w, w.rD.p, upS, upS, w.cf.p, w.cb.p, gate);
        fillSentinel(w.y, N); fillSentinel(w.z, N);   // re-arm both ready-flag vectors (see dicPrecondition): the previous solve's stop path left one dirty
    }
    B200_CHECK_CUDA(cudaEventSynchronize(w.ev[0]));
    bool stopped = w.hsc[0].stop != 0;
    const int BATCH = (int)std::max(1.0, rt().opt("pcg_batch", 4));   // iterations enqueued per host poll (two batches in flight)
    int slot = 0, enq = 0;
    const bool kt = timers().enabled && rt().opt("pcg_kernel_timers", 0) > 0;   // per-kernel-class events (roofline pass)
#define KT(id) std::unique_ptr<ScopedTimer> _kt(kt ? new ScopedTimer(id) : nullptr)
    // single rank: the four kernels of an iteration are chained by programmatic dependent launch (option pcg_pdl)
    const bool pdl = !multi && !kt && rt().opt("pcg_pdl", 1) > 0;
    // multi-rank: only the sweeps start early (the forward sweep's predecessor is the peer-memory all-reduce kernel, which triggers at once)
    const bool pdlSweeps = !kt && rt().opt("pcg_pdl", 1) > 0;
    if (multi && !w.peerTried) { w.peerTried = true; w.peer = peerHaloGet(m); }   // collective: every rank reaches its first multi-rank solve
    PeerHalo* peer = halo ? w.peer : nullptr;
    bool firstIteration = true;
    // peer-memory halo: k_halo_push (NVLink stores + system-scope fences, ~10 us) runs on a side stream BESIDE the local Amul instead of in
    // front of it; the next p-update waits for it (it reads pA), which by then is long finished
    const bool sidePush = peer && !kt && rt().opt("p2p_push_stream", 1) > 0;
    if (sidePush && !w.pushStream) {
        // (highest priority: its few small blocks are dispatched ahead of the pending CTAs of the big kernels whenever an SM has room --
        //  the neighbours' interface updates wait for this rank's push)
        int prLo = 0, prHi = 0;
        B200_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
        B200_CHECK_CUDA(cudaStreamCreateWithPriority(&w.pushStream, cudaStreamNonBlocking, prHi));
        B200_CHECK_CUDA(cudaEventCreateWithFlags(&w.evP, cudaEventDisableTiming));
        B200_CHECK_CUDA(cudaEventCreateWithFlags(&w.evPush, cudaEventDisableTiming));
    }
    bool pushPending = false;
    auto enqueueIteration = [&]() {
        // (an empty rank still runs the sweeps: they carry the iteration's residual sum / convergence test and the wArA sum)
        { KT(T_DIC_FWD);   // + the previous iteration's x / rA update, residual sum and convergence test (fused)
          // (the first sweep of a solve follows k_dic_coeffs, which writes the operand blocks the prologue stages: no early start)
          launchFwd(m, w, w.rD.p, w.rA.p, sc, PcgFuse{x, w.rA.p, w.pA.p, w.wA.p, sc, part, cnt, hist, multi ? 0 : 1}, pdlSweeps && !firstIteration); }
        firstIteration = false;
        // multi-rank: ONE all-reduce carries the previous iteration's sum|rA| and this iteration's sum(wA*rA); the convergence
        // test therefore runs after the backward sweep (one wasted sweep at the end of a solve).  Over peer memory it rides in the
        // tail of the backward sweep (fold1); otherwise a stand-alone kernel / NCCL + k_pcg_check
        const P2PFold fold1 = (multi && !kt) ? p2pFoldNext() : P2PFold{};
        { KT(T_DIC_BWD); launchBwd(m, w, w.rA.p, sc, part, cnt, multi ? &sc->wArAnew : &sc->wArA, pdlSweeps, nullptr, fold1, hist); }
        if (multi && !fold1.on) { KT(T_ALLREDUCE);
          if (!allReduceSumThenPcgCheck(&sc->sumMagR, 2, sc, hist)) { allReduce(&sc->sumMagR, 2, RED_SUM); B200_LAUNCH(k_pcg_check, 1, 32, st, sc, hist); } }
        if (pushPending) { B200_CHECK_CUDA(cudaStreamWaitEvent(st, w.evPush, 0)); pushPending = false; }
        { KT(T_PCG_VEC);
          B200_LAUNCH_PDL(pdl || (pdlSweeps && commUsesPeerMemory()), k_pcg_p_update, std::max(1, gridFor((N + 1) / 2)), BLOCK, 0, st, N, w.z.p, w.pA.p, (const PcgScalars*)sc); }
        P2PFold fold2 = P2PFold{};
        if (halo) {
            // Amul of the local block with its dot product, then the interface update (+ dot correction) over the boundary rows only.
            // peer: the halo goes over NVLink peer memory -- push -> Amul (overlaps the transfer) -> wait inside k_iface_rows; else NCCL
            { KT(T_HALO);
              if (sidePush) {
                  B200_CHECK_CUDA(cudaEventRecord(w.evP, st));
                  B200_CHECK_CUDA(cudaStreamWaitEvent(w.pushStream, w.evP, 0));
                  peerHaloPush(peer, m, w.pA, false, w.pushStream);
                  B200_CHECK_CUDA(cudaEventRecord(w.evPush, w.pushStream));
                  pushPending = true;
              } else if (peer) peerHaloPush(peer, m, w.pA, pdlSweeps);
              else haloPack(m, w.pA, w.sendBuf); }
            { KT(T_AMUL); B200_LAUNCH_AMUL(1, pdlSweeps && peer != nullptr, mv, diag, upS, upS, (const double*)w.pA.p, w.wA.p, sc, part, cnt); }
            // ... and gSumProd(wA, pA) rides in the tail of the interface update (fold2)
            fold2 = !kt ? p2pFoldNext() : P2PFold{};
            { KT(T_HALO);
              if (!peer) haloExchange(m, w.sendBuf, w.recvBuf);
              // (not itself launched early: measured at 2 GPUs, CTAs that spin on the neighbours' flags beside the Amul tail cost 10 us per iteration;
              //  its pdlTrigger lets the next forward sweep's prologue overlap it)
              B200_LAUNCH_PDL(false, k_iface_rows, std::max(1, gridFor(m.nBRows)), BLOCK, 0, st, mv, bou,
                              peer ? peerHaloRecv(peer) : (const double*)w.recvBuf.p, w.wA.p, (const double*)w.pA.p, m.nBRows, (const int*)w.bRowOfSlot.p, sc, part, cnt,
                              peer ? peerHaloWait(peer) : PeerHaloWait{}, fold2); }
        } else {
            KT(T_AMUL);
            B200_LAUNCH_AMUL(1, pdl, mv, diag, upS, upS, (const double*)w.pA.p, w.wA.p, sc, part, cnt);
        }
        if (multi && !fold2.on) { KT(T_ALLREDUCE); allReduce(&sc->wApA, 1, RED_SUM); }
    };
#undef KT
    // two batches in flight: while the host waits for batch i's flag, batch i+1 is already queued
    int pendingSlots = 0;
    while (!stopped) {
        while (pendingSlots < 2 && enq < ctl.maxIter + BATCH) {
            for (int i = 0; i < BATCH; i++) enqueueIteration();
            enq += BATCH;
            int sl = (slot + pendingSlots) & 1;
            B200_CHECK_CUDA(cudaMemcpyAsync(&w.hsc[sl], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
            B200_CHECK_CUDA(cudaEventRecord(w.ev[sl], st));
            pendingSlots++;
        }
        if (pendingSlots == 0) break;
        B200_CHECK_CUDA(cudaEventSynchronize(w.ev[slot]));
        stopped = w.hsc[slot].stop != 0;
        h = w.hsc[slot];
        slot ^= 1; pendingSlots--;
    }
    // drain anything still queued (no-op iterations) and read the final scalars
    if (pushPending) { B200_CHECK_CUDA(cudaStreamWaitEvent(st, w.evPush, 0)); pushPending = false; }   // the last side-stream push reads pA
    B200_CHECK_CUDA(cudaMemcpyAsync(&w.hsc[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    syncStream();
    h = w.hsc[0];
    b200_solver_perf perf;
    perf.initialResidual = h.initialResidual; perf.finalResidual = h.finalResidual; perf.nIterations = h.nIter;
    perf.converged = h.converged; perf.singular = h.singular;
    if (resHist) {
        resHist->resize(h.nIter + 1);
        B200_CHECK_CUDA(cudaMemcpyAsync(resHist->data(), w.hist.p, 8 * (size_t)(h.nIter + 1), cudaMemcpyDeviceToHost, st));
        syncStream();
    }
    return perf;
}

// ================================================================================================ b200PBiCG + DILU
// [UPSTREAM DILUPreconditioner.C]  calcReciprocalD: rD[u] -= upper*lower/rD[l] (faces ascending), rD = 1/rD
//   precondition  : wA = rD*rA; forward (losort order)  wA[u] -= rD[u]*lower*wA[l]; backward (faces descending) wA[l] -= rD[l]*upper*wA[u]
//   preconditionT : wT = rD*rT; forward (faces asc)     wT[u] -= rD[u]*upper*wT[l]; backward (losort desc)      wT[l] -= rD[l]*lower*wT[u]
// Per row the updates arrive in the same order as in DIC (faces with u == row ascending; faces with l == row descending), so the
// closed-tile sweeps of dic_kernels.cuh apply unchanged with other coefficient columns: bitwise the sequential loops.
static void diluEnsure(DeviceMesh& m, PcgWorkspace& w) {
    if (w.pT.n >= (size_t)std::max(m.N, 1)) return;
    w.pT.alloc(m.N); w.wT.alloc(m.N); w.rT.alloc(m.N); w.cfT.alloc(m.fwdBlkSize); w.cbT.alloc(m.bwdBlkSize); w.dotScal.alloc(8);
    if (m.fwdBlkSize) B200_CHECK_CUDA(cudaMemcpyAsync(w.cfT.p, m.dFwdBlk0.p, 8 * m.fwdBlkSize, cudaMemcpyDeviceToDevice, rt().stream));
    if (m.bwdBlkSize) B200_CHECK_CUDA(cudaMemcpyAsync(w.cbT.p, m.dBwdBlk0.p, 8 * m.bwdBlkSize, cudaMemcpyDeviceToDevice, rt().stream));
}
void diluCalcReciprocalD(DeviceMesh& m, const double* diag, const double* upS, const double* loS, double* rD) {
    PcgWorkspace& w = pcgWorkspace(m);
    ScopedTimer t(T_DIC_RD);
    if (!m.N) return;
    w.keptValid = false;
    launchCalcRD(m, w, diag, upS, loS, rD, w.cf.p);
}
// one application of the preconditioner with the operand blocks cf / cb (already filled): out = M^-1 rA
static void sweepApply(DeviceMesh& m, PcgWorkspace& w, const double* rD, const double* rA, const double* cf, const double* cb, double* out) {
    fillSentinel(w.y, m.N); fillSentinel(w.z, m.N);
    { ScopedTimer t(T_DIC_FWD); launchFwd(m, w, rD, rA, nullptr, PcgFuse{}, false, cf); }
    { ScopedTimer t(T_DIC_BWD); launchBwd(m, w, rA, nullptr, nullptr, nullptr, nullptr, false, cb); }
    B200_CHECK_CUDA(cudaMemcpyAsync(out, w.z.p, 8 * (size_t)m.N, cudaMemcpyDeviceToDevice, rt().stream));
}
void diluPrecondition(DeviceMesh& m, const double* rD, const double* upS, const double* loS, const double* rA, double* wA, bool transpose) {
    PcgWorkspace& w = pcgWorkspace(m);
    if (!m.N) return;
    w.keptValid = false;
    { ScopedTimer t(T_DIC_RD); launchCoeffs(m, w, rD, transpose ? upS : loS, transpose ? loS : upS, w.cf.p, w.cb.p); }
    sweepApply(m, w, rD, rA, w.cf.p, w.cb.p, wA);
}

// rA = b - wA ; rT = rA ; partial sums of x and |rA|
static __global__ void k_pbicg_resid(int N, const double* __restrict__ b, const double* __restrict__ wA, const double* __restrict__ x,
                                     double* __restrict__ rA, double* __restrict__ rT, PcgScalars* sc, double* partials, unsigned* counter) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (row < N) { double r = b[row] - wA[row]; rA[row] = r; rT[row] = r; v[0] = x[row]; v[1] = fabs(r); }
    gridReduce<2, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<2, OpSum>(gridDim.x, partials, counter, &sc->sumX, &sc->sumMagR);
}
// pA = wA + beta*pA ; pT = wT + beta*pT  (first iteration: copies)
static __global__ void k_pbicg_p(int N, const double* __restrict__ wA, const double* __restrict__ wT, double* __restrict__ pA, double* __restrict__ pT,
                                 int first, double beta) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= N) return;
    if (first) { pA[row] = wA[row]; pT[row] = wT[row]; }
    else { pA[row] = wA[row] + beta * pA[row]; pT[row] = wT[row] + beta * pT[row]; }
}
// psi += alpha*pA ; rA -= alpha*wA ; rT -= alpha*wT ; sum|rA|
static __global__ void k_pbicg_update(int N, double alpha, const double* __restrict__ pA, const double* __restrict__ wA, const double* __restrict__ wT,
                                      double* __restrict__ x, double* __restrict__ rA, double* __restrict__ rT, double* out, double* partials,
                                      unsigned* counter) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (row < N) {
        x[row] = x[row] + alpha * pA[row];
        double r = rA[row] - alpha * wA[row];
        rA[row] = r; rT[row] = rT[row] - alpha * wT[row];
        v[0] = fabs(r);
    }
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpSum>(gridDim.x, partials, counter, out);
}

// [UPSTREAM PBiCG.C solve]  The momentum predictor converges in a handful of iterations, so the scalars (wArT, wApT, alpha, beta, residual)
// are read back every iteration and the loop is upstream's, line by line; vectors, sweeps and products stay on the device.
b200_solver_perf pbicgSolveDevice(DeviceMesh& m, const double* diag, const double* upS, const double* loS, const double* bou, const double* intc,
                                  double* x, const double* b, const b200_solver_controls& ctl, std::vector<double>* resHist) {
    PcgWorkspace& w = pcgWorkspace(m);
    ScopedTimer tt(T_PCG_SOLVE);
    diluEnsure(m, w);
    cudaStream_t st = rt().stream;
    const int N = m.N, grid = std::max(1, gridFor(N));
    const bool multi = rt().nRanks > 1;
    MeshView mv = m.view();
    PcgScalars* sc = w.sc.p;
    double* part = w.partials.p; unsigned* cnt = w.counters.p;
    double* ds = w.dotScal.p;
    auto readScal = [&](const double* p, int n, double* h) {
        B200_CHECK_CUDA(cudaMemcpyAsync(h, p, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
        syncStream();
    };
    auto gsumDev = [&](double* p, int n) { if (multi) allReduce(p, n, RED_SUM); };
    auto dot = [&](const double* a, const double* c) {
        B200_LAUNCH(k_dot, grid, BLOCK, st, N, a, c, ds, part, cnt, (const PcgScalars*)nullptr);
        gsumDev(ds, 1);
        double h; readScal(ds, 1, &h); return h;
    };
    // wA = A psi ; wT = A^T psi (only its interface part differs from what the iteration needs: upstream computes it and so do we)
    amulDevice(m, diag, upS, loS, x, w.wA, bou);
    B200_LAUNCH(k_pbicg_resid, grid, BLOCK, st, N, b, (const double*)w.wA.p, (const double*)x, w.rA.p, w.rT.p, sc, part, cnt);
    gsumDev(&sc->sumX, 2);
    // normFactor [UPSTREAM lduMatrix::solver::normFactor]: sumA -> pA (scratch), xRef = gAverage(psi)
    double hN = (double)N;
    B200_CHECK_CUDA(cudaMemcpyAsync(&sc->nGlobal, &hN, 8, cudaMemcpyHostToDevice, st));
    gsumDev(&sc->nGlobal, 1);
    if (N) B200_LAUNCH(k_sumA, grid, BLOCK, st, mv, diag, upS, loS, (multi && bou && meshHasProcessorPatches(m)) ? bou : (const double*)nullptr, w.pA.p);
    B200_LAUNCH(k_pcg_init2, grid, BLOCK, st, N, (const double*)w.wA.p, b, (const double*)w.pA.p, sc, part, cnt);
    gsumDev(&sc->nfSum, 1);
    double h2[2];   // {sumMagR, nfSum}: sumMagR sits at offset 1, nfSum at offset 3 of PcgScalars
    PcgScalars hs;
    B200_CHECK_CUDA(cudaMemcpyAsync(&hs, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    syncStream();
    (void)h2;
    const double normFactor = hs.nfSum + 1.0e-20;
    b200_solver_perf perf;
    perf.initialResidual = perf.finalResidual = hs.sumMagR / normFactor;
    perf.nIterations = 0; perf.singular = 0; perf.converged = 0;
    if (resHist) { resHist->clear(); resHist->push_back(perf.initialResidual); }
    auto converged = [&]() { return perf.finalResidual < ctl.tolerance || (ctl.relTol > 0 && perf.finalResidual < ctl.relTol * perf.initialResidual); };
    auto stop = [&]() { return perf.nIterations >= ctl.minIter && (perf.nIterations >= ctl.maxIter || converged()); };
    if (!stop()) {
        // DILU: rD once, then the operand blocks of M^-1 (cf, cb) and of M^-T (cfT, cbT)
        w.keptValid = false;
        if (N) {
            ScopedTimer t(T_DIC_RD);
            launchCalcRD(m, w, diag, upS, loS, w.rD.p, w.cf.p);
            launchCoeffs(m, w, w.rD.p, loS, upS, w.cf.p, w.cb.p);
            launchCoeffs(m, w, w.rD.p, upS, loS, w.cfT.p, w.cbT.p);
        }
        double wArT = 1.0e20, wArTold;
        do {
            wArTold = wArT;
            sweepApply(m, w, w.rD.p, w.rA.p, w.cf.p, w.cb.p, w.wA.p);
            sweepApply(m, w, w.rD.p, w.rT.p, w.cfT.p, w.cbT.p, w.wT.p);
            wArT = dot(w.wA.p, w.rT.p);
            { ScopedTimer t(T_PCG_VEC);
              B200_LAUNCH(k_pbicg_p, grid, BLOCK, st, N, (const double*)w.wA.p, (const double*)w.wT.p, w.pA.p, w.pT.p, perf.nIterations == 0 ? 1 : 0,
                          perf.nIterations == 0 ? 0.0 : wArT / wArTold); }
            amulDevice(m, diag, upS, loS, w.pA, w.wA, bou);
            amulDevice(m, diag, loS, upS, w.pT, w.wT, intc);       // Tmul: lower and upper swap roles, interfaces use interfaceIntCoeffs
            const double wApT = dot(w.wA.p, w.pT.p);
            if (fabs(wApT) / normFactor <= 1.0e-300) { perf.singular = 1; break; }   // checkSingularity
            const double alpha = wArT / wApT;
            { ScopedTimer t(T_PCG_VEC);
              B200_LAUNCH(k_pbicg_update, grid, BLOCK, st, N, alpha, (const double*)w.pA.p, (const double*)w.wA.p, (const double*)w.wT.p, x, w.rA.p, w.rT.p,
                          ds, part, cnt); }
            gsumDev(ds, 1);
            double s; readScal(ds, 1, &s);
            perf.finalResidual = s / normFactor;
            perf.nIterations++;
            if (resHist) resHist->push_back(perf.finalResidual);
        } while (!stop());
    }
    perf.converged = converged() ? 1 : 0;
    return perf;
}

}  // namespace b200
