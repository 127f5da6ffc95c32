// pcg.cuh -- b200PCG: device-resident DIC-preconditioned CG on the cached lduAddressing.
#pragma once
#include "device_mesh.cuh"
#include "ldu_kernels.cuh"
#include "comm.cuh"

namespace b200 {


struct PcgWorkspace {
    DevBuf<double> pA, wA, rA, y, z, rD, Dt, partials, hist;
    DevBuf<double> cf, cb;                     // per-solve DIC coefficients rD[row]*upper in the sweep tiles' ELL layout
    // b200PBiCG + DILU (momentum predictor): the transposed system's vectors and sweep operand blocks (allocated on first use)
    DevBuf<double> pT, wT, rT, cfT, cbT, loS, intc, dotScal;
    DevBuf<double> diag, upS, x, b, bou;       // staging for the host-pointer API
    DevBuf<double> sendBuf, recvBuf;           // processor-patch halo
    DevBuf<PcgScalars> sc;
    // DIC reuse: foam-extend builds a new solver (and calcReciprocalD) per solve, but the PISO correctors of one time step
    // solve the SAME pd matrix (rUA is fixed when momentumPredictor is off; also every non-orthogonal corrector).  The last
    // solve's diag / upper are kept; a solve whose coefficients are bitwise the same keeps rD and the sweep operand blocks.
    DevBuf<double> keptDiag, keptUp;
    DevBuf<PcgScalars> gate;                   // gate->stop = 1: coefficients unchanged, the calcReciprocalD kernels return at once
    bool keptValid = false;
    PeerHalo* peer = nullptr;                  // halo swap of Amul over NVLink peer memory (comm.cuh); null: NCCL
    bool peerTried = false;
    DevBuf<unsigned> counters;                 // [0] reduce counter, [1] sweep ticket, [2] sweep done
    DevBuf<int> bRowOfSlot;
    PcgScalars* hsc = nullptr;                 // pinned mirror (2 slots)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t pushStream = nullptr;         // side stream of the peer-memory halo push (runs beside the local Amul)
    cudaEvent_t evP = nullptr, evPush = nullptr;
    int histCap = 0;
    int sweepGrid = 0;
    double rowsPerLevel = 1.0;
    ~PcgWorkspace();
};

PcgWorkspace& pcgWorkspace(DeviceMesh& m);

// device-pointer forms (row order / slot order); used by the region3d step and by the host API wrappers
void amulDevice(DeviceMesh& m, const double* diag, const double* upS, const double* loS, const double* psi, double* Apsi,
                const double* bou /*[B] or null*/);
void dicCalcReciprocalD(DeviceMesh& m, const double* diag, const double* upS, double* rD);
void dicPrecondition(DeviceMesh& m, const double* rD, const double* upS, const double* rA, double* wA);
// x in/out (rows); bou = interface coefficients of processor patches ([B], may be null)
b200_solver_perf pcgSolveDevice(DeviceMesh& m, const double* diag, const double* upS, const double* bou, double* x,
                                const double* b, const b200_solver_controls& ctl, std::vector<double>* resHist);
// b200PBiCG + DILU [UPSTREAM PBiCG.C solve; DILUPreconditioner.C]: asymmetric systems (momentum predictor, [REF region3d/UEqn.H:19-34]).
// bou / intc: interfaceBouCoeffs (Amul) and interfaceIntCoeffs (Tmul) of processor patches ([B], may be null)
void diluCalcReciprocalD(DeviceMesh& m, const double* diag, const double* upS, const double* loS, double* rD);
void diluPrecondition(DeviceMesh& m, const double* rD, const double* upS, const double* loS, const double* rA, double* wA, bool transpose);
b200_solver_perf pbicgSolveDevice(DeviceMesh& m, const double* diag, const double* upS, const double* loS, const double* bou, const double* intc,
                                  double* x, const double* b, const b200_solver_controls& ctl, std::vector<double>* resHist);

}  // namespace b200
