// region3d.cuh -- device-resident replay of /root/reference region3d/solve3d.H:40-72 (laminar, momentumPredictor off):
// alphaEqnSubCycle.H -> interface.correct() -> rho -> UEqn.H assembly -> nCorr x pEqn.H -> continuityErrs -> p.
#pragma once
#include <memory>
#include "device_mesh.cuh"
#include "fv_kernels.cuh"
#include "pcg.cuh"

namespace b200 {

struct DevBC {
    int nc = 1, B = 0;
    std::vector<int> types;
    DevBuf<int> dType;
    DevBuf<double> rv, rg, vf;  // SoA: component k at rv + k*B
    void upload(const DeviceMesh& m, const b200_bc* bc, int nc);
    void setTypes(const DeviceMesh& m, const std::vector<int>& t, int nc);  // zero-valued data
    BCView view() const { return BCView{dType.p, rv.p, rg.p, vf.p}; }
    bool fixes(int patch) const { return types[patch] == B200_BC_FIXED_VALUE || types[patch] == B200_BC_MIXED; }
};

// scratch of the finite-volume operators, cached per mesh
struct FvWorkspace {
    DevBuf<double> phiBD, lamA, lamB, flux, psiMaxn, psiMinn, sumPhip, mSumPhim, lamP0, lamM0, lamP1, lamM1;
    DevBuf<double> gx, gy, gz, gb, partials, scal, lamPn, lamMn, haloSend;
    DevBuf<double> stg[16];  // staging for the host-pointer API
    DevBuf<unsigned> counter;
    DevBuf<double> invT;
    bool haveInvT = false;
};
FvWorkspace& fvWorkspace(DeviceMesh& m);
const double* reconTensor(DeviceMesh& m);

// MULES::explicitSolve on device arrays (row/slot order).  phiPsi: in high-order flux, out limited flux.
void mulesExplicitDevice(DeviceMesh& m, const BCView& bc, double* psi, const double* psib, const double* psi0,
                         const double* phi, double* phiPsi, double deltaT, double psiMax, double psiMin, int includeOwn,
                         int nLimiterIter, double* lambdaOut);
// limiter only: phiBD/phiCorr given, lambda out
void mulesLimiterDevice(DeviceMesh& m, const BCView& bc, const double* psi, const double* psib, const double* psi0,
                        const double* phiBD, const double* phiCorr, double deltaT, double psiMax, double psiMin,
                        int includeOwn, int nLimiterIter, double* lambdaOut);

struct Region3D {
    DeviceMesh* m = nullptr;
    b200_region3d_controls ctl;
    DevBC bcAlpha, bcU, bcPd, bcPcorr;
    // cell fields
    DevBuf<double> alpha, alpha0, alphaStart, pd, rho, rho0, K, rUA, p, gh, uDiag, pdDiag, pdSource, pcorr;
    DevBuf<double> Ux, Uy, Uz, U0x, U0y, U0z, Hx, Hy, Hz, uSx, uSy, uSz;
    // boundary fields
    DevBuf<double> alphab, pdb, rhob, rho0b, Ub, U0b, uIC, uBC, ic, bcv, pcorrb, Kb, rUAb, haloSend;
    // face fields
    DevBuf<double> phi, phi0, rhoPhi, rhoPhiSum, nHatf, ghf, phiAlpha, phiU, rUAf, muf, uLower, uUpper, pdUpper, oneF;
    DevBuf<double> momFlux, momDiag, momSrc, corrP;   // momentum predictor: face flux of the explicit pressure-gradient terms, per-component system
    DevBuf<double> ffc, corrA, corrR;   // non-orthogonal correction: faceFluxCorrection of the pd / pcorr laplacian; corrected-snGrad parts of alpha1, rho
    DevBuf<double> scal;  // device scalars: [0] maxPhic, [1..4] adjustPhi sums, [5..8] alpha stats, [9..11] contErr sums
    double* hscal = nullptr;
    double deltaN = 0;
    std::vector<b200_solver_perf> perfs;
    std::vector<double> resHist, alphaStats;
    std::vector<size_t> pendingStats;
    bool keepHist = true;
    // capture (debug / bench): host copies, in foam-extend's ORIGINAL cell / face order, of every linear system handed to the solver
    // and of every MULES::explicitSolve input -- what the two plug points would receive from an unmodified shallowInterFoam
    bool capture = false;
    struct CapturedSolve { std::vector<double> diag, upper, source, x0, bou; double tol, relTol; };
    struct CapturedMules { std::vector<double> psi, psib, psi0, phi, phiPsi; double deltaT; };
    std::vector<CapturedSolve> capSolves;
    std::vector<CapturedMules> capMules;
    bool pdNeedRef = false;      // pd.needReference(): no patch of pd fixes the value, on any rank  [UPSTREAM GeometricField::needReference]
    int refRow = -1;             // device row of pdRefCell (-1: not on this rank / none)
    double sumLocalContErr = 0, globalContErr = 0, cumulativeContErr = 0;

    ~Region3D();
    void create(DeviceMesh* mesh, const b200_region3d_controls& c, const b200_bc* a, const b200_bc* u, const b200_bc* p,
                const double* alpha1, const double* U, const double* pdHost);
    void calculateK();
    void alphaEqn(double dt);
    void assembleUEqn();
    void momentumPredictor();                               // [REF region3d/UEqn.H:19-34] solve(UEqn == reconstruct(...)) with b200PBiCG + DILU
    void pEqn(int corr);
    void solvePressureLike(const DevBC& bc, const double* gammaf, double* psi, double* psib, double tol, double relTol,
                           bool subtractFlux);
    void correctPhi();
    void step();
    void resolveStats();
    bool nonOrth() const { return ctl.nonOrthCorrected != 0 && !m->orthogonal; }   // `corrected` schemes act only on a non-orthogonal mesh
    // out[f] = (gammaf*magSf)_f * (corrVec & interpolate(grad(vf)))_f   (gammaf null: the bare correction of the corrected snGrad)
    void nonOrthCorrection(const double* vf, const double* vfb, const double* gammaf, double* out);
    void courant(double* out3);                             // {meanCoNum, CoNum, velMag} [REF include/CourantNoMultiRegion.H:78-92]
    void refreshDerived();                                  // rho and interface.correct() from the current alpha1 (restart)
    bool coupled() const;                                   // nRanks > 1 and the mesh has processor patches
    void haloScalar(const double* cellField, double* bdst);  // patchNeighbourField of processor faces -> bdst[b]
    void evalPdBC(const DevBC& bc, const double* psi, double* psib);
    void evalAlphaBC();
    void evalUBC();
};

}  // namespace b200
