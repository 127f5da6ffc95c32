/* b200foam.h -- C-ABI of libb200foam.so, the B200-native drop-in for shallowInterFoam's 3D-region hot path.
 *
 * Reference = mintgen/shallowInterFoam (/root/reference), built on foam-extend-3.1.  The hot path is
 *   region3d/alphaEqn.H + alphaEqnSubCycle.H   (MULES-limited alpha1 advection)
 *   region3d/pEqn.H (+ correctPhi.H)           (pd assembly + lduMatrix PCG/DIC solve)
 * whose arithmetic lives in foam-extend-3.1 (Make/options:22-31).  Every entry point below names the
 * reference call site [REF file:line] and the upstream interface it replaces [UPSTREAM].
 *
 * Conventions (foam-extend-3.1 ABI): scalar = double, label = int32_t.  All pointers are HOST pointers
 * owned by the caller; nothing is retained past the call except what b200_ldu_get_or_create /
 * b200_mesh_create copy into their device caches.  Cell fields [nCells]; vectors are interleaved xyz
 * (Foam::vector) [3*nCells]; face fields [nInternalFaces + nBoundaryFaces] (internal faces in LDU face
 * order, then boundary faces patch by patch); boundary fields [nBoundaryFaces].
 * Every function returns 0 on success or a negative b200_status; b200_last_error() gives the message
 * (the foam-extend adapter turns it into FatalErrorIn(...) << abort(FatalError)).  Not thread-safe:
 * one single-threaded process per rank, as in foam-extend.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with B200_ERR_NO_DEVICE.
 */
#ifndef B200FOAM_H
#define B200FOAM_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B200_OK = 0,
    B200_ERR_NO_DEVICE = -1,
    B200_ERR_CUDA = -2,
    B200_ERR_NCCL = -3,
    B200_ERR_ARG = -4,
    B200_ERR_UNSUPPORTED = -5, /* e.g. asymmetric matrix given to b200PCG, cyclic/ggi interfaces */
    B200_ERR_STATE = -6
} b200_status;

typedef struct b200_ldu b200_ldu_t;         /* device cache of one lduAddressing (+ interfaces)      */
typedef struct b200_mesh b200_mesh_t;       /* b200_ldu + fvMesh geometry                            */
typedef struct b200_region3d b200_region3d_t; /* device-resident fields + controls of one 3D region  */
typedef struct b200_polymesh b200_polymesh_t; /* host polyMesh produced by the in-repo generators     */

/* [UPSTREAM lduInterface / processorLduInterface: faceCells(), neighbProcNo()] */
typedef struct {
    int32_t nFaces;
    const int32_t* faceCells; /* [nFaces] */
    int32_t neighbRank;       /* -1: not a processor patch */
    int32_t tag;
} b200_patch_desc;

/* [UPSTREAM lduMatrix::solver controls read from the fvSolution sub-dictionary: tolerance, relTol, minIter, maxIter]
 * [REF region3d/pEqn.H:34,38 mesh.solutionDict().solver("pd"|"pdFinal"); correctPhi.H:45 "pcorr"] */
typedef struct {
    double tolerance; /* 1e-6 */
    double relTol;    /* 0    */
    int32_t minIter;  /* 0    */
    int32_t maxIter;  /* 1000 */
} b200_solver_controls;

/* [UPSTREAM lduSolverPerformance] solverName is "b200DICPCG" */
typedef struct {
    double initialResidual;
    double finalResidual;
    int32_t nIterations;
    int32_t converged;
    int32_t singular;
} b200_solver_perf;

/* patch-field kinds understood by the device-side boundary algebra [UPSTREAM fvPatchField hierarchy].
 * The sif* coupling BCs of the reference are mixed BCs evaluated on the host
 * [REF derivedFvPatchFields/sifAlpha1/sifAlpha1FvPatchScalarField.C:342-588 updateCoeffs]; the host uploads
 * their refValue/refGrad/valueFraction through b200_region3d_set_patch_mixed. */
enum { B200_BC_FIXED_VALUE = 0, B200_BC_ZERO_GRADIENT = 1, B200_BC_MIXED = 2, B200_BC_PROCESSOR = 3, B200_BC_CALCULATED = 4 };

typedef struct {
    const int32_t* type;         /* [nPatches]                         */
    const double* refValue;      /* [nBoundaryFaces * nCmpt]           */
    const double* refGrad;       /* [nBoundaryFaces * nCmpt]           */
    const double* valueFraction; /* [nBoundaryFaces]                   */
} b200_bc;

/* ---------------------------------------------------------------- runtime */
/* One process per GPU.  ncclId: the 128-byte ncclUniqueId made by rank 0 with b200_nccl_unique_id and broadcast
 * by the host application (Pstream / MPI / torch.distributed); ignored when nRanks == 1. */
int b200_init(int device, int rank, int nRanks, const void* ncclId, size_t ncclIdBytes);
int b200_nccl_unique_id(void* out128);
int b200_finalize(void);
const char* b200_last_error(void);
int b200_device_count(void);
/* tuning knobs (defaults are the measured best): "sweep_grid" = CTAs of the data-flow DIC sweeps (0 = auto) */
int b200_set_option(const char* name, double value);
/* number of kernels this library launched since process start (bench.py's gpu_launches) */
long long b200_kernel_launch_count(void);
/* CUDA-event timing of kernel classes since the last reset: out[0..n) milliseconds, names via b200_timer_name */
int b200_timers_reset(void);
int b200_timers_get(double* outMs, long long* outCalls, int cap);
const char* b200_timer_name(int i);

/* ---------------------------------------------------------------- lduAddressing cache
 * [UPSTREAM lduAddressing: lowerAddr(), upperAddr(), patchAddr(i)]  The solver object foam-extend builds per solve
 * is a temporary (fvMatrix::solve), so device state is cached here, keyed by the address of the lduAddressing. */
int b200_ldu_get_or_create(const void* key, int32_t nCells, int32_t nFaces, const int32_t* lower, const int32_t* upper,
                           int32_t nPatches, const b200_patch_desc* patches, b200_ldu_t** out);
int b200_ldu_release(b200_ldu_t* ldu);
/* drop the cache entry of `key` (mesh motion / topology change: the lduAddressing object was rebuilt at the same address).
 * A hit in b200_ldu_get_or_create is also validated against a hash of lower/upper/faceCells and rebuilt when it differs. */
int b200_ldu_invalidate(const void* key);
/* renumbering used on the device: rowCell[row] = original cell of device row `row`; level[cell] = forward DIC level */
int b200_ldu_get_renumbering(const b200_ldu_t* ldu, int32_t* rowCell, int32_t* cellLevel, int32_t* nLevels);
/* DIC tiles of the device row order: rows are ordered (tile, level, original cell); tileStart[nTiles+1].
 * Pass tileStart = NULL to query the sizes only. */
int b200_ldu_get_tiling(const b200_ldu_t* ldu, int32_t* bandLevels, int32_t* tileRows, int32_t* nTiles, int32_t* tileStart);
/* debug: with option "sweep_trace" > 0 the forward sweep stamps %globaltimer per tile (ticket, staged, halo ready, done);
 * out[8*nTiles] ns (5 stamps used), tileLevel[nTiles] = level of the tile in the tile graph */
int b200_debug_get_sweep_trace(const b200_ldu_t* ldu, unsigned long long* out, int32_t cap, int32_t* tileLevel);
/* derived addressing, for bit-exact parity with lduAddressing::ownerStartAddr/losortAddr/losortStartAddr */
int b200_ldu_get_addressing(const b200_ldu_t* ldu, int32_t* ownerStart, int32_t* losort, int32_t* losortStart);

/* ---------------------------------------------------------------- a1  lduMatrix::Amul
 * [REF region3d/pEqn.H:34,38; correctPhi.H:45 (inside PCG)] [UPSTREAM lduMatrixATmul.C Amul]
 * lower == NULL or lower == upper => symmetric.  bouCoeffs/psiNbr: [nBoundaryFaces] or NULL (processor interfaces;
 * when NULL and nRanks > 1 the halo is exchanged over NCCL). */
int b200_amul(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, const double* psi,
              double* Apsi, const double* bouCoeffs, const double* psiNbr);
/* [UPSTREAM lduMatrixATmul.C sumA] */
int b200_sumA(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, double* sumA,
              const double* bouCoeffs);

/* ---------------------------------------------------------------- a2/a3  DICPreconditioner
 * [UPSTREAM DICPreconditioner.C calcReciprocalD / precondition] level-scheduled, bitwise the sequential result */
int b200_dic_calc_reciprocal_d(b200_ldu_t* ldu, const double* diag, const double* upper, double* rD);
int b200_dic_precondition(b200_ldu_t* ldu, const double* rD, const double* upper, const double* rA, double* wA);

/* ---------------------------------------------------------------- a4  PCG::solve  ("solver b200PCG; preconditioner DIC;")
 * [REF region3d/pEqn.H:34,38; correctPhi.H:45] [UPSTREAM PCG.C solve, lduMatrixSolver.C normFactor/stop]
 * x: initial guess in, solution out.  resHist (optional, capacity histCap) receives the normalised residual after
 * every iteration (entry 0 = initial residual). */
int b200_pcg_solve(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower,
                   const double* bouCoeffs, double* x, const double* b, const b200_solver_controls* ctl,
                   b200_solver_perf* perf, double* resHist, int32_t histCap);

/* ---------------------------------------------------------------- f1  PBiCG::solve + DILUPreconditioner  ("solver b200PBiCG; preconditioner DILU;")
 * [REF region3d/UEqn.H:19-34  solve(UEqn == fvc::reconstruct(...)), momentumPredictor defaults to true: readRegion3dPISOControls.H:9-10]
 * [UPSTREAM PBiCG.C solve; DILUPreconditioner.C calcReciprocalD / precondition / preconditionT; lduMatrixATmul.C Tmul]
 * Asymmetric matrices (lower != upper).  bouCoeffs / intCoeffs: interfaceBouCoeffs (Amul) and interfaceIntCoeffs (Tmul) of the processor
 * patches ([nBoundaryFaces] or NULL).  One call per vector component, as fvMatrix<vector>::solve does. */
int b200_pbicg_solve(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, const double* bouCoeffs,
                     const double* intCoeffs, double* x, const double* b, const b200_solver_controls* ctl, b200_solver_perf* perf,
                     double* resHist, int32_t histCap);
int b200_dilu_calc_reciprocal_d(b200_ldu_t* ldu, const double* diag, const double* upper, const double* lower, double* rD);
/* transpose != 0: preconditionT */
int b200_dilu_precondition(b200_ldu_t* ldu, const double* rD, const double* upper, const double* lower, const double* rA, double* wA,
                           int32_t transpose);

/* ---------------------------------------------------------------- mesh (lduAddressing + fvMesh geometry) */
typedef struct {
    int32_t nCells, nInternalFaces, nBoundaryFaces, nPatches;
    const int32_t* owner;      /* [nInternalFaces + nBoundaryFaces] */
    const int32_t* neighbour;  /* [nInternalFaces] */
    const int32_t* patchStart; /* [nPatches] absolute face index */
    const int32_t* patchSize;  /* [nPatches] */
    const int32_t* patchNeighbRank; /* [nPatches] -1 = physical patch */
    const double* Sf;          /* [3*(nIF+nBF)] */
    const double* magSf;       /* [nIF+nBF] */
    const double* Cf;          /* [3*(nIF+nBF)] */
    const double* C;           /* [3*nCells] */
    const double* V;           /* [nCells] */
    const double* weights;     /* [nIF+nBF] */
    const double* deltaCoeffs; /* [nIF+nBF] */
} b200_mesh_desc;
int b200_mesh_create(const b200_mesh_desc* desc, b200_mesh_t** out);
int b200_mesh_release(b200_mesh_t* mesh);
b200_ldu_t* b200_mesh_ldu(b200_mesh_t* mesh);
/* [UPSTREAM surfaceInterpolation::makeCorrectionVectors] non-orthogonality measure asin(min(sum(magSf*|corrVec|)/sum(magSf), 1)) in degrees over
 * the internal faces of all ranks, and mesh.orthogonal() (< 0.1 degrees: the `corrected` schemes are no-ops) */
int b200_mesh_non_orthogonality(b200_mesh_t* mesh, double* degrees, int32_t* orthogonal);
/* [UPSTREAM correctedSnGrad::correction; gaussLaplacianScheme::fvmLaplacian faceFluxCorrection] out[nIF+nBF] = (gammaf*magSf) *
 * (corrVec & linearInterpolate(fvc::grad(vf))) with the Gauss-linear gradient; gammaf == NULL: the bare snGrad correction */
int b200_non_orth_correction(b200_mesh_t* mesh, const double* gammaf, const double* vf, const double* vfBoundary, double* out);

/* ---------------------------------------------------------------- a13 building blocks (host-pointer forms)
 * [REF region3d/pEqn.H:3,16-20,47] [UPSTREAM linear interpolation, snGradScheme, fvcSurfaceIntegrate, gaussGrad, fvcReconstruct] */
int b200_interpolate(b200_mesh_t* mesh, const double* vf, const double* vfBoundary, int32_t nCmpt, double* out);
int b200_sngrad(b200_mesh_t* mesh, const b200_bc* bc, const double* vf, const double* vfBoundary, double* out);
int b200_surface_integrate(b200_mesh_t* mesh, const double* ssf, double* out);
int b200_grad_gauss(b200_mesh_t* mesh, const b200_bc* bc, const double* vf, const double* vfBoundary, double* grad,
                    double* gradBoundary);
int b200_reconstruct(b200_mesh_t* mesh, const double* ssf, double* out);

/* ---------------------------------------------------------------- a7/a8  limited face fluxes
 * [REF region3d/alphaEqn.H:16-27] [UPSTREAM fvc::flux -> gaussConvectionScheme::flux; LimitedScheme<vanLeer,NVDTVD>;
 * interfaceCompression]  scheme: 0 vanLeer (Gauss-linear gradient computed internally), 1 interfaceCompression,
 * 2 upwind, 3 linear */
enum { B200_SCHEME_VANLEER = 0, B200_SCHEME_INTERFACE_COMPRESSION = 1, B200_SCHEME_UPWIND = 2, B200_SCHEME_LINEAR = 3 };
int b200_flux_limited(b200_mesh_t* mesh, int32_t scheme, const b200_bc* bc, const double* faceFlux, const double* vf,
                      const double* vfBoundary, double* out);

/* ---------------------------------------------------------------- a5/a6  MULES
 * [REF region3d/alphaEqn.H:31  MULES::explicitSolve(alpha1, phi, phiAlpha, 1, 0)] [UPSTREAM MULESTemplates.C]
 * The adapter calls psi.correctBoundaryConditions() on the host before and after (virtual BC dispatch -- sifAlpha1 --
 * stays host C++) and passes the evaluated boundary values.  phiPsi: high-order flux in, limited flux out.
 * lambdaOut (optional) [nIF+nBF] receives the limiter. */
int b200_mules_explicit_solve(b200_mesh_t* mesh, const b200_bc* bc, double* psi, const double* psiBoundary,
                              const double* psi0, const double* phi, double* phiPsi, double deltaT, double psiMax,
                              double psiMin, double* lambdaOut);
int b200_mules_limiter(b200_mesh_t* mesh, const b200_bc* bc, const double* psi, const double* psiBoundary,
                       const double* psi0, const double* phiBD, const double* phiCorr, double deltaT, double psiMax,
                       double psiMin, int32_t nLimiterIter, double* lambda);

/* ---------------------------------------------------------------- a10/a11/a12  pressure assembly
 * [REF region3d/pEqn.H:25-28,43] [UPSTREAM gaussLaplacianScheme::fvmLaplacianUncorrected, fvMatrix::negSumDiag,
 * operator==(fvMatrix, fvc::div(phi)), fvScalarMatrix::solve addBoundaryDiag/addBoundarySource, fvMatrix::flux] */
int b200_laplacian_assemble(b200_mesh_t* mesh, const b200_bc* bc, const double* gammaf, const double* psiBoundary,
                            const double* phi /* NULL: no == fvc::div(phi) */, double* diag, double* upper,
                            double* source, double* internalCoeffs, double* boundaryCoeffs);
int b200_matrix_flux(b200_mesh_t* mesh, const double* upper, const double* lower, const double* internalCoeffs,
                     const double* boundaryCoeffs, const double* psi, double* flux);
/* a14 [REF region3d/alphaEqn.H:36-40] out3 = {weightedAverage(V), min, max} */
int b200_alpha_stats(b200_mesh_t* mesh, const double* alpha, const double* alphaBoundary, double* out3);

/* ---------------------------------------------------------------- region3d: the device-resident 3D step
 * Replays region3d/solve3d.H:40-72 (laminar, momentumPredictor off) with every field resident in HBM:
 * alphaEqnSubCycle.H -> interface.correct() -> rho -> UEqn.H assembly -> nCorr x pEqn.H -> continuityErrs -> p.
 * [REF region3d/readRegion3dPISOControls.H:1-15, alphaEqnSubCycle.H:1-9, create3dFields.H] */
typedef struct {
    double deltaT;
    int32_t nCorr, nNonOrthCorr, nAlphaCorr, nAlphaSubCycles;
    double cAlpha;
    double rho1, rho2, nu1, nu2, sigma;
    double g[3];
    double pdTol, pdRelTol, pdFinalTol, pdFinalRelTol, pcorrTol, pcorrRelTol;
    int32_t maxIter, minIter;
    int32_t adjustPhi;       /* adjustPhi(phiU,U,p): p is `calculated` => needReference() [REF pEqn.H:14]          */
    int32_t mulesIncludeOwn; /* SURVEY.md Appendix C item 2                                                      */
    int32_t divUScheme;      /* div(rhoPhi,U): B200_SCHEME_UPWIND | B200_SCHEME_LINEAR                            */
    /* [REF region3d/readRegion3dPISOControls.H:9-10  momentumPredictor (default true); region3d/UEqn.H:19-34]
     * 1: solve(UEqn == reconstruct((sigmaK snGrad(alpha1) - ghf snGrad(rho) - snGrad(pd)) magSf)) with b200PBiCG + DILU            */
    int32_t momentumPredictor;
    /* laplacianSchemes / snGradSchemes `corrected` (1) or `uncorrected` (0): explicit non-orthogonal correction of
     * fvm::laplacian(rUAf, pd) (source term + faceFluxCorrection in pdEqn.flux()) [REF region3d/pEqn.H:23-45]                     */
    int32_t nonOrthCorrected;
    /* [REF region3d/pEqn.H:30  pdEqn.setReference(pdRefCell, pdRefValue); create3dFields.H:227-233 setRefCell]
     * applied only when no pd patch fixes the value (pd.needReference()); pdRefCell < 0: none                                     */
    int32_t pdRefCell;
    int32_t reserved0;
    double pdRefValue;
    double UTol, URelTol;    /* fvSolution solvers{U{tolerance; relTol}} of the momentum predictor                                 */
} b200_region3d_controls;

int b200_region3d_create(b200_mesh_t* mesh, const b200_region3d_controls* ctl, const b200_bc* bcAlpha,
                         const b200_bc* bcU, const b200_bc* bcPd, const double* alpha1, const double* U,
                         const double* pd, b200_region3d_t** out);
int b200_region3d_release(b200_region3d_t* r);
/* [REF shallowInterFoam.C:85-90 -> region3d/correctPhi.H] */
int b200_region3d_correct_phi(b200_region3d_t* r);
/* nSteps device-resident time steps (value path: inputs already in HBM) */
int b200_region3d_step(b200_region3d_t* r, int32_t nSteps);
/* end-to-end step with HOST buffers: H2D of alpha1,U,pd,phi; one step; D2H of alpha1,U,pd,phi,p (what
 * runTime.write() and the host-side sif BCs / Courant number read) */
int b200_region3d_step_host(b200_region3d_t* r, double* alpha1, double* U, double* pd, double* phi, double* p);
int b200_region3d_get_fields(b200_region3d_t* r, double* alpha1, double* U, double* pd, double* phi, double* p,
                             double* rho);
int b200_region3d_set_fields(b200_region3d_t* r, const double* alpha1, const double* U, const double* pd,
                             const double* phi);
/* host-evaluated mixed BC upload (sif* coupling): field 0 = alpha1, 1 = U, 2 = pd */
int b200_region3d_set_patch_mixed(b200_region3d_t* r, int32_t field, int32_t patch, const double* refValue,
                                  const double* refGrad, const double* valueFraction);
/* patch-internal gather for host BCs (KBs, not whole fields) [REF sifAlpha1FvPatchScalarField.C:431-435] */
int b200_region3d_get_patch_internal(b200_region3d_t* r, int32_t field, int32_t patch, double* out);
/* log of the solves since the last clear: perfs [nSolves], residual histories concatenated, alpha stats [n][3] */
int b200_region3d_num_solves(b200_region3d_t* r);
int b200_region3d_get_perfs(b200_region3d_t* r, b200_solver_perf* out, int32_t cap);
int b200_region3d_get_res_hist(b200_region3d_t* r, double* out, int32_t cap);
int b200_region3d_get_alpha_stats(b200_region3d_t* r, double* out, int32_t cap);
int b200_region3d_get_cont_errs(b200_region3d_t* r, double* out3);
/* [REF include/CourantNoMultiRegion.H:78-92] out3 = {meanCoNum, CoNum, velMag} of the current phi, reduced on the device
 * (gMax / gSum over ranks): the host never needs the face field */
int b200_region3d_get_courant(b200_region3d_t* r, double* out3);
/* restart from a saved state: set_fields + the fields derived from alpha1 that a step expects to find (rho, interface.correct()) */
int b200_region3d_restore(b200_region3d_t* r, const double* alpha1, const double* U, const double* pd, const double* phi);
int b200_region3d_clear_log(b200_region3d_t* r);
/* capture (bench / debugging): with `on`, every later step keeps host copies -- in foam-extend's original cell / face order -- of each
 * linear system handed to the solver (after addBoundaryDiag/Source: what fvScalarMatrix::solve passes to lduMatrix::solver::solve) and
 * of each MULES::explicitSolve input: exactly what the two plug points receive from an unmodified shallowInterFoam.  kind 0 = solves,
 * 1 = MULES calls.  bench.py replays them through b200_pcg_solve / b200_mules_explicit_solve from pageable host arrays ("e2e_plugin"). */
int b200_region3d_capture(b200_region3d_t* r, int32_t on);
int b200_region3d_num_captured(b200_region3d_t* r, int32_t kind);
int b200_region3d_get_captured_solve(b200_region3d_t* r, int32_t i, double* diag, double* upper, double* source, double* x0, double* bou,
                                     double* tolRelTol);
int b200_region3d_get_captured_mules(b200_region3d_t* r, int32_t i, double* psi, double* psib, double* psi0, double* phi, double* phiPsi,
                                     double* deltaT);

/* ---------------------------------------------------------------- host mesh tools (integers bit-exact)
 * [UPSTREAM blockMesh (single hex block ordering), subsetMesh, primitiveMesh geometry, surfaceInterpolation
 * weights/deltaCoeffs, simpleGeomDecomp, decomposePar processor-patch construction] */
int b200_polymesh_box(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz, b200_polymesh_t** out);
/* remove cells whose centre has x in [x0,x1)*lx and z < zfrac*lz ("weir"); exposed faces -> last patch "weir" */
int b200_polymesh_box_weir(int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz, double x0, double x1,
                           double zfrac, b200_polymesh_t** out);
/* skewed mesh (BASELINE configs[3]): moves every vertex that is not on a boundary face by (2u-1)*frac*(dx,dy,dz) with u from the
 * 64-bit LCG x <- x*6364136223846793005 + 1442695040888963407 (u = (x >> 11) * 2^-53; three draws per vertex, vertices in index order,
 * x0 = seed).  Addressing is untouched; faces become non-planar and non-orthogonal (what moveDynamicMesh / a skewed blockMesh gives). */
int b200_polymesh_jitter(b200_polymesh_t* pm, double frac, double dx, double dy, double dz, uint64_t seed);
/* non-orthogonal (but not skewed) mesh: affine shear of every point, x += sxy*y + sxz*z, y += syz*z */
int b200_polymesh_shear(b200_polymesh_t* pm, double sxy, double sxz, double syz);
/* polyhedral mesh by agglomeration (BASELINE configs[3], irregular LDU addressing): cells with equal group[cell] merge.  New labels by
 * first appearance; faces between merged cells vanish; the rest are flipped to owner < neighbour and sorted (owner, neighbour, original
 * face); boundary faces and patches keep their order.  Two cells may share several faces. */
int b200_polymesh_agglomerate(const b200_polymesh_t* pm, const int64_t* group, b200_polymesh_t** out);
int b200_polymesh_release(b200_polymesh_t* pm);
int b200_polymesh_sizes(const b200_polymesh_t* pm, int32_t* nCells, int32_t* nInternalFaces, int32_t* nBoundaryFaces,
                        int32_t* nPatches, int32_t* nPoints);
int b200_polymesh_get(const b200_polymesh_t* pm, double* points, int32_t* faces4, int32_t* owner, int32_t* neighbour,
                      int32_t* patchStart, int32_t* patchSize, int32_t* patchNeighbRank);
const char* b200_polymesh_patch_name(const b200_polymesh_t* pm, int32_t patch);
int b200_polymesh_geometry(const b200_polymesh_t* pm, double* Sf, double* magSf, double* Cf, double* C, double* V,
                           double* weights, double* deltaCoeffs);
/* simple decomposition: procOfCell[nCells] */
int b200_decompose_simple(const b200_polymesh_t* pm, int32_t npx, int32_t npy, int32_t npz, int32_t* procOfCell);
/* sub-mesh of `rank` (decomposePar layout); cellMap/faceMap/faceFlip may be NULL */
int b200_polymesh_submesh(const b200_polymesh_t* pm, const int32_t* procOfCell, int32_t nRanks, int32_t rank,
                          b200_polymesh_t** out);
int b200_polymesh_submesh_maps(const b200_polymesh_t* sub, int32_t* cellMap, int32_t* faceMap, int32_t* faceFlip);
/* device mesh straight from a host polyMesh (geometry computed on the host) */
int b200_mesh_create_from_polymesh(const b200_polymesh_t* pm, b200_mesh_t** out);

#ifdef __cplusplus
}
#endif
#endif /* B200FOAM_H */
