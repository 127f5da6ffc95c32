/*---------------------------------------------------------------------------*\
  b200PBiCG.C -- run-time-selectable lduMatrix::solver for ASYMMETRIC matrices: the momentum predictor of
  [REF /root/reference region3d/UEqn.H:19-34] (`U { solver b200PBiCG; preconditioner DILU; tolerance 1e-6; relTol 0; }`).
  Replaces [UPSTREAM foam-extend-3.1 src/foam/matrices/lduMatrix/solvers/PBiCG/PBiCG.C + preconditioners/DILUPreconditioner].
  fvMatrix<vector>::solve calls it once per velocity component with that component's diagonal and source.
\*---------------------------------------------------------------------------*/
#include "b200PBiCG.H"
#include "b200LduCache.H"

#ifndef B200_MINIFOAM
#include "addToRunTimeSelectionTable.H"
#endif

namespace Foam
{
    defineTypeNameAndDebug(b200PBiCG, 0);

    // `solver b200PBiCG;` resolves through lduMatrix::solver::New's asymMatrix table
    lduSolver::addasymMatrixConstructorToTable<b200PBiCG>
        addb200PBiCGAsymMatrixConstructorToTable_;
}

Foam::b200PBiCG::b200PBiCG
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& coupleBouCoeffs,
    const FieldField<Field, scalar>& coupleIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& dict
)
:
    lduMatrix::solver(fieldName, matrix, coupleBouCoeffs, coupleIntCoeffs, interfaces, dict)
{}

Foam::lduSolverPerformance Foam::b200PBiCG::solve
(
    scalarField& x,
    const scalarField& b,
    const direction cmpt
) const
{
    b200::start();
    b200::requirePreconditioner(dict(), "DILU", "b200PBiCG::solve");

    List<b200_patch_desc> patches;
    const label nBnd = b200::describeInterfaces(interfaces_, patches);
    b200_ldu_t* ldu = b200::lduCache(matrix_, x.size(), patches);

    // Amul uses interfaceBouCoeffs, Tmul interfaceIntCoeffs [UPSTREAM lduMatrixATmul.C]
    scalarField bou(nBnd, 0.0), intc(nBnd, 0.0);
    b200::flattenCoupled(interfaces_, coupleBouCoeffs_, bou);
    b200::flattenCoupled(interfaces_, coupleIntCoeffs_, intc);

    b200_solver_controls ctl;
    ctl.tolerance = tolerance_; ctl.relTol = relTolerance_; ctl.minIter = minIter_; ctl.maxIter = maxIter_;
    b200_solver_perf perf;

    if (b200_pbicg_solve(ldu, matrix_.diag().begin(), matrix_.upper().begin(), matrix_.lower().begin(),
                         nBnd ? bou.begin() : NULL, nBnd ? intc.begin() : NULL, x.begin(), b.begin(), &ctl, &perf, NULL, 0) != B200_OK)
    {
        FatalErrorIn("b200PBiCG::solve") << b200_last_error() << abort(FatalError);
    }

    lduSolverPerformance solverPerf("b200DILUPBiCG", fieldName());
    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;
}
