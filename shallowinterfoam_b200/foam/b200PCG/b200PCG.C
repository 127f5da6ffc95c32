/*---------------------------------------------------------------------------*\
  b200PCG.C -- see b200PCG.H.  Host C++ only: every loop runs in libb200foam.so on the GPU.
\*---------------------------------------------------------------------------*/
#include "b200PCG.H"
#include "b200LduCache.H"

#ifndef B200_MINIFOAM
#include "addToRunTimeSelectionTable.H"
#endif

namespace Foam
{
    defineTypeNameAndDebug(b200PCG, 0);

    // symmetric matrices only (PCG): `solver b200PCG;` resolves through lduMatrix::solver::New's symMatrix table
    lduSolver::addsymMatrixConstructorToTable<b200PCG>
        addb200PCGSymMatrixConstructorToTable_;
}

Foam::b200PCG::b200PCG
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

Foam::lduSolverPerformance Foam::b200PCG::solve
(
    scalarField& x,
    const scalarField& b,
    const direction cmpt
) const
{
    b200::start();
    b200::requirePreconditioner(dict(), "DIC", "b200PCG::solve");

    if (!matrix_.symmetric())
    {
        FatalErrorIn("b200PCG::solve") << "b200PCG handles symmetric matrices only (use b200PBiCG)" << abort(FatalError);
    }

    // the solver object is a per-solve temporary: device state is cached in the library, keyed by the lduAddressing
    List<b200_patch_desc> patches;
    const label nBnd = b200::describeInterfaces(interfaces_, patches);
    b200_ldu_t* ldu = b200::lduCache(matrix_, x.size(), patches);

    // interface (boundary) coefficients, patch by patch, in the library's boundary-face order
    scalarField bou(nBnd, 0.0);
    b200::flattenCoupled(interfaces_, coupleBouCoeffs_, bou);

    b200_solver_controls ctl;
    ctl.tolerance = tolerance_; ctl.relTol = relTolerance_; ctl.minIter = minIter_; ctl.maxIter = maxIter_;
    b200_solver_perf perf;

    if (b200_pcg_solve(ldu, matrix_.diag().begin(), matrix_.upper().begin(), NULL, nBnd ? bou.begin() : NULL,
                       x.begin(), b.begin(), &ctl, &perf, NULL, 0) != B200_OK)
    {
        FatalErrorIn("b200PCG::solve") << b200_last_error() << abort(FatalError);
    }

    lduSolverPerformance solverPerf("b200DICPCG", fieldName());
    solverPerf.initialResidual() = perf.initialResidual;
    solverPerf.finalResidual() = perf.finalResidual;
    solverPerf.nIterations() = perf.nIterations;
    solverPerf.converged() = perf.converged;
    solverPerf.singular() = perf.singular;
    return solverPerf;   // fvScalarMatrix::solve prints it: "b200DICPCG:  Solving for pd, Initial residual = ..."
}
