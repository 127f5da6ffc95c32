/*---------------------------------------------------------------------------*\
  b200PCG.C -- see b200PCG.H.  Host C++ only: every loop runs in libb200foam.so on the GPU.
\*---------------------------------------------------------------------------*/
#include "b200PCG.H"
#include "b200foam.h"

#ifndef B200_MINIFOAM
#include "processorLduInterface.H"
#include "addToRunTimeSelectionTable.H"
#include "Pstream.H"
#endif

namespace Foam
{
    defineTypeNameAndDebug(b200PCG, 0);

    // symmetric matrices only (PCG): `solver b200PCG;` resolves through lduMatrix::solver::New's symMatrix table
    lduSolver::addsymMatrixConstructorToTable<b200PCG>
        addb200PCGSymMatrixConstructorToTable_;
}

namespace
{
    //- one-off runtime start-up: one GPU per MPI rank, NCCL id broadcast over Pstream
    void b200Start()
    {
        static bool started = false;
        if (started) return;
        char id[128] = {0};
        if (Foam::Pstream::parRun())
        {
            if (Foam::Pstream::master()) b200_nccl_unique_id(id);
            Foam::List<char> lid(128);
            forAll(lid, i) lid[i] = id[i];
            Foam::Pstream::scatter(lid);
            forAll(lid, i) id[i] = lid[i];
        }
        const int nDev = b200_device_count();
        if (nDev < 1 || b200_init
            (
                Foam::Pstream::parRun() ? Foam::Pstream::myProcNo() % nDev : 0,
                Foam::Pstream::parRun() ? Foam::Pstream::myProcNo() : 0,
                Foam::Pstream::parRun() ? Foam::Pstream::nProcs() : 1,
                id, sizeof(id)
            ) != B200_OK)
        {
            FatalErrorIn("b200PCG") << "libb200foam: " << b200_last_error() << Foam::abort(Foam::FatalError);
        }
        started = true;
    }
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
    b200Start();

    if (!matrix_.symmetric())
    {
        FatalErrorIn("b200PCG::solve") << "b200PCG handles symmetric matrices only" << abort(FatalError);
    }

    const lduAddressing& addr = matrix_.lduAddr();

    // interfaces -> b200_patch_desc (processor patches only; cyclic/ggi are rejected by the library)
    const label nPatches = interfaces_.size();
    List<b200_patch_desc> patches(nPatches);
    label nBnd = 0;
    forAll(interfaces_, patchI)
    {
        patches[patchI].nFaces = 0; patches[patchI].faceCells = NULL; patches[patchI].neighbRank = -1; patches[patchI].tag = patchI;
        if (interfaces_.set(patchI))
        {
            const lduInterface& ifc = interfaces_[patchI].coupledInterface();
            patches[patchI].nFaces = ifc.faceCells().size();
            patches[patchI].faceCells = ifc.faceCells().begin();
            if (isA<processorLduInterface>(ifc))
            {
                patches[patchI].neighbRank = refCast<const processorLduInterface>(ifc).neighbProcNo();
            }
            else
            {
                FatalErrorIn("b200PCG::solve") << "only processor interfaces are supported" << abort(FatalError);
            }
            nBnd += patches[patchI].nFaces;
        }
    }

    // the solver object is a per-solve temporary: device state is cached in the library, keyed by the lduAddressing
    b200_ldu_t* ldu = NULL;
    if (b200_ldu_get_or_create(&addr, x.size(), addr.lowerAddr().size(), addr.lowerAddr().begin(), addr.upperAddr().begin(),
                               nPatches, patches.begin(), &ldu) != B200_OK)
    {
        FatalErrorIn("b200PCG::solve") << b200_last_error() << abort(FatalError);
    }

    // interface (boundary) coefficients, patch by patch, in the library's boundary-face order
    scalarField bou(nBnd, 0.0);
    label off = 0;
    forAll(interfaces_, patchI)
    {
        if (interfaces_.set(patchI))
        {
            const scalarField& pc = coupleBouCoeffs_[patchI];
            forAll(pc, i) bou[off + i] = pc[i];
            off += pc.size();
        }
    }

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
