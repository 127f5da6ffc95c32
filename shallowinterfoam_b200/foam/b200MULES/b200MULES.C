/*---------------------------------------------------------------------------*\
  b200MULES.C -- host adapter: flattens the foam-extend fields into the C-ABI layout ([internal faces][boundary faces patch by
  patch]), keeps the virtual BC dispatch (psi.correctBoundaryConditions(): sifAlpha1 et al.) on the host, and calls
  b200_mules_explicit_solve.  Compiled with wmake where foam-extend-3.1 exists; in this repository against foam/miniFoam.H (foam/Makefile) and run by tests/test_adapter.py.
\*---------------------------------------------------------------------------*/
#include "b200MULES.H"
#include "b200Runtime.H"
#ifndef B200_MINIFOAM
#include "fvMesh.H"
#include "processorFvPatch.H"
#include "HashTable.H"
#endif

namespace
{
    using namespace Foam;

    //- device mesh of an fvMesh, cached for the life of the process
    b200_mesh_t* deviceMesh(const fvMesh& mesh)
    {
        static HashTable<b200_mesh_t*, const fvMesh*, Hash<const fvMesh*> > cache;
        if (cache.found(&mesh)) return cache[&mesh];

        const label nIF = mesh.nInternalFaces();
        label nBF = 0;
        forAll(mesh.boundary(), p) nBF += mesh.boundary()[p].size();
        labelList owner(nIF + nBF), pStart(mesh.boundary().size()), pSize(mesh.boundary().size()), pRank(mesh.boundary().size());
        vectorField Sf(nIF + nBF), Cf(nIF + nBF);
        scalarField magSf(nIF + nBF), w(nIF + nBF), dc(nIF + nBF);
        for (label f = 0; f < nIF; f++)
        {
            owner[f] = mesh.owner()[f]; Sf[f] = mesh.Sf()[f]; Cf[f] = mesh.Cf()[f]; magSf[f] = mesh.magSf()[f];
            w[f] = mesh.weights()[f]; dc[f] = mesh.deltaCoeffs()[f];
        }
        label off = nIF;
        forAll(mesh.boundary(), p)
        {
            const fvPatch& fp = mesh.boundary()[p];
            pStart[p] = off; pSize[p] = fp.size();
            pRank[p] = isA<processorFvPatch>(fp) ? refCast<const processorFvPatch>(fp).neighbProcNo() : -1;
            forAll(fp, i)
            {
                owner[off + i] = fp.faceCells()[i]; Sf[off + i] = fp.Sf()[i]; Cf[off + i] = fp.Cf()[i];
                magSf[off + i] = fp.magSf()[i]; w[off + i] = fp.weights()[i]; dc[off + i] = fp.deltaCoeffs()[i];
            }
            off += fp.size();
        }
        b200_mesh_desc d;
        d.nCells = mesh.nCells(); d.nInternalFaces = nIF; d.nBoundaryFaces = nBF; d.nPatches = mesh.boundary().size();
        d.owner = owner.begin(); d.neighbour = mesh.neighbour().begin();
        d.patchStart = pStart.begin(); d.patchSize = pSize.begin(); d.patchNeighbRank = pRank.begin();
        d.Sf = reinterpret_cast<const double*>(Sf.begin()); d.magSf = magSf.begin();
        d.Cf = reinterpret_cast<const double*>(Cf.begin());
        d.C = reinterpret_cast<const double*>(mesh.C().internalField().begin()); d.V = mesh.V().field().begin();
        d.weights = w.begin(); d.deltaCoeffs = dc.begin();
        b200_mesh_t* m = NULL;
        if (b200_mesh_create(&d, &m) != B200_OK)
        {
            FatalErrorIn("b200::MULES") << b200_last_error() << abort(FatalError);
        }
        cache.insert(&mesh, m);
        return m;
    }

    void flatten(const surfaceScalarField& s, scalarField& out)
    {
        const label nIF = s.internalField().size();
        forAll(s.internalField(), f) out[f] = s.internalField()[f];
        label off = nIF;
        forAll(s.boundaryField(), p) { forAll(s.boundaryField()[p], i) out[off + i] = s.boundaryField()[p][i]; off += s.boundaryField()[p].size(); }
    }
}

void Foam::b200::MULES::explicitSolve
(
    volScalarField& psi,
    const surfaceScalarField& phi,
    surfaceScalarField& phiPsi,
    const scalar psiMax,
    const scalar psiMin
)
{
    b200::start();                                       // MULES may be reached before any b200PCG solve (stock solver for pd / pcorr)
    const fvMesh& mesh = psi.mesh();
    psi.correctBoundaryConditions();                     // virtual BC dispatch stays on the host (sifAlpha1::updateCoeffs)

    b200_mesh_t* dm = deviceMesh(mesh);
    const label nIF = mesh.nInternalFaces();
    label nBF = 0;
    forAll(mesh.boundary(), p) nBF += mesh.boundary()[p].size();

    // patch kinds: processor patches carry patchNeighbourField in psiBoundary
    labelList types(mesh.boundary().size());
    scalarField psiB(nBF), zero(nBF, 0.0);
    label off = 0;
    forAll(psi.boundaryField(), p)
    {
        const fvPatchScalarField& pf = psi.boundaryField()[p];
        types[p] = pf.coupled() ? B200_BC_PROCESSOR : B200_BC_CALCULATED;   // values are already evaluated by the host BCs
        tmp<scalarField> tv = pf.coupled() ? pf.patchNeighbourField() : tmp<scalarField>(new scalarField(pf));
        forAll(pf, i) psiB[off + i] = tv()[i];
        off += pf.size();
    }
    b200_bc bc; bc.type = types.begin(); bc.refValue = zero.begin(); bc.refGrad = zero.begin(); bc.valueFraction = zero.begin();

    scalarField phiFlat(nIF + nBF), phiPsiFlat(nIF + nBF);
    flatten(phi, phiFlat); flatten(phiPsi, phiPsiFlat);

    if (b200_mules_explicit_solve(dm, &bc, psi.internalField().begin(), psiB.begin(), psi.oldTime().internalField().begin(),
                                  phiFlat.begin(), phiPsiFlat.begin(), mesh.time().deltaT().value(), psiMax, psiMin, NULL) != B200_OK)
    {
        FatalErrorIn("b200::MULES::explicitSolve") << b200_last_error() << abort(FatalError);
    }

    forAll(phiPsi.internalField(), f) phiPsi.internalField()[f] = phiPsiFlat[f];
    off = nIF;
    forAll(phiPsi.boundaryField(), p) { forAll(phiPsi.boundaryField()[p], i) phiPsi.boundaryField()[p][i] = phiPsiFlat[off + i]; off += phiPsi.boundaryField()[p].size(); }

    psi.correctBoundaryConditions();
}
