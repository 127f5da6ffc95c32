/*---------------------------------------------------------------------------*\
  test_adapters.C -- drives the UNCHANGED adapter sources (b200PCG.C, b200PBiCG.C, b200MULES.C, compiled against miniFoam.H) through
  libb200foam.so the way foam-extend would: lduMatrix::solver::New(dict)->solve(x, b) via the run-time selection tables, and
  b200::MULES::explicitSolve(psi, phi, phiPsi, 1, 0) on stand-in vol/surface fields.  extern-C entry points for tests/test_adapter.py,
  which builds the inputs with numpy, compares the results with the CPU oracle and checks the error paths.  TEST INFRASTRUCTURE.
\*---------------------------------------------------------------------------*/
#include "miniFoam.H"
#include "b200MULES.H"
#include <cstring>

namespace Foam { FatalErrorType FatalError; }

using namespace Foam;

static std::string g_err;

namespace
{
struct ProcIface : public lduInterface, public processorLduInterface
{
    labelList fc_; int nbr_;
    ProcIface(const label* fc, label n, int nbr) : fc_(n), nbr_(nbr) { for (label i = 0; i < n; i++) fc_[i] = fc[i]; }
    const unallocLabelList& faceCells() const { return fc_; }
    int myProcNo() const { return 0; }
    int neighbProcNo() const { return nbr_; }
};
struct OtherIface : public lduInterface      // a coupled interface that is NOT a processor interface (cyclic, ggi ...)
{
    labelList fc_;
    const unallocLabelList& faceCells() const { return fc_; }
};
}

extern "C" const char* mini_last_error() { return g_err.c_str(); }

// solverName: "b200PCG" | "b200PBiCG" (looked up in the sym / asym table according to `lower`); preconditioner as given.
// nIface coupled patches described by (ifaceSize[p], faceCells concatenated, ifaceRank[p]); ifaceRank < 0 = a non-processor interface.
// perf5 = {initialResidual, finalResidual, nIterations, converged, singular}.  Returns 0, or 1 with mini_last_error() set.
extern "C" int mini_solve(const char* solverName, const char* preconditioner, int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
                          const double* diag, const double* upper, const double* lower /*null: symmetric*/, int nIface,
                          const int* ifaceSize, const int* ifaceFaceCells, const int* ifaceRank, const double* bouCoeffs, const double* intCoeffs,
                          double tolerance, double relTol, int maxIter, double* x, const double* b, double* perf5, char* solverNameOut)
{
    try
    {
        labelList l(nFaces), u(nFaces);
        for (int f = 0; f < nFaces; f++) { l[f] = lowerAddr[f]; u[f] = upperAddr[f]; }
        lduAddressing addr(nCells, l, u);
        scalarField d(nCells), up(nFaces), lo(nFaces);
        for (int c = 0; c < nCells; c++) d[c] = diag[c];
        for (int f = 0; f < nFaces; f++) { up[f] = upper[f]; lo[f] = lower ? lower[f] : upper[f]; }
        std::unique_ptr<lduMatrix> A(lower ? new lduMatrix(addr, d, up, lo) : new lduMatrix(addr, d, up));

        FieldField<Field, scalar> bou(nIface), intc(nIface);
        lduInterfaceFieldPtrsList ifaces(nIface);
        std::vector<std::unique_ptr<lduInterface> > ifs; std::vector<std::unique_ptr<lduInterfaceField> > iffs;
        int off = 0;
        for (int p = 0; p < nIface; p++)
        {
            if (ifaceRank[p] >= 0) ifs.emplace_back(new ProcIface(ifaceFaceCells + off, ifaceSize[p], ifaceRank[p]));
            else { OtherIface* o = new OtherIface(); o->fc_.setSize(ifaceSize[p]); for (int i = 0; i < ifaceSize[p]; i++) o->fc_[i] = ifaceFaceCells[off + i]; ifs.emplace_back(o); }
            iffs.emplace_back(new lduInterfaceField(*ifs.back()));
            ifaces.set(p, iffs.back().get());
            scalarField* pb = new scalarField(ifaceSize[p]); scalarField* pi = new scalarField(ifaceSize[p]);
            for (int i = 0; i < ifaceSize[p]; i++) { (*pb)[i] = bouCoeffs[off + i]; (*pi)[i] = intCoeffs[off + i]; }
            bou.set(p, pb); intc.set(p, pi);
            off += ifaceSize[p];
        }
        dictionary dict;
        dict.add("solver", solverName);
        if (preconditioner && *preconditioner) dict.add("preconditioner", preconditioner);
        { std::ostringstream os; os.precision(17); os << tolerance; dict.add("tolerance", os.str()); }
        { std::ostringstream os; os.precision(17); os << relTol; dict.add("relTol", os.str()); }
        { std::ostringstream os; os << maxIter; dict.add("maxIter", os.str()); }

        scalarField xs(nCells), bs(nCells);
        for (int c = 0; c < nCells; c++) { xs[c] = x[c]; bs[c] = b[c]; }
        // exactly what fvMatrix::solve does: a temporary solver object per solve, selected by name at run time
        std::unique_ptr<lduMatrix::solver> s(lduMatrix::solver::New("pd", *A, bou, intc, ifaces, dict));
        lduSolverPerformance sp = s->solve(xs, bs);
        for (int c = 0; c < nCells; c++) x[c] = xs[c];
        perf5[0] = sp.initialResidual(); perf5[1] = sp.finalResidual(); perf5[2] = sp.nIterations(); perf5[3] = sp.converged(); perf5[4] = sp.singular();
        if (solverNameOut) std::strncpy(solverNameOut, sp.solverName().c_str(), 63);
        return 0;
    }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// MULES::explicitSolve(psi, phi, phiPsi, 1, 0) on a stand-in fvMesh.  Face arrays are [internal faces][boundary faces patch by patch];
// patchRank[p] >= 0: processor patch (psiBoundary there = patchNeighbourField).  The host BC dispatch is a callback that counts its calls
// and re-imposes psiBoundary (fixed values).  Outputs: psi (in/out), phiPsi (in: high-order flux, out: limited flux), nCorrectBCs.
extern "C" int mini_mules(int nCells, int nIF, int nPatches, const int* patchStart, const int* patchSize, const int* patchRank,
                          const int* owner /*[nIF+nBF]*/, const int* neighbour, const double* Sf, const double* magSf, const double* Cf,
                          const double* C, const double* V, const double* weights, const double* deltaCoeffs, double deltaT,
                          double* psi, const double* psiBoundary, const double* psi0, const double* phi, double* phiPsi, int* nCorrectBCs)
{
    try
    {
        fvMesh mesh(deltaT);
        mesh.nCells_ = nCells;
        mesh.owner_.setSize(nIF); mesh.neighbour_.setSize(nIF); mesh.Sf_.setSize(nIF); mesh.Cf_.setSize(nIF);
        mesh.magSf_.setSize(nIF); mesh.weights_.setSize(nIF); mesh.deltaCoeffs_.setSize(nIF);
        auto vec = [](const double* a, int i) { vector v; v.x_ = a[3 * i]; v.y_ = a[3 * i + 1]; v.z_ = a[3 * i + 2]; return v; };
        for (int f = 0; f < nIF; f++)
        {
            mesh.owner_[f] = owner[f]; mesh.neighbour_[f] = neighbour[f]; mesh.Sf_[f] = vec(Sf, f); mesh.Cf_[f] = vec(Cf, f);
            mesh.magSf_[f] = magSf[f]; mesh.weights_[f] = weights[f]; mesh.deltaCoeffs_[f] = deltaCoeffs[f];
        }
        mesh.C_.f_.setSize(nCells); mesh.V_.f_.setSize(nCells);
        for (int c = 0; c < nCells; c++) { mesh.C_.f_[c] = vec(C, c); mesh.V_.f_[c] = V[c]; }
        volScalarField a(mesh), aOld(mesh);
        surfaceScalarField sphi, sphiPsi;
        a.internal_.setSize(nCells); aOld.internal_.setSize(nCells);
        for (int c = 0; c < nCells; c++) { a.internal_[c] = psi[c]; aOld.internal_[c] = psi0[c]; }
        sphi.internal_.setSize(nIF); sphiPsi.internal_.setSize(nIF);
        for (int f = 0; f < nIF; f++) { sphi.internal_[f] = phi[f]; sphiPsi.internal_[f] = phiPsi[f]; }
        for (int p = 0; p < nPatches; p++)
        {
            const int s = patchStart[p], n = patchSize[p];
            fvPatch* fp = patchRank[p] >= 0 ? new processorFvPatch() : new fvPatch();
            if (patchRank[p] >= 0) static_cast<processorFvPatch*>(fp)->neighbProcNo_ = patchRank[p];
            fp->faceCells_.setSize(n); fp->Sf_.setSize(n); fp->Cf_.setSize(n); fp->magSf_.setSize(n); fp->weights_.setSize(n); fp->deltaCoeffs_.setSize(n);
            std::shared_ptr<fvPatchScalarField> pf(new fvPatchScalarField(n, patchRank[p] >= 0));
            std::shared_ptr<scalarField> pphi(new scalarField(n)), pphiPsi(new scalarField(n));
            for (int i = 0; i < n; i++)
            {
                const int f = s + i;
                fp->faceCells_[i] = owner[f]; fp->Sf_[i] = vec(Sf, f); fp->Cf_[i] = vec(Cf, f); fp->magSf_[i] = magSf[f];
                fp->weights_[i] = weights[f]; fp->deltaCoeffs_[i] = deltaCoeffs[f];
                (*pf)[i] = psiBoundary[f - nIF]; pf->pnf_[i] = psiBoundary[f - nIF];
                (*pphi)[i] = phi[f]; (*pphiPsi)[i] = phiPsi[f];
            }
            mesh.boundary_.emplace_back(fp);
            a.boundary_.l_.push_back(pf); sphi.boundary_.l_.push_back(pphi); sphiPsi.boundary_.l_.push_back(pphiPsi);
        }
        a.old_.reset(new volScalarField(mesh)); a.old_->internal_ = aOld.internal_;
        Foam::b200::MULES::explicitSolve(a, sphi, sphiPsi, 1, 0);
        for (int c = 0; c < nCells; c++) psi[c] = a.internal_[c];
        for (int f = 0; f < nIF; f++) phiPsi[f] = sphiPsi.internal_[f];
        for (int p = 0; p < nPatches; p++) for (int i = 0; i < patchSize[p]; i++) phiPsi[patchStart[p] + i] = sphiPsi.boundary_[p][i];
        if (nCorrectBCs) *nCorrectBCs = a.nCorrectBCs_;
        return 0;
    }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}
