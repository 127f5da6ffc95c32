// region3d.cu -- MULES driver + device-resident 3D-region step (see region3d.cuh).
#include "region3d.cuh"
#include <cmath>
#include "comm.cuh"

namespace b200 {

// ================================================================================================ BC upload
__global__ void k_aos_to_soa_impl(const double* __restrict__ src, double* __restrict__ dst, int n, int nc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int k = 0; k < nc; k++) dst[(size_t)k * n + i] = src[(size_t)nc * i + k];
}
__global__ void k_soa_to_aos_impl(const double* __restrict__ src, double* __restrict__ dst, int n, int nc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int k = 0; k < nc; k++) dst[(size_t)nc * i + k] = src[(size_t)k * n + i];
}

void DevBC::upload(const DeviceMesh& m, const b200_bc* bc, int nc_) {
    nc = nc_; B = m.B;
    int P = (int)m.patches.size();
    types.assign(bc->type, bc->type + P);
    dType.upload(types);
    size_t n = (size_t)nc * B;
    rv.alloc(n); rg.alloc(n); vf.alloc(B);
    if (B) {
        DevBuf<double> tmp(n);
        tmp.upload(bc->refValue, n);
        B200_LAUNCH(k_aos_to_soa_impl, gridFor(B), BLOCK, rt().stream, (const double*)tmp.p, rv.p, B, nc);
        tmp.upload(bc->refGrad, n);
        B200_LAUNCH(k_aos_to_soa_impl, gridFor(B), BLOCK, rt().stream, (const double*)tmp.p, rg.p, B, nc);
        vf.upload(bc->valueFraction, B);
        syncStream();
    }
}
void DevBC::setTypes(const DeviceMesh& m, const std::vector<int>& t, int nc_) {
    nc = nc_; B = m.B; types = t;
    dType.upload(types);
    rv.alloc((size_t)nc * B); rg.alloc((size_t)nc * B); vf.alloc(B);
    rv.zero(); rg.zero(); vf.zero();
    syncStream();
}

// ================================================================================================ workspace
FvWorkspace& fvWorkspace(DeviceMesh& m) {
    if (!m.fvHolder) {
        FvWorkspace* w = new FvWorkspace();
        m.fvHolder = std::shared_ptr<void>(w, [](void* p) { delete (FvWorkspace*)p; });
        size_t N = m.N, nf = m.nFaceField();
        w->phiBD.alloc(nf); w->lamA.alloc(nf); w->lamB.alloc(nf); w->flux.alloc(nf); w->flux.zero();
        w->psiMaxn.alloc(N); w->psiMinn.alloc(N); w->sumPhip.alloc(N); w->mSumPhim.alloc(N);
        w->lamP0.alloc(N); w->lamM0.alloc(N); w->lamP1.alloc(N); w->lamM1.alloc(N);
        w->gx.alloc(N); w->gy.alloc(N); w->gz.alloc(N); w->gb.alloc(3 * (size_t)m.B); w->gb.zero();
        w->lamPn.alloc(m.B); w->lamMn.alloc(m.B); w->haloSend.alloc(m.B);
        w->partials.alloc(8 * (size_t)(gridFor(std::max<long long>(m.N, m.B)) + 2));
        w->scal.alloc(16); w->counter.alloc(2); w->counter.zero();
        syncStream();
    }
    return *(FvWorkspace*)m.fvHolder.get();
}
const double* reconTensor(DeviceMesh& m) {
    FvWorkspace& w = fvWorkspace(m);
    if (!w.haveInvT) {
        w.invT.alloc(6 * (size_t)m.N);
        if (m.N) B200_LAUNCH(k_recon_tensor, gridFor(m.N), BLOCK, rt().stream, m.view(), w.invT.p);
        w.haveInvT = true;
    }
    return w.invT.p;
}

// ================================================================================================ MULES
static void mulesCore(DeviceMesh& m, const BCView& bc, const double* psi, const double* psib, const double* psi0,
                      const double* phiBD, const double* phiCorr, double deltaT, double psiMax, double psiMin,
                      int includeOwn, int nIter, double* psiOut, double* fluxOut, double* lambdaOut) {
    FvWorkspace& w = fvWorkspace(m);
    MeshView mv = m.view();
    cudaStream_t st = rt().stream;
    int grid = gridFor(m.N);
    if (!m.N) return;
    const double* nul = nullptr;
    const bool halo = rt().nRanks > 1 && meshHasProcessorPatches(m);
    const double* lpn = halo ? w.lamPn.p : nul; const double* lmn = halo ? w.lamMn.p : nul;
    B200_LAUNCH(k_mules_bounds, grid, BLOCK, st, mv, bc, psi, psib, nul, psi0, phiBD, phiCorr, deltaT, psiMax, psiMin, includeOwn,
                w.psiMaxn.p, w.psiMinn.p, w.sumPhip.p, w.mSumPhim.p);
    // iteration it (1-based) consumes cell lambdas of it-1 and face lambdas of it-2, writes face lambdas of it-1
    double *lpPrev = nullptr, *lmPrev = nullptr, *lpNew = w.lamP0.p, *lmNew = w.lamM0.p;
    double *facePrev = nullptr, *faceOut = w.lamA.p;
    for (int it = 1; it <= nIter; it++) {
        if (it == 1)
            B200_LAUNCH(k_mules_iter<true>, grid, BLOCK, st, mv, phiCorr, nul, faceOut, nul, nul, nul, nul, mv.bCoupled,
                        (const double*)w.psiMaxn.p, (const double*)w.psiMinn.p, (const double*)w.sumPhip.p, (const double*)w.mSumPhim.p, lpNew, lmNew);
        else {
            B200_LAUNCH(k_mules_iter<false>, grid, BLOCK, st, mv, phiCorr, (const double*)facePrev, faceOut, (const double*)lpPrev,
                        (const double*)lmPrev, lpn, lmn, mv.bCoupled, (const double*)w.psiMaxn.p, (const double*)w.psiMinn.p,
                        (const double*)w.sumPhip.p, (const double*)w.mSumPhim.p, lpNew, lmNew);
            facePrev = faceOut;
            faceOut = (faceOut == w.lamA.p) ? w.lamB.p : w.lamA.p;
        }
        if (halo) {   // the other side of a processor face limits with its own cell: swap lambdap/lambdam (syncFaceList(minEqOp))
            haloPack(m, lpNew, w.haloSend); haloExchange(m, w.haloSend, w.lamPn);
            haloPack(m, lmNew, w.haloSend); haloExchange(m, w.haloSend, w.lamMn);
        }
        lpPrev = lpNew; lmPrev = lmNew;
        lpNew = (lpNew == w.lamP0.p) ? w.lamP1.p : w.lamP0.p;
        lmNew = (lmNew == w.lamM0.p) ? w.lamM1.p : w.lamM0.p;
    }
    if (nIter == 0) {  // lambda = 1 everywhere: cell lambdas of "iteration 0" are 1
        fillValue(w.lamP0, m.N, 1.0); fillValue(w.lamM0, m.N, 1.0);
        lpPrev = w.lamP0.p; lmPrev = w.lamM0.p;
        if (halo) { fillValue(w.lamPn, m.B, 1.0); fillValue(w.lamMn, m.B, 1.0); }
    }
    B200_LAUNCH(k_mules_update, grid, BLOCK, st, mv, phiBD, (double*)nullptr, phiCorr, (const double*)facePrev, (const double*)lpPrev,
                (const double*)lmPrev, lpn, lmn, mv.bCoupled, psi0, deltaT, psiOut, lambdaOut, fluxOut);
}

void mulesExplicitDevice(DeviceMesh& m, const BCView& bc, double* psi, const double* psib, const double* psi0,
                         const double* phi, double* phiPsi, double deltaT, double psiMax, double psiMin, int includeOwn,
                         int nLimiterIter, double* lambdaOut) {
    ScopedTimer t(T_MULES);
    FvWorkspace& w = fvWorkspace(m);
    int n = (int)std::max<long long>(m.N, m.B);
    if (n) B200_LAUNCH(k_mules_bd_corr, gridFor(n), BLOCK, rt().stream, m.view(), phi, (const double*)psi, psib, phiPsi, w.phiBD.p);
    // phiPsi now holds phiCorr.  The final pass rebuilds every face flux on the fly for both of its rows, so the limited
    // flux is written to scratch (never in place: the neighbour row may still have to read phiCorr) and copied back.
    mulesCore(m, bc, psi, psib, psi0, w.phiBD, phiPsi, deltaT, psiMax, psiMin, includeOwn, nLimiterIter, psi, w.flux.p, lambdaOut);
    B200_CHECK_CUDA(cudaMemcpyAsync(phiPsi, w.flux.p, 8 * m.nFaceField(), cudaMemcpyDeviceToDevice, rt().stream));
}

void mulesLimiterDevice(DeviceMesh& m, const BCView& bc, const double* psi, const double* psib, const double* psi0,
                        const double* phiBD, const double* phiCorr, double deltaT, double psiMax, double psiMin,
                        int includeOwn, int nLimiterIter, double* lambdaOut) {
    ScopedTimer t(T_MULES);
    mulesCore(m, bc, psi, psib, psi0, phiBD, phiCorr, deltaT, psiMax, psiMin, includeOwn, nLimiterIter, nullptr, nullptr, lambdaOut);
}


// ================================================================================================ step kernels
// phi = linearInterpolate(U) & Sf  [REF region3d/create3dFields.H:99-114]
static __global__ void k_init_phi(MeshView mv, const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
                                  const double* __restrict__ Ub, double* __restrict__ phi) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double ax = Ux[row], ay = Uy[row], az = Uz[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double w = mv.w[pos];
            double fx = w * (ax - Ux[u]) + Ux[u], fy = w * (ay - Uy[u]) + Uy[u], fz = w * (az - Uz[u]) + Uz[u];
            phi[pos] = fx * mv.Sfx[pos] + fy * mv.Sfy[pos] + fz * mv.Sfz[pos];
        END_FACES
    }
    if (row < mv.B) {
        int p = mv.Fs + row, B = mv.B, c = mv.bFaceRow[row];
        double fx = bndFaceValue(mv, row, Ux[c], Ub[row]), fy = bndFaceValue(mv, row, Uy[c], Ub[B + row]), fz = bndFaceValue(mv, row, Uz[c], Ub[2 * B + row]);
        phi[p] = fx * mv.Sfx[p] + fy * mv.Sfy[p] + fz * mv.Sfz[p];
    }
}
// rho = alpha*rho1 + (1-alpha)*rho2 (cells and boundary) [REF region3d/alphaEqnSubCycle.H:35]
static __global__ void k_rho(int N, int B, const double* __restrict__ a, const double* __restrict__ ab, double rho1, double rho2,
                             double* __restrict__ rho, double* __restrict__ rhob) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) rho[i] = a[i] * rho1 + (1.0 - a[i]) * rho2;
    if (i < B) rhob[i] = ab[i] * rho1 + (1.0 - ab[i]) * rho2;
}
// out = a*x (+ out) over n entries
static __global__ void k_scale(int n, double a, const double* __restrict__ x, double* __restrict__ out, int accumulate) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = accumulate ? out[i] + a * x[i] : a * x[i];
}
// rhoPhi = phiAlpha*(rho1-rho2) + phi*rho2 [REF region3d/alphaEqn.H:33]
static __global__ void k_rhophi(int n, const double* __restrict__ phiAlpha, const double* __restrict__ phi, double rho1, double rho2,
                                double* __restrict__ rhoPhi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rhoPhi[i] = phiAlpha[i] * (rho1 - rho2) + phi[i] * rho2;
}
// gh = g & C ; ghf = g & Cf [REF region3d/create3dFields.H:188-206]
static __global__ void k_gh(MeshView mv, const double* __restrict__ Cfx, const double* __restrict__ Cfy, const double* __restrict__ Cfz,
                            double g0, double g1, double g2, double* __restrict__ gh, double* __restrict__ ghf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mv.N) gh[i] = g0 * mv.Cx[i] + g1 * mv.Cy[i] + g2 * mv.Cz[i];
    if (i < mv.Fs + mv.B) ghf[i] = g0 * Cfx[i] + g1 * Cfy[i] + g2 * Cfz[i];
}
// nHatf = (interpolate(grad(alpha1))/(|.| + deltaN)) & Sf [UPSTREAM interfaceProperties::calculateK]
static __global__ void k_nhatf(MeshView mv, const double* __restrict__ gx, const double* __restrict__ gy, const double* __restrict__ gz,
                               const double* __restrict__ gb, double deltaN, double* __restrict__ nHatf) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double ax = gx[row], ay = gy[row], az = gz[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double w = mv.w[pos];
            double fx = w * (ax - gx[u]) + gx[u], fy = w * (ay - gy[u]) + gy[u], fz = w * (az - gz[u]) + gz[u];
            double den = sqrt(fx * fx + fy * fy + fz * fz) + deltaN;
            nHatf[pos] = (fx / den) * mv.Sfx[pos] + (fy / den) * mv.Sfy[pos] + (fz / den) * mv.Sfz[pos];
        END_FACES
    }
    if (row < mv.B) {
        int p = mv.Fs + row, B = mv.B;
        int c = mv.bFaceRow[row];
        double fx = bndFaceValue(mv, row, gx[c], gb[row]), fy = bndFaceValue(mv, row, gy[c], gb[B + row]), fz = bndFaceValue(mv, row, gz[c], gb[2 * B + row]);
        double den = sqrt(fx * fx + fy * fy + fz * fz) + deltaN;
        nHatf[p] = (fx / den) * mv.Sfx[p] + (fy / den) * mv.Sfy[p] + (fz / den) * mv.Sfz[p];
    }
}
// muf = alpha1f*rho1*nu1 + (1-alpha1f)*rho2*nu2, alpha1f = clamp(interpolate(alpha1)) [UPSTREAM twoPhaseMixture::muf]
static __global__ void k_muf(MeshView mv, const double* __restrict__ alpha, const double* __restrict__ alphab, double rho1, double rho2,
                             double nu1, double nu2, double* __restrict__ muf) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double a = alpha[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double n = alpha[u];
            double af = fmin(fmax(mv.w[pos] * (a - n) + n, 0.0), 1.0);
            muf[pos] = af * rho1 * nu1 + (1.0 - af) * rho2 * nu2;
        END_FACES
    }
    if (row < mv.B) { double af = fmin(fmax(bndFaceValue(mv, row, alpha[mv.bFaceRow[row]], alphab[row]), 0.0), 1.0); muf[mv.Fs + row] = af * rho1 * nu1 + (1.0 - af) * rho2 * nu2; }
}
// UEqn = ddt(rho,U) + div(rhoPhi,U) - laplacian(muEff,U) - (grad(U) & grad(muEff))  [REF region3d/UEqn.H:8-15]
// one row pass: diag, final lower/upper of the owner faces, source (incl. the explicit gradient term)
static __global__ void __launch_bounds__(BLOCK) k_ueqn_row(MeshView mv, int upwind, double rDeltaT, const double* __restrict__ rho,
        const double* __restrict__ rho0, const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
        const double* __restrict__ Ub, const double* __restrict__ U0x, const double* __restrict__ U0y, const double* __restrict__ U0z,
        const double* __restrict__ rhoPhi, const double* __restrict__ muf, double* __restrict__ uDiag, double* __restrict__ uLower,
        double* __restrict__ uUpper, double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    double V = mv.V[row];
    double ux = Ux[row], uy = Uy[row], uz = Uz[row];
    double divDiag = 0.0, lapDiag = 0.0;
    double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gm[3] = {0, 0, 0};
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double flux = rhoPhi[pos], w = mv.w[pos];
        double wc = upwind ? (flux >= 0 ? 1.0 : 0.0) : w;
        double lo0 = -wc * flux, up0 = lo0 + flux;
        double lapU = mv.dc[pos] * (muf[pos] * mv.magSf[pos]);
        divDiag = divDiag - up0; lapDiag = lapDiag - lapU;
        double S[3] = {mv.Sfx[pos], mv.Sfy[pos], mv.Sfz[pos]};
        double uf[3] = {w * (Ux[l] - ux) + ux, w * (Uy[l] - uy) + uy, w * (Uz[l] - uz) + uz};
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) T[3 * i + j] = T[3 * i + j] - S[i] * uf[j];
            gm[i] = gm[i] - S[i] * muf[pos];
        }
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double flux = rhoPhi[pos], w = mv.w[pos];
        double wc = upwind ? (flux >= 0 ? 1.0 : 0.0) : w;
        double lo0 = -wc * flux, up0 = lo0 + flux;
        double lapU = mv.dc[pos] * (muf[pos] * mv.magSf[pos]);
        divDiag = divDiag - lo0; lapDiag = lapDiag - lapU;
        uLower[pos] = lo0 - lapU; uUpper[pos] = up0 - lapU;
        double S[3] = {mv.Sfx[pos], mv.Sfy[pos], mv.Sfz[pos]};
        double uf[3] = {w * (ux - Ux[u]) + Ux[u], w * (uy - Uy[u]) + Uy[u], w * (uz - Uz[u]) + Uz[u]};
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) T[3 * i + j] = T[3 * i + j] + S[i] * uf[j];
            gm[i] = gm[i] + S[i] * muf[pos];
        }
    END_FACES
    FOR_BND_FACES(mv, s, lane, b)
        int p = mv.Fs + b, B = mv.B;
        double S[3] = {mv.Sfx[p], mv.Sfy[p], mv.Sfz[p]};
        double ub[3] = {bndFaceValue(mv, b, ux, Ub[b]), bndFaceValue(mv, b, uy, Ub[B + b]), bndFaceValue(mv, b, uz, Ub[2 * B + b])};
#pragma unroll
        for (int i = 0; i < 3; i++) {
#pragma unroll
            for (int j = 0; j < 3; j++) T[3 * i + j] = T[3 * i + j] + S[i] * ub[j];
            gm[i] = gm[i] + S[i] * muf[p];
        }
    END_BND
#pragma unroll
    for (int q = 0; q < 9; q++) T[q] = T[q] / V;
    double g0 = gm[0] / V, g1 = gm[1] / V, g2 = gm[2] / V;
    double ddtDiag = rDeltaT * rho[row] * V;
    uDiag[row] = (ddtDiag + divDiag) - lapDiag;
    double r0 = rDeltaT * rho0[row];
    double s0 = r0 * U0x[row] * V, s1 = r0 * U0y[row] * V, s2 = r0 * U0z[row] * V;
    s0 += V * (T[0] * g0 + T[1] * g1 + T[2] * g2);
    s1 += V * (T[3] * g0 + T[4] * g1 + T[5] * g2);
    s2 += V * (T[6] * g0 + T[7] * g1 + T[8] * g2);
    sx[row] = s0; sy[row] = s1; sz[row] = s2;
}
// boundary coefficients of UEqn: internalCoeffs = div - lap, boundaryCoeffs = div - lap  (per component, SoA [3B])
static __global__ void k_ueqn_boundary(MeshView mv, BCView bc, int upwind, const double* __restrict__ rhoPhi, const double* __restrict__ muf,
                                       const double* __restrict__ Ub, double* __restrict__ uIC, double* __restrict__ uBC) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B) return;
    int p = mv.Fs + b, t = bc.type[mv.bPatch[b]], B = mv.B;
    double pf = rhoPhi[p], pg = muf[p] * mv.magSf[p], dc = mv.dc[p];
    if (t == B200_BC_PROCESSOR) {   // coupledFvPatchField: value coeffs (w, 1-w), gradient coeffs (-dc, +dc)
        double w = upwind ? (pf >= 0 ? 1.0 : 0.0) : mv.w[p];   // the convection scheme's own weights on the coupled patch
        for (int k = 0; k < 3; k++) {
            int idx = k * B + b;
            double dIC = pf * (1.0 * w), dBC = -pf * (1.0 * (1.0 - w));
            double lIC = pg * (-1.0 * dc), lBC = -pg * dc;
            uIC[idx] = dIC - lIC; uBC[idx] = dBC - lBC;
        }
        return;
    }
    for (int k = 0; k < 3; k++) {
        int idx = k * B + b;
        double dIC = pf * bcValueIntCoeff(bc, t, b);
        double dBC = -pf * bcValueBouCoeff(bc, t, b, idx, dc, Ub[idx]);
        double lIC = pg * bcGradIntCoeff(bc, t, b, dc);
        double lBC = -pg * bcGradBouCoeff(bc, t, b, idx, dc, Ub[idx]);
        uIC[idx] = dIC - lIC;
        uBC[idx] = dBC - lBC;
    }
}
// momentum predictor [REF region3d/UEqn.H:23-31]: out = (interpolate(sigmaK)*snGrad(alpha1) - ghf*snGrad(rho) - snGrad(pd))*magSf
static __global__ void __launch_bounds__(BLOCK) k_mom_flux(MeshView mv, BCView bcA, BCView bcP, double sigma, const double* __restrict__ K,
        const double* __restrict__ Kb, const double* __restrict__ alpha, const double* __restrict__ alphab, const double* __restrict__ rho,
        const double* __restrict__ rhob, const double* __restrict__ pd, const double* __restrict__ pdb, const double* __restrict__ ghf,
        const double* __restrict__ corrA, const double* __restrict__ corrR, const double* __restrict__ corrP, double* __restrict__ out) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double sKl = sigma * K[row], al = alpha[row], rl = rho[row], pl = pd[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double sKu = sigma * K[u], dc = mv.dc[pos];
            double sKf = mv.w[pos] * (sKl - sKu) + sKu;
            double sga = dc * (alpha[u] - al), sgr = dc * (rho[u] - rl), sgp = dc * (pd[u] - pl);
            if (corrA) { sga = sga + corrA[pos]; sgr = sgr + corrR[pos]; sgp = sgp + corrP[pos]; }
            out[pos] = ((sKf * sga - ghf[pos] * sgr) - sgp) * mv.magSf[pos];
        END_FACES
    }
    if (row < mv.B) {
        int b = row, p = mv.Fs + b, c = mv.bFaceRow[b], t = bcA.type[mv.bPatch[b]], tp = bcP.type[mv.bPatch[b]];
        double dc = mv.dc[p];
        double sKf = sigma * K[c];
        if (mv.bCoupled[b]) { double sKn = sigma * Kb[b]; sKf = mv.w[p] * (sKf - sKn) + sKn; }
        double sga = bcSnGrad(bcA, t, b, b, dc, alpha[c], alphab[b]);
        double sgr = bcSnGrad(bcA, t, b, b, dc, rho[c], rhob[b]);
        double sgp = bcSnGrad(bcP, tp, b, b, dc, pd[c], pdb[b]);
        if (corrA) { sga = sga + corrA[p]; sgr = sgr + corrR[p]; sgp = sgp + corrP[p]; }
        out[p] = ((sKf * sga - ghf[p] * sgr) - sgp) * mv.magSf[p];
    }
}
// component k of the system fvMatrix<vector>::solve hands to the solver [UPSTREAM fvMatrixSolve.C]:
//   diag = UEqn.diag + internalCoeffs.component(k) (addBoundaryDiag) ; source = (UEqn.source + V*R).component(k) + boundaryCoeffs (non-coupled)
static __global__ void __launch_bounds__(BLOCK) k_mom_system(MeshView mv, int k, const double* __restrict__ uDiag, const double* __restrict__ uS,
        const double* __restrict__ Rk, const double* __restrict__ uIC, const double* __restrict__ uBC, double* __restrict__ diagK, double* __restrict__ srcK) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    int B = mv.B;
    double d = uDiag[row], sc = uS[row] + mv.V[row] * Rk[row];
    FOR_BND_FACES(mv, s, lane, b)
        d = d + uIC[k * B + b];
        if (!mv.bCoupled[b]) sc = sc + uBC[k * B + b];
    END_BND
    diagK[row] = d; srcK[row] = sc;
}
// rUA = 1/UEqn.A() ; Unew = rUA*UEqn.H()  [REF region3d/pEqn.H:2,5] [UPSTREAM fvMatrix::A(), H(); lduMatrix::H]
static __global__ void __launch_bounds__(BLOCK) k_ua_h(MeshView mv, const double* __restrict__ uDiag, const double* __restrict__ uLower,
        const double* __restrict__ uUpper, const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
        const double* __restrict__ uIC, const double* __restrict__ uBC, const double* __restrict__ Ub, const double* __restrict__ Ux,
        const double* __restrict__ Uy, const double* __restrict__ Uz, double* __restrict__ rUA, double* __restrict__ Hx, double* __restrict__ Hy, double* __restrict__ Hz) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    int B = mv.B;
    double V = mv.V[row];
    double a = uDiag[row];
    double bd[3] = {0, 0, 0}, av = 0.0;
    FOR_BND_FACES(mv, s, lane, b)
        double i0 = uIC[b], i1 = uIC[B + b], i2 = uIC[2 * B + b];
        a = a + (i0 + i1 + i2) / 3.0;
        bd[0] = bd[0] + i0; bd[1] = bd[1] + i1; bd[2] = bd[2] + i2;
    END_BND
    (void)av;
    bd[0] = -bd[0]; bd[1] = -bd[1]; bd[2] = -bd[2];
    FOR_BND_FACES(mv, s, lane, b)
        double c = (uIC[b] + uIC[B + b] + uIC[2 * B + b]) / 3.0;
        bd[0] = bd[0] + c; bd[1] = bd[1] + c; bd[2] = bd[2] + c;
    END_BND
    double A = a / V;
    double ra = 1.0 / A;
    double h0 = bd[0] * Ux[row], h1 = bd[1] * Uy[row], h2 = bd[2] * Uz[row];
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double c = uLower[pos];
        l0 = l0 - c * Ux[l]; l1 = l1 - c * Uy[l]; l2 = l2 - c * Uz[l];
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double c = uUpper[pos];
        l0 = l0 - c * Ux[u]; l1 = l1 - c * Uy[u]; l2 = l2 - c * Uz[u];
    END_FACES
    h0 = h0 + (l0 + sx[row]); h1 = h1 + (l1 + sy[row]); h2 = h2 + (l2 + sz[row]);
    FOR_BND_FACES(mv, s, lane, b)
        if (!mv.bCoupled[b]) { h0 = h0 + uBC[b]; h1 = h1 + uBC[B + b]; h2 = h2 + uBC[2 * B + b]; }
        else { h0 = h0 + uBC[b] * Ub[b]; h1 = h1 + uBC[B + b] * Ub[B + b]; h2 = h2 + uBC[2 * B + b] * Ub[2 * B + b]; }   // boundaryCoeffs*patchNeighbourField
    END_BND
    h0 = h0 / V; h1 = h1 / V; h2 = h2 / V;
    rUA[row] = ra;
    Hx[row] = ra * h0; Hy[row] = ra * h1; Hz[row] = ra * h2;
}
// U = rUA*H assignment on the boundary: fixedValue patches keep their value, all others take the product (= U of the cell)
static __global__ void k_ub_assign(MeshView mv, BCView bc, const double* __restrict__ Ux, const double* __restrict__ Uy,
                                   const double* __restrict__ Uz, double* __restrict__ Ub) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B) return;
    int t = bc.type[mv.bPatch[b]];
    if (t == B200_BC_FIXED_VALUE || t == B200_BC_PROCESSOR) return;   // processor faces: halo swap of the new U
    int c = mv.bFaceRow[b], B = mv.B;
    Ub[b] = Ux[c]; Ub[B + b] = Uy[c]; Ub[2 * B + b] = Uz[c];
}
// phiU = (interpolate(U) & Sf) + ddtPhiCorr(rUA, rho, U, phi) ; rUAf = interpolate(rUA)  [REF region3d/pEqn.H:3,7-12]
static __global__ void __launch_bounds__(BLOCK) k_phiU(MeshView mv, BCView bcU, double rDeltaT, const double* __restrict__ rUA,
        const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz, const double* __restrict__ Ub,
        const double* __restrict__ U0x, const double* __restrict__ U0y, const double* __restrict__ U0z, const double* __restrict__ U0b,
        const double* __restrict__ rho0, const double* __restrict__ rho0b, const double* __restrict__ phi0, const double* __restrict__ rUAb,
        double* __restrict__ phiU, double* __restrict__ rUAf) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double rl = rUA[row], rArl = rl * rho0[row];
        double ux = Ux[row], uy = Uy[row], uz = Uz[row], ox = U0x[row], oy = U0y[row], oz = U0z[row];
        double rx = rArl * ox, ry = rArl * oy, rz = rArl * oz;
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double w = mv.w[pos], ru = rUA[u], rAru = ru * rho0[u];
            double S0 = mv.Sfx[pos], S1 = mv.Sfy[pos], S2 = mv.Sfz[pos];
            rUAf[pos] = w * (rl - ru) + ru;
            double fx = w * (ux - Ux[u]) + Ux[u], fy = w * (uy - Uy[u]) + Uy[u], fz = w * (uz - Uz[u]) + Uz[u];
            double rArf = w * (rArl - rAru) + rAru;
            double qx = rAru * U0x[u], qy = rAru * U0y[u], qz = rAru * U0z[u];
            double gx = w * (rx - qx) + qx, gy = w * (ry - qy) + qy, gz = w * (rz - qz) + qz;
            double px = w * (ox - U0x[u]) + U0x[u], py = w * (oy - U0y[u]) + U0y[u], pz = w * (oz - U0z[u]) + U0z[u];
            double p0 = phi0[pos];
            double phiCorr = p0 - (px * S0 + py * S1 + pz * S2);
            double coeff = 1.0 - fmin(fabs(phiCorr) / (fabs(p0) + FV_SMALL), 1.0);
            double ddt = coeff * rDeltaT * (rArf * p0 - (gx * S0 + gy * S1 + gz * S2));
            phiU[pos] = (fx * S0 + fy * S1 + fz * S2) + ddt;
        END_FACES
    }
    if (row < mv.B) {
        int b = row, p = mv.Fs + b, c = mv.bFaceRow[b], B = mv.B;
        double S0 = mv.Sfx[p], S1 = mv.Sfy[p], S2 = mv.Sfz[p];
        if (mv.bCoupled[b]) {   // processor face: the internal-face formula with the neighbour rank's cell as N
            double w = mv.w[p], rl = rUA[c], ru = rUAb[b], rArl = rl * rho0[c], rAru = ru * rho0b[b];
            rUAf[p] = w * (rl - ru) + ru;
            double fx = w * (Ux[c] - Ub[b]) + Ub[b], fy = w * (Uy[c] - Ub[B + b]) + Ub[B + b], fz = w * (Uz[c] - Ub[2 * B + b]) + Ub[2 * B + b];
            double rArf = w * (rArl - rAru) + rAru;
            double lx = rArl * U0x[c], ly = rArl * U0y[c], lz = rArl * U0z[c];
            double qx = rAru * U0b[b], qy = rAru * U0b[B + b], qz = rAru * U0b[2 * B + b];
            double gx = w * (lx - qx) + qx, gy = w * (ly - qy) + qy, gz = w * (lz - qz) + qz;
            double px = w * (U0x[c] - U0b[b]) + U0b[b], py = w * (U0y[c] - U0b[B + b]) + U0b[B + b], pz = w * (U0z[c] - U0b[2 * B + b]) + U0b[2 * B + b];
            double p0 = phi0[p];
            double phiCorr = p0 - (px * S0 + py * S1 + pz * S2);
            double coeff = 1.0 - fmin(fabs(phiCorr) / (fabs(p0) + FV_SMALL), 1.0);
            double ddt = coeff * rDeltaT * (rArf * p0 - (gx * S0 + gy * S1 + gz * S2));
            phiU[p] = (fx * S0 + fy * S1 + fz * S2) + ddt;
            return;
        }
        double rb = rUA[c];
        rUAf[p] = rb;
        double rArb = rb * rho0b[b];
        double gx = rArb * U0b[b], gy = rArb * U0b[B + b], gz = rArb * U0b[2 * B + b];
        double p0 = phi0[p];
        double phiCorr = p0 - (U0b[b] * S0 + U0b[B + b] * S1 + U0b[2 * B + b] * S2);
        double coeff = 1.0 - fmin(fabs(phiCorr) / (fabs(p0) + FV_SMALL), 1.0);
        if (bcFixes(bcU.type[mv.bPatch[b]])) coeff = 0.0;
        double ddt = coeff * rDeltaT * (rArb * p0 - (gx * S0 + gy * S1 + gz * S2));
        phiU[p] = (Ub[b] * S0 + Ub[B + b] * S1 + Ub[2 * B + b] * S2) + ddt;
    }
}
// adjustPhi sums [UPSTREAM adjustPhi.C]: out = {massIn, fixedMassOut, adjustableMassOut} over non-coupled boundary faces
static __global__ void k_adjust_bsums(MeshView mv, BCView bcU, const double* __restrict__ phiU, double* out, double* partials, unsigned* counter) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    if (b < mv.B && !mv.bCoupled[b]) {
        double f = phiU[mv.Fs + b];
        if (f < 0.0) v[0] = -f;
        else if (bcFixes(bcU.type[mv.bPatch[b]])) v[1] = f;
        else v[2] = f;
    }
    gridReduce<3, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<3, OpSum>(gridDim.x, partials, counter, out, out + 1, out + 2);
}
static __global__ void k_abs_sum_faces(MeshView mv, const double* __restrict__ f, double* out, double* partials, unsigned* counter) {
    ROW_PROLOGUE(mv)
    double v[1] = {0.0};
    if (row < mv.N) { FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; v[0] = v[0] + fabs(f[pos]); END_FACES }
    gridReduce<1, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpSum>(gridDim.x, partials, counter, out);
}
static __global__ void k_adjust_scale(MeshView mv, BCView bcU, const double* __restrict__ sums /*massIn,fixedOut,adjOut,absSum*/,
                                      double* __restrict__ phiU) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B || mv.bCoupled[b] || bcFixes(bcU.type[mv.bPatch[b]])) return;
    double massIn = sums[0], fixedMassOut = sums[1], adjustableMassOut = sums[2], totalFlux = FV_VSMALL + sums[3];
    double magAdj = fabs(adjustableMassOut), massCorr = 1.0;
    if (magAdj > FV_VSMALL && magAdj / totalFlux > FV_SMALL) massCorr = (massIn - fixedMassOut) / adjustableMassOut;
    double f = phiU[mv.Fs + b];
    if (f > 0.0) phiU[mv.Fs + b] = f * massCorr;
}
// phi = phiU + (interpolate(sigmaK)*snGrad(alpha1) - ghf*snGrad(rho))*rUAf*magSf  [REF region3d/pEqn.H:16-20]
static __global__ void __launch_bounds__(BLOCK) k_phi(MeshView mv, BCView bcA, double sigma, const double* __restrict__ K, const double* __restrict__ Kb,
        const double* __restrict__ alpha, const double* __restrict__ alphab, const double* __restrict__ rho, const double* __restrict__ rhob,
        const double* __restrict__ ghf, const double* __restrict__ rUAf, const double* __restrict__ phiU, double* __restrict__ phi,
        const double* __restrict__ corrA, const double* __restrict__ corrR /*corrected snGrad: explicit non-orthogonal parts, or null*/) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double sKl = sigma * K[row], al = alpha[row], rl = rho[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double sKu = sigma * K[u], dc = mv.dc[pos];
            double sKf = mv.w[pos] * (sKl - sKu) + sKu;
            double sga = dc * (alpha[u] - al), sgr = dc * (rho[u] - rl);
            if (corrA) { sga = sga + corrA[pos]; sgr = sgr + corrR[pos]; }
            phi[pos] = phiU[pos] + (sKf * sga - ghf[pos] * sgr) * rUAf[pos] * mv.magSf[pos];
        END_FACES
    }
    if (row < mv.B) {
        int b = row, p = mv.Fs + b, c = mv.bFaceRow[b], t = bcA.type[mv.bPatch[b]];
        double dc = mv.dc[p];
        double sKf = sigma * K[c];
        if (mv.bCoupled[b]) { double sKn = sigma * Kb[b]; sKf = mv.w[p] * (sKf - sKn) + sKn; }
        double sga = bcSnGrad(bcA, t, b, b, dc, alpha[c], alphab[b]);
        double sgr = bcSnGrad(bcA, t, b, b, dc, rho[c], rhob[b]);
        if (corrA) { sga = sga + corrA[p]; sgr = sgr + corrR[p]; }
        phi[p] = phiU[p] + (sKf * sga - ghf[p] * sgr) * rUAf[p] * mv.magSf[p];
    }
}
// continuityErrs.H: sums of V*|div phi|, V*div phi, V
static __global__ void k_cont_err(MeshView mv, const double* __restrict__ phi, double* out, double* partials, unsigned* counter) {
    ROW_PROLOGUE(mv)
    double v[3] = {0.0, 0.0, 0.0};
    if (row < mv.N) {
        double a = 0.0;
        FOR_NBR_FACES(mv, s, lane, pos, l) (void)l; a = a - phi[pos]; END_FACES
        FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; a = a + phi[pos]; END_FACES
        FOR_BND_FACES(mv, s, lane, b) a = a + phi[mv.Fs + b]; END_BND
        double V = mv.V[row], d = a / V;
        v[0] = V * fabs(d); v[1] = V * d; v[2] = V;
    }
    gridReduce<3, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<3, OpSum>(gridDim.x, partials, counter, out, out + 1, out + 2);
}
// p = pd + rho*gh [REF region3d/solve3d.H:56]
static __global__ void k_p(int N, const double* __restrict__ pd, const double* __restrict__ rho, const double* __restrict__ gh, double* __restrict__ p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) p[i] = pd[i] + rho[i] * gh[i];
}

// ================================================================================================ Region3D
Region3D::~Region3D() { if (hscal) cudaFreeHost(hscal); }

bool Region3D::coupled() const { return rt().nRanks > 1 && meshHasProcessorPatches(*m); }
void Region3D::haloScalar(const double* cellField, double* bdst) {
    if (!coupled()) return;
    haloPack(*m, cellField, haloSend.p);
    haloExchange(*m, haloSend.p, bdst);
}
void Region3D::evalPdBC(const DevBC& bc, const double* psi, double* psib) {
    if (m->B) B200_LAUNCH(k_bc_evaluate, gridFor(m->B), BLOCK, rt().stream, m->view(), bc.view(), 1, psi, (const double*)nullptr, (const double*)nullptr, psib);
    haloScalar(psi, psib);
}
void Region3D::evalAlphaBC() {
    if (m->B) B200_LAUNCH(k_bc_evaluate, gridFor(m->B), BLOCK, rt().stream, m->view(), bcAlpha.view(), 1, (const double*)alpha.p,
                          (const double*)nullptr, (const double*)nullptr, alphab.p);
    haloScalar(alpha, alphab);
}
void Region3D::evalUBC() {
    if (m->B) B200_LAUNCH(k_bc_evaluate, gridFor(m->B), BLOCK, rt().stream, m->view(), bcU.view(), 3, (const double*)Ux.p, (const double*)Uy.p,
                          (const double*)Uz.p, Ub.p);
    haloScalar(Ux, Ub.p); haloScalar(Uy, Ub.p + m->B); haloScalar(Uz, Ub.p + 2 * (size_t)m->B);
}

void Region3D::create(DeviceMesh* mesh, const b200_region3d_controls& c, const b200_bc* a, const b200_bc* u, const b200_bc* pbc,
                      const double* alpha1, const double* U, const double* pdHost) {
    m = mesh; ctl = c;
    B200_REQUIRE(m->hasGeometry, B200_ERR_ARG, "b200_region3d_create needs a mesh with geometry");
    B200_REQUIRE(c.nCorr >= 1 && c.nAlphaCorr >= 1 && c.nAlphaSubCycles >= 1 && c.deltaT > 0, B200_ERR_ARG, "bad PISO controls");
    size_t N = m->N, B = m->B, nf = m->nFaceField();
    bcAlpha.upload(*m, a, 1); bcU.upload(*m, u, 3); bcPd.upload(*m, pbc, 1);
    std::vector<int> tp(m->patches.size());
    for (size_t p = 0; p < tp.size(); p++)
        tp[p] = bcPd.fixes((int)p) ? B200_BC_FIXED_VALUE : (bcPd.types[p] == B200_BC_PROCESSOR ? B200_BC_PROCESSOR : B200_BC_ZERO_GRADIENT);
    bcPcorr.setTypes(*m, tp, 1);
    for (DevBuf<double>* b : {&alpha, &alpha0, &alphaStart, &pd, &rho, &rho0, &K, &rUA, &p, &gh, &uDiag, &pdDiag, &pdSource, &pcorr,
                              &Ux, &Uy, &Uz, &U0x, &U0y, &U0z, &Hx, &Hy, &Hz, &uSx, &uSy, &uSz}) { b->alloc(N); b->zero(); }
    for (DevBuf<double>* b : {&alphab, &pdb, &rhob, &rho0b, &ic, &bcv, &pcorrb, &Kb, &rUAb, &haloSend}) { b->alloc(B); b->zero(); }
    for (DevBuf<double>* b : {&Ub, &U0b, &uIC, &uBC}) { b->alloc(3 * B); b->zero(); }
    for (DevBuf<double>* b : {&phi, &phi0, &rhoPhi, &rhoPhiSum, &nHatf, &ghf, &phiAlpha, &phiU, &rUAf, &muf, &uLower, &uUpper, &pdUpper, &oneF}) { b->alloc(nf); b->zero(); }
    if (nonOrth()) for (DevBuf<double>* b : {&ffc, &corrA, &corrR}) { b->alloc(nf); b->zero(); }
    scal.alloc(64); scal.zero();
    B200_CHECK_CUDA(cudaMallocHost((void**)&hscal, 64 * sizeof(double)));
    fillValue(oneF, nf, 1.0);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    m->uploadCells(alpha1, alpha); m->uploadCellVec(U, Ux, Uy, Uz); m->uploadCells(pdHost, pd);
    evalAlphaBC(); evalUBC();
    evalPdBC(bcPd, pd, pdb);
    int nmax = (int)std::max<size_t>(std::max(N, B), 1);
    B200_LAUNCH(k_init_phi, gridFor(nmax), BLOCK, st, mv, (const double*)Ux.p, (const double*)Uy.p, (const double*)Uz.p, (const double*)Ub.p, phi.p);
    B200_LAUNCH(k_rho, gridFor(nmax), BLOCK, st, (int)N, (int)B, (const double*)alpha.p, (const double*)alphab.p, c.rho1, c.rho2, rho.p, rhob.p);
    B200_LAUNCH(k_scale, gridFor((long long)nf), BLOCK, st, (int)nf, c.rho1, (const double*)phi.p, rhoPhi.p, 0);
    B200_LAUNCH(k_gh, gridFor((long long)std::max(nf, N)), BLOCK, st, mv, (const double*)m->dCfx.p, (const double*)m->dCfy.p, (const double*)m->dCfz.p,
                c.g[0], c.g[1], c.g[2], gh.p, ghf.p);
    reconTensor(*m);
    double sV = 0;
    for (size_t i = 0; i < N; i++) sV += m->hostV[i];
    double nGlobal = (double)N;
    if (rt().nRanks > 1) {   // average(mesh.V()) is a global average (gAverage)
        double h2[2] = {sV, nGlobal};
        B200_CHECK_CUDA(cudaMemcpyAsync(scal.p + 40, h2, 16, cudaMemcpyHostToDevice, st));
        allReduce(scal.p + 40, 2, RED_SUM);
        B200_CHECK_CUDA(cudaMemcpyAsync(h2, scal.p + 40, 16, cudaMemcpyDeviceToHost, st));
        syncStream();
        sV = h2[0]; nGlobal = h2[1];
    }
    deltaN = 1e-8 / std::pow(sV / nGlobal, 1.0 / 3.0);
    // pd.needReference(): true when no patch fixes the value on any rank; the reference cell label is local (every rank that holds a
    // cell with that label applies it, which is what setRefCell's label form does under decomposition) [REF create3dFields.H:227-233]
    {
        double anyFixes = 0.0;
        for (size_t p = 0; p < m->patches.size(); p++) if (bcPd.fixes((int)p) && m->patches[p].size > 0) anyFixes = 1.0;
        if (rt().nRanks > 1) {
            B200_CHECK_CUDA(cudaMemcpyAsync(scal.p + 42, &anyFixes, 8, cudaMemcpyHostToDevice, st));
            allReduce(scal.p + 42, 1, RED_MAX);
            B200_CHECK_CUDA(cudaMemcpyAsync(&anyFixes, scal.p + 42, 8, cudaMemcpyDeviceToHost, st));
            syncStream();
        }
        pdNeedRef = anyFixes == 0.0;
        refRow = (c.pdRefCell >= 0 && c.pdRefCell < m->N) ? m->cellRow[c.pdRefCell] : -1;
    }
    calculateK();   // interfaceProperties constructor [REF region3d/create3dFields.H:254-263]
    syncStream();
}

void Region3D::calculateK() {
    FvWorkspace& w = fvWorkspace(*m);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    ScopedTimer t(T_FLUX);
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    B200_LAUNCH(k_grad_scalar, gridFor(N), BLOCK, st, mv, (const double*)alpha.p, (const double*)alphab.p, w.gx.p, w.gy.p, w.gz.p);
    if (B) B200_LAUNCH(k_grad_boundary, gridFor(B), BLOCK, st, mv, bcAlpha.view(), (const double*)alpha.p, (const double*)alphab.p,
                       (const double*)w.gx.p, (const double*)w.gy.p, (const double*)w.gz.p, w.gb.p);
    haloScalar(w.gx, w.gb.p); haloScalar(w.gy, w.gb.p + B); haloScalar(w.gz, w.gb.p + 2 * (size_t)B);
    B200_LAUNCH(k_nhatf, gridFor(nmax), BLOCK, st, mv, (const double*)w.gx.p, (const double*)w.gy.p, (const double*)w.gz.p, (const double*)w.gb.p, deltaN, nHatf.p);
    B200_LAUNCH(k_surface_integrate, gridFor(N), BLOCK, st, mv, (const double*)nHatf.p, K.p, -1.0);
    haloScalar(K, Kb);
}

// [REF region3d/alphaEqn.H:1-41]
void Region3D::alphaEqn(double dt) {
    FvWorkspace& w = fvWorkspace(*m);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    size_t nf = m->nFaceField();
    { ScopedTimer t(T_FLUX);
      B200_LAUNCH(k_phic_max, gridFor(nmax), BLOCK, st, mv, (const double*)phi.p, scal.p, w.partials.p, w.counter.p); }
    if (rt().nRanks > 1) allReduce(scal.p, 1, RED_MAX);
    for (int aCorr = 0; aCorr < ctl.nAlphaCorr; aCorr++) {
        { ScopedTimer t(T_FLUX);
          B200_LAUNCH(k_grad_scalar, gridFor(N), BLOCK, st, mv, (const double*)alpha.p, (const double*)alphab.p, w.gx.p, w.gy.p, w.gz.p);
          haloScalar(w.gx, w.gb.p); haloScalar(w.gy, w.gb.p + B); haloScalar(w.gz, w.gb.p + 2 * (size_t)B);
          B200_LAUNCH(k_alpha_flux, gridFor(nmax), BLOCK, st, mv, (const double*)phi.p, (const double*)alpha.p, (const double*)alphab.p,
                      (const double*)w.gx.p, (const double*)w.gy.p, (const double*)w.gz.p, (const double*)w.gb.p, (const double*)nHatf.p, ctl.cAlpha, (const double*)scal.p, phiAlpha.p); }
        evalAlphaBC();
        if (capture) {
            capMules.emplace_back();
            CapturedMules& c = capMules.back();
            c.psi.resize(N); c.psib.resize(B); c.psi0.resize(N); c.phi.resize(m->F + B); c.phiPsi.resize(m->F + B); c.deltaT = dt;
            m->downloadCells(alpha, c.psi.data()); if (B) m->downloadBoundary(alphab, c.psib.data()); m->downloadCells(alpha0, c.psi0.data());
            m->downloadFaces(phi, c.phi.data(), true); m->downloadFaces(phiAlpha, c.phiPsi.data(), true);
            syncStream();
        }
        mulesExplicitDevice(*m, bcAlpha.view(), alpha, alphab, alpha0, phi, phiAlpha, dt, 1.0, 0.0, ctl.mulesIncludeOwn, 3, nullptr);
        evalAlphaBC();
        B200_LAUNCH(k_rhophi, gridFor((long long)nf), BLOCK, st, (int)nf, (const double*)phiAlpha.p, (const double*)phi.p, ctl.rho1, ctl.rho2, rhoPhi.p);
    }
    // Info: weightedAverage / min / max [REF region3d/alphaEqn.H:36-40]
    int slot = 16 + 4 * (int)(alphaStats.size() / 3 % 8);
    B200_LAUNCH(k_alpha_stats, gridFor(nmax), BLOCK, st, mv, (const double*)alpha.p, (const double*)alphab.p, scal.p + slot, w.partials.p, w.counter.p);
    if (rt().nRanks > 1) { allReduce(scal.p + slot, 2, RED_SUM); allReduce(scal.p + slot + 2, 1, RED_MIN); allReduce(scal.p + slot + 3, 1, RED_MAX); }
    B200_CHECK_CUDA(cudaMemcpyAsync(hscal + slot, scal.p + slot, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    alphaStats.push_back((double)slot); alphaStats.push_back(0); alphaStats.push_back(0);  // resolved after the next sync
    pendingStats.push_back(alphaStats.size() - 3);
}

void Region3D::resolveStats() {
    for (size_t i : pendingStats) {
        int slot = (int)alphaStats[i];
        alphaStats[i] = hscal[slot] / hscal[slot + 1]; alphaStats[i + 1] = hscal[slot + 2]; alphaStats[i + 2] = hscal[slot + 3];
    }
    pendingStats.clear();
}

// [REF region3d/UEqn.H:1-17]
void Region3D::assembleUEqn() {
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    ScopedTimer t(T_UEQN);
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    double rDeltaT = 1.0 / ctl.deltaT;
    B200_LAUNCH(k_muf, gridFor(nmax), BLOCK, st, mv, (const double*)alpha.p, (const double*)alphab.p, ctl.rho1, ctl.rho2, ctl.nu1, ctl.nu2, muf.p);
    B200_LAUNCH(k_ueqn_row, gridFor(N), BLOCK, st, mv, ctl.divUScheme == B200_SCHEME_UPWIND ? 1 : 0, rDeltaT, (const double*)rho.p, (const double*)rho0.p,
                (const double*)Ux.p, (const double*)Uy.p, (const double*)Uz.p, (const double*)Ub.p, (const double*)U0x.p, (const double*)U0y.p,
                (const double*)U0z.p, (const double*)rhoPhi.p, (const double*)muf.p, uDiag.p, uLower.p, uUpper.p, uSx.p, uSy.p, uSz.p);
    if (B) B200_LAUNCH(k_ueqn_boundary, gridFor(B), BLOCK, st, mv, bcU.view(), ctl.divUScheme == B200_SCHEME_UPWIND ? 1 : 0, (const double*)rhoPhi.p, (const double*)muf.p, (const double*)Ub.p, uIC.p, uBC.p);
}

// [UPSTREAM correctedSnGrad::correction]: Gauss-linear gradient of vf (boundary values vfb), interpolated to the faces, dotted with the
// non-orthogonal correction vectors; processor faces take the neighbour rank's cell gradient through the halo swap
void Region3D::nonOrthCorrection(const double* vf, const double* vfb, const double* gammaf, double* out) {
    FvWorkspace& w = fvWorkspace(*m);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    B200_LAUNCH(k_grad_scalar, gridFor(N), BLOCK, st, mv, vf, vfb, w.gx.p, w.gy.p, w.gz.p);
    haloScalar(w.gx, w.gb.p); haloScalar(w.gy, w.gb.p + B); haloScalar(w.gz, w.gb.p + 2 * (size_t)B);
    B200_LAUNCH(k_nonorth_ffc, gridFor(nmax), BLOCK, st, mv, gammaf, (const double*)w.gx.p, (const double*)w.gy.p, (const double*)w.gz.p,
                (const double*)w.gb.p, out);
}

void Region3D::solvePressureLike(const DevBC& bc, const double* gammaf, double* psi, double* psib, double tol, double relTol,
                                 bool subtractFlux) {
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    PcgWorkspace& pw = pcgWorkspace(*m);
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    { ScopedTimer t(T_ASSEMBLY);
      const bool corr = nonOrth();   // laplacian `corrected`: explicit faceFluxCorrection from the CURRENT psi (re-assembled per non-orthogonal corrector)
      if (corr) nonOrthCorrection(psi, psib, gammaf, ffc.p);
      B200_LAUNCH(k_lap_row, gridFor(N), BLOCK, st, mv, gammaf, (const double*)phi.p, pdDiag.p, pdUpper.p, pdSource.p, corr ? (const double*)ffc.p : (const double*)nullptr);
      // pdEqn.setReference(pdRefCell, pdRefValue) [REF region3d/pEqn.H:30; correctPhi.H:43]
      if (pdNeedRef && refRow >= 0) B200_LAUNCH(k_set_reference, 1, 32, st, refRow, ctl.pdRefValue, pdDiag.p, pdSource.p);
      if (B) {
          B200_LAUNCH(k_lap_boundary, gridFor(B), BLOCK, st, mv, bc.view(), gammaf, (const double*)psib, ic.p, bcv.p);
          B200_LAUNCH(k_add_boundary, gridFor(m->nBRows), BLOCK, st, mv, (const double*)ic.p, (const double*)bcv.p, pdDiag.p, pdSource.p, m->nBRows,
                      (const int*)pw.bRowOfSlot.p);
      } }
    b200_solver_controls sc{tol, relTol, ctl.minIter, ctl.maxIter};
    if (capture) {
        capSolves.emplace_back();
        CapturedSolve& c = capSolves.back();
        c.diag.resize(N); c.upper.resize(m->F); c.source.resize(N); c.x0.resize(N); c.bou.resize(B); c.tol = tol; c.relTol = relTol;
        m->downloadCells(pdDiag, c.diag.data()); m->downloadFaces(pdUpper, c.upper.data(), false); m->downloadCells(pdSource, c.source.data());
        m->downloadCells(psi, c.x0.data()); if (B) m->downloadBoundary(bcv, c.bou.data());
        syncStream();
    }
    std::vector<double> hist;
    b200_solver_perf perf = pcgSolveDevice(*m, pdDiag, pdUpper, coupled() ? bcv.p : nullptr, psi, pdSource, sc, keepHist ? &hist : nullptr);
    perfs.push_back(perf);
    if (keepHist) resHist.insert(resHist.end(), hist.begin(), hist.end());
    { ScopedTimer t(T_ASSEMBLY);
      evalPdBC(bc, psi, psib);
      if (subtractFlux)
          B200_LAUNCH(k_matrix_flux<true>, gridFor(nmax), BLOCK, st, mv, (const double*)pdUpper.p, (const double*)pdUpper.p, (const double*)ic.p,
                      (const double*)bcv.p, (const double*)psi, (const double*)psib, phi.p, nonOrth() ? (const double*)ffc.p : (const double*)nullptr); }
}

// [REF region3d/UEqn.H:19-34]
void Region3D::momentumPredictor() {
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    if (!momFlux.p) { momFlux.alloc(m->nFaceField()); momDiag.alloc(N); momSrc.alloc(N); if (nonOrth()) corrP.alloc(m->nFaceField()); }
    { ScopedTimer t(T_UEQN);
      const bool corr = nonOrth();
      if (corr) { nonOrthCorrection(alpha, alphab, nullptr, corrA.p); nonOrthCorrection(rho, rhob, nullptr, corrR.p); nonOrthCorrection(pd, pdb, nullptr, corrP.p); }
      B200_LAUNCH(k_mom_flux, gridFor(nmax), BLOCK, st, mv, bcAlpha.view(), bcPd.view(), ctl.sigma, (const double*)K.p, (const double*)Kb.p,
                  (const double*)alpha.p, (const double*)alphab.p, (const double*)rho.p, (const double*)rhob.p, (const double*)pd.p, (const double*)pdb.p,
                  (const double*)ghf.p, corr ? (const double*)corrA.p : (const double*)nullptr, corr ? (const double*)corrR.p : (const double*)nullptr,
                  corr ? (const double*)corrP.p : (const double*)nullptr, momFlux.p);
      B200_LAUNCH(k_reconstruct<0>, gridFor(N), BLOCK, st, mv, reconTensor(*m), (const double*)momFlux.p, (const double*)nullptr, (const double*)nullptr,
                  (const double*)nullptr, Hx.p, Hy.p, Hz.p); }
    b200_solver_controls sc{ctl.UTol, ctl.URelTol, ctl.minIter, ctl.maxIter};
    double* Uk[3] = {Ux.p, Uy.p, Uz.p}; const double* Rk[3] = {Hx.p, Hy.p, Hz.p}; const double* Sk[3] = {uSx.p, uSy.p, uSz.p};
    for (int k = 0; k < 3; k++) {
        { ScopedTimer t(T_UEQN);
          B200_LAUNCH(k_mom_system, gridFor(N), BLOCK, st, mv, k, (const double*)uDiag.p, Sk[k], Rk[k], (const double*)uIC.p, (const double*)uBC.p, momDiag.p, momSrc.p); }
        std::vector<double> hist;
        b200_solver_perf perf = pbicgSolveDevice(*m, momDiag, uUpper, uLower, coupled() ? uBC.p + (size_t)k * B : nullptr, coupled() ? uIC.p + (size_t)k * B : nullptr,
                                                 Uk[k], momSrc, sc, keepHist ? &hist : nullptr);
        perfs.push_back(perf);
        if (keepHist) resHist.insert(resHist.end(), hist.begin(), hist.end());
    }
}

// [REF include/CourantNoMultiRegion.H:78-92]
void Region3D::courant(double* out3) {
    FvWorkspace& w = fvWorkspace(*m);
    cudaStream_t st = rt().stream;
    int nmax = std::max(std::max(m->N, m->B), 1);
    B200_LAUNCH(k_courant, gridFor(nmax), BLOCK, st, m->view(), (const double*)phi.p, scal.p + 44, w.partials.p, w.counter.p);
    if (rt().nRanks > 1) { allReduce(scal.p + 44, 2, RED_SUM); allReduce(scal.p + 46, 2, RED_MAX); }
    double h[4];
    B200_CHECK_CUDA(cudaMemcpyAsync(h, scal.p + 44, 32, cudaMemcpyDeviceToHost, st));
    syncStream();
    // mesh.nInternalFaces() == 0: the reference leaves meanCoNum = -GREAT, CoNum unchanged (0), velMag = 0
    bool any = h[1] > 0.0;
    out3[0] = any ? (h[0] / h[1]) * ctl.deltaT : -1.0e15;
    out3[1] = any ? h[2] * ctl.deltaT : 0.0;
    out3[2] = any ? h[3] : 0.0;
}

void Region3D::refreshDerived() {
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    B200_LAUNCH(k_rho, gridFor(nmax), BLOCK, rt().stream, N, B, (const double*)alpha.p, (const double*)alphab.p, ctl.rho1, ctl.rho2, rho.p, rhob.p);
    calculateK();
    syncStream();
}

// [REF region3d/correctPhi.H]
void Region3D::correctPhi() {
    pcorr.zero(); pcorrb.zero();
    for (int nonOrth = 0; nonOrth <= ctl.nNonOrthCorr; nonOrth++)
        solvePressureLike(bcPcorr, oneF, pcorr, pcorrb, ctl.pcorrTol, ctl.pcorrRelTol, nonOrth == ctl.nNonOrthCorr);
    syncStream();
}

// [REF region3d/pEqn.H:1-49]
void Region3D::pEqn(int corr) {
    FvWorkspace& w = fvWorkspace(*m);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    int N = m->N, B = m->B, nmax = std::max(std::max(N, B), 1);
    double rDeltaT = 1.0 / ctl.deltaT;
    { ScopedTimer t(T_UEQN);
      B200_LAUNCH(k_ua_h, gridFor(N), BLOCK, st, mv, (const double*)uDiag.p, (const double*)uLower.p, (const double*)uUpper.p, (const double*)uSx.p,
                  (const double*)uSy.p, (const double*)uSz.p, (const double*)uIC.p, (const double*)uBC.p, (const double*)Ub.p, (const double*)Ux.p,
                  (const double*)Uy.p, (const double*)Uz.p, rUA.p, Hx.p, Hy.p, Hz.p);
      std::swap(Ux, Hx); std::swap(Uy, Hy); std::swap(Uz, Hz);   // U = rUA*UEqn.H()
      if (B) B200_LAUNCH(k_ub_assign, gridFor(B), BLOCK, st, mv, bcU.view(), (const double*)Ux.p, (const double*)Uy.p, (const double*)Uz.p, Ub.p);
      haloScalar(Ux, Ub.p); haloScalar(Uy, Ub.p + B); haloScalar(Uz, Ub.p + 2 * (size_t)B); haloScalar(rUA, rUAb); }
    { ScopedTimer t(T_ASSEMBLY);
      B200_LAUNCH(k_phiU, gridFor(nmax), BLOCK, st, mv, bcU.view(), rDeltaT, (const double*)rUA.p, (const double*)Ux.p, (const double*)Uy.p,
                  (const double*)Uz.p, (const double*)Ub.p, (const double*)U0x.p, (const double*)U0y.p, (const double*)U0z.p, (const double*)U0b.p,
                  (const double*)rho0.p, (const double*)rho0b.p, (const double*)phi0.p, (const double*)rUAb.p, phiU.p, rUAf.p);
      if (ctl.adjustPhi) {   // adjustPhi(phiU, U, p) [REF region3d/pEqn.H:14]
          B200_LAUNCH(k_adjust_bsums, gridFor(std::max(B, 1)), BLOCK, st, mv, bcU.view(), (const double*)phiU.p, scal.p + 1, w.partials.p, w.counter.p);
          B200_LAUNCH(k_abs_sum_faces, gridFor(std::max(N, 1)), BLOCK, st, mv, (const double*)phiU.p, scal.p + 4, w.partials.p, w.counter.p);
          if (rt().nRanks > 1) allReduce(scal.p + 1, 4, RED_SUM);
          if (B) B200_LAUNCH(k_adjust_scale, gridFor(B), BLOCK, st, mv, bcU.view(), (const double*)(scal.p + 1), phiU.p);
      }
      const bool corr = nonOrth();   // snGrad `corrected` [REF region3d/pEqn.H:18-19]
      if (corr) { nonOrthCorrection(alpha, alphab, nullptr, corrA.p); nonOrthCorrection(rho, rhob, nullptr, corrR.p); }
      B200_LAUNCH(k_phi, gridFor(nmax), BLOCK, st, mv, bcAlpha.view(), ctl.sigma, (const double*)K.p, (const double*)Kb.p, (const double*)alpha.p, (const double*)alphab.p,
                  (const double*)rho.p, (const double*)rhob.p, (const double*)ghf.p, (const double*)rUAf.p, (const double*)phiU.p, phi.p,
                  corr ? (const double*)corrA.p : (const double*)nullptr, corr ? (const double*)corrR.p : (const double*)nullptr); }
    for (int nonOrth = 0; nonOrth <= ctl.nNonOrthCorr; nonOrth++) {
        bool fin = (corr == ctl.nCorr - 1 && nonOrth == ctl.nNonOrthCorr);
        solvePressureLike(bcPd, rUAf, pd, pdb, fin ? ctl.pdFinalTol : ctl.pdTol, fin ? ctl.pdFinalRelTol : ctl.pdRelTol,
                          nonOrth == ctl.nNonOrthCorr);
    }
    { ScopedTimer t(T_ASSEMBLY);
      const double* invT = reconTensor(*m);
      B200_LAUNCH(k_reconstruct<1>, gridFor(N), BLOCK, st, mv, invT, (const double*)phi.p, (const double*)phiU.p, (const double*)rUAf.p,
                  (const double*)rUA.p, Ux.p, Uy.p, Uz.p);
      evalUBC(); }
}

// [REF region3d/solve3d.H:40-72]
void Region3D::step() {
    FvWorkspace& w = fvWorkspace(*m);
    MeshView mv = m->view();
    cudaStream_t st = rt().stream;
    ScopedTimer tstep(T_STEP);
    size_t N = m->N, B = m->B, nf = m->nFaceField();
    int nmax = (int)std::max<size_t>(std::max(N, B), 1);
    auto d2d = [&](DevBuf<double>& dst, const DevBuf<double>& src, size_t n) {
        if (n) B200_CHECK_CUDA(cudaMemcpyAsync(dst.p, src.p, 8 * n, cudaMemcpyDeviceToDevice, st));
    };
    // runTime++ : old-time fields
    d2d(alpha0, alpha, N); d2d(U0x, Ux, N); d2d(U0y, Uy, N); d2d(U0z, Uz, N); d2d(U0b, Ub, 3 * B);
    d2d(rho0, rho, N); d2d(rho0b, rhob, B); d2d(phi0, phi, nf);
    // alphaEqnSubCycle.H [REF region3d/alphaEqnSubCycle.H:11-35]
    if (ctl.nAlphaSubCycles > 1) {
        double dt = ctl.deltaT / ctl.nAlphaSubCycles;
        B200_LAUNCH(k_scale, gridFor((long long)nf), BLOCK, st, (int)nf, 0.0, (const double*)rhoPhi.p, rhoPhiSum.p, 0);
        d2d(alphaStart, alpha, N);
        for (int sc = 0; sc < ctl.nAlphaSubCycles; sc++) {
            d2d(alpha0, alpha, N);
            alphaEqn(dt);
            B200_LAUNCH(k_scale, gridFor((long long)nf), BLOCK, st, (int)nf, dt / ctl.deltaT, (const double*)rhoPhi.p, rhoPhiSum.p, 1);
        }
        d2d(alpha0, alphaStart, N);
        d2d(rhoPhi, rhoPhiSum, nf);
    } else alphaEqn(ctl.deltaT);
    calculateK();
    B200_LAUNCH(k_rho, gridFor(nmax), BLOCK, st, (int)N, (int)B, (const double*)alpha.p, (const double*)alphab.p, ctl.rho1, ctl.rho2, rho.p, rhob.p);
    assembleUEqn();
    if (ctl.momentumPredictor) momentumPredictor();
    evalUBC();
    for (int corr = 0; corr < ctl.nCorr; corr++) pEqn(corr);
    B200_LAUNCH(k_cont_err, gridFor((long long)std::max<size_t>(N, 1)), BLOCK, st, mv, (const double*)phi.p, scal.p + 9, w.partials.p, w.counter.p);
    if (rt().nRanks > 1) allReduce(scal.p + 9, 3, RED_SUM);
    B200_CHECK_CUDA(cudaMemcpyAsync(hscal + 9, scal.p + 9, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200_LAUNCH(k_p, gridFor((long long)std::max<size_t>(N, 1)), BLOCK, st, (int)N, (const double*)pd.p, (const double*)rho.p, (const double*)gh.p, p.p);
    syncStream();
    resolveStats();
    sumLocalContErr = ctl.deltaT * (hscal[9] / hscal[11]);
    globalContErr = ctl.deltaT * (hscal[10] / hscal[11]);
    cumulativeContErr += globalContErr;
}

}  // namespace b200
