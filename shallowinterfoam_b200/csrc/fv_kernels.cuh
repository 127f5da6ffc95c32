// fv_kernels.cuh -- sm_100a kernels of the alpha1 advection (limited face fluxes + MULES) and the pd assembly
// (interpolate / snGrad / laplacian coefficients / div / flux / reconstruct) plus UEqn A()/H().
// All kernels are row- or face-gather kernels on the owner-slot SELL layout (device_mesh.cuh): coalesced SoA
// streams, neighbour values through L1/L2, no atomics, per-cell summation in upstream's face order
// ([nbr-side faces asc] + [owner faces asc] + [boundary faces asc]) => bitwise the sequential foam-extend loops.
#pragma once
#include "device_mesh.cuh"
#include "ldu_kernels.cuh"

namespace b200 {

#define FV_VSMALL 1.0e-300
#define FV_SMALL 1.0e-15

// ---- face iteration of one row (thread = row; s = slice = warp; lane) ----------------------------------------
// (uniform slice widths, device_mesh.cu: the slice offsets are arithmetic -- two dependent loads less per row)
#define FOR_NBR_FACES(mv, s, lane, pos, l)                                                              \
    for (int _j = 0, _nb0 = (mv).uniNw > 0 ? (s) * 32 * (mv).uniNw : (mv).nbrOff[s],                    \
             _nbw = (mv).uniNw > 0 ? (mv).uniNw : ((mv).nbrOff[(s) + 1] - _nb0) >> 5; _j < _nbw; _j++) { \
        int _pk = (mv).loPk[_nb0 + (_j << 5) + (lane)];                                                 \
        if (_pk < 0) break;                                                                             \
        int l;                                                                                          \
        int pos = facePosOf(mv, _pk, l);
#define FOR_OWN_FACES(mv, s, lane, pos, u)                                                                \
    for (int _k = 0, _so0 = (mv).uniOw > 0 ? (s) * 32 * (mv).uniOw : (mv).sliceOff[s],                    \
             _ow = (mv).uniOw > 0 ? (mv).uniOw : ((mv).sliceOff[(s) + 1] - _so0) >> 5; _k < _ow; _k++) {  \
        int pos = _so0 + (_k << 5) + (lane);                                                              \
        int u = (mv).upIdx[pos];                                                                          \
        if (u < 0) break;
#define FOR_BND_FACES(mv, s, lane, b)                                                      \
    if (((mv).bMask[s] >> (lane)) & 1u) {                                                  \
        int _bi = (mv).bBase[s] + __popc((mv).bMask[s] & ((1u << (lane)) - 1u));           \
        for (int _q = (mv).bRowStart[_bi]; _q < (mv).bRowStart[_bi + 1]; _q++) {            \
            int b = (mv).bRowFaces[_q];
#define END_FACES }
#define END_BND }}

// L1 prefetch of the row's index lines (uniform slices: their addresses are arithmetic), so that the face loops below find them in
// L1 instead of paying one DRAM round trip per face; PREFETCH_OWN_FACES does the same for a face field at the row's owner slots
#ifdef B200_EMU
__device__ __forceinline__ void prefetchL1(const void*) {}
#else
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif
#define PREFETCH_OWN_FACES(mv, s, lane, arr)                                                                       \
    if ((mv).uniOw > 0 && (arr)) { for (int _k = 0; _k < (mv).uniOw; _k++) prefetchL1((arr) + (((s) * (mv).uniOw + _k) << 5) + (lane)); }
#define ROW_PROLOGUE(mv)                                               \
    int row = blockIdx.x * blockDim.x + threadIdx.x;                   \
    int s = row >> 5, lane = row & 31;                                 \
    (void)s; (void)lane;                                               \
    if (row < (mv).N && (mv).uniOw > 0) {                                                                          \
        for (int _k = 0; _k < (mv).uniOw; _k++) prefetchL1((mv).upIdx + ((s * (mv).uniOw + _k) << 5) + lane);      \
        for (int _k = 0; _k < (mv).uniNw; _k++) prefetchL1((mv).loPk + ((s * (mv).uniNw + _k) << 5) + lane);       \
    }

// ---- boundary-condition algebra [UPSTREAM fvPatchField / fixedValue / zeroGradient / mixedFvPatchField] ---------
struct BCView {
    const int* type;    // [nPatches]
    const double* rv;   // [nc*B] SoA: component k at rv + k*B
    const double* rg;   // [nc*B]
    const double* vf;   // [B]
};
__device__ __forceinline__ bool bcFixes(int t) { return t == B200_BC_FIXED_VALUE || t == B200_BC_MIXED; }
__device__ __forceinline__ double bcSnGrad(const BCView& bc, int t, int b, int idx, double dc, double pif, double bv) {
    if (t == B200_BC_ZERO_GRADIENT) return 0.0;
    if (t == B200_BC_MIXED) { double fr = bc.vf[b]; return fr * (bc.rv[idx] - pif) * dc + (1.0 - fr) * bc.rg[idx]; }
    return dc * (bv - pif);
}
__device__ __forceinline__ double bcValueIntCoeff(const BCView& bc, int t, int b) {
    return t == B200_BC_FIXED_VALUE ? 0.0 : t == B200_BC_ZERO_GRADIENT ? 1.0 : t == B200_BC_MIXED ? 1.0 * (1.0 - bc.vf[b]) : 0.0;
}
__device__ __forceinline__ double bcValueBouCoeff(const BCView& bc, int t, int b, int idx, double dc, double bv) {
    if (t == B200_BC_FIXED_VALUE) return bv;
    if (t == B200_BC_MIXED) { double fr = bc.vf[b]; return fr * bc.rv[idx] + (1.0 - fr) * bc.rg[idx] / dc; }
    return 0.0;
}
__device__ __forceinline__ double bcGradIntCoeff(const BCView& bc, int t, int b, double dc) {
    if (t == B200_BC_FIXED_VALUE) return -1.0 * dc;
    if (t == B200_BC_MIXED) return -1.0 * bc.vf[b] * dc;
    return 0.0;
}
__device__ __forceinline__ double bcGradBouCoeff(const BCView& bc, int t, int b, int idx, double dc, double bv) {
    if (t == B200_BC_FIXED_VALUE) return dc * bv;
    if (t == B200_BC_MIXED) { double fr = bc.vf[b]; return fr * dc * bc.rv[idx] + (1.0 - fr) * bc.rg[idx]; }
    return 0.0;
}

// face value on a boundary face: the patch value, or -- processor patch -- the linear interpolate with the neighbour cell
// (vfb holds patchNeighbourField there) [UPSTREAM surfaceInterpolationScheme::interpolate coupled branch]
__device__ __forceinline__ double bndFaceValue(const MeshView& mv, int b, double own, double bv) {
    return mv.bCoupled[b] ? mv.w[mv.Fs + b] * (own - bv) + bv : bv;
}

// evaluate(): nc components, SoA boundary arrays (component k at bval + k*B); internal SoA pointers i0,i1,i2
static __global__ void k_bc_evaluate(MeshView mv, BCView bc, int nc, const double* __restrict__ i0,
                                     const double* __restrict__ i1, const double* __restrict__ i2, double* bval) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B) return;
    int t = bc.type[mv.bPatch[b]], c = mv.bFaceRow[b];
    if (t == B200_BC_PROCESSOR || t == B200_BC_CALCULATED) return;
    double dc = mv.dc[mv.Fs + b];
    for (int k = 0; k < nc; k++) {
        double pif = (k == 0 ? i0 : k == 1 ? i1 : i2)[c];
        int idx = k * mv.B + b;
        double v;
        if (t == B200_BC_FIXED_VALUE) v = bc.rv[idx];
        else if (t == B200_BC_ZERO_GRADIENT) v = pif;
        else { double fr = bc.vf[b]; v = fr * bc.rv[idx] + (1.0 - fr) * (pif + bc.rg[idx] / dc); }
        bval[idx] = v;
    }
}

// ---- a13: linear interpolate (nc = 1 or 3), snGrad, surfaceIntegrate ------------------------------------------
static __global__ void k_interpolate(MeshView mv, const double* __restrict__ vf, const double* __restrict__ vfb,
                                     double* __restrict__ out) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double a = vf[row];
        FOR_OWN_FACES(mv, s, lane, pos, u) double n = vf[u]; out[pos] = mv.w[pos] * (a - n) + n; END_FACES
    }
    if (row < mv.B) out[mv.Fs + row] = bndFaceValue(mv, row, vf[mv.bFaceRow[row]], vfb[row]);
}
static __global__ void k_sngrad(MeshView mv, BCView bc, const double* __restrict__ vf, const double* __restrict__ vfb,
                                double* __restrict__ out) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double a = vf[row];
        FOR_OWN_FACES(mv, s, lane, pos, u) out[pos] = mv.dc[pos] * (vf[u] - a); END_FACES
    }
    if (row < mv.B) {
        int b = row, t = bc.type[mv.bPatch[b]];
        out[mv.Fs + b] = bcSnGrad(bc, t, b, b, mv.dc[mv.Fs + b], vf[mv.bFaceRow[b]], vfb[b]);
    }
}
// out = sign * (sum_owner - sum_nbr + sum_boundary) / V
static __global__ void k_surface_integrate(MeshView mv, const double* __restrict__ ssf, double* __restrict__ out, double sign) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    double a = 0.0;
    FOR_NBR_FACES(mv, s, lane, pos, l) (void)l; a = a - ssf[pos]; END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; a = a + ssf[pos]; END_FACES
    FOR_BND_FACES(mv, s, lane, b) a = a + ssf[mv.Fs + b]; END_BND
    a = a / mv.V[row];
    out[row] = sign < 0 ? -a : a;
}

// ---- Gauss linear gradient of a scalar [UPSTREAM gaussGrad.C] ---------------------------------------------------
static __global__ void k_grad_scalar(MeshView mv, const double* __restrict__ vf, const double* __restrict__ vfb,
                                     double* __restrict__ gx, double* __restrict__ gy, double* __restrict__ gz) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, mv.w)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfx)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfy)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfz)
    double a = vf[row], x = 0.0, y = 0.0, z = 0.0;
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double o = vf[l];
        double sf = mv.w[pos] * (o - a) + a;
        x = x - mv.Sfx[pos] * sf; y = y - mv.Sfy[pos] * sf; z = z - mv.Sfz[pos] * sf;
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double n = vf[u];
        double sf = mv.w[pos] * (a - n) + n;
        x = x + mv.Sfx[pos] * sf; y = y + mv.Sfy[pos] * sf; z = z + mv.Sfz[pos] * sf;
    END_FACES
    FOR_BND_FACES(mv, s, lane, b)
        double v = bndFaceValue(mv, b, a, vfb[b]); int p = mv.Fs + b;
        x = x + mv.Sfx[p] * v; y = y + mv.Sfy[p] * v; z = z + mv.Sfz[p] * v;
    END_BND
    double V = mv.V[row];
    gx[row] = x / V; gy[row] = y / V; gz[row] = z / V;
}
// boundary value of the gradient: zero-gradient copy with the normal component replaced by the patch snGrad
static __global__ void k_grad_boundary(MeshView mv, BCView bc, const double* __restrict__ vf, const double* __restrict__ vfb,
                                       const double* __restrict__ gx, const double* __restrict__ gy,
                                       const double* __restrict__ gz, double* __restrict__ gb /*[3B] SoA*/) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B) return;
    int c = mv.bFaceRow[b], p = mv.Fs + b, t = bc.type[mv.bPatch[b]];
    double ms = mv.magSf[p];
    double n0 = mv.Sfx[p] / ms, n1 = mv.Sfy[p] / ms, n2 = mv.Sfz[p] / ms;
    if (t == B200_BC_PROCESSOR) return;   // filled by the halo swap of the neighbour's cell gradient
    double g0 = gx[c], g1 = gy[c], g2 = gz[c];
    {
        double sn = bcSnGrad(bc, t, b, b, mv.dc[p], vf[c], vfb[b]);
        double nd = n0 * g0 + n1 * g1 + n2 * g2;
        g0 += n0 * (sn - nd); g1 += n1 * (sn - nd); g2 += n2 * (sn - nd);
    }
    gb[b] = g0; gb[mv.B + b] = g1; gb[2 * mv.B + b] = g2;
}

// ---- a7/a8: limited schemes [UPSTREAM LimitedScheme<vanLeer,NVDTVD>, interfaceCompression] ---------------------
__device__ __forceinline__ double vanLeerLimiter(double phiP, double phiN, double gPx, double gPy, double gPz,
                                                 double gNx, double gNy, double gNz, double dx, double dy, double dz,
                                                 double faceFlux) {
    double gradf = phiN - phiP, gradcf;
    if (faceFlux > 0) gradcf = dx * gPx + dy * gPy + dz * gPz;
    else gradcf = dx * gNx + dy * gNy + dz * gNz;
    double r;
    if (fabs(gradcf) >= 1000 * fabs(gradf)) {
        double sgc = gradcf >= 0 ? 1.0 : -1.0, sgf = gradf >= 0 ? 1.0 : -1.0;
        r = 2 * 1000 * sgc * sgf - 1;
    } else r = 2 * (gradcf / gradf) - 1;
    return (r + fabs(r)) / (1 + fabs(r));
}
__device__ __forceinline__ double compressionLimiter(double aP, double aN) {
    double vP = fmin(fmax(aP, 0.0), 1.0), vN = fmin(fmax(aN, 0.0), 1.0);
    double wP = 1.0 - 4.0 * vP * (1.0 - vP), wN = 1.0 - 4.0 * vN * (1.0 - vN);
    return 1.0 - fmax(wP * wP, wN * wN);
}
__device__ __forceinline__ double limitedFaceFlux(double lim, double w, double flux, double vP, double vN) {
    double pos = flux >= 0 ? 1.0 : 0.0;
    double wgt = lim * w + (1.0 - lim) * pos;
    return flux * (wgt * (vP - vN) + vN);
}
// generic fvc::flux(faceFlux, vf, scheme): out = faceFlux*interpolate(vf)
static __global__ void k_flux_limited(MeshView mv, int scheme, const double* __restrict__ faceFlux,
                                      const double* __restrict__ vf, const double* __restrict__ vfb,
                                      const double* __restrict__ gx, const double* __restrict__ gy,
                                      const double* __restrict__ gz, const double* __restrict__ gbn /*[3B] neighbour gradient on processor faces*/,
                                      double* __restrict__ out) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        PREFETCH_OWN_FACES(mv, s, lane, faceFlux)
        PREFETCH_OWN_FACES(mv, s, lane, mv.w)
        double a = vf[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double n = vf[u], ff = faceFlux[pos], lim;
            if (scheme == B200_SCHEME_VANLEER)
                lim = vanLeerLimiter(a, n, gx[row], gy[row], gz[row], gx[u], gy[u], gz[u], mv.Cx[u] - mv.Cx[row],
                                     mv.Cy[u] - mv.Cy[row], mv.Cz[u] - mv.Cz[row], ff);
            else if (scheme == B200_SCHEME_INTERFACE_COMPRESSION) lim = compressionLimiter(a, n);
            else if (scheme == B200_SCHEME_UPWIND) lim = 0.0;
            else lim = 1.0;
            out[pos] = limitedFaceFlux(lim, mv.w[pos], ff, a, n);
        END_FACES
    }
    if (row < mv.B) {
        int b = row, p = mv.Fs + b;
        double ff = faceFlux[p], n = vfb[b];
        if (!mv.bCoupled[b]) out[p] = ff * n;
        else {   // processor face: same formula as an internal face, own cell = P, neighbour-rank cell = N
            int c = mv.bFaceRow[b];
            double a = vf[c], lim;
            if (scheme == B200_SCHEME_VANLEER)
                lim = vanLeerLimiter(a, n, gx[c], gy[c], gz[c], gbn[b], gbn[mv.B + b], gbn[2 * mv.B + b], mv.Cnx[b] - mv.Cx[c],
                                     mv.Cny[b] - mv.Cy[c], mv.Cnz[b] - mv.Cz[c], ff);
            else if (scheme == B200_SCHEME_INTERFACE_COMPRESSION) lim = compressionLimiter(a, n);
            else if (scheme == B200_SCHEME_UPWIND) lim = 0.0;
            else lim = 1.0;
            out[p] = limitedFaceFlux(lim, mv.w[p], ff, a, n);
        }
    }
}

// max(|phi/magSf|) over internal + boundary faces [REF region3d/alphaEqn.H:6-8]
static __global__ void k_phic_max(MeshView mv, const double* __restrict__ phi, double* out, double* partials, unsigned* counter) {
    ROW_PROLOGUE(mv)
    double m = 0.0;
    if (row < mv.N) { FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; m = fmax(m, fabs(phi[pos] / mv.magSf[pos])); END_FACES }
    if (row < mv.B) m = fmax(m, fabs(phi[mv.Fs + row] / mv.magSf[mv.Fs + row]));
    double v[1] = {m};
    gridReduce<1, OpMax>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    gridFinalize<1, OpMax>(gridDim.x, partials, counter, out);
}

// phiAlpha = flux(phi, alpha1, vanLeer) + flux(-flux(-phir, 1-alpha1, iC), alpha1, iC)  [REF region3d/alphaEqn.H:6-28]
static __global__ void __launch_bounds__(BLOCK) k_alpha_flux(MeshView mv, const double* __restrict__ phi,
                                                              const double* __restrict__ alpha, const double* __restrict__ alphab,
                                                              const double* __restrict__ gx, const double* __restrict__ gy,
                                                              const double* __restrict__ gz, const double* __restrict__ gbn,
                                                              const double* __restrict__ nHatf, double cAlpha, const double* __restrict__ maxPhic,
                                                              double* __restrict__ phiAlpha) {
    ROW_PROLOGUE(mv)
    double mx = *maxPhic;
    if (row < mv.N) {
        PREFETCH_OWN_FACES(mv, s, lane, phi)
        PREFETCH_OWN_FACES(mv, s, lane, mv.w)
        PREFETCH_OWN_FACES(mv, s, lane, mv.magSf)
        PREFETCH_OWN_FACES(mv, s, lane, nHatf)
        double a = alpha[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double n = alpha[u], ff = phi[pos], w = mv.w[pos];
            double phic = fmin(cAlpha * fabs(ff / mv.magSf[pos]), mx);
            double phir = phic * nHatf[pos];
            double lim = vanLeerLimiter(a, n, gx[row], gy[row], gz[row], gx[u], gy[u], gz[u], mv.Cx[u] - mv.Cx[row],
                                        mv.Cy[u] - mv.Cy[row], mv.Cz[u] - mv.Cz[row], ff);
            double f1 = limitedFaceFlux(lim, w, ff, a, n);
            double oa = 1.0 - a, on = 1.0 - n, mphir = -phir;
            double t1 = limitedFaceFlux(compressionLimiter(oa, on), w, mphir, oa, on);
            t1 = -t1;
            double t2 = limitedFaceFlux(compressionLimiter(a, n), w, t1, a, n);
            phiAlpha[pos] = f1 + t2;
        END_FACES
    }
    if (row < mv.B) {
        int p = mv.Fs + row;
        double ff = phi[p], ab = alphab[row];
        double phic = fmin(cAlpha * fabs(ff / mv.magSf[p]), mx);
        double phir = phic * nHatf[p];
        if (!mv.bCoupled[row]) {
            double f1 = ff * ab;
            double t1 = (-phir) * (1.0 - ab);
            t1 = -t1;
            phiAlpha[p] = f1 + t1 * ab;
        } else {
            int b = row, c = mv.bFaceRow[b];
            double a = alpha[c], n = ab, w = mv.w[p];
            double lim = vanLeerLimiter(a, n, gx[c], gy[c], gz[c], gbn[b], gbn[mv.B + b], gbn[2 * mv.B + b], mv.Cnx[b] - mv.Cx[c],
                                        mv.Cny[b] - mv.Cy[c], mv.Cnz[b] - mv.Cz[c], ff);
            double f1 = limitedFaceFlux(lim, w, ff, a, n);
            double oa = 1.0 - a, on = 1.0 - n, mphir = -phir;
            double t1 = limitedFaceFlux(compressionLimiter(oa, on), w, mphir, oa, on);
            t1 = -t1;
            phiAlpha[p] = f1 + limitedFaceFlux(compressionLimiter(a, n), w, t1, a, n);
        }
    }
}

// ---- a5/a6: MULES [UPSTREAM MULESTemplates.C explicitSolve / limiter] -------------------------------------------
// phiBD = upwind(phi).flux(psi) ; phiCorr = phiPsi - phiBD (in place in phiPsi)
static __global__ void k_mules_bd_corr(MeshView mv, const double* __restrict__ phi, const double* __restrict__ psi,
                                       const double* __restrict__ psib, double* __restrict__ phiPsi, double* __restrict__ phiBD) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        PREFETCH_OWN_FACES(mv, s, lane, phi)
        PREFETCH_OWN_FACES(mv, s, lane, phiPsi)
        double a = psi[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double ff = phi[pos];
            double bd = ff * (ff >= 0 ? a : psi[u]);
            phiBD[pos] = bd; phiPsi[pos] = phiPsi[pos] - bd;
        END_FACES
    }
    if (row < mv.B) {
        int p = mv.Fs + row;
        double ff = phi[p];
        double bd = ff * ((mv.bCoupled[row] && ff >= 0) ? psi[mv.bFaceRow[row]] : psib[row]);
        phiBD[p] = bd; phiPsi[p] = phiPsi[p] - bd;
    }
}
// local extrema and admissible in/out-flows
static __global__ void __launch_bounds__(BLOCK) k_mules_bounds(MeshView mv, BCView bc, const double* __restrict__ psi,
                                                                const double* __restrict__ psib, const double* __restrict__ psiNbr,
                                                                const double* __restrict__ psi0, const double* __restrict__ phiBD,
                                                                const double* __restrict__ phiCorr, double deltaT, double psiMax,
                                                                double psiMin, int includeOwn, double* __restrict__ psiMaxn,
                                                                double* __restrict__ psiMinn, double* __restrict__ sumPhip,
                                                                double* __restrict__ mSumPhim) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, phiBD)
    PREFETCH_OWN_FACES(mv, s, lane, phiCorr)
    double mxn = psiMin, mnn = psiMax, sBD = 0.0, sP = FV_VSMALL, sM = FV_VSMALL;
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double o = psi[l];
        mxn = fmax(mxn, o); mnn = fmin(mnn, o);
        sBD = sBD - phiBD[pos];
        double pc = phiCorr[pos];
        if (pc > 0.0) sM = sM + pc; else sP = sP - pc;
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double n = psi[u];
        mxn = fmax(mxn, n); mnn = fmin(mnn, n);
        sBD = sBD + phiBD[pos];
        double pc = phiCorr[pos];
        if (pc > 0.0) sP = sP + pc; else sM = sM - pc;
    END_FACES
    FOR_BND_FACES(mv, s, lane, b)
        double pv = (bc.type[mv.bPatch[b]] == B200_BC_PROCESSOR && psiNbr) ? psiNbr[b] : psib[b];
        mxn = fmax(mxn, pv); mnn = fmin(mnn, pv);
        sBD = sBD + phiBD[mv.Fs + b];
        double pc = phiCorr[mv.Fs + b];
        if (pc > 0.0) sP = sP + pc; else sM = sM - pc;
    END_BND
    double me = psi[row];
    if (includeOwn) { mxn = fmax(mxn, me); mnn = fmin(mnn, me); }
    mxn = fmin(mxn, psiMax); mnn = fmax(mnn, psiMin);
    double rDT = 1.0 / deltaT, V = mv.V[row], p0 = psi0[row];
    psiMaxn[row] = V * (rDT * mxn - p0 / deltaT) + sBD;
    psiMinn[row] = V * (p0 / deltaT - rDT * mnn) - sBD;
    sumPhip[row] = sP; mSumPhim[row] = sM;
}
// lambda of one face after the previous limiter iteration, rebuilt on the fly:
//   lambda = min(lambdaBefore, phiCorr > 0 ? min(lambdap[own], lambdam[nei]) : min(lambdam[own], lambdap[nei]))
__device__ __forceinline__ double faceLambda(double before, double pc, double lpOwn, double lmOwn, double lpNei, double lmNei) {
    return pc > 0.0 ? fmin(before, fmin(lpOwn, lmNei)) : fmin(before, fmin(lmOwn, lpNei));
}
// One limiter iteration, fused: rebuilds the face lambdas of iteration it-1 (from lamFacePrev = lambdas of it-2 and the
// cell lambdap/lambdam of it-1), sums lambda*phiCorr, and produces this iteration's cell lambdap/lambdam.
// FIRST: all lambdas are 1.  lamFacePrev == nullptr means "1".  Writes the rebuilt lambdas of the row's owner + boundary faces.
template <bool FIRST>
static __global__ void __launch_bounds__(BLOCK) k_mules_iter(MeshView mv, const double* __restrict__ phiCorr,
                                                              const double* __restrict__ lamFacePrev, double* __restrict__ lamFaceOut,
                                                              const double* __restrict__ lamPprev, const double* __restrict__ lamMprev,
                                                              const double* __restrict__ lamPnbr, const double* __restrict__ lamMnbr,
                                                              const int* __restrict__ bCoupled,
                                                              const double* __restrict__ psiMaxn, const double* __restrict__ psiMinn,
                                                              const double* __restrict__ sumPhip, const double* __restrict__ mSumPhim,
                                                              double* __restrict__ lamPnew, double* __restrict__ lamMnew) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, phiCorr)
    if (!FIRST) { PREFETCH_OWN_FACES(mv, s, lane, lamFacePrev) }
    double sP = 0.0, sM = 0.0, lpMe = 1.0, lmMe = 1.0;
    if (!FIRST) { lpMe = lamPprev[row]; lmMe = lamMprev[row]; }
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double pc = phiCorr[pos], lam = 1.0;
        if (!FIRST) lam = faceLambda(lamFacePrev ? lamFacePrev[pos] : 1.0, pc, lamPprev[l], lamMprev[l], lpMe, lmMe);
        double q = lam * pc;
        if (q > 0.0) sM = sM + q; else sP = sP - q;
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double pc = phiCorr[pos], lam = 1.0;
        if (!FIRST) { lam = faceLambda(lamFacePrev ? lamFacePrev[pos] : 1.0, pc, lpMe, lmMe, lamPprev[u], lamMprev[u]); lamFaceOut[pos] = lam; }
        double q = lam * pc;
        if (q > 0.0) sP = sP + q; else sM = sM - q;
    END_FACES
    FOR_BND_FACES(mv, s, lane, b)
        int p = mv.Fs + b;
        double pc = phiCorr[p], lam = 1.0;
        if (!FIRST) {
            double before = lamFacePrev ? lamFacePrev[p] : 1.0;
            lam = pc > 0.0 ? fmin(before, lpMe) : fmin(before, lmMe);
            if (bCoupled[b] && lamPnbr)  // syncFaceList(minEqOp): the other side limits with its own cell
                lam = fmin(lam, pc > 0.0 ? lamMnbr[b] : lamPnbr[b]);
            lamFaceOut[p] = lam;
        }
        double q = lam * pc;
        if (q > 0.0) sP = sP + q; else sM = sM - q;
    END_BND
    lamMnew[row] = fmax(fmin((sP + psiMaxn[row]) / mSumPhim[row], 1.0), 0.0);
    lamPnew[row] = fmax(fmin((sM + psiMinn[row]) / sumPhip[row], 1.0), 0.0);
}
// final pass: last face lambdas on the fly, phiPsi = phiBD + lambda*phiCorr, psi = (psi0/dt - div(phiPsi))/(1/dt)
static __global__ void __launch_bounds__(BLOCK) k_mules_update(MeshView mv, const double* __restrict__ phiBD,
                                                                double* __restrict__ phiPsi /*in: phiCorr, out: limited flux*/,
                                                                const double* __restrict__ phiCorrRO /* == phiPsi unless a copy is used */,
                                                                const double* __restrict__ lamFacePrev,
                                                                const double* __restrict__ lamP, const double* __restrict__ lamM,
                                                                const double* __restrict__ lamPnbr, const double* __restrict__ lamMnbr,
                                                                const int* __restrict__ bCoupled, const double* __restrict__ psi0,
                                                                double deltaT, double* __restrict__ psi, double* __restrict__ lambdaOut,
                                                                double* __restrict__ fluxOut) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, phiCorrRO)
    PREFETCH_OWN_FACES(mv, s, lane, phiBD)
    PREFETCH_OWN_FACES(mv, s, lane, lamFacePrev)
    double lpMe = lamP[row], lmMe = lamM[row], a = 0.0;
    FOR_NBR_FACES(mv, s, lane, pos, l)
        double pc = phiCorrRO[pos];
        double lam = faceLambda(lamFacePrev ? lamFacePrev[pos] : 1.0, pc, lamP[l], lamM[l], lpMe, lmMe);
        a = a - (phiBD[pos] + lam * pc);
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        double pc = phiCorrRO[pos];
        double lam = faceLambda(lamFacePrev ? lamFacePrev[pos] : 1.0, pc, lpMe, lmMe, lamP[u], lamM[u]);
        double fl = phiBD[pos] + lam * pc;
        a = a + fl;
        if (fluxOut) fluxOut[pos] = fl;
        if (lambdaOut) lambdaOut[pos] = lam;
    END_FACES
    FOR_BND_FACES(mv, s, lane, b)
        int p = mv.Fs + b;
        double pc = phiCorrRO[p], before = lamFacePrev ? lamFacePrev[p] : 1.0;
        double lam = pc > 0.0 ? fmin(before, lpMe) : fmin(before, lmMe);
        if (bCoupled[b] && lamPnbr) lam = fmin(lam, pc > 0.0 ? lamMnbr[b] : lamPnbr[b]);
        double fl = phiBD[p] + lam * pc;
        a = a + fl;
        if (fluxOut) fluxOut[p] = fl;
        if (lambdaOut) lambdaOut[p] = lam;
    END_BND
    if (psi) psi[row] = (psi0[row] / deltaT - a / mv.V[row]) / (1.0 / deltaT);
    (void)phiPsi;
}

// ---- a10/a11/a12: pressure assembly ---------------------------------------------------------------------------
// upper = dc*(gamma*magSf) ; diag = -sum(upper) ; source = V*(sum(+-phi)/V)  [UPSTREAM fvmLaplacianUncorrected, negSumDiag, ==div]
// ffc (optional): faceFluxCorrection of the `corrected` laplacian; source -= V*div(ffc) before the == div(phi) term is added
static __global__ void __launch_bounds__(BLOCK) k_lap_row(MeshView mv, const double* __restrict__ gammaf, const double* __restrict__ phi,
                                                           double* __restrict__ diag, double* __restrict__ upper, double* __restrict__ source,
                                                           const double* __restrict__ ffc) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, gammaf)
    PREFETCH_OWN_FACES(mv, s, lane, mv.dc)
    PREFETCH_OWN_FACES(mv, s, lane, mv.magSf)
    PREFETCH_OWN_FACES(mv, s, lane, phi)
    double d = 0.0, a = 0.0, cc = 0.0;
    FOR_NBR_FACES(mv, s, lane, pos, l)
        (void)l;
        d = d - mv.dc[pos] * (gammaf[pos] * mv.magSf[pos]);
        if (phi) a = a - phi[pos];
        if (ffc) cc = cc - ffc[pos];
    END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u)
        (void)u;
        double up = mv.dc[pos] * (gammaf[pos] * mv.magSf[pos]);
        upper[pos] = up;
        d = d - up;
        if (phi) a = a + phi[pos];
        if (ffc) cc = cc + ffc[pos];
    END_FACES
    double src = 0.0;
    if (ffc) {   // fvm.source() -= mesh.V()*fvc::div(faceFluxCorrection) [UPSTREAM gaussLaplacianScheme::fvmLaplacian]
        FOR_BND_FACES(mv, s, lane, b) cc = cc + ffc[mv.Fs + b]; END_BND
        double V = mv.V[row];
        src = src - V * (cc / V);
    }
    if (phi) {
        FOR_BND_FACES(mv, s, lane, b) a = a + phi[mv.Fs + b]; END_BND
        double V = mv.V[row];
        src = src + V * (a / V);
    }
    diag[row] = d; source[row] = src;
}
static __global__ void k_lap_boundary(MeshView mv, BCView bc, const double* __restrict__ gammaf, const double* __restrict__ pb,
                                      double* __restrict__ ic, double* __restrict__ bcv) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mv.B) return;
    int p = mv.Fs + b, t = bc.type[mv.bPatch[b]];
    double pg = gammaf[p] * mv.magSf[p], dc = mv.dc[p];
    if (t == B200_BC_PROCESSOR) {   // coupledFvPatchField: gradientInternalCoeffs = -deltaCoeffs, gradientBoundaryCoeffs = +deltaCoeffs
        ic[b] = pg * (-1.0 * dc); bcv[b] = -pg * dc;
    } else {
        ic[b] = pg * bcGradIntCoeff(bc, t, b, dc);
        bcv[b] = -pg * bcGradBouCoeff(bc, t, b, b, dc, pb[b]);
    }
}
// fvScalarMatrix::solve prologue: diag[faceCells] += internalCoeffs ; source[faceCells] += boundaryCoeffs (non-coupled)
static __global__ void k_add_boundary(MeshView mv, const double* __restrict__ ic, const double* __restrict__ bcv,
                                      double* __restrict__ diag, double* __restrict__ source, int nBRows, const int* __restrict__ bRowOfSlot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nBRows) return;
    int row = bRowOfSlot[i];
    double d = diag[row], sc = source[row];
    for (int q = mv.bRowStart[i]; q < mv.bRowStart[i + 1]; q++) {
        int b = mv.bRowFaces[q];
        d = d + ic[b];
        if (!mv.bCoupled[b]) sc = sc + bcv[b];
    }
    diag[row] = d; source[row] = sc;
}
// flux = upper*psi[u] - lower*psi[l] ; boundary: internalCoeffs*psi_c - boundaryCoeffs.  SUB: phi -= flux
template <bool SUB>
static __global__ void k_matrix_flux(MeshView mv, const double* __restrict__ upper, const double* __restrict__ lower,
                                     const double* __restrict__ ic, const double* __restrict__ bcv, const double* __restrict__ psi,
                                     const double* __restrict__ psib /*patchNeighbourField on processor faces (may be null without them)*/,
                                     double* __restrict__ out, const double* __restrict__ ffc /*faceFluxCorrection or null*/) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        double a = psi[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            double fl = upper[pos] * psi[u] - lower[pos] * a;
            if (ffc) fl = fl + ffc[pos];      // fieldFlux += *faceFluxCorrectionPtr_ [UPSTREAM fvMatrix::flux]
            if (SUB) out[pos] = out[pos] - fl; else out[pos] = fl;
        END_FACES
    }
    if (row < mv.B) {
        int b = row, p = mv.Fs + b;
        double fl = ic[b] * psi[mv.bFaceRow[b]] - (mv.bCoupled[b] ? bcv[b] * psib[b] : bcv[b]);
        if (ffc) fl = fl + ffc[p];
        if (SUB) out[p] = out[p] - fl; else out[p] = fl;
    }
}

// ---- non-orthogonal correction [UPSTREAM correctedSnGrad::correction, gaussLaplacianScheme::fvmLaplacian] -------
// corr_f = corrVec_f & linearInterpolate(grad(vf))_f  (grad = the case's gradScheme: Gauss linear); physical boundary faces: 0;
// processor faces: the internal-face formula with the neighbour-rank cell's gradient (gbn, halo).
__device__ __forceinline__ double nonOrthCorr(const MeshView& mv, int pos, double gPx, double gPy, double gPz, double gNx, double gNy, double gNz) {
    const double w = mv.w[pos];
    const double fx = w * (gPx - gNx) + gNx, fy = w * (gPy - gNy) + gNy, fz = w * (gPz - gNz) + gNz;
    return (mv.cvx[pos] * fx + mv.cvy[pos] * fy) + mv.cvz[pos] * fz;
}
// out = scale_f * corr_f with scale = gammaf*magSf (faceFluxCorrection of fvm::laplacian(gamma, vf)); gammaf == null: scale = 1
static __global__ void __launch_bounds__(BLOCK) k_nonorth_ffc(MeshView mv, const double* __restrict__ gammaf, const double* __restrict__ gx,
                                                               const double* __restrict__ gy, const double* __restrict__ gz,
                                                               const double* __restrict__ gbn, double* __restrict__ out) {
    ROW_PROLOGUE(mv)
    if (row < mv.N) {
        const double ax = gx[row], ay = gy[row], az = gz[row];
        FOR_OWN_FACES(mv, s, lane, pos, u)
            const double c = nonOrthCorr(mv, pos, ax, ay, az, gx[u], gy[u], gz[u]);
            out[pos] = gammaf ? (gammaf[pos] * mv.magSf[pos]) * c : c;
        END_FACES
    }
    if (row < mv.B) {
        const int b = row, p = mv.Fs + b;
        double v = 0.0;
        if (mv.bCoupled[b]) {
            const int c0 = mv.bFaceRow[b];
            const double c = nonOrthCorr(mv, p, gx[c0], gy[c0], gz[c0], gbn[b], gbn[mv.B + b], gbn[2 * mv.B + b]);
            v = gammaf ? (gammaf[p] * mv.magSf[p]) * c : c;
        }
        out[p] = v;
    }
}

// ---- fvc::reconstruct [UPSTREAM fvcReconstruct.C; SymmTensorI.H inv] -------------------------------------------
static __global__ void k_recon_tensor(MeshView mv, double* __restrict__ invT /*6 SoA arrays of N*/) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    double T[6] = {0, 0, 0, 0, 0, 0};
#define RT_ADD(p)                                                                            \
    {                                                                                        \
        double sx = mv.Sfx[p], sy = mv.Sfy[p], sz = mv.Sfz[p], ms = mv.magSf[p];             \
        T[0] += sx * sx / ms; T[1] += sx * sy / ms; T[2] += sx * sz / ms;                    \
        T[3] += sy * sy / ms; T[4] += sy * sz / ms; T[5] += sz * sz / ms;                    \
    }
    FOR_NBR_FACES(mv, s, lane, pos, l) (void)l; RT_ADD(pos) END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; RT_ADD(pos) END_FACES
    FOR_BND_FACES(mv, s, lane, b) RT_ADD(mv.Fs + b) END_BND
#undef RT_ADD
    double xx = T[0], xy = T[1], xz = T[2], yy = T[3], yz = T[4], zz = T[5];
    double det = xx * yy * zz + xy * yz * xz + xz * xy * yz - xx * yz * yz - xy * xy * zz - xz * yy * xz;
    size_t N = mv.N;
    invT[row] = (yy * zz - yz * yz) / det; invT[N + row] = (xz * yz - xy * zz) / det; invT[2 * N + row] = (xy * yz - xz * yy) / det;
    invT[3 * N + row] = (xx * zz - xz * xz) / det; invT[4 * N + row] = (xy * xz - xx * yz) / det; invT[5 * N + row] = (xx * yy - xy * xy) / det;
}
// rec = invT & surfaceSum((Sf/magSf)*ssf), ssf = (phi - phiU)/rUAf when phiU != null else ssf = phi.
// MODE 0: out = rec ; MODE 1: out += rUA*rec  (U += rUA*fvc::reconstruct(...), [REF region3d/pEqn.H:47])
template <int MODE>
static __global__ void __launch_bounds__(BLOCK) k_reconstruct(MeshView mv, const double* __restrict__ invT, const double* __restrict__ phi,
                                                               const double* __restrict__ phiU, const double* __restrict__ rUAf,
                                                               const double* __restrict__ rUA, double* __restrict__ ox,
                                                               double* __restrict__ oy, double* __restrict__ oz) {
    ROW_PROLOGUE(mv)
    if (row >= mv.N) return;
    PREFETCH_OWN_FACES(mv, s, lane, phi)
    PREFETCH_OWN_FACES(mv, s, lane, phiU)
    PREFETCH_OWN_FACES(mv, s, lane, rUAf)
    PREFETCH_OWN_FACES(mv, s, lane, mv.magSf)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfx)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfy)
    PREFETCH_OWN_FACES(mv, s, lane, mv.Sfz)
    double sx = 0.0, sy = 0.0, sz = 0.0;
#define RC_ADD(p)                                                                  \
    {                                                                              \
        double v = phiU ? (phi[p] - phiU[p]) / rUAf[p] : phi[p];                   \
        double ms = mv.magSf[p];                                                   \
        sx += (mv.Sfx[p] / ms) * v; sy += (mv.Sfy[p] / ms) * v; sz += (mv.Sfz[p] / ms) * v; \
    }
    FOR_NBR_FACES(mv, s, lane, pos, l) (void)l; RC_ADD(pos) END_FACES
    FOR_OWN_FACES(mv, s, lane, pos, u) (void)u; RC_ADD(pos) END_FACES
    FOR_BND_FACES(mv, s, lane, b) RC_ADD(mv.Fs + b) END_BND
#undef RC_ADD
    size_t N = mv.N;
    double t0 = invT[row], t1 = invT[N + row], t2 = invT[2 * N + row], t3 = invT[3 * N + row], t4 = invT[4 * N + row], t5 = invT[5 * N + row];
    double r0 = t0 * sx + t1 * sy + t2 * sz, r1 = t1 * sx + t3 * sy + t4 * sz, r2 = t2 * sx + t4 * sy + t5 * sz;
    if (MODE == 0) { ox[row] = r0; oy[row] = r1; oz[row] = r2; }
    else { double ra = rUA[row]; ox[row] += ra * r0; oy[row] += ra * r1; oz[row] += ra * r2; }
}

static __global__ void k_soa_to_aos(const double* __restrict__ src, double* __restrict__ dst, int n, int nc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) for (int k = 0; k < nc; k++) dst[(size_t)nc * i + k] = src[(size_t)k * n + i];
}
static __global__ void k_gather_patch(const int* __restrict__ rows, int n, const double* __restrict__ f, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = f[rows[i]];
}

// ---- a14: alpha statistics [REF region3d/alphaEqn.H:36-40] ------------------------------------------------------
static __global__ void k_alpha_stats(MeshView mv, const double* __restrict__ a, const double* __restrict__ ab, double* out /*[4]: sumVa,sumV,min,max*/,
                                     double* partials, unsigned* counter) {
    ROW_PROLOGUE(mv)
    double v[2] = {0.0, 0.0}, mn = 1.0e300, mx = -1.0e300;
    if (row < mv.N) { double x = a[row], V = mv.V[row]; v[0] = V * x; v[1] = V; mn = x; mx = x; }
    if (row < mv.B) { double x = ab[row]; mn = fmin(mn, x); mx = fmax(mx, x); }
    gridReduce<2, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    double w[2] = {mn, mx};
    gridReduce<2, OpMin, OpMax>(w, blockIdx.x, gridDim.x, partials + 2 * (size_t)gridDim.x, counter, nullptr);
    // two finalize passes share the counter: do them as one 4-value pass by hand
    __shared__ double sh[32];
    __shared__ int isLast;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); unsigned t = atomicInc(counter, gridDim.x - 1); isLast = (t == gridDim.x - 1); }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const volatile double* vp = partials;
    double a0 = 0.0, a1 = 0.0, a2 = 1.0e300, a3 = -1.0e300;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) {
        a0 += vp[2 * (size_t)j]; a1 += vp[2 * (size_t)j + 1];
        a2 = fmin(a2, vp[2 * (size_t)gridDim.x + 2 * (size_t)j]); a3 = fmax(a3, vp[2 * (size_t)gridDim.x + 2 * (size_t)j + 1]);
    }
    a0 = blockReduce<OpSum>(a0, sh); a1 = blockReduce<OpSum>(a1, sh); a2 = blockReduce<OpMin>(a2, sh); a3 = blockReduce<OpMax>(a3, sh);
    if (threadIdx.x == 0) { out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3; }
}

// ---- a11: fvMatrix::setReference [REF region3d/pEqn.H:30] [UPSTREAM fvMatrix.C setReference]:
//   source[celli] += diag[celli]*value ; diag[celli] += diag[celli]   (before addBoundaryDiag / addBoundarySource)
static __global__ void k_set_reference(int row, double value, double* __restrict__ diag, double* __restrict__ source) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { double d = diag[row]; source[row] = source[row] + d * value; diag[row] = d + d; }
}

// ---- f3: Courant number [REF include/CourantNoMultiRegion.H:78-92]: SfUfbyDelta = deltaCoeffs*|phi| ;
//   out = {sum(SfUfbyDelta) over internal faces, sum(magSf) over internal faces, max(SfUfbyDelta/magSf), max(|phi|/magSf)}
//   (max(GeometricField) takes internal and boundary faces, sum(GeometricField) the internal field only [UPSTREAM GeometricFieldFunctions])
static __global__ void k_courant(MeshView mv, const double* __restrict__ phi, double* out /*[4]*/, double* partials, unsigned* counter) {
    ROW_PROLOGUE(mv)
    double v[2] = {0.0, 0.0}, m0 = 0.0, m1 = 0.0;
    if (row < mv.N) {
        FOR_OWN_FACES(mv, s, lane, pos, u)
            (void)u;
            double mp = fabs(phi[pos]), ms = mv.magSf[pos], sd = mv.dc[pos] * mp;
            v[0] = v[0] + sd; v[1] = v[1] + ms;
            m0 = fmax(m0, sd / ms); m1 = fmax(m1, mp / ms);
        END_FACES
    }
    if (row < mv.B) {
        int p = mv.Fs + row;
        double mp = fabs(phi[p]), ms = mv.magSf[p];
        m0 = fmax(m0, mv.dc[p] * mp / ms); m1 = fmax(m1, mp / ms);
    }
    gridReduce<2, OpSum>(v, blockIdx.x, gridDim.x, partials, counter, nullptr);
    double w[2] = {m0, m1};
    gridReduce<2, OpMax>(w, blockIdx.x, gridDim.x, partials + 2 * (size_t)gridDim.x, counter, nullptr);
    __shared__ double sh[32];
    __shared__ int isLast;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); unsigned t = atomicInc(counter, gridDim.x - 1); isLast = (t == gridDim.x - 1); }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    const volatile double* vp = partials;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) {
        a0 += vp[2 * (size_t)j]; a1 += vp[2 * (size_t)j + 1];
        a2 = fmax(a2, vp[2 * (size_t)gridDim.x + 2 * (size_t)j]); a3 = fmax(a3, vp[2 * (size_t)gridDim.x + 2 * (size_t)j + 1]);
    }
    a0 = blockReduce<OpSum>(a0, sh); a1 = blockReduce<OpSum>(a1, sh); a2 = blockReduce<OpMax>(a2, sh); a3 = blockReduce<OpMax>(a3, sh);
    if (threadIdx.x == 0) { out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3; }
}

}  // namespace b200
