// device_mesh.cu -- build the renumbered owner-slot SELL-32 layout on the host, upload it, and convert fields
// between foam-extend's host layout and the device layout.
#include "device_mesh.cuh"
#include <cmath>
#include <algorithm>
#include <algorithm>

namespace b200 {

// ------------------------------------------------------------------------------------------------ kernels
__global__ void k_fill_bits(unsigned long long* p, size_t n, unsigned long long bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = bits;
}
__global__ void k_gather(const double* __restrict__ src, const int* __restrict__ idx, double* __restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { int j = idx[i]; dst[i] = j >= 0 ? src[j] : 0.0; }
}
__global__ void k_gather_vec(const double* __restrict__ srcXYZ, const int* __restrict__ idx, double* __restrict__ dx,
                             double* __restrict__ dy, double* __restrict__ dz, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int j = idx[i];
        dx[i] = j >= 0 ? srcXYZ[3 * (size_t)j] : 0.0;
        dy[i] = j >= 0 ? srcXYZ[3 * (size_t)j + 1] : 0.0;
        dz[i] = j >= 0 ? srcXYZ[3 * (size_t)j + 2] : 0.0;
    }
}
__global__ void k_scatter_vec_to_aos(const double* __restrict__ dx, const double* __restrict__ dy,
                                     const double* __restrict__ dz, const int* __restrict__ idx,
                                     double* __restrict__ dstXYZ, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = original cell; idx[i] = row
    if (i < n) { int r = idx[i]; dstXYZ[3 * (size_t)i] = dx[r]; dstXYZ[3 * (size_t)i + 1] = dy[r]; dstXYZ[3 * (size_t)i + 2] = dz[r]; }
}
__global__ void k_copy_vec_b(const double* __restrict__ srcXYZ, double* __restrict__ dx, double* __restrict__ dy,
                             double* __restrict__ dz, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { dx[i] = srcXYZ[3 * (size_t)i]; dy[i] = srcXYZ[3 * (size_t)i + 1]; dz[i] = srcXYZ[3 * (size_t)i + 2]; }
}

double nonOrthogonalityDegrees(double sumMagSfCorr, double sumMagSf) {
    if (!(sumMagSf > 0.0)) return 0.0;
    return std::asin(std::min(sumMagSfCorr / sumMagSf, 1.0)) * 180.0 / 3.14159265358979323846;
}

void fillSentinel(double* p, size_t n) {
    if (n) B200_LAUNCH(k_fill_bits, gridFor((long long)n), BLOCK, rt().stream, (unsigned long long*)p, n, SENTINEL_BITS);
}
void fillValue(double* p, size_t n, double v) {
    unsigned long long bits; memcpy(&bits, &v, 8);
    if (n) B200_LAUNCH(k_fill_bits, gridFor((long long)n), BLOCK, rt().stream, (unsigned long long*)p, n, bits);
}

// -------------------------------------------------------------------------------------------------- build
void DeviceMesh::build(int32_t nCells, int32_t nFaces, const int32_t* lo, const int32_t* up,
                       const std::vector<Patch>& pt, const std::vector<int32_t>& fc) {
    N = nCells; F = nFaces; patches = pt; faceCells = fc; B = (int)fc.size();
    lower.assign(lo, lo + F); upper.assign(up, up + F);
    for (int f = 0; f < F; f++)
        B200_REQUIRE(lower[f] >= 0 && upper[f] < N, B200_ERR_ARG, "lduAddressing label out of range");
    for (int b = 0; b < B; b++) B200_REQUIRE(fc[b] >= 0 && fc[b] < N, B200_ERR_ARG, "faceCells label out of range");
    int bandLevels = std::min(SWEEP_MAX_STEPS, std::max(1, (int)rt().opt("sweep_band_levels", 16)));
    int lanes = std::max(32, std::min(SWEEP_THREADS, (int)rt().opt("sweep_threads", 256))) / 32 * 32;
    int tileRows = std::max(8, (int)rt().opt("sweep_tile_rows", 4096));
    int smemBudget = std::min(227 * 1024, std::max(4096, (int)rt().opt("sweep_smem_bytes", 225000)));
    try {
        buildSweepPlan(N, lower.data(), upper.data(), F, bandLevels, tileRows, smemBudget, lanes, plan);
    } catch (const std::exception& e) { throw Error(B200_ERR_ARG, e.what()); }
    level = plan.level; rowCell = plan.rowCell; cellRow = plan.cellRow; nLevels = plan.nLevels;
    nTiles = plan.nTiles;
    sweepSmem = std::max(plan.smemFwd, plan.smemBwd);
    lduDerived(N, lower, upper, ownerStart, losort, losortStart);
    levelStart.assign(nLevels + 1, 0);
    for (int c = 0; c < N; c++) levelStart[level[c] + 1]++;
    for (int l = 0; l < nLevels; l++) levelStart[l + 1] += levelStart[l];

    nSlices = (N + 31) / 32;
    std::vector<int> sliceOff(nSlices + 1, 0), nbrOff(nSlices + 1, 0);
    int maxOw = 1, maxNw = 0;
    std::vector<int> wOwn(nSlices, 0), wNbr(nSlices, 0);
    long long sumOw = 0, sumNw = 0;
    for (int s = 0; s < nSlices; s++) {
        int wo = 0, wn = 0;
        for (int r = s * 32; r < std::min(N, s * 32 + 32); r++) {
            int c = rowCell[r];
            wo = std::max(wo, ownerStart[c + 1] - ownerStart[c]);
            wn = std::max(wn, losortStart[c + 1] - losortStart[c]);
        }
        maxOw = std::max(maxOw, wo); maxNw = std::max(maxNw, wn);
        wOwn[s] = wo; wNbr[s] = wn; sumOw += wo; sumNw += wn;
    }
    // Uniform widths: when padding every slice to the widest one costs < 1/8 more slots (hex, prism and tet meshes: only the slices
    // made of boundary cells are narrower), slot positions become arithmetic -- pos = ((row/32)*width + k)*32 + row%32 -- and the
    // gather kernels need no offset look-ups (two dependent loads less per row).  sliceOff / nbrOff stay valid either way.
    uniOw = uniNw = 0;
    if (rt().opt("uniform_slices", 1) > 0 && maxOw <= 4 && maxNw <= 4 && maxNw >= 1 &&
        (long long)nSlices * maxOw * 8 <= sumOw * 9 && (long long)nSlices * maxNw * 8 <= sumNw * 9) {
        uniOw = maxOw; uniNw = maxNw;
        for (int s = 0; s < nSlices; s++) { wOwn[s] = maxOw; wNbr[s] = maxNw; }
    }
    for (int s = 0; s < nSlices; s++) {
        sliceOff[s + 1] = sliceOff[s] + 32 * wOwn[s];
        nbrOff[s + 1] = nbrOff[s] + 32 * wNbr[s];
    }
    Fs = sliceOff[nSlices]; Ls = nbrOff[nSlices];
    KB = 1;
    while ((1 << KB) < maxOw) KB++;
    B200_REQUIRE(((long long)N << KB) < (1ll << 31), B200_ERR_UNSUPPORTED, "mesh too large for packed 32-bit face references");
    std::vector<int> upIdx(Fs, -1), slotFace(Fs, -1), loPk(Ls, -1);
    facePosHost.assign(F, -1);
    for (int r = 0; r < N; r++) {
        int c = rowCell[r], s = r >> 5, lane = r & 31;
        for (int f = ownerStart[c]; f < ownerStart[c + 1]; f++) {
            int k = f - ownerStart[c], pos = sliceOff[s] + k * 32 + lane;
            upIdx[pos] = cellRow[upper[f]]; slotFace[pos] = f; facePosHost[f] = pos;
        }
        for (int j = losortStart[c]; j < losortStart[c + 1]; j++) {
            int f = losort[j], l = lower[f];
            int q = nbrOff[s] + (j - losortStart[c]) * 32 + lane;
            loPk[q] = (cellRow[l] << KB) | (f - ownerStart[l]);
        }
    }
    {   // sweep-plan upload: descriptors, operand-block templates (source / own-row offsets), face slots, slot rows, halo rows
        std::vector<SweepTile> tf(nTiles), tb(nTiles);
        size_t blkF = 0, blkB = 0;
        const int W = plan.lanes;
        for (int t = 0; t < nTiles; t++) {
            SweepTile d; memset(&d, 0, sizeof(d));
            d.row0 = plan.tileStart[t]; d.nRows = plan.tileStart[t + 1] - d.row0;
            d.virt = plan.chunked ? 1 : 0;
            tf[t] = d; tb[t] = d;
            tf[t].slot0 = plan.fwdStepStart[t] * W; tf[t].nSteps = plan.fwdStepStart[t + 1] - plan.fwdStepStart[t];
            tb[t].slot0 = plan.bwdStepStart[t] * W; tb[t].nSteps = plan.bwdStepStart[t + 1] - plan.bwdStepStart[t];
            tf[t].d = plan.fwdD[t]; tf[t].halo = plan.fwdHalo[t]; tf[t].nHalo = plan.fwdHalo[t + 1] - plan.fwdHalo[t];
            tb[t].d = plan.bwdD[t]; tb[t].halo = plan.bwdHalo[t]; tb[t].nHalo = plan.bwdHalo[t + 1] - plan.bwdHalo[t];
            B200_REQUIRE(tf[t].d <= 3 && tb[t].d <= 3, B200_ERR_UNSUPPORTED, "sweep plan entry with more than 3 sources (internal error)");
            tf[t].blk = (int)blkF; tb[t].blk = (int)blkB;
            blkF += (size_t)(tf[t].d + (tf[t].d ? 1 : 0)) * tf[t].nSteps * W; blkB += (size_t)(tb[t].d + (tb[t].d ? 1 : 0)) * tb[t].nSteps * W;
            B200_REQUIRE(blkF < ((size_t)1 << 31) && blkB < ((size_t)1 << 31), B200_ERR_UNSUPPORTED, "mesh too large for 32-bit sweep block offsets");
        }
        fwdBlkSize = blkF; bwdBlkSize = blkB;
        for (int dir = 0; dir < 2; dir++) {
            std::vector<SweepTile>& tl = dir == 0 ? tf : tb;
            const std::vector<uint16_t>& idx = dir == 0 ? plan.fwdIdx : plan.bwdIdx;
            const std::vector<uint16_t>& own = dir == 0 ? plan.fwdOwn : plan.bwdOwn;
            const std::vector<int32_t>& face = dir == 0 ? plan.fwdFace : plan.bwdFace;
            const std::vector<int32_t>& ell = dir == 0 ? plan.fwdEll : plan.bwdEll;
            size_t total = dir == 0 ? blkF : blkB;
            std::vector<double> blk0(total, 0.0);
            std::vector<int> pos(total, -1), slotRow(own.size(), -1);
            for (int t = 0; t < nTiles; t++) {
                SweepTile& d = tl[t];
                const size_t S = (size_t)d.nSteps * W;
                for (size_t sl = 0; sl < S; sl++) {
                    const uint16_t o = own[d.slot0 + sl];
                    if (o != SWEEP_OWN_NONE) slotRow[d.slot0 + sl] = d.row0 + (o & SWEEP_OWN_ROW);
                }
                d.stepWarps[0] = d.stepWarps[1] = 0xffffffffu;
                if (d.nSteps <= 16) {
                    d.stepWarps[0] = d.stepWarps[1] = 0u;
                    for (int k = 0; k < d.nSteps; k++) {
                        int last = -1;
                        for (int i = 0; i < W; i++) if (own[d.slot0 + (size_t)k * W + i] != SWEEP_OWN_NONE) last = i;
                        const unsigned nw = (unsigned)std::min(15, (last + 32) / 32);   // (15 = every warp: 512+ lanes do not fit 4 bits)
                        d.stepWarps[k >> 3] |= nw << (4 * (k & 7));
                    }
                }
                if (!d.virt && d.nSteps == 16) {   // hybrid publish (dic_kernels.cuh bulkPlan): rows of an unchunked tile are contiguous in step order
                    int first4 = 0, first12 = 0;
                    for (size_t sl = 0; sl < S; sl++) if (own[d.slot0 + sl] != SWEEP_OWN_NONE) { if (sl < (size_t)4 * W) first4++; if (sl < (size_t)12 * W) first12++; }
                    if (dir == 0) { d.bulkLo = 0; d.bulkHi = first12; } else { d.bulkLo = first4; d.bulkHi = d.nRows; }
                }
                if (d.d == 0) continue;
                uint16_t* offs = reinterpret_cast<uint16_t*>(blk0.data() + d.blk + (size_t)d.d * S);
                const uint16_t padOff = (uint16_t)(8 * (d.nRows + d.nHalo)), trashOff = (uint16_t)(8 * (d.nRows + d.nHalo + 1));
                for (size_t e = 0; e < S * 4; e++) offs[e] = padOff;
                for (int j = 0; j < d.d; j++)
                    for (size_t sl = 0; sl < S; sl++) {
                        size_t e = (size_t)ell[t] + (size_t)j * S + sl;
                        offs[sl * 4 + j] = (uint16_t)(8 * idx[e]);
                        pos[d.blk + (size_t)j * S + sl] = face[e] >= 0 ? facePosHost[face[e]] : -1;
                    }
                for (size_t sl = 0; sl < S; sl++) {   // lane 3 of the offset word: where the entry's own row lives in the value array (+ chunk flags)
                    const uint16_t o = own[d.slot0 + sl];
                    offs[sl * 4 + 3] = o == SWEEP_OWN_NONE ? trashOff
                        : (uint16_t)(8 * (o & SWEEP_OWN_ROW) + ((o & SWEEP_OWN_PARTIAL) ? OWN_PARTIAL : 0u) + ((o & SWEEP_OWN_CARRY) ? OWN_CARRY : 0u));
                }
            }
            (dir == 0 ? dFwdBlk0 : dBwdBlk0).upload(blk0);
            (dir == 0 ? dFwdPos : dBwdPos).upload(pos);
            (dir == 0 ? dFwdSlotRow : dBwdSlotRow).upload(slotRow);
            syncStream();
        }
        dTilesF.upload(tf); dTilesB.upload(tb);
        dFwdHaloRow.upload(plan.fwdHaloRow); dBwdHaloRow.upload(plan.bwdHaloRow);
        dTrace.alloc(8 * (size_t)nTiles + 8); dTrace.zero();
        syncStream();   // the host vectors above go out of scope
        plan.fwdFace.clear(); plan.fwdFace.shrink_to_fit(); plan.bwdFace.clear(); plan.bwdFace.shrink_to_fit();
        plan.fwdIdx.clear(); plan.fwdIdx.shrink_to_fit(); plan.bwdIdx.clear(); plan.bwdIdx.shrink_to_fit();
        plan.fwdOwn.clear(); plan.fwdOwn.shrink_to_fit(); plan.bwdOwn.clear(); plan.bwdOwn.shrink_to_fit();
    }
    // boundary CSR by row
    std::vector<int> cnt(N, 0);
    for (int b = 0; b < B; b++) cnt[cellRow[fc[b]]]++;
    std::vector<unsigned> bMask(nSlices, 0u);
    std::vector<int> bBase(nSlices, 0), bRowStart(1, 0), bRowFaces(B), rowSlot(N, -1);
    nBRows = 0;
    for (int s = 0; s < nSlices; s++) {
        bBase[s] = nBRows;
        for (int r = s * 32; r < std::min(N, s * 32 + 32); r++)
            if (cnt[r] > 0) { bMask[s] |= (1u << (r & 31)); rowSlot[r] = nBRows++; bRowStart.push_back(bRowStart.back() + cnt[r]); }
    }
    std::vector<int> cur(bRowStart.begin(), bRowStart.end() - 1), bFaceRow(B), bPatch(B, 0), bCoupled(B, 0);
    for (int b = 0; b < B; b++) { int r = cellRow[fc[b]]; bFaceRow[b] = r; bRowFaces[cur[rowSlot[r]]++] = b; }
    for (size_t p = 0; p < patches.size(); p++)
        for (int f = patches[p].start; f < patches[p].start + patches[p].size; f++) {
            B200_REQUIRE(f - F >= 0 && f - F < B, B200_ERR_ARG, "patch range outside the boundary faces");
            bPatch[f - F] = (int)p; bCoupled[f - F] = patches[p].neighbRank >= 0 ? 1 : 0;
        }
    dSliceOff.upload(sliceOff); dNbrOff.upload(nbrOff); dUpIdx.upload(upIdx); dLoPk.upload(loPk);
    dSlotFace.upload(slotFace); dFacePos.upload(facePosHost); dRowCell.upload(rowCell); dCellRow.upload(cellRow);
    dBMask.upload(bMask); dBBase.upload(bBase); dBRowStart.upload(bRowStart); dBRowFaces.upload(bRowFaces);
    dBFaceRow.upload(bFaceRow); dBPatch.upload(bPatch); dBCoupled.upload(bCoupled);
    syncStream();
}

void DeviceMesh::setGeometry(const double* Sf, const double* magSf, const double* Cf, const double* C,
                             const double* V, const double* w, const double* dc) {
    size_t nf = nFaceField();
    hostV.assign(V, V + N);
    dV.alloc(N); dCx.alloc(N); dCy.alloc(N); dCz.alloc(N);
    dSfx.alloc(nf); dSfy.alloc(nf); dSfz.alloc(nf); dCfx.alloc(nf); dCfy.alloc(nf); dCfz.alloc(nf);
    dMagSf.alloc(nf); dW.alloc(nf); dDc.alloc(nf);
    dCnx.alloc(B); dCny.alloc(B); dCnz.alloc(B); dCnx.zero(); dCny.zero(); dCnz.zero();
    uploadCells(V, dV); uploadCellVec(C, dCx, dCy, dCz);
    uploadFaceVec(Sf, dSfx, dSfy, dSfz); uploadFaceVec(Cf, dCfx, dCfy, dCfz);
    uploadFaces(magSf, dMagSf, true); uploadFaces(w, dW, true); uploadFaces(dc, dDc, true);
    syncStream();
    // non-orthogonal correction vectors (internal faces; boundary faces zero, processor faces in coupleGeometry) and the
    // non-orthogonality measure, both in the sequential face order of upstream's makeCorrectionVectors
    {
        std::vector<double> cv(3 * ((size_t)F + B), 0.0);
        double sMC = 0.0, sM = 0.0;
        for (int f = 0; f < F; f++) {
            const int o = lower[f], n = upper[f];
            double c[3];
            for (int k = 0; k < 3; k++) {
                const double d = C[3 * (size_t)n + k] - C[3 * (size_t)o + k];
                c[k] = Sf[3 * (size_t)f + k] / magSf[f] - d * dc[f];
                cv[3 * (size_t)f + k] = c[k];
            }
            sMC += magSf[f] * std::sqrt((c[0] * c[0] + c[1] * c[1]) + c[2] * c[2]);
            sM += magSf[f];
        }
        nonOrthSum[0] = sMC; nonOrthSum[1] = sM;
        orthogonal = nonOrthogonalityDegrees(sMC, sM) < 0.1;       // single rank; coupleGeometry() makes it global
        dCvx.alloc(nf); dCvy.alloc(nf); dCvz.alloc(nf);
        uploadFaceVec(cv.data(), dCvx, dCvy, dCvz);
        syncStream();
    }
    hasGeometry = true;
}

MeshView DeviceMesh::view() const {
    MeshView v;
    v.N = N; v.F = F; v.B = B; v.nSlices = nSlices; v.KB = KB; v.Fs = Fs; v.uniOw = uniOw; v.uniNw = uniNw;
    v.sliceOff = dSliceOff; v.nbrOff = dNbrOff; v.upIdx = dUpIdx; v.loPk = dLoPk;
    v.bMask = dBMask; v.bBase = dBBase; v.bRowStart = dBRowStart; v.bRowFaces = dBRowFaces;
    v.bFaceRow = dBFaceRow; v.bPatch = dBPatch; v.bCoupled = dBCoupled;
    v.V = dV; v.Cx = dCx; v.Cy = dCy; v.Cz = dCz; v.Sfx = dSfx; v.Sfy = dSfy; v.Sfz = dSfz;
    v.magSf = dMagSf; v.w = dW; v.dc = dDc; v.Cnx = dCnx; v.Cny = dCny; v.Cnz = dCnz;
    v.cvx = dCvx; v.cvy = dCvy; v.cvz = dCvz;
    return v;
}

// ------------------------------------------------------------------------------------ layout conversion
void DeviceMesh::uploadCells(const double* host, double* dev) {
    double* st = stage(stageA, N);
    hostCopyIn(st, host, 8 * (size_t)N);
    if (N) B200_LAUNCH(k_gather, gridFor(N), BLOCK, rt().stream, st, dRowCell.p, dev, N);
}
void DeviceMesh::downloadCells(const double* dev, double* host) {
    double* st = stage(stageA, N);
    if (N) B200_LAUNCH(k_gather, gridFor(N), BLOCK, rt().stream, dev, dCellRow.p, st, N);
    hostCopyOut(host, st, 8 * (size_t)N);
    syncStream();
}
void DeviceMesh::uploadCellVec(const double* hostXYZ, double* dx, double* dy, double* dz) {
    double* st = stage(stageA, 3 * (size_t)N);
    hostCopyIn(st, hostXYZ, 24 * (size_t)N);
    if (N) B200_LAUNCH(k_gather_vec, gridFor(N), BLOCK, rt().stream, st, dRowCell.p, dx, dy, dz, N);
}
void DeviceMesh::downloadCellVec(const double* dx, const double* dy, const double* dz, double* hostXYZ) {
    double* st = stage(stageA, 3 * (size_t)N);
    if (N) B200_LAUNCH(k_scatter_vec_to_aos, gridFor(N), BLOCK, rt().stream, dx, dy, dz, dCellRow.p, st, N);
    hostCopyOut(hostXYZ, st, 24 * (size_t)N);
    syncStream();
}
void DeviceMesh::uploadFaces(const double* host, double* dev, bool withBoundary) {
    size_t n = (size_t)F + (withBoundary ? B : 0);
    double* st = stage(stageA, n);
    hostCopyIn(st, host, 8 * n);
    if (Fs) B200_LAUNCH(k_gather, gridFor(Fs), BLOCK, rt().stream, st, dSlotFace.p, dev, Fs);
    if (withBoundary && B)
        B200_CHECK_CUDA(cudaMemcpyAsync(dev + Fs, st + F, 8 * (size_t)B, cudaMemcpyDeviceToDevice, rt().stream));
}
void DeviceMesh::downloadFaces(const double* dev, double* host, bool withBoundary) {
    size_t n = (size_t)F + (withBoundary ? B : 0);
    double* st = stage(stageA, n);
    if (F) B200_LAUNCH(k_gather, gridFor(F), BLOCK, rt().stream, dev, dFacePos.p, st, F);
    if (withBoundary && B)
        B200_CHECK_CUDA(cudaMemcpyAsync(st + F, dev + Fs, 8 * (size_t)B, cudaMemcpyDeviceToDevice, rt().stream));
    hostCopyOut(host, st, 8 * n);
    syncStream();
}
void DeviceMesh::uploadFaceVec(const double* hostXYZ, double* dx, double* dy, double* dz) {
    size_t n = (size_t)F + B;
    double* st = stage(stageA, 3 * n);
    hostCopyIn(st, hostXYZ, 24 * n);
    if (Fs) B200_LAUNCH(k_gather_vec, gridFor(Fs), BLOCK, rt().stream, st, dSlotFace.p, dx, dy, dz, Fs);
    if (B) B200_LAUNCH(k_copy_vec_b, gridFor(B), BLOCK, rt().stream, st + 3 * (size_t)F, dx + Fs, dy + Fs, dz + Fs, B);
}
void DeviceMesh::uploadBoundary(const double* host, double* dev, int nc) {
    if (B) hostCopyIn(dev, host, 8 * (size_t)B * nc);
}
void DeviceMesh::downloadBoundary(const double* dev, double* host, int nc) {
    if (B) hostCopyOut(host, dev, 8 * (size_t)B * nc);
    syncStream();
}

}  // namespace b200
