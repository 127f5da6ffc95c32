"""CPU ORACLE (test infrastructure only; PARITY UNPINNED) -- N-rank emulation of the decomposed lduMatrix solve:
processor-patch halo swaps as explicit arrays, rank-local (block-Jacobi) DIC exactly like foam-extend under MPI,
gSum* = per-rank sequential sums added in rank order.  [UPSTREAM PCG.C, processorFvPatchField::updateInterfaceMatrix]"""
import numpy as np
from . import foam_oracle as fo


def halo(subs, fields):
    """patchNeighbourField for every rank: out[r][b] = fields[q][faceCell of the matching face on rank q]"""
    out = [np.zeros(s["mesh"].B) for s in subs]
    for r, s in enumerate(subs):
        for (name, start, size, q) in s["patches"]:
            if q < 0 or size == 0:
                continue
            t = subs[q]
            pq = [p for p in t["patches"] if p[3] == r][0]
            cells_q = t["mesh"].owner[pq[1]:pq[1] + pq[2]]
            out[r][start - s["mesh"].F:start - s["mesh"].F + size] = fields[q][cells_q]
    return out


def amul_multi(subs, diags, uppers, bous, psis):
    nbr = halo(subs, psis)
    return [fo.amul(s["mesh"], diags[r], uppers[r], uppers[r], psis[r], bous[r], nbr[r]) for r, s in enumerate(subs)]


def pcg_solve_multi(subs, diags, uppers, bous, x0s, bs, tol=1e-6, relTol=0.0, minIter=0, maxIter=1000):
    R = len(subs)
    x = [np.array(v, dtype=float) for v in x0s]
    gsum = lambda parts: float(sum(parts))          # rank-ordered reduce of per-rank sequential sums
    seq = lambda a: float(np.add.reduce(a)) if False else float(fo.np.cumsum(a)[-1]) if a.size else 0.0   # sequential sum
    wA = amul_multi(subs, diags, uppers, bous, x)
    rA = [bs[r] - wA[r] for r in range(R)]
    sumA = [fo.sumA(subs[r]["mesh"], diags[r], uppers[r], uppers[r], bous[r]) for r in range(R)]
    nTot = sum(s["mesh"].N for s in subs)
    xRef = gsum([seq(x[r]) for r in range(R)]) / nTot
    nf = gsum([seq(np.abs(wA[r] - xRef * sumA[r]) + np.abs(bs[r] - xRef * sumA[r])) for r in range(R)]) + 1e-20
    res = gsum([seq(np.abs(rA[r])) for r in range(R)]) / nf
    init = res
    hist = [res]
    nIter = 0

    def stop():
        if nIter < minIter:
            return False
        return nIter >= maxIter or res < tol or (relTol > 0 and res < relTol * init)

    if not stop():
        rD = [fo.dic_calc_rd(subs[r]["mesh"], diags[r], uppers[r]) for r in range(R)]
        wArA = 1e20
        pA = [np.zeros_like(v) for v in x]
        while True:
            wArAold = wArA
            wA = [fo.dic_precondition(subs[r]["mesh"], rD[r], uppers[r], rA[r]) for r in range(R)]
            wArA = gsum([seq(wA[r] * rA[r]) for r in range(R)])
            if nIter == 0:
                pA = [w.copy() for w in wA]
            else:
                beta = wArA / wArAold
                pA = [wA[r] + beta * pA[r] for r in range(R)]
            wA = amul_multi(subs, diags, uppers, bous, pA)
            wApA = gsum([seq(wA[r] * pA[r]) for r in range(R)])
            alpha = wArA / wApA
            for r in range(R):
                x[r] = x[r] + alpha * pA[r]
                rA[r] = rA[r] - alpha * wA[r]
            res = gsum([seq(np.abs(rA[r])) for r in range(R)]) / nf
            nIter += 1
            hist.append(res)
            if stop():
                break
    return x, dict(initialResidual=init, finalResidual=res, nIterations=nIter), np.array(hist)
