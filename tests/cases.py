"""Shared helpers for the parity tests: oracle mesh from a product PolyMesh, random SPD LDU matrices, channel case."""
import numpy as np
from oracle import foam_oracle as fo
from oracle import mesh_oracle as mo


def oracle_mesh_of(pm):
    """OrcMesh fed with the SAME arrays the product uses (after test_mesh proves they equal the oracle's own)."""
    g = pm.geometry()
    kinds = [1 if r >= 0 else 0 for r in pm.patchNeighbRank]
    return fo.Mesh(pm.owner, pm.neighbour, pm.patches, g, pm.nCells, kinds)


def random_spd(m, seed=1, shift=(0.01, 0.1)):
    rng = np.random.default_rng(seed)
    upper = -rng.uniform(0.5, 1.5, m.F)
    diag = np.zeros(m.N)
    np.add.at(diag, m.owner[:m.F], -upper)
    np.add.at(diag, m.neighbour, -upper)
    diag += rng.uniform(shift[0], shift[1], m.N)
    return diag, upper


def channel_case(pm, lx, ly, lz, fo_mod=None):
    """Initial fields + BCs of the synthetic channel: patches inlet/outlet/walls/atmosphere[/weir]."""
    g = pm.geometry()
    C = g["C"]; F = pm.nInternalFaces; B = pm.nBoundaryFaces
    Cfb = g["Cf"][F:]
    alpha = np.where(C[:, 2] < 0.5 * lz + 0.05 * lz * np.sin(2 * np.pi * C[:, 0] / lx), 1.0, 0.0)
    U = np.zeros((pm.nCells, 3)); U[:, 0] = alpha
    pd = np.zeros(pm.nCells)
    FV, ZG = 0, 1
    names = pm.patchNames
    tA, tU, tP = [], [], []
    for n in names:
        if n == "inlet": tA.append(FV); tU.append(FV); tP.append(ZG)
        elif n == "outlet": tA.append(ZG); tU.append(ZG); tP.append(FV)
        elif n == "atmosphere": tA.append(ZG); tU.append(ZG); tP.append(FV)
        else: tA.append(ZG); tU.append(FV); tP.append(ZG)   # walls, weir
    rvA = np.zeros(B); rvU = np.zeros((B, 3))
    sl = pm.patch_slice("inlet")
    rvA[sl] = np.where(Cfb[sl, 2] < 0.5 * lz, 1.0, 0.0)
    rvU[sl, 0] = rvA[sl]
    return dict(alpha=alpha, U=U, pd=pd, tA=tA, tU=tU, tP=tP, rvA=rvA, rvU=rvU)


def default_controls(cls, deltaT=0.005):
    c = cls()
    c.deltaT = deltaT; c.nCorr = 3; c.nNonOrthCorr = 0; c.nAlphaCorr = 1; c.nAlphaSubCycles = 2; c.cAlpha = 1.0
    c.rho1 = 1000.0; c.rho2 = 1.0; c.nu1 = 1e-6; c.nu2 = 1.48e-5; c.sigma = 0.07
    c.g[0] = 0.0; c.g[1] = 0.0; c.g[2] = -9.81
    c.pdTol = 1e-7; c.pdRelTol = 0.05; c.pdFinalTol = 1e-7; c.pdFinalRelTol = 0.0; c.pcorrTol = 1e-10; c.pcorrRelTol = 0.0
    c.maxIter = 1000; c.minIter = 0; c.adjustPhi = 1; c.mulesIncludeOwn = 0; c.divUScheme = 2
    return c
