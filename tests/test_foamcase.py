"""OpenFOAM on-disk formats: the case writer round-trips bit-exactly, the decomposed layout is
consistent with decomposePar's addressing files, and the log parser reads lduSolverPerformance::print lines."""
import os

import numpy as np
import pytest

import bench
from shallowinterfoam_b200 import capi, foamcase as fc


@pytest.fixture(scope="module")
def hostlib():
    return capi.load()   # host tools need no device


def _case(hostlib, dims, lengths, weir):
    pm = hostlib.polymesh_box(*dims, *lengths, weir=weir)
    g = pm.geometry()
    F = pm.nInternalFaces
    slices = {n: pm.patch_slice(n) for n in pm.patchNames}
    cs = bench.channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, lengths)
    ctl = bench.set_controls(capi.Region3dControls())
    return pm, cs, ctl


@pytest.mark.parametrize("dims,weir", [((6, 4, 5), None), ((12, 5, 7), (0.45, 0.6, 0.3))])
def test_case_round_trip_bit_exact(hostlib, tmp_path, dims, weir):
    lengths = (5.0, 1.0, 2.0)
    pm, cs, ctl = _case(hostlib, dims, lengths, weir)
    case = str(tmp_path / "case")
    fc.write_case(case, pm, cs, ctl, bench.DELTA_T, 5)
    m = fc.read_polymesh(os.path.join(case, "constant", "region3d", "polyMesh"))
    assert m["nCells"] == pm.nCells
    assert np.array_equal(m["points"], pm.points)              # repr() round-trips every double
    assert np.array_equal(np.array(m["faces"]), pm.faces)
    assert np.array_equal(m["owner"], pm.owner) and np.array_equal(m["neighbour"], pm.neighbour)
    assert [(n, s, z) for (n, s, z, t, q) in m["patches"]] == pm.patches
    types = {n: t for (n, s, z, t, q) in m["patches"]}
    assert types["inlet"] == "patch" and types["outlet"] == "patch" and types["atmosphere"] == "patch"
    assert all(t == "wall" for n, t in types.items() if n not in ("inlet", "outlet", "atmosphere"))
    for name, key in (("alpha1", "alpha"), ("U", "U"), ("pd", "pd")):
        v = fc.read_field(os.path.join(case, "0", "region3d", name), pm.nCells)
        assert np.array_equal(v, cs[key])
    txt = open(os.path.join(case, "0", "region3d", "U")).read()
    assert "fixedValue" in txt and "zeroGradient" in txt and "volVectorField" in txt
    # the dictionaries the reference reads (regionProperties.C, readRegion3dPISOControls.H, alphaEqnSubCycle.H:1-9)
    rp = open(os.path.join(case, "constant", "regionProperties")).read()
    assert "region3dNames   (region3d);" in rp and "region2dNames   ();" in rp
    sol = open(os.path.join(case, "system", "region3d", "fvSolution")).read()
    for key in ("nCorrectors     3;", "nAlphaSubCycles 2;", "nAlphaCorr      1;", "momentumPredictor no;", "preconditioner DIC", "pdFinal"):
        assert key in sol
    sch = open(os.path.join(case, "system", "region3d", "fvSchemes")).read()
    assert "div(phi,alpha)  Gauss vanLeer;" in sch and "div(phirb,alpha) Gauss interfaceCompression;" in sch
    for f in ("g", "transportProperties", "turbulenceProperties"):
        assert os.path.exists(os.path.join(case, "constant", "region3d", f))


def test_decomposed_layout_matches_decomposepar_addressing(hostlib, tmp_path):
    dims, lengths = (8, 4, 6), (5.0, 1.0, 2.0)
    pm, cs, ctl = _case(hostlib, dims, lengths, None)
    case = str(tmp_path / "case")
    proc = fc.write_decomposed(case, pm, (2, 1, 2))
    seen = np.zeros(pm.nCells, dtype=int)
    for r in range(4):
        d = os.path.join(case, "processor%d" % r, "constant", "region3d", "polyMesh")
        m = fc.read_polymesh(d)
        cpa = np.array(open(os.path.join(d, "cellProcAddressing")).read().split("(")[-1].split(")")[0].split(), dtype=int)
        fpa = np.array(open(os.path.join(d, "faceProcAddressing")).read().split("(")[-1].split(")")[0].split(), dtype=int)
        assert len(cpa) == m["nCells"] and np.all(proc[cpa] == r)
        seen[cpa] += 1
        gf = np.abs(fpa) - 1                     # decomposePar: global face + 1, negative = flipped
        # every local face's owner cell maps to the global face's owner (or neighbour when flipped)
        own_g = np.where(fpa > 0, pm.owner[gf], -1)
        nI = pm.nInternalFaces
        for lf in range(len(fpa)):
            c = cpa[m["owner"][lf]]
            if fpa[lf] > 0:
                assert c == pm.owner[gf[lf]]
            else:
                assert gf[lf] < nI and c == pm.neighbour[gf[lf]]
        # processor patches come last, name both ranks, and pair up with the neighbour's patch of the same size
        for (name, start, size, typ, q) in m["patches"]:
            if typ == "processor":
                assert name == "procBoundary%dto%d" % (r, q)
                other = fc.read_polymesh(os.path.join(case, "processor%d" % q, "constant", "region3d", "polyMesh"))
                match = [p for p in other["patches"] if p[3] == "processor" and p[4] == r]
                assert len(match) == 1 and match[0][2] == size
                # same global faces in the same order on both sides
                fq = np.array(open(os.path.join(case, "processor%d" % q, "constant", "region3d", "polyMesh", "faceProcAddressing")).read()
                              .split("(")[-1].split(")")[0].split(), dtype=int)
                assert np.array_equal(np.abs(fpa[start:start + size]), np.abs(fq[match[0][1]:match[0][1] + size]))
    assert np.all(seen == 1)


def test_log_parser_reads_solver_performance_lines():
    log = """Time = 0.002

MULES: Solving for alpha1
Liquid phase volume fraction = 0.5  Min(alpha1) = 0  Max(alpha1) = 1
DICPCG:  Solving for pd, Initial residual = 1, Final residual = 0.0456789, No Iterations 12
DICPCG:  Solving for pd, Initial residual = 0.0123456, Final residual = 0.000567, No Iterations 9
DICPCG:  Solving for pd, Initial residual = 0.00234, Final residual = 9.87654e-08, No Iterations 131
time step continuity errors : sum local = 1e-10, global = 1e-12, cumulative = 1e-12
Time = 0.004

PCG:  Solving for pcorr, Initial residual = 1, Final residual = 1e-8, No Iterations 40
b200PCG:  Solving for pd, Initial residual = 0.5, Final residual = 0.02, No Iterations 11
"""
    r = fc.parse_log(log, "pd")
    assert [x[4] for x in r] == [12, 9, 131, 11]
    assert r[0][0] == "0.002" and r[3][0] == "0.004" and r[3][1] == "b200PCG"
    assert r[2][3] == 9.87654e-08
    assert fc.parse_log(log, "pcorr")[0][4] == 40
    line = fc.format_log_line("b200PCG", "pd", 0.5, 0.02, 11)
    assert fc.parse_log(line + "\n", "pd")[0][2:] == (0.5, 0.02, 11)


def test_check_command_against_a_foam_extend_style_run(lib, oracle, tmp_path, capsys):
    """f4, the `check` leg: `foamcase check CASE` replays the case through the library and compares it with foam-extend's log and
    written fields.  No foam-extend exists here, so the stand-in for its output is produced by the CPU ORACLE in foam-extend's own
    formats (lduSolverPerformance::print lines under `Time = ...` headers; ASCII volScalarField / volVectorField files in the last time
    directory): what passes here is the whole pinning pipeline -- case writer, log parser, field reader, comparison, verdict -- and a
    doctored log / field must make it fail."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cases import oracle_mesh_of
    dims, lengths, steps = (20, 6, 10), (5.0, 1.0, 2.0), 3
    case = str(tmp_path / "case")
    argv = [case, "--dims"] + [str(d) for d in dims] + ["--lengths"] + [str(x) for x in lengths] + ["--steps", str(steps)]
    assert fc.main(["write"] + argv, lib=lib) == 0
    # "foam-extend": the oracle, serial, same controls as the case's dictionaries
    pm = lib.polymesh_box(*dims, *lengths)
    m = oracle_mesh_of(pm)
    g = pm.geometry(); F = pm.nInternalFaces
    cs = bench.channel_fields(pm.patchNames, g["C"], g["Cf"][F:], {n: pm.patch_slice(n) for n in pm.patchNames}, lengths)
    ro = oracle.Region(m, bench.set_controls(oracle.OrcControls()), oracle.BC(m, cs["tA"], 1, cs["rvA"]), oracle.BC(m, cs["tU"], 3, cs["rvU"]),
                       oracle.BC(m, cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    ro.correct_phi()
    log = "Create time\n\n" + fc.format_log_line("DICPCG", "pcorr", *ro.perfs()[0][:3]) + "\n"
    ro.clear_log()
    dt = bench.DELTA_T
    for s_ in range(steps):
        ro.step()
        log += "\nTime = %s\n\nMULES: Solving for alpha1\n" % fc._fmt(dt * (s_ + 1))
        for p in ro.perfs()[-3:]:
            log += fc.format_log_line("DICPCG", "pd", p[0], p[1], p[2]) + "\n"
        log += "ExecutionTime = 0 s\n"
    open(os.path.join(case, "log"), "w").write(log)
    fo_ = ro.fields()
    tdir = os.path.join(case, fc._fmt(dt * steps).rstrip("0").rstrip("."), "region3d")
    patches = pm.patches
    zg = [fc.ZG] * pm.nPatches
    for name, arr in (("alpha1", fo_["alpha1"]), ("pd", fo_["pd"]), ("U", fo_["U"])):
        fc.write_field(os.path.join(tdir, name), name, arr, patches, F, zg, None, os.path.basename(os.path.dirname(tdir)) + "/region3d")
    capsys.readouterr()
    assert fc.main(["check"] + argv, lib=lib) == 0
    out = capsys.readouterr().out
    assert "PARITY PINNED" in out and out.count(" ok") >= 3 * steps + 3 and "DIFFERS" not in out
    # a log whose iteration count is off by one, and a field that is off in the 6th digit, must both be caught
    bad = log.replace("No Iterations %d" % ro.perfs()[-1][2], "No Iterations %d" % (ro.perfs()[-1][2] + 1), 1)
    open(os.path.join(case, "log"), "w").write(bad)
    assert fc.main(["check"] + argv, lib=lib) == 1 and "DIFFERS" in capsys.readouterr().out
    open(os.path.join(case, "log"), "w").write(log)
    fc.write_field(os.path.join(tdir, "pd"), "pd", fo_["pd"] * (1 + 1e-6), patches, F, zg, None, "x/region3d")
    assert fc.main(["check"] + argv, lib=lib) == 1 and "parity NOT pinned" in capsys.readouterr().out
