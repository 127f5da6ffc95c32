"""foamcase.py -- OpenFOAM/foam-extend on-disk formats for the 3D region.

Host-side case I/O only (no compute): writes the meshes and fields this repo generates as a foam-extend-3.1 multi-region case
that the UNMODIFIED reference solver can run, and reads back what foam-extend writes (field files, the solver log), so that anyone
with a foam-extend-3.1 install can pin the parity of this repo against the real thing with two commands:

    python -m shallowinterfoam_b200.foamcase write  CASE [--dims 50 10 20] [--steps 5]     # here (no GPU needed)
    shallowInterFoam -case CASE > CASE/log                                                 # where foam-extend-3.1 lives
    python -m shallowinterfoam_b200.foamcase check  CASE                                   # on a GPU box: our run vs foam-extend's

Layout written (the reference reads exactly these: shallowInterFoam.C:63-132, region3d/create3dMeshes.H, create3dFields.H:27-95,
include/regionProperties.C, region3d/readRegion3dPISOControls.H, alphaEqnSubCycle.H:1-9):

    constant/regionProperties                       region2dNames (); region3dNames (<region>);
    constant/<region>/polyMesh/{points,faces,owner,neighbour,boundary}
    constant/<region>/{g,transportProperties,turbulenceProperties,RASProperties}
    0/<region>/{alpha1,U,pd}
    system/{controlDict,fvSchemes,fvSolution}       (top-level copies: foam-extend wants them for the default region)
    system/<region>/{fvSchemes,fvSolution,decomposeParDict}
    processor<r>/constant/<region>/polyMesh/...     (write_decomposed: decomposePar `simple` layout incl. *ProcAddressing)

File format: ASCII FoamFile headers + `List<T>` bodies, labels as int32, scalars with 17 significant digits (round-trips doubles).
"""
import os
import re
import sys

import numpy as np

BANNER = "/*--------------------------------*- C++ -*----------------------------------*\\\n" \
         "| foam-extend-3.1 case written by b200-shallowInterFoam-hotpath (foamcase.py)  |\n" \
         "\\*---------------------------------------------------------------------------*/\n"


def _header(cls, obj, location=None, note=None):
    s = BANNER + "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       %s;\n" % cls
    if note:
        s += "    note        \"%s\";\n" % note
    if location:
        s += "    location    \"%s\";\n" % location
    return s + "    object      %s;\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n" % obj


def _fmt(x):
    return repr(float(x))   # shortest string that round-trips the double


def _write(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def _scalar_list(a):
    return "%d\n(\n%s\n)\n" % (len(a), "\n".join(_fmt(v) for v in a))


def _label_list(a):
    return "%d\n(\n%s\n)\n" % (len(a), "\n".join(str(int(v)) for v in a))


def _vector_list(a):
    return "%d\n(\n%s\n)\n" % (len(a), "\n".join("(%s %s %s)" % (_fmt(v[0]), _fmt(v[1]), _fmt(v[2])) for v in a))


# ------------------------------------------------------------------------------------------------------ polyMesh
PATCH_TYPE = {"inlet": "patch", "outlet": "patch", "atmosphere": "patch"}   # everything else: wall (processor patches by neighbRank)


def write_polymesh(mesh_dir, points, faces, owner, neighbour, patches, n_cells, proc_of_patch=None, my_rank=0):
    """points [nP,3]; faces [nF,4] (quads, blockMesh vertex order); owner [nF]; neighbour [nInternal];
    patches: list of (name, start, size); proc_of_patch: neighbour rank per patch or -1 (processor patches)."""
    nF, nI = len(owner), len(neighbour)
    loc = mesh_dir[mesh_dir.index("constant"):] if "constant" in mesh_dir else "constant/polyMesh"
    note = "nPoints: %d nCells: %d nFaces: %d nInternalFaces: %d" % (len(points), n_cells, nF, nI)
    _write(os.path.join(mesh_dir, "points"), _header("vectorField", "points", loc) + _vector_list(points))
    body = "%d\n(\n%s\n)\n" % (nF, "\n".join("4(%d %d %d %d)" % tuple(int(v) for v in f) for f in faces))
    _write(os.path.join(mesh_dir, "faces"), _header("faceList", "faces", loc) + body)
    _write(os.path.join(mesh_dir, "owner"), _header("labelList", "owner", loc, note) + _label_list(owner))
    _write(os.path.join(mesh_dir, "neighbour"), _header("labelList", "neighbour", loc, note) + _label_list(neighbour))
    s = _header("polyBoundaryMesh", "boundary", loc) + "%d\n(\n" % len(patches)
    for k, (name, start, size) in enumerate(patches):
        q = -1 if proc_of_patch is None else int(proc_of_patch[k])
        if q >= 0:
            s += "    %s\n    {\n        type            processor;\n        nFaces          %d;\n        startFace       %d;\n" \
                 "        myProcNo        %d;\n        neighbProcNo    %d;\n    }\n" % (name, size, start, my_rank, q)
        else:
            s += "    %s\n    {\n        type            %s;\n        nFaces          %d;\n        startFace       %d;\n    }\n" % (
                name, PATCH_TYPE.get(name, "wall"), size, start)
    _write(os.path.join(mesh_dir, "boundary"), s + ")\n")


# ------------------------------------------------------------------------------------------------------ fields
FV, ZG, COUPLED = 0, 1, 3   # BC codes of the C-ABI (include/b200foam.h b200_bc: fixedValue, zeroGradient, processor)
DIMS = {"alpha1": "[0 0 0 0 0 0 0]", "U": "[0 1 -1 0 0 0 0]", "pd": "[1 -1 -2 0 0 0 0]", "p": "[1 -1 -2 0 0 0 0]"}


def write_field(path, name, internal, patches, n_internal_faces, bc_types, ref_value, location):
    """vol<Scalar|Vector>Field file.  bc_types[k] in {FV, ZG, COUPLED}; ref_value [B] or [B,3] boundary-face values (fixedValue)."""
    vec = internal.ndim == 2
    s = _header("volVectorField" if vec else "volScalarField", name, location)
    s += "dimensions      %s;\n\ninternalField   nonuniform List<%s> " % (DIMS[name], "vector" if vec else "scalar")
    s += (_vector_list(internal) if vec else _scalar_list(internal)).rstrip("\n") + ";\n\nboundaryField\n{\n"
    for k, (pname, start, size) in enumerate(patches):
        sl = slice(start - n_internal_faces, start - n_internal_faces + size)
        s += "    %s\n    {\n" % pname
        t = bc_types[k]
        if t == FV:
            v = ref_value[sl]
            s += "        type            fixedValue;\n        value           nonuniform List<%s> %s;\n" % (
                "vector" if vec else "scalar", (_vector_list(v) if vec else _scalar_list(v)).rstrip("\n"))
        elif t == ZG:
            s += "        type            zeroGradient;\n"
        else:
            s += "        type            processor;\n        value           uniform %s;\n" % ("(0 0 0)" if vec else "0")
        s += "    }\n"
    _write(path, s + "}\n")


_LIST_RE = re.compile(r"internalField\s+nonuniform\s+List<(scalar|vector)>\s*(\d+)\s*\(", re.S)
_UNIFORM_RE = re.compile(r"internalField\s+uniform\s+(\([^)]*\)|[^;\s]+)\s*;")


def read_field(path, n_cells=None):
    """internalField of an ASCII vol field written by foam-extend (or by write_field) -> ndarray [N] or [N,3]."""
    txt = open(path).read()
    m = _LIST_RE.search(txt)
    if m:
        kind, n = m.group(1), int(m.group(2))
        i = m.end()
        if kind == "scalar":
            j = txt.index(")", i)
            a = np.array(txt[i:j].split(), dtype=float)
            assert a.size == n, "%s: %d values, header says %d" % (path, a.size, n)
            return a
        a = np.empty((n, 3))
        for k, mm in enumerate(re.finditer(r"\(\s*([^\s()]+)\s+([^\s()]+)\s+([^\s()]+)\s*\)", txt[i:])):
            if k >= n:
                break
            a[k] = [float(mm.group(1)), float(mm.group(2)), float(mm.group(3))]
        return a
    m = _UNIFORM_RE.search(txt)
    if not m:
        raise ValueError("%s: no internalField" % path)
    v = m.group(1)
    if v.startswith("("):
        return np.tile(np.array(v.strip("()").split(), dtype=float), (n_cells or 1, 1))
    return np.full(n_cells or 1, float(v))


# ------------------------------------------------------------------------------------------------------ solver log
_RES_RE = re.compile(r"^(\w+):\s+Solving for (\w+), Initial residual = (\S+), Final residual = (\S+), No Iterations (\d+)", re.M)
_TIME_RE = re.compile(r"^Time = (\S+)\s*$", re.M)


def parse_log(path_or_text, field="pd"):
    """lduSolverPerformance::print lines of a foam-extend log -> list of (time, solverName, initialResidual, finalResidual, nIterations)
    for `field` (pcorr solves of correctPhi carry field name pcorr), in order of appearance."""
    txt = open(path_or_text).read() if os.path.exists(path_or_text) else path_or_text
    times = [(m.start(), m.group(1)) for m in _TIME_RE.finditer(txt)]
    out = []
    for m in _RES_RE.finditer(txt):
        if m.group(2) != field:
            continue
        t = None
        for pos, val in times:
            if pos < m.start():
                t = val
            else:
                break
        out.append((t, m.group(1), float(m.group(3).rstrip(",")), float(m.group(4).rstrip(",")), int(m.group(5))))
    return out


def format_log_line(solver, field, init, final, n):
    """the line fvMatrix::solve prints (lduSolverPerformance::print), for writing our own log in the same format"""
    return "%s:  Solving for %s, Initial residual = %s, Final residual = %s, No Iterations %d" % (solver, field, "%g" % init, "%g" % final, n)


# ------------------------------------------------------------------------------------------------------ dictionaries
def _dict(cls, obj, location, body):
    return _header(cls, obj, location) + body + "\n"


def write_dictionaries(case, region, ctl, delta_t, n_steps, solver="PCG", n_procs=(1, 1, 1), libs=()):
    """system/ + constant/ dictionaries of the SURVEY.md 8(d) set-up.  solver = "PCG" for stock foam-extend, "b200PCG" (+ libs
    ("libb200Foam.so")) for the plug-in run."""
    c = ctl
    libs_line = ("libs (%s);\n" % " ".join('"%s"' % l for l in libs)) if libs else ""
    control = "application     shallowInterFoam;\nstartFrom       startTime;\nstartTime       0;\nstopAt          endTime;\nendTime         %s;\n" \
              "deltaT          %s;\nwriteControl    timeStep;\nwriteInterval   %d;\npurgeWrite      0;\nwriteFormat     ascii;\nwritePrecision  17;\n" \
              "writeCompression uncompressed;\ntimeFormat      general;\ntimePrecision   10;\nrunTimeModifiable no;\nadjustTimeStep  no;\nmaxCo           0.5;\n" \
              "maxDeltaT       1;\n%s" % (_fmt(delta_t * n_steps), _fmt(delta_t), n_steps, libs_line)
    schemes = "ddtSchemes { default Euler; }\ngradSchemes { default Gauss linear; }\n" \
              "divSchemes\n{\n    div(rho*phi,U)  Gauss %s;\n    div(phi,alpha)  Gauss vanLeer;\n    div(phirb,alpha) Gauss interfaceCompression;\n}\n" \
              "laplacianSchemes { default Gauss linear uncorrected; }\ninterpolationSchemes { default linear; }\nsnGradSchemes { default uncorrected; }\n" \
              "fluxRequired { default no; pd; pcorr; alpha1; }\n" % {0: "upwind", 1: "linear", 2: "limitedLinearV 1"}.get(int(c.divUScheme), "upwind")
    sol = lambda tol, rel: "{ solver %s; preconditioner DIC; tolerance %s; relTol %s; minIter %d; maxIter %d; }" % (
        solver, _fmt(tol), _fmt(rel), int(c.minIter), int(c.maxIter))
    solution = "solvers\n{\n    pcorr   %s\n    pd      %s\n    pdFinal %s\n    U       { solver PBiCG; preconditioner DILU; tolerance 1e-06; relTol 0; }\n}\n" \
               "PISO\n{\n    momentumPredictor no;\n    nCorrectors     %d;\n    nNonOrthogonalCorrectors %d;\n    nAlphaCorr      %d;\n    nAlphaSubCycles %d;\n" \
               "    cAlpha          %s;\n    pdRefCell       0;\n    pdRefValue      0;\n}\n" % (
                   sol(c.pcorrTol, c.pcorrRelTol), sol(c.pdTol, c.pdRelTol), sol(c.pdFinalTol, c.pdFinalRelTol), int(c.nCorr), int(c.nNonOrthCorr),
                   int(c.nAlphaCorr), int(c.nAlphaSubCycles), _fmt(c.cAlpha))
    decomp = "numberOfSubdomains %d;\nmethod          simple;\nsimpleCoeffs\n{\n    n               (%d %d %d);\n    delta           0.001;\n}\n" % (
        n_procs[0] * n_procs[1] * n_procs[2], n_procs[0], n_procs[1], n_procs[2])
    _write(os.path.join(case, "system", "controlDict"), _dict("dictionary", "controlDict", "system", control))
    for d in ("system", os.path.join("system", region)):
        _write(os.path.join(case, d, "fvSchemes"), _dict("dictionary", "fvSchemes", d, schemes))
        _write(os.path.join(case, d, "fvSolution"), _dict("dictionary", "fvSolution", d, solution))
        _write(os.path.join(case, d, "decomposeParDict"), _dict("dictionary", "decomposeParDict", d, decomp))
    _write(os.path.join(case, "constant", "regionProperties"),
           _dict("dictionary", "regionProperties", "constant", "region2dNames   ();\nregion3dNames   (%s);\n" % region))
    cr = os.path.join("constant", region)
    _write(os.path.join(case, cr, "g"), _header("uniformDimensionedVectorField", "g", cr) +
           "dimensions      [0 1 -2 0 0 0 0];\nvalue           (%s %s %s);\n" % (_fmt(c.g[0]), _fmt(c.g[1]), _fmt(c.g[2])))
    phase = lambda nm, nu, rho: "%s\n{\n    transportModel  Newtonian;\n    nu              nu [0 2 -1 0 0 0 0] %s;\n    rho             rho [1 -3 0 0 0 0 0] %s;\n}\n" % (
        nm, _fmt(nu), _fmt(rho))
    _write(os.path.join(case, cr, "transportProperties"), _dict("dictionary", "transportProperties", cr,
           phase("phase1", c.nu1, c.rho1) + phase("phase2", c.nu2, c.rho2) + "sigma           sigma [1 0 -2 0 0 0 0] %s;\n" % _fmt(c.sigma)))
    _write(os.path.join(case, cr, "turbulenceProperties"), _dict("dictionary", "turbulenceProperties", cr, "simulationType  laminar;\n"))
    _write(os.path.join(case, cr, "RASProperties"), _dict("dictionary", "RASProperties", cr, "RASModel        laminar;\nturbulence      off;\nprintCoeffs     off;\n"))


# ------------------------------------------------------------------------------------------------------ whole cases
def write_case(case, pm, cs, ctl, delta_t, n_steps, region="region3d", solver="PCG", libs=()):
    """pm: capi.PolyMesh (or any object with points/faces/owner/neighbour/patches/nCells/nInternalFaces);
    cs: dict from bench.channel_fields (alpha, U, pd, tA, tU, tP, rvA, rvU)."""
    mesh_dir = os.path.join(case, "constant", region, "polyMesh")
    write_polymesh(mesh_dir, pm.points, pm.faces, pm.owner, pm.neighbour, pm.patches, pm.nCells)
    nI = pm.nInternalFaces
    B = len(pm.owner) - nI
    loc = "0/" + region
    write_field(os.path.join(case, "0", region, "alpha1"), "alpha1", cs["alpha"], pm.patches, nI, cs["tA"], cs["rvA"], loc)
    write_field(os.path.join(case, "0", region, "U"), "U", cs["U"], pm.patches, nI, cs["tU"], cs["rvU"], loc)
    write_field(os.path.join(case, "0", region, "pd"), "pd", cs["pd"], pm.patches, nI, cs["tP"], np.zeros(B), loc)
    write_dictionaries(case, region, ctl, delta_t, n_steps, solver=solver, libs=libs)


def write_decomposed(case, pm, n_procs, region="region3d"):
    """processor<r>/constant/<region>/polyMesh for the `simple` decomposition (decomposePar layout): sub-mesh + processor patches +
    cellProcAddressing / faceProcAddressing (face index + 1, negative when the face is flipped) / boundaryProcAddressing."""
    n = n_procs[0] * n_procs[1] * n_procs[2]
    proc = pm.decompose_simple(n_procs)
    names = pm.patchNames
    for r in range(n):
        sub = pm.submesh(proc, n, r)
        d = os.path.join(case, "processor%d" % r, "constant", region, "polyMesh")
        write_polymesh(d, sub.points, sub.faces, sub.owner, sub.neighbour, sub.patches, sub.nCells, sub.patchNeighbRank, r)
        loc = "constant/%s/polyMesh" % region
        _write(os.path.join(d, "cellProcAddressing"), _header("labelIOList", "cellProcAddressing", loc) + _label_list(sub.cellMap))
        fpa = np.where(sub.faceFlip != 0, -(sub.faceMap + 1), sub.faceMap + 1)
        _write(os.path.join(d, "faceProcAddressing"), _header("labelIOList", "faceProcAddressing", loc) + _label_list(fpa))
        bpa = [names.index(nm) if nm in names else -1 for nm in sub.patchNames]
        _write(os.path.join(d, "boundaryProcAddressing"), _header("labelIOList", "boundaryProcAddressing", loc) + _label_list(bpa))
        sub.release()
    return proc


def read_polymesh(mesh_dir):
    """owner / neighbour / points / faces / boundary of an ASCII polyMesh (what write_polymesh wrote, or blockMesh's own output)"""
    def body(name):
        txt = open(os.path.join(mesh_dir, name)).read()
        txt = txt[txt.index("}", txt.index("FoamFile")) + 1:]
        txt = re.sub(r"//.*", "", txt)
        i = txt.index("(")
        return int(txt[:i].split()[-1]), txt[i + 1:txt.rindex(")")]
    n, b = body("owner"); owner = np.array(b.split(), dtype=np.int32); assert owner.size == n
    n, b = body("neighbour"); neighbour = np.array(b.split(), dtype=np.int32); assert neighbour.size == n
    n, b = body("points"); points = np.array(re.sub(r"[()]", " ", b).split(), dtype=float).reshape(-1, 3); assert len(points) == n
    n, b = body("faces")
    faces = [np.array(m.group(2).split(), dtype=np.int32) for m in re.finditer(r"(\d+)\(([^)]*)\)", b)]
    assert len(faces) == n
    _, b = body("boundary")
    patches = []
    for m in re.finditer(r"(\w+)\s*\{([^}]*)\}", b):
        d = dict(re.findall(r"(\w+)\s+([^;]+);", m.group(2)))
        patches.append((m.group(1), int(d["startFace"]), int(d["nFaces"]), d["type"].strip(), int(d.get("neighbProcNo", -1))))
    return dict(points=points, faces=faces, owner=owner, neighbour=neighbour, patches=patches, nCells=int(owner.max()) + 1)


# ------------------------------------------------------------------------------------------------------ command line
def _bench_case(dims, lengths, weir, lib=None):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from . import capi
    lib = lib or capi.load()
    pm = lib.polymesh_box(*dims, *lengths, weir=weir)
    g = pm.geometry()
    F = pm.nInternalFaces
    slices = {n: pm.patch_slice(n) for n in pm.patchNames}
    cs = bench.channel_fields(pm.patchNames, g["C"], g["Cf"][F:], slices, lengths)
    ctl = bench.set_controls(capi.Region3dControls())
    return lib, pm, cs, ctl, bench.DELTA_T


def main(argv=None, lib=None):
    """lib: an already loaded (and, for `check`, initialised) B200Lib -- the tests pass theirs; the command line loads the product library"""
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("cmd", choices=["write", "check"])
    ap.add_argument("case")
    ap.add_argument("--dims", type=int, nargs=3, default=[50, 10, 20])
    ap.add_argument("--lengths", type=float, nargs=3, default=[5.0, 1.0, 2.0])
    ap.add_argument("--weir", type=float, nargs=3, default=None, help="x0 x1 zfrac of the removed block (fractions)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--region", default="region3d")
    ap.add_argument("--procs", type=int, nargs=3, default=None, help="also write processor*/ for this `simple` decomposition")
    ap.add_argument("--solver", default="PCG")
    a = ap.parse_args(argv)
    given = lib is not None
    lib, pm, cs, ctl, dt = _bench_case(tuple(a.dims), tuple(a.lengths), tuple(a.weir) if a.weir else None, lib)
    if a.cmd == "write":
        write_case(a.case, pm, cs, ctl, dt, a.steps, a.region, a.solver, ("libb200Foam.so",) if a.solver == "b200PCG" else ())
        if a.procs:
            write_decomposed(a.case, pm, tuple(a.procs), a.region)
        print("wrote %s: %d cells, %d faces, %d patches; run `shallowInterFoam -case %s > %s/log`" % (
            a.case, pm.nCells, len(pm.owner), pm.nPatches, a.case, a.case))
        return 0
    # check: our device run of the same case against foam-extend's log and written fields
    if not given:
        lib.init(0, 0, 1)
    dm = pm.device_mesh()
    reg = dm.region3d(ctl, dm.bc(cs["tA"], 1, cs["rvA"]), dm.bc(cs["tU"], 3, cs["rvU"]), dm.bc(cs["tP"], 1), cs["alpha"], cs["U"], cs["pd"])
    reg.correct_phi(); reg.clear_log()
    reg.step(a.steps)
    ours = reg.perfs()
    ref = parse_log(os.path.join(a.case, "log"), "pd")
    ok = True
    print("%d pd solves in the foam-extend log, %d here" % (len(ref), len(ours)))
    for k, (r, o) in enumerate(zip(ref, ours)):
        rel = abs(o[0] - r[2]) / max(abs(r[2]), 1e-300)
        same = o[2] == r[4] and rel < 1e-5      # the log prints 6 significant digits
        ok &= same
        print("solve %3d  t=%s  foam-extend: init %.6g final %.6g it %d | b200: init %.6g final %.6g it %d  %s" % (
            k, r[0], r[2], r[3], r[4], o[0], o[1], o[2], "ok" if same else "DIFFERS"))
    tdir = os.path.join(a.case, _fmt(dt * a.steps).rstrip("0").rstrip("."), a.region)
    if os.path.isdir(tdir):
        f = reg.fields()
        for name, key in (("alpha1", "alpha1"), ("pd", "pd"), ("U", "U")):
            p = os.path.join(tdir, name)
            if os.path.exists(p):
                v = read_field(p, pm.nCells)
                err = np.abs(v - f[key]).max() / max(np.abs(v).max(), 1e-300)
                good = err <= 1e-8
                ok &= good
                print("%-6s Linf relative difference %.3e  %s" % (name, err, "ok" if good else "DIFFERS"))
    else:
        print("no time directory %s: fields not compared" % tdir)
    print("PARITY PINNED" if ok and ref else "parity NOT pinned")
    return 0 if ok and ref else 1


if __name__ == "__main__":
    sys.exit(main())
