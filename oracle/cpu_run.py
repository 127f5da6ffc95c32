"""CPU ORACLE (test infrastructure / reported baseline only; PARITY UNPINNED) -- driver of the CPU arm of bench.py.

run_window() replays, with the oracle restatement of foam-extend-3.1, the window every bench leg times: fresh fields ->
correctPhi -> `warm` steps -> `steps` timed steps of the synthetic channel, serial (n_proc = 1) or decomposed n_proc ways with one
thread per `simple` sub-domain (oracle/foam_oracle.cpp orc_world_*: rank-local DIC, processor-patch halo swaps, rank-ordered global
sums -- the semantics of `mpirun -np n_proc shallowInterFoam -parallel` [REF shallowInterFoam.C:66-67]).
Only bench.py's cpu_baseline leg and `--impl reference` call this."""
import time

import numpy as np

from . import foam_oracle as fo, mesh_oracle as mo


def build_serial(dims, lengths, weir, channel_fields, set_controls):
    mesh = mo.block_mesh_box(*dims, *lengths)
    if weir:
        mesh = mo.subset_mesh(mesh, mo.weir_keep_mask(*dims, *weir))
    g = mo.geometry(mesh)
    m = fo.Mesh(mesh["owner"], mesh["neighbour"], mesh["patches"], g, mesh["nCells"])
    names = [p[0] for p in mesh["patches"]]
    slices = {n: m.patch_slice(n) for n in names}
    cs = channel_fields(names, g["C"], g["Cf"][m.F:], slices, lengths)
    ctl = set_controls(fo.OrcControls())
    r = fo.Region(m, ctl, fo.BC(m, cs["tA"], 1, cs["rvA"]), fo.BC(m, cs["tU"], 3, cs["rvU"]), fo.BC(m, cs["tP"], 1),
                  cs["alpha"], cs["U"], cs["pd"])
    return m.N, r


def run_window(dims, lengths, weir, channel_fields, set_controls, n_proc, warm, steps, budget_s):
    t_setup = time.time()
    if n_proc > 1 and hasattr(fo, "World"):
        n_cells, r = fo.World.build(dims, lengths, weir, channel_fields, set_controls, n_proc)
        cores = r.n_ranks
        how = "%d-way `simple` decomposition, one thread per sub-domain (rank-local DIC, halo swaps, rank-ordered sums)" % cores
    else:
        n_cells, r = build_serial(dims, lengths, weir, channel_fields, set_controls)
        cores = 1
        how = "serial oracle"
    r.correct_phi()
    t_setup = time.time() - t_setup
    # warm-up steps of the window (untimed, but they are CPU work too: if they alone would blow the budget the window starts earlier
    # and the sample says so)
    t0 = time.time(); warm_done = 0
    for _ in range(warm):
        if warm_done and (time.time() - t0) / warm_done * (warm_done + 1) > 2.0 * budget_s:
            break
        r.step(); warm_done += 1
    t_warm = time.time() - t0
    r.clear_log()
    t0 = time.time(); done = 0
    for _ in range(steps):
        r.step(); done += 1
        dt = time.time() - t0
        if done < steps and dt / done * (done + 1) > budget_s:
            break
    dt = time.time() - t0
    iters = sum(p[2] for p in r.perfs())
    sample = "%s on %d host core(s): correctPhi + %d of %d warm-up steps (%.1f s, untimed) + %d of %d timed steps (%.1f s) of the same " \
             "%d-cell workload; %d PCG iterations in the timed steps; set-up %.1f s untimed" % (
                 how, cores, warm_done, warm, t_warm, done, steps, dt, n_cells, iters, t_setup)
    return {"value": n_cells * done / dt, "cells": n_cells, "steps": done, "warmup": warm_done, "ms_per_step": 1e3 * dt / done,
            "pcg_iterations_per_step": iters / done, "us_per_pcg_iteration": 1e6 * dt / max(iters, 1), "cores": cores, "sample": sample}
