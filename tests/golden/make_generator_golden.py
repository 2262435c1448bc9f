#!/usr/bin/env python
"""Generates tests/golden/generator_1clj_2662.npz and generator_water_432.npz: what the UNMODIFIED reference binary (oracle/_ref/{mixed,double}/MarDyn) builds
from `<generator name="CubicGridGenerator">` -- io/CubicGridGeneratorInternal.cpp:62-123, ParticleCellBase::initCubicGrid
(particleContainer/ParticleCellBase.cpp:51-159: one std::mt19937(0) per cell), IOHelpers::removeMomentum incl. the rescaling to the
target temperature (io/IOHelpers.cpp:8-115) -- and where the stock Simulation loop takes that state in 20 steps:

  id / r / v of every molecule after 0 steps (final checkpoint of a `--steps 0` run, 17+ digits),
  the ResultWriter rows (15 digits) of a 20-step run with the reference's default long-range correction (Homogeneous).

Pins ls1b200_generate_cubic_grid(LS1B200_GEN_REFERENCE_VELOCITIES) and its host mirror (generators.cubic_grid_reference), and --
through the trace -- that the 16 M-molecule state bench.py generates on the device is the state `bench.py --impl reference` runs.

Run in the build container only (needs oracle/_ref, i.e. /root/reference); the .npz file is committed.

  python tests/golden/make_generator_golden.py
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ls1_mardyn_b200 import workloads  # noqa: E402
from ls1_mardyn_b200.ascii_reader import read_molecules_only  # noqa: E402
from ls1_mardyn_b200.xml_config import simulation_xml  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref")
# workload, lattice cells per dimension (N = 2 n^3), steps, file. The water-like model starts "exploding" by the reference's own
# definition: its generator puts the whole thermal energy, (3 + rotDOF) T / 2 per molecule, into the translation, so beta_trans =
# 0.5^0.4 < MIN_BETA and Domain's explosion heuristics (Domain.cpp:259-314) run from the first evaluation on: beta_trans = 1,
# beta_rot^0.0091, per-molecule limits, the warning / error counters (the heuristics switch themselves off at step 80 of this run).
CASES = [("lj", 11, 20, "generator_1clj_2662.npz"), ("water", 6, 100, "generator_water_432.npz")]


def run(prec, wl, n_per_dim, steps, td, longrange="homogeneous"):
    N = 2 * n_per_dim ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    open(os.path.join(td, "config.xml"), "w").write(simulation_xml(wl.ensemble, [L, L, L], wl.temperature, wl.dt, steps, wl.rc, generator_density=N / L ** 3 * (1 + 1e-9),
                                                                  longrange=longrange))
    p = subprocess.run([os.path.join(REFBIN, prec, "MarDyn"), "--steps", str(steps), "config.xml"], cwd=td, capture_output=True, text=True,
                       env=dict(os.environ, OMP_NUM_THREADS="4"))
    if p.returncode != 0:
        raise RuntimeError((p.stdout + p.stderr)[-4000:])
    rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
    res = np.array([[float(x) for x in r] for r in rows])
    ids, _, r, v, q, Lang = read_molecules_only(os.path.join(td, "config.restart.dat"))
    return L, res, ids, r, v, q, Lang, (p.stdout + p.stderr).count("Explosion!")


def main():
    for name, n_per_dim, steps, fname in CASES:
        wl = workloads.get(name)
        out = dict(n_per_dim=n_per_dim, steps=steps, rho=wl.rho, temperature=wl.temperature, rc=wl.rc, dt=wl.dt, workload=np.array(name),
                   columns=np.array(("step", "time", "U_pot", "U_pot_avg", "U_kin", "U_kin_avg", "p", "p_avg", "beta_trans", "beta_rot", "c_v", "N")))
        for prec in ("mixed", "double"):
            with tempfile.TemporaryDirectory() as td:
                L, _, ids, r, v, q, Lang, _ = run(prec, wl, n_per_dim, 0, td)
            with tempfile.TemporaryDirectory() as td:
                _, res, _, _, _, _, _, nexp = run(prec, wl, n_per_dim, steps, td)
            with tempfile.TemporaryDirectory() as td:
                _, res0, _, _, _, _, _, _ = run(prec, wl, n_per_dim, steps, td, longrange="none")
            assert len(ids) == 2 * n_per_dim ** 3 and res.shape[0] >= steps
            out["length"] = L
            out[prec + "_id"] = ids; out[prec + "_r"] = r; out[prec + "_v"] = v; out[prec + "_q"] = q; out[prec + "_L"] = Lang
            out[prec + "_trace"] = res; out[prec + "_trace_nolrc"] = res0; out[prec + "_explosions"] = nexp
            print(name, prec, "N", len(ids), "distinct velocities", len(np.unique(np.round(v, 12), axis=0)), "U_pot", res[0, 2], "->", res[-1, 2],
                  "LRC dU", res[0, 2] - res0[0, 2], "explosion warnings", nexp, "beta_trans", res[1, 8], res[-1, 8], "beta_rot", res[1, 9], res[-1, 9])
        np.savez_compressed(os.path.join(GOLD, fname), **out)


if __name__ == "__main__":
    main()
