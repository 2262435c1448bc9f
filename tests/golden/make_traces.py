#!/usr/bin/env python
"""Generates tests/golden/trace_*.npz: per-step U_pot / U_kin / p / beta_trans / beta_rot of the UNMODIFIED reference binary
(oracle/_ref/{mixed,double}/MarDyn: stock Simulation loop, Leapfrog incl. rigid-body rotation, Domain::calculateGlobalValues,
VelocityScalingThermostat, ResultWriter with 15 digits) over 100 steps, started from the phase spaces of the periodic golden
fixtures. These pin what the force fixtures cannot: FullMolecule::upd_preF / upd_postF with quaternions and angular momentum
(src/molecules/FullMolecule.cpp:331-386), sum I w^2, beta_rot, T and p (src/Domain.cpp:157-252).

Run in the build container only (needs oracle/_ref, i.e. /root/reference); the .npz files are committed.

  python tests/golden/make_traces.py
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ls1_mardyn_b200.ascii_reader import PhaseSpace, write_phase_space  # noqa: E402
from ls1_mardyn_b200.xml_config import simulation_xml  # noqa: E402
from helpers import load  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref")

# fixture (initial state), time step, thermostat target, steps. The rotating fixtures start with L = 0 and random orientations;
# the torques spin the molecules up within a few steps, so beta_rot != 1 from the second step on.
CASES = [
    ("periodic_1clj_1458", 0.002, 0.85, 100, None),
    ("periodic_water_432", 0.0166879, 9.437e-4, 100, None),      # dt / 4: keeps beta >= 0.9, i.e. outside Domain's "explosion" heuristics (Domain.cpp:259-314)
    ("periodic_multi_250", 0.0667516, 8.55e-4, 100, None),
    # ... and INSIDE them: the 1CLJ liquid at T = 0.85 with a thermostat target of 0.34 -- beta_trans = 0.4^0.4 = 0.69 < MIN_BETA, so the
    # reference caps every molecule above 10 T_target = 4 kT of translational energy (a few per cent of them per step) with its force,
    # holds beta_trans at 1, raises a warning after 40 steps and an error after 80, from where the plain factor applies
    ("periodic_1clj_1458", 0.002, 0.34, 100, "trace_1clj_1458_explosion"),
]

COLS = ("step", "time", "U_pot", "U_pot_avg", "U_kin", "U_kin_avg", "p", "p_avg", "beta_trans", "beta_rot", "c_v", "N")


def run_reference(prec, g, dt, temperature, steps, threads=4):
    ens = g["ensemble"]
    n = len(g["in_id"])
    ps = PhaseSpace(np.array(g["length"]), float(temperature), float(g["epsilon_rf"]), ens, g["in_id"], g["in_cid"], g["in_r"], g["in_v"], g["in_q"], g["in_L"])
    with tempfile.TemporaryDirectory() as td:
        write_phase_space(os.path.join(td, "start.inp"), ps)
        xml = simulation_xml(ens, g["length"], temperature, dt, steps, float(g["cutoff"]), epsilon_rf=float(g["epsilon_rf"]),
                             phasespace_file="start.inp", longrange="none")
        open(os.path.join(td, "config.xml"), "w").write(xml)
        p = subprocess.run([os.path.join(REFBIN, prec, "MarDyn"), "--steps", str(steps), "--final-checkpoint=0", "config.xml"], cwd=td,
                           capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
        if p.returncode != 0:
            raise RuntimeError((p.stdout + p.stderr)[-4000:])
        rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
    res = np.array([[float(x) for x in r] for r in rows])
    assert res.shape[1] == len(COLS) and res.shape[0] >= steps and int(res[0, 11]) == n, res.shape
    return res


def traces():
    for name, dt, temperature, steps, outname in CASES:
        g = load(name)
        out = dict(dt=dt, temperature=temperature, steps=steps, fixture=np.array(name), columns=np.array(COLS))
        for prec in ("mixed", "double"):
            out[prec] = run_reference(prec, g, dt, temperature, steps)
        np.savez_compressed(os.path.join(GOLD, (outname or "trace_" + name[len("periodic_"):]) + ".npz"), **out)
        d = out["double"]
        print("wrote trace of", name, "rows", d.shape[0], "U_pot", d[0, 2], "->", d[-1, 2], "beta_rot range", d[:, 9].min(), d[:, 9].max())


if __name__ == "__main__":
    if len(sys.argv) > 1:      # only the named outputs, e.g. trace_1clj_1458_explosion
        CASES = [c for c in CASES if (c[4] or "trace_" + c[0][len("periodic_"):]) in sys.argv[1:]]
    traces()
