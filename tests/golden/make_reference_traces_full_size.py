#!/usr/bin/env python
"""Generates tests/golden/reference_trace_<workload>_full.npz (lj: reference_trace_lj16M.npz): the ResultWriter rows (15 digits) of
the UNMODIFIED reference binary (oracle/_ref/mixed/MarDyn, OpenMP) on BASELINE.json's single-GPU configurations AT FULL SIZE, started
from the reference's own CubicGridGenerator (the XML is the one bench.py's reference arm runs):

  lj     configs[1]: 16 000 000 1CLJ molecules, rc = 2.5 sigma, 100 steps      (~9 minutes on 8 cores)
  water  configs[2]:  4 000 752 water-like molecules (1 LJ + 3 charges), 20 steps
  co2    configs[3]:  8 039 358 2CLJQ molecules (2 LJ + quadrupole), rc = 5 sigma, 20 steps
  argon  configs[0]:    101 306 1CLJ molecules (examples/Argon state point), 100 steps

ls1b200_generate_cubic_grid(LS1B200_GEN_REFERENCE_VELOCITIES) builds the same initial state on the device (for the rotating models:
all orientations (1, 0, 0, 0), L = 0, the whole thermal energy in the translation -- EqualVelocityAssigner.h:18), so the CUDA path is
held against the reference at the sizes the benchmark is quoted on: tests/test_gpu_parity.py::test_benchmark_state_*_vs_reference_binary
and the `trace.vs_reference_fixture` entry of bench.py's JSON line.

Run in the build container only (needs oracle/_ref, i.e. /root/reference); the .npz files (a few KB each) are committed.

  python tests/golden/make_reference_traces_full_size.py [lj|lj:double|lj:rmm|water|co2|argon ...]
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ls1_mardyn_b200 import workloads  # noqa: E402
from ls1_mardyn_b200.xml_config import simulation_xml  # noqa: E402

STEPS = {"lj": 100, "water": 20, "co2": 20, "argon": 100}
FILES = {"lj": "reference_trace_lj16M.npz", "water": "reference_trace_water_full.npz", "co2": "reference_trace_co2_full.npz", "argon": "reference_trace_argon_full.npz"}
COLS = ("step", "time", "U_pot", "U_pot_avg", "U_kin", "U_kin_avg", "p", "p_avg", "beta_trans", "beta_rot", "c_v", "N")


def make(name, prec="mixed"):
    steps = STEPS[name]
    wl = workloads.get(name)
    n = wl.n_per_dim
    N = 2 * n ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    exe = os.path.join(ROOT, "oracle", "_ref", prec, "MarDyn")
    with tempfile.TemporaryDirectory() as td:
        if name == "lj":
            text = bench.XML.format(dt=wl.dt, steps=steps, T=wl.temperature, L=repr(L), rho=repr(N / L ** 3 * (1 + 1e-9)), rc=wl.rc)
        else:      # what bench.run_reference_binary writes for the other workloads, with the ResultWriter on
            text = simulation_xml(wl.ensemble, [L, L, L], wl.temperature, wl.dt, steps, wl.rc, generator_density=N / L ** 3 * (1 + 1e-9), longrange="homogeneous")
        if prec == "rmm":      # the reference's reduced-memory build (all FP32: MoleculeRMM, VCP1CLJRMM, LeapfrogRMM)
            text = text.replace('integrator type="Leapfrog"', 'integrator type="LeapfrogRMM"')
            steps = min(steps, 40)
        open(os.path.join(td, "config.xml"), "w").write(text)
        p = subprocess.run([exe, "--steps", str(steps), "--final-checkpoint=0", "config.xml"], cwd=td, capture_output=True, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
        if p.returncode != 0:
            raise RuntimeError((p.stdout + p.stderr)[-4000:])
        rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
    res = np.array([[float(x) for x in r] for r in rows])
    assert res.shape[0] >= steps + 1 and int(res[0, 11]) == N, (res.shape, res[0, 11], N)
    out = FILES[name] if prec == "mixed" else FILES[name].replace(".npz", "_" + prec + ".npz")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", out), trace=res, precision=np.array(prec), n_per_dim=n, molecules=N, length=L, steps=steps,
                        rho=wl.rho, temperature=wl.temperature, rc=wl.rc, dt=wl.dt, workload=np.array(name), columns=np.array(COLS))
    print(name, "rows", res.shape[0], "U_pot / N", res[0, 2] / N, "->", res[-1, 2] / N, "U_kin", res[-1, 4], "p", res[-1, 6], "beta_rot", res[1, 9], res[-1, 9])


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["lj", "lj:double", "lj:rmm", "water", "co2", "argon"]):      # "lj:double": the reference's double build (MARDYN_DPDP); "lj:rmm": its reduced-memory build
        make(*w.split(":"))
