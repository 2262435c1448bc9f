#!/usr/bin/env python
"""Generates tests/golden/reference_trace_lj16M.npz: the ResultWriter rows (15 digits) of the UNMODIFIED reference binary
(oracle/_ref/mixed/MarDyn, OpenMP) on BASELINE.json configs[1] AT FULL SIZE -- 16 000 000 1CLJ molecules from its own
CubicGridGenerator, rc = 2.5 sigma, dt = 0.002, NVT velocity scaling -- for 100 steps (the XML is bench.py's reference-arm template).

ls1b200_generate_cubic_grid(LS1B200_GEN_REFERENCE_VELOCITIES) builds the same initial state on the device, so the CUDA path can be
held against the reference at the size the benchmark is quoted on: tests/test_gpu_parity.py::test_benchmark_state_16M_vs_reference_binary
and the `trace.vs_reference_fixture` entry of bench.py's JSON line.

Run in the build container only (needs oracle/_ref, i.e. /root/reference; ~10 minutes on 8 cores); the .npz file is committed.

  python tests/golden/make_reference_trace_16m.py [steps]
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


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    wl = workloads.get("lj")
    n = wl.n_per_dim
    N = 2 * n ** 3; L = (N / wl.rho) ** (1.0 / 3.0)
    exe = os.path.join(ROOT, "oracle", "_ref", "mixed", "MarDyn")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "config.xml"), "w").write(bench.XML.format(dt=wl.dt, steps=steps, T=wl.temperature, L=repr(L), rho=repr(N / L ** 3 * (1 + 1e-9)), rc=wl.rc))
        p = subprocess.run([exe, "--steps", str(steps), "--final-checkpoint=0", "config.xml"], cwd=td, capture_output=True, text=True,
                           env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)))
        if p.returncode != 0:
            raise RuntimeError((p.stdout + p.stderr)[-4000:])
        rows = [l.split() for l in open(os.path.join(td, "res.res")) if l.strip() and not l.lstrip().startswith("#")]
    res = np.array([[float(x) for x in r] for r in rows])
    assert res.shape[0] >= steps + 1 and int(res[0, 11]) == N, res.shape
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_trace_lj16M.npz"), trace=res, n_per_dim=n, molecules=N, length=L, steps=steps,
                        rho=wl.rho, temperature=wl.temperature, rc=wl.rc, dt=wl.dt,
                        columns=np.array(("step", "time", "U_pot", "U_pot_avg", "U_kin", "U_kin_avg", "p", "p_avg", "beta_trans", "beta_rot", "c_v", "N")))
    print("rows", res.shape[0], "U_pot / N", res[0, 2] / N, "->", res[-1, 2] / N, "T", 2 * res[-1, 4] / (3 * N), "p", res[-1, 6])


if __name__ == "__main__":
    main()
