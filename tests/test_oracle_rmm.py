"""The oracle of the NEXT hot-path row (the reference's reduced-memory mode: all-FP32 1CLJ, force fused with the
velocity update) pinned before any CUDA is written for it: oracle/rmm.py (numpy restatement of VCP1CLJRMM + LeapfrogRMM +
VelocityCellProcessorRMM + the thermostat / global values) against the per-step ResultWriter output of the reference built in RMM
configuration (tests/golden/rmm_1clj_2000.npz, made by tests/golden/make_golden_rmm.py from oracle/_ref/rmm/MarDyn)."""
import os

import numpy as np

from oracle.rmm import OracleRMM

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rmm_1clj_2000.npz")


def test_rmm_trace_vs_reference():
    g = np.load(GOLD)
    steps = int(g["steps"])
    o = OracleRMM(g["length"], float(g["cutoff"]), float(g["dt"]), float(g["temperature"]))
    o.upload(g["id"], g["r"], g["v"])
    rows = o.simulate(steps)
    assert len(rows) == steps + 1 == len(g["upot"]) and int(g["n"][0]) == o.n == 2000
    N, V, T = o.n, float(np.prod(g["length"])), float(g["temperature"])
    tol = 1e-5                                         # all-FP32 sums in a different order than the reference's SIMD lanes
    for k in ("upot", "ukin", "beta_trans"):
        a = np.array([r[k] for r in rows]); b = g[k]
        assert np.abs(a - b).max() <= tol * np.abs(b).max(), (k, np.abs(a - b).max() / np.abs(b).max())
    a = np.array([r["p"] for r in rows]); b = g["p"]
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), (N / V) * T)      # p = rho T + rho W / 3N is a near cancellation
    # what the mode is about: FP32 state, 32 bytes per molecule (position, velocity, id)
    assert o.r.dtype == np.float32 and o.v.dtype == np.float32 and o.r.nbytes + o.v.nbytes + o.id.nbytes == 32 * N


def test_rmm_thermostat_holds_the_temperature_and_momentum():
    g = np.load(GOLD)
    o = OracleRMM(g["length"], float(g["cutoff"]), float(g["dt"]), float(g["temperature"]))
    o.upload(g["id"], g["r"], g["v"])
    p0 = o.v.astype(np.float64).sum(axis=0)
    rows = o.simulate(5)
    # velocity scaling with the exponent 0.4 (Domain.cpp:248) relaxes T towards the target; pair forces conserve momentum
    assert abs((o.v.astype(np.float64) ** 2).sum() / (3 * o.n) - float(g["temperature"])) < 2e-4
    assert np.abs(o.v.astype(np.float64).sum(axis=0) - p0 * np.prod([r["beta_trans"] for r in rows[1:]])).max() < 5e-3
    assert np.all((o.r >= 0) & (o.r.astype(np.float64) < g["length"]))
