"""The reduced-memory mode on the device (ls1b200_rmm_*, csrc/rmm.cu) against (a) the per-step ResultWriter output of
the REFERENCE built in RMM configuration (tests/golden/rmm_1clj_2000.npz, made by tests/golden/make_golden_rmm.py from
oracle/_ref/rmm/MarDyn) and (b) the numpy restatement oracle/rmm.py. Needs a B200: run with -m gpu.
All arithmetic of the mode is FP32 (VCP1CLJRMM.cpp, PositionCellProcessorRMM.h); sums are taken in a different order than the
reference's SIMD lanes, hence 1e-5 (the tolerance of the mixed build) -- the same bound tests/test_oracle_rmm.py holds the oracle to."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rmm_1clj_2000.npz")


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ls1_mardyn_b200 import capi
    capi.lib()
    return capi


def _ctx(capi, g):
    from ls1_mardyn_b200.components import lj_fluid
    c = capi.Context([0, 0, 0], list(g["length"]), float(g["cutoff"]), precision="mixed", geometry="xml", periodic=(1, 1, 1))
    c.set_components(lj_fluid(rc=float(g["cutoff"])))
    c.rmm_upload(g["id"], g["r"], g["v"])
    return c


def test_rmm_trace_vs_reference_build(capi):
    """40 steps of U_pot, U_kin, p, beta_trans against the reference's own RMM binary."""
    from oracle.rmm import OracleRMM
    g = np.load(GOLD)
    steps, dt, T = int(g["steps"]), float(g["dt"]), float(g["temperature"])
    o = OracleRMM(g["length"], float(g["cutoff"]), dt, T); o.upload(g["id"], g["r"], g["v"])
    uc, wc = o._corrections()                      # Domain's homogeneous long-range correction: host side, a constant of (N, V, rc)
    c = _ctx(capi, g)
    assert capi.lib().ls1b200_rmm_state_bytes(c.h) == 32 * len(g["id"])
    rows = [c.rmm_start(dt, T, uc, wc)] + list(c.rmm_simulate(steps, dt, T, uc, wc))
    rows = np.array(rows)
    N, V = len(g["id"]), float(np.prod(g["length"]))
    tol = 1e-5
    for name, col in (("upot", rows[:, 0]), ("ukin", 0.5 * rows[:, 2]), ("beta_trans", rows[:, 6])):
        b = g[name]
        err = np.abs(col - b).max() / np.abs(b).max()
        assert err <= tol, (name, err)
    err = np.abs(rows[:, 5] - g["p"]).max() / max(np.abs(g["p"]).max(), (N / V) * T)      # p = rho T + rho W / 3N is a near cancellation
    assert err <= tol, ("p", err)


def test_rmm_state_vs_oracle(capi):
    """Positions and velocities after 10 steps against oracle/rmm.py, molecule by molecule; split calls == one call."""
    from oracle.rmm import OracleRMM
    g = np.load(GOLD)
    dt, T = float(g["dt"]), float(g["temperature"])
    o = OracleRMM(g["length"], float(g["cutoff"]), dt, T); o.upload(g["id"], g["r"], g["v"])
    uc, wc = o._corrections()
    o.simulate(10)
    c = _ctx(capi, g)
    c.rmm_start(dt, T, uc, wc); c.rmm_simulate(4, dt, T, uc, wc); c.rmm_simulate(6, dt, T, uc, wc)
    ids, r, v = c.rmm_download()
    oo = np.argsort(o.id)
    assert np.array_equal(ids, o.id[oo])
    L = np.asarray(g["length"])
    d = r - o.r[oo].astype(np.float64); d -= L * np.round(d / L)
    assert np.abs(d).max() <= 2e-4 and np.abs(v - o.v[oo].astype(np.float64)).max() <= 2e-3      # FP32 trajectories, sums in a different order
    c2 = _ctx(capi, g)
    c2.rmm_start(dt, T, uc, wc); c2.rmm_simulate(10, dt, T, uc, wc)
    i2, r2, v2 = c2.rmm_download()
    assert np.array_equal(r2, r) and np.array_equal(v2, v)                                         # deterministic, and the pending thermostat factor survives a call boundary


def test_rmm_rejects_what_the_mode_does_not_cover(capi):
    from ls1_mardyn_b200.capi import Ls1B200Error
    from helpers import load
    g = load("periodic_water_432")
    c = capi.Context([0, 0, 0], list(g["length"]), float(g["cutoff"]), precision="mixed", geometry="xml", periodic=(1, 1, 1))
    c.set_components(g["ensemble"])
    with pytest.raises(Ls1B200Error):
        c.rmm_upload(g["in_id"], g["in_r"], g["in_v"])

