"""CPU-side checks of bench.py's host logic: the reference arm's JSON line (--impl reference runs the UNMODIFIED reference binary,
oracle/_ref, on the host cores) and the comparison of a run with the committed full-size reference traces. No GPU needed."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_and_trace():
    """`bench.py --impl reference` on a small lattice: one JSON line with the contract's keys, `impl`, a cpu_baseline describing the run,
    an e2e object without copies, and the physical state after the K-th step -- which equals the committed generator fixture (the bench's
    XML template is the configuration that fixture was generated with)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mixed", "MarDyn")):
        pytest.skip("oracle/_ref not built (build container only)")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-per-dim", "11", "--steps", "20", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "MMUPS" and d["steps"] == 20 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MMUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["execution"]["molecules_run"] == 2 * 11 ** 3
    g = np.load(os.path.join(ROOT, "tests", "golden", "generator_1clj_2662.npz"))
    build = d["execution"]["build"]
    r = g["mixed_trace"][20]
    tol = 1e-12 if "_ref/mixed/" in build else 1e-6      # the AVX-512 build of the reference rounds differently
    t = d["trace"]
    assert t["at_step"] == 20 and t["molecules"] == 2662
    assert abs(t["upot_per_molecule"] - r[2] / r[11]) <= tol * abs(r[2] / r[11])
    assert abs(t["ukin_per_molecule"] - r[4] / r[11]) <= tol * abs(r[4] / r[11])
    assert abs(t["p"] - r[6]) <= tol * abs(r[6])


def test_comparison_with_the_full_size_fixtures():
    """vs_reference_fixture: step records built from a fixture's own rows deviate by nothing; a perturbed series by what was put in;
    runs of another size or workload are not compared."""
    import bench
    for (wl, prec), fname in bench.REFERENCE_TRACES.items():
        fx = os.path.join(ROOT, "tests", "golden", fname)
        assert os.path.exists(fx), fname
        g = np.load(fx)
        ref = g["trace"]; n = int(g["molecules"])
        assert ref.shape[0] == int(g["steps"]) + 1 and int(ref[0, 11]) == n
        m = ref.shape[0] - 1
        series = np.zeros((m, 8))      # upot, virial, summv2, sum_iw2, T, p, beta_trans, beta_rot (ls1b200_step_record)
        series[:, 0] = ref[1:, 2]; series[:, 2] = 2.0 * ref[1:, 4]; series[:, 5] = ref[1:, 6]; series[:, 6] = ref[1:, 8]; series[:, 7] = ref[1:, 9]
        out = bench.vs_reference_fixture(wl, prec, n, series)
        assert out["steps_compared"] == m
        for k in ("U_pot_rel_max", "T_rel_max", "p_rel_max", "beta_trans_abs_max", "beta_rot_abs_max"):
            assert out[k] <= 1e-15, (fname, k, out[k])
        series[3, 0] *= 1.0 + 1e-7
        out = bench.vs_reference_fixture(wl, prec, n, series[:10])
        assert out["steps_compared"] == 10 and 0.5e-7 < out["U_pot_rel_max"] < 2e-7
        assert bench.vs_reference_fixture(wl, prec, n + 2, series) is None
    assert bench.vs_reference_fixture("lj", "single", 16000000, np.zeros((5, 8))) is None
