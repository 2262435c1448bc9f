"""Parity of the CUDA path (libls1b200.so through its C ABI) against (a) the golden vectors produced by the unmodified
reference build and (b) the CPU oracle on seeded synthetic inputs. Needs a B200: run with -m gpu.

Mirrors the reference's own VectorizedCellProcessorTest (known answers U0/F0, the 11 differential Vectorization*
fixtures) and LinkedCellsTest (cell assignment, halo), plus multi-step T / U / p traces.
Tolerances (north_star): bit-exact cell binning; F, M, Vi, Upot, Virial within 1e-10 relative (double) / 1e-5 (mixed)."""
import numpy as np
import pytest

from helpers import ABS_SCALE, TOL, check_trace_vs_reference, macro_close, golden_names, load, make_sim, ref_by_id, rel_close

pytestmark = pytest.mark.gpu

FORCE = golden_names("force_")
PERIODIC = golden_names("periodic_")


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ls1_mardyn_b200 import capi
    capi.lib()
    return capi


def test_known_answers_U0_F0(capi):
    g = load("force_ForceCalculationTestU0")
    for prec, tol in (("double", 1e-8), ("mixed", 1e-4)):
        s = make_sim(capi.Context, g, prec); s.update(); m = s.traverse_cells()
        F = s.download("F")
        assert abs(m.upot) < tol and abs(m.virial - 96.0) < tol * 96
        expect = np.array([[-24, -24, 0], [24, -24, 0], [-24, 24, 0], [24, 24, 0]], dtype=float)
        assert np.allclose(F, expect, atol=tol * 24)
    g = load("force_ForceCalculationTestF0")
    for prec, tol in (("double", 1e-8), ("mixed", 1e-4)):
        s = make_sim(capi.Context, g, prec); s.update(); m = s.traverse_cells()
        assert abs(m.upot + 4.0) < tol * 4 and np.abs(s.download("F")).max() < max(tol, 1e-7) * 24


@pytest.mark.parametrize("name", FORCE + PERIODIC)
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_forces_vs_reference_golden(capi, name, prec):
    g = load(name)
    s = make_sim(capi.Context, g, prec)
    s.update()
    m = s.traverse_cells()
    ref = ref_by_id(g, prec)
    tol = TOL[prec]
    assert np.array_equal(s.download("CELL"), ref["cell"]), "cell assignment differs from the reference"
    for k, fld in (("F", "F"), ("M", "M"), ("Vi", "VI")):
        ok, worst = rel_close(s.download(fld), ref[k], tol, ABS_SCALE.get(name))
        assert ok, f"{name} {prec} {k}: worst err/bound = {worst:.3g}"
    assert macro_close(m, g, prec, ref, ABS_SCALE.get(name))
    assert np.array_equal(s.download("ID"), g["in_id"]) and np.array_equal(s.download("CID"), g["in_cid"])


def _lattice_case(n_per_dim, seed, jitter=0.1, ens=None, rho=0.7, T=0.85, rc=2.5):
    from ls1_mardyn_b200 import generators
    from ls1_mardyn_b200.components import lj_fluid
    ens = ens or lj_fluid(rc=rc)
    return generators.lattice(ens, n_per_dim, rho, T, seed=seed, jitter=jitter, rc=rc)


def _mk(cls, ps, rc, prec, periodic=(1, 1, 1)):
    s = cls([0, 0, 0], list(ps.length), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=periodic)
    s.set_components(ps.ensemble)
    s.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    return s


@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_1clj_lattice_vs_oracle(capi, prec):
    """16k-molecule periodic 1CLJ liquid: bit-exact cells, forces / virial / energy against the CPU oracle."""
    from oracle import oracle as orc
    ps = _lattice_case(20, seed=11)
    a = _mk(capi.Context, ps, 2.5, prec); b = _mk(orc.OracleSim, ps, 2.5, prec)
    a.update(); b.update()
    assert np.array_equal(a.download("CELL"), b.download("CELL"))
    assert np.array_equal(a.download("R"), b.download("R"))   # periodic wrap is bit-exact
    ma, mb = a.traverse_cells(), b.traverse_cells()
    tol = TOL[prec]
    for f in ("F", "VI"):
        ok, worst = rel_close(a.download(f), b.download(f), tol)
        assert ok, (f, worst)
    assert abs(ma.upot - mb.upot) <= tol * abs(mb.upot) and abs(ma.virial - mb.virial) <= tol * max(abs(mb.virial), np.abs(b.download("VI")).sum())
    # FlopCounter counts (pair set identity): integers, must agree exactly
    fa, fb = a.count_flops(), b.count_flops()
    for k, _ in fa._fields_:
        assert getattr(fa, k) == getattr(fb, k), k


def test_multicomponent_1clj_fast_path(capi):
    """Two single-centre LJ components (Lorentz-Berthelot mixing): the 1CLJ fast path with table lookups."""
    from oracle import oracle as orc
    from ls1_mardyn_b200.components import Component, Ensemble, LJcenter
    ens = Ensemble([Component(ljcenters=[LJcenter((0, 0, 0), 1.0, 1.0, 1.0, 2.5, False)]),
                    Component(ljcenters=[LJcenter((0, 0, 0), 2.0, 0.7, 1.15, 2.5, True)])],
                   xi=np.array([[1, 0.9], [1, 1]]), eta=np.array([[1, 1.05], [1, 1]]))
    ps = _lattice_case(8, seed=5, ens=ens, rho=0.6)
    for prec in ("double", "mixed"):
        a = _mk(capi.Context, ps, 2.5, prec); b = _mk(orc.OracleSim, ps, 2.5, prec)
        a.update(); b.update(); ma, mb = a.traverse_cells(), b.traverse_cells()
        for f in ("F", "VI"):
            ok, worst = rel_close(a.download(f), b.download(f), TOL[prec])
            assert ok, (prec, f, worst)
        assert abs(ma.upot - mb.upot) <= TOL[prec] * abs(mb.upot)


@pytest.mark.parametrize("kernel", ["tiled", "global", "mma", "tc", "v3", "v3:4", "v3:8:2", "pf"])
def test_1clj_kernel_variants_vs_oracle(capi, kernel):
    """The alternative 1CLJ force kernels (LS1B200_LJ_KERNEL): the global-memory v1 kernel, the two tensor-core prefilter
    experiments (mma.sync TF32 / tcgen05 FP16 with TMEM; both fall back to the tiled kernel in the double build) and the round-2
    experiment v3 (molecule-granular chunks; "v3:S[:A]" = sorted by S fine x-bins per cell with x-trimmed candidate ranges, lane
    groups padded to A). Same bar as
    the production kernel: forces, virial, energy against the CPU oracle, single- and multi-component."""
    import os
    from oracle import oracle as orc
    from ls1_mardyn_b200.components import Component, Ensemble, LJcenter
    two = Ensemble([Component(ljcenters=[LJcenter((0, 0, 0), 1.0, 1.0, 1.0, 2.5, False)]),
                    Component(ljcenters=[LJcenter((0, 0, 0), 2.0, 0.7, 1.15, 2.5, True)])],
                   xi=np.array([[1, 0.9], [1, 1]]), eta=np.array([[1, 1.05], [1, 1]]))
    for case, ps in enumerate((_lattice_case(20, seed=11), _lattice_case(8, seed=5, ens=two, rho=0.6), _lattice_case(14, seed=3, jitter=0.15))):
        for prec in ("mixed", "double"):
            parts = kernel.split(":")
            os.environ["LS1B200_LJ_KERNEL"] = parts[0]
            if len(parts) > 1:
                os.environ["LS1B200_FINE_S"] = parts[1]
            if len(parts) > 2:
                os.environ["LS1B200_V3_ALIGN"] = parts[2]
            if parts[0] == "pf":
                os.environ["LS1B200_PF_CHECK"] = "1"      # the prefilter kernel verifies that every window translation is exact
            try:
                a = _mk(capi.Context, ps, 2.5, prec)
            finally:
                for k in ("LS1B200_LJ_KERNEL", "LS1B200_FINE_S", "LS1B200_V3_ALIGN", "LS1B200_PF_CHECK"):
                    os.environ.pop(k, None)
            b = _mk(orc.OracleSim, ps, 2.5, prec)
            a.update(); b.update(); ma, mb = a.traverse_cells(), b.traverse_cells()
            tol = TOL[prec]
            for f in ("F", "VI"):
                ok, worst = rel_close(a.download(f), b.download(f), tol)
                assert ok, (kernel, case, prec, f, worst)
            assert abs(ma.upot - mb.upot) <= tol * abs(mb.upot), (kernel, case, prec, ma.upot, mb.upot)
            assert abs(ma.virial - mb.virial) <= tol * max(abs(mb.virial), np.abs(b.download("VI")).sum()), (kernel, case, prec)


@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_leapfrog_and_thermostat_steps_vs_oracle(capi, prec):
    """Phase-by-phase: pre -> update -> traverse -> post -> scale, 5 steps, state compared after every phase."""
    from oracle import oracle as orc
    g = load("periodic_water_432")
    a = make_sim(capi.Context, g, prec); b = make_sim(orc.OracleSim, g, prec)
    for s in (a, b):
        s.update(); s.traverse_cells()
    dt = 0.0667516
    tol = 1e-9 if prec == "double" else 1e-4   # mixed: the state integrates forces that agree to 1e-5
    for step in range(5):
        for s in (a, b):
            s.leapfrog_pre(dt)
        for f in ("R", "V", "Q", "L"):
            ok, worst = rel_close(a.download(f), b.download(f), tol); assert ok, (step, "pre", f, worst)
        for s in (a, b):
            s.update()
        assert np.array_equal(a.download("CELL"), b.download("CELL"))
        ma, mb = a.traverse_cells(), b.traverse_cells()
        for f in ("F", "M", "VI"):
            ok, worst = rel_close(a.download(f), b.download(f), 10 * tol); assert ok, (step, "force", f, worst)
        ka, kb = a.leapfrog_post(dt), b.leapfrog_post(dt)
        assert ka.n_molecules == kb.n_molecules and ka.rot_dof == kb.rot_dof
        assert abs(ka.summv2 - kb.summv2) <= tol * kb.summv2 and abs(ka.sum_iw2 - kb.sum_iw2) <= 10 * tol * kb.sum_iw2
        for s in (a, b):
            s.scale_velocities(0.999, 1.002)
        for f in ("V", "L"):
            ok, worst = rel_close(a.download(f), b.download(f), tol); assert ok, (step, "scale", f, worst)


@pytest.mark.parametrize("name,dt,nsteps", [("periodic_1clj_1458", 0.002, 100), ("periodic_water_432", 0.0166879, 40), ("periodic_multi_250", 0.0667516, 40)])      # water: the time step of tests/golden/make_traces.py (at the full step this lattice start heats by a factor 3 within 10 steps: the reference itself then switches to its "explosion" heuristics, Domain.cpp:260-306, and FP32 summation order decides the 7th digit)
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_simulate_trace_vs_oracle(capi, name, dt, nsteps, prec):
    """Fused multi-step loop (NVT velocity scaling): T, U_pot, p per step against the oracle's loop."""
    from oracle import oracle as orc
    g = load(name)
    a = make_sim(capi.Context, g, prec); b = make_sim(orc.OracleSim, g, prec)
    for s in (a, b):
        s.update(); s.traverse_cells()
    T = float(g["temperature"])
    ra = a.simulate(nsteps, dt, T); rb = b.simulate(nsteps, dt, T)
    cols = dict(upot=0, virial=1, summv2=2, temperature=4, pressure=5, beta_trans=6)
    tol = 1e-9 if prec == "double" else 1e-6
    for k, i in cols.items():
        scale = np.abs(rb[:, i]).max()
        # p = rho*T + rho*W/3N is a near cancellation: compare against the size of its terms (hard parts)
        if k == "pressure":
            N = len(g["in_id"]); V = float(np.prod(g["length"]))
            scale = max(scale, (N / V) * np.abs(rb[:, 4]).max())
        err = np.abs(ra[:, i] - rb[:, i]).max() / scale
        assert err <= tol, f"{name} {prec} {k}: {err:.3g}"
    ok, worst = rel_close(a.download("R"), b.download("R"), 1e-7 if prec == "double" else 1e-4, scale=1.0)
    assert ok, worst


@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_fused_step_switches(capi, prec):
    """The fused step of ls1b200_simulate (k_pre_key, kinetic sum in the force epilogue, lazy permutation, chunk list beside the sort)
    against the same loop with each of its parts switched off (LS1B200_FUSED / LS1B200_LAZY / LS1B200_CHUNKS_EARLY = 0): same
    floating-point sequence per molecule; only the order of the kinetic sum differs between fused and phase-by-phase."""
    import os
    from ls1_mardyn_b200 import generators
    ps = generators.lj_lattice(24, seed=9, jitter=0.1); ps.v *= 2.0
    dt, T, n = 0.002, 0.85, 30

    def run(env):
        os.environ.update(env)
        try:
            c = capi.Context([0, 0, 0], list(ps.length), 2.5, precision=prec)
        finally:
            for k in env:
                os.environ.pop(k, None)
        c.set_components(ps.ensemble); c.upload(ps.id, ps.cid, ps.r, ps.v)
        c.update(); c.traverse_cells()
        rec = c.simulate(n, dt, T)
        o = np.argsort(c.download("ID"))
        return rec, c.download("R")[o], c.download("V")[o], c.download("F")[o]

    ref = run({})
    for env in ({"LS1B200_FUSED": "0"}, {"LS1B200_LAZY": "0"}, {"LS1B200_CHUNKS_EARLY": "0"}):
        got = run(env)
        for i in (0, 1, 2, 4, 5, 6):
            err = np.abs(got[0][:, i] - ref[0][:, i]).max() / np.abs(ref[0][:, i]).max()
            assert err <= 1e-12, (env, prec, i, err)
        for k in (1, 2, 3):
            assert np.allclose(got[k], ref[k], rtol=1e-11, atol=1e-11 * np.abs(ref[k]).max()), (env, prec, k)


def test_simulate_equals_phase_calls(capi):
    """ls1b200_simulate is the same arithmetic as the individual phase calls (thermostat applied on load)."""
    g = load("periodic_1clj_1458")
    a = make_sim(capi.Context, g, "double"); b = make_sim(capi.Context, g, "double")
    for s in (a, b):
        s.update(); s.traverse_cells()
    dt, T, n = 0.002, 0.85, 6
    ra = a.simulate(n, dt, T)
    N = len(g["in_id"])
    for i in range(n):
        b.leapfrog_pre(dt); b.update(); m = b.traverse_cells(); k = b.leapfrog_post(dt)
        bt = (3.0 * N * T / k.summv2) ** 0.4
        assert abs(m.upot - ra[i, 0]) <= 1e-12 * abs(m.upot) and abs(bt - ra[i, 6]) <= 1e-13
        b.scale_velocities(bt, 1.0)
    assert np.allclose(a.download("V"), b.download("V"), rtol=1e-12, atol=1e-14)
    assert np.allclose(a.download("R"), b.download("R"), rtol=1e-13, atol=1e-13)


def test_async_download_matches_download(capi):
    """ls1b200_download_begin / _end (side-stream copy overlapping the next kernels) returns what ls1b200_download returns."""
    ps = _lattice_case(12, seed=4, jitter=0.1)
    a = _mk(capi.Context, ps, 2.5, "mixed")
    a.update()
    r_async = np.empty((len(ps.id), 3))
    a.download_begin("R", r_async)              # positions while the forces are computed
    a.traverse_cells()
    a.download_end()
    assert np.array_equal(r_async, a.download("R"))
    f_async = np.empty((len(ps.id), 3))
    a.download_begin("F", f_async); a.download_end()
    assert np.array_equal(f_async, a.download("F"))
    with pytest.raises(capi.Ls1B200Error):
        a.download_begin("F", f_async); a.download_begin("V", f_async)      # one download in flight
    a.download_end()


def test_update_is_deterministic_and_idempotent(capi):
    ps = _lattice_case(12, seed=3, jitter=0.2)
    a = _mk(capi.Context, ps, 2.5, "mixed"); b = _mk(capi.Context, ps, 2.5, "mixed")
    a.update(); a.traverse_cells(); b.update(); b.update(); b.update(); b.traverse_cells()
    assert np.array_equal(a.download("CELL"), b.download("CELL"))
    assert np.array_equal(a.download("F"), b.download("F"))          # bit-identical: order inside a cell is deterministic
    assert np.array_equal(a.download("R"), b.download("R"))


def test_edge_cases(capi):
    from ls1_mardyn_b200.components import lj_fluid
    ens = lj_fluid()
    # empty container
    s = capi.Context([0, 0, 0], [10, 10, 10], 2.5); s.set_components(ens)
    s.upload(np.zeros(0, np.uint64), np.zeros(0, np.uint32), np.zeros((0, 3)), np.zeros((0, 3)))
    s.update(); m = s.traverse_cells(); assert m.upot == 0 and m.virial == 0
    # a single molecule, and two molecules interacting only through the periodic image
    s = capi.Context([0, 0, 0], [10, 10, 10], 2.5, precision="double"); s.set_components(ens)
    s.upload(np.array([7], np.uint64), np.zeros(1, np.uint32), np.array([[0.1, 5, 5]]), np.zeros((1, 3)))
    s.update(); m = s.traverse_cells(); assert m.upot == 0 and np.all(s.download("F") == 0)
    s.upload(np.array([1, 2], np.uint64), np.zeros(2, np.uint32), np.array([[0.5, 5, 5], [9.5, 5, 5]]), np.zeros((2, 3)))
    s.update(); m = s.traverse_cells()
    r = 1.0; u = 4 * (r ** -12 - r ** -6); f = 24 * (2 * r ** -12 - r ** -6) / r
    F = s.download("F")
    assert abs(m.upot - u) < 1e-12 and abs(F[0, 0] - f) < 1e-10 * f and abs(F[1, 0] + f) < 1e-10 * f
    assert abs(m.virial - f * r) < 1e-10 * f
    # molecules that left the box are wrapped by update() exactly like DomainDecompBase::handleDomainLeavingParticles
    s.upload(np.array([1, 2], np.uint64), np.zeros(2, np.uint32), np.array([[-0.25, 5, 5], [10.0, 11.5, -1e-17]]), np.zeros((2, 3)))
    s.update(); R = s.download("R")
    assert R[0, 0] == 9.75 and R[1, 0] == 0.0 and R[1, 1] == 1.5 and 0 <= R[1, 2] < 10
    # errors are reported, never thrown past the ABI
    with pytest.raises(capi.Ls1B200Error):
        capi.Context([0, 0, 0], [1, 1, 1], 2.5)
    with pytest.raises(capi.Ls1B200Error):
        s2 = capi.Context([0, 0, 0], [10, 10, 10], 2.5); s2.update()


def test_kernels_were_launched(capi):
    assert capi.kernel_launch_count() > 0


def _sig_ensembles():
    """One single-component ensemble per compile-time site signature of force_multi_tiled.cu (plus one that has none)."""
    from ls1_mardyn_b200 import workloads
    from ls1_mardyn_b200.components import Charge, Component, Dipole, Ensemble, LJcenter, Quadrupole
    lj = lambda z, m=1.0: LJcenter((0.0, 0.0, z), m, 1.0, 1.0, 0.0, False)
    return {
        "water_1lj3q": (workloads.water_like(), 4.935e-3, 9.437e-4, 14.9),
        "co2_2cljq": (workloads.co2_2cljq(), 0.0137, 293.44, 14.9235 / 2),
        "2cljd": (Ensemble([Component(ljcenters=[lj(-0.3), lj(0.3)], dipoles=[Dipole((0, 0, 0), (0, 0, 1), 1.2)])]), 0.5, 1.2, 2.5),
        "stockmayer": (Ensemble([Component(ljcenters=[lj(0.0)], dipoles=[Dipole((0, 0, 0), (0, 0, 1), 1.5)], I_override=(0.1, 0.1, 0.1))]), 0.6, 1.2, 2.5),
        "2clj": (Ensemble([Component(ljcenters=[lj(-0.25), lj(0.25)])]), 0.5, 1.2, 2.5),
        "3lj3q": (Ensemble([Component(ljcenters=[lj(-0.4), lj(0.0, 2.0), lj(0.4)],
                                       charges=[Charge((0, 0, -0.4), 0.0, -0.3), Charge((0, 0, 0.0), 0.0, 0.6), Charge((0, 0, 0.4), 0.0, -0.3)])]), 0.4, 1.5, 2.5),
        "generic_1lj1q1d1quad": (Ensemble([Component(ljcenters=[lj(0.0)], charges=[Charge((0.2, 0, 0), 0.5, 0.4), Charge((-0.2, 0, 0), 0.5, -0.4)],
                                                     dipoles=[Dipole((0, 0.1, 0), (0, 1, 0), 0.8)], quadrupoles=[Quadrupole((0, 0, 0.1), (0, 0, 1), 0.9)])]), 0.4, 1.5, 2.5),
    }


@pytest.mark.parametrize("name", ["water_1lj3q", "co2_2cljq", "2cljd", "stockmayer", "2clj", "3lj3q", "generic_1lj1q1d1quad"])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_site_signature_kernels_vs_oracle(capi, name, prec):
    """Every specialisation of the tiled multi-site kernel (and the generic one) against the CPU oracle: randomly oriented
    molecules on a jittered periodic lattice; F, M, Vi per molecule, Upot, virial."""
    import os
    from oracle import oracle as orc
    from ls1_mardyn_b200 import generators
    ens, rho, T, rc = _sig_ensembles()[name]
    ps = generators.lattice(ens, 6, rho, T, seed=17, jitter=0.12, rc=rc)      # 432 molecules
    b = _mk(orc.OracleSim, ps, rc, prec); b.update(); mb = b.traverse_cells()
    tol = TOL[prec]
    # specialised; specialised with the site records of the windows staged in shared memory (where the signature allows it); specialised
    # with two neighbours prefetched; generic
    for kernel in ("", "stage", "pd2", "generic"):
        os.environ["LS1B200_MULTI_KERNEL"] = kernel if kernel == "generic" else ""
        if kernel == "stage":
            os.environ["LS1B200_MULTI_STAGE"] = "1"
        if kernel == "pd2":
            os.environ["LS1B200_MULTI_PD2"] = "1"
        try:
            a = _mk(capi.Context, ps, rc, prec)
        finally:
            for k in ("LS1B200_MULTI_KERNEL", "LS1B200_MULTI_STAGE", "LS1B200_MULTI_PD2"):
                os.environ.pop(k, None)
        a.update(); ma = a.traverse_cells()
        assert np.array_equal(a.download("CELL"), b.download("CELL"))
        for f in ("F", "M", "VI"):
            ok, worst = rel_close(a.download(f), b.download(f), tol)
            assert ok, (name, prec, kernel or "specialised", f, worst)
        terms = np.abs(b.download("VI")).sum()
        assert abs(ma.upot - mb.upot) <= tol * max(abs(mb.upot), terms) and abs(ma.virial - mb.virial) <= tol * max(abs(mb.virial), terms)


@pytest.mark.parametrize("name", ["1clj_1458", "water_432", "multi_250", "1clj_1458_explosion"])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_trace_vs_reference_binary(capi, name, prec):
    """ls1b200_simulate (k_leapfrog_pre / _post with quaternions, k_globals, thermostat on load) against 100 steps of the
    UNMODIFIED reference binary's ResultWriter output (tests/golden/trace_*.npz, made by tests/golden/make_traces.py):
    U_pot, U_kin, p, beta_trans, beta_rot at 1e-9 (double) / 1e-6 (mixed). water_432 and multi_250 are rotating multi-site
    ensembles with beta_rot != 1."""
    out = check_trace_vs_reference(capi.Context, name, prec)
    print(name, prec, {k: "%.2e" % v for k, v in out.items()})


@pytest.mark.parametrize("prec", ["mixed", "double"])
def test_device_cubic_grid_generator(capi, prec):
    """ls1b200_generate_cubic_grid (the reference's CubicGridGenerator on the device) against its host mirror
    (generators.cubic_grid_reference: same lattice arithmetic as ParticleCellBase::initCubicGrid, same counter-based uniform numbers):
    ids and positions bit-identical, velocities to rounding of sin / cos / acos, |v| = sqrt(3 T / m), total momentum removed; a
    2 x 2 x 1 decomposition generates exactly the molecules of the single domain, block by block; the generated state then runs
    through update / traverse_cells like an uploaded one (forces against the oracle fed with the downloaded state)."""
    from ls1_mardyn_b200 import generators
    from ls1_mardyn_b200.components import lj_fluid
    from ls1_mardyn_b200.decomposition import DomainDecomposition
    from oracle import oracle as orc
    n, rho, T, rc = 11, 0.7, 0.85, 2.5
    N = 2 * n ** 3; L = (N / rho) ** (1.0 / 3.0); box = [L, L, L]
    ens = lj_fluid(rc=rc)
    ctx = capi.Context([0, 0, 0], box, rc, precision=prec, geometry="xml", periodic=(1, 1, 1))
    ctx.set_components(ens)
    assert ctx.generate_cubic_grid(n, box, T, seed=7, remove_momentum=False) == N
    ids, r, v = ctx.download("ID"), ctx.download("R"), ctx.download("V")
    rid, rr, rv = generators.cubic_grid_reference(n, box, [0, 0, 0], box, T, 1.0, seed=7, mixed=(prec == "mixed"))
    o = np.argsort(ids)
    assert np.array_equal(ids[o], rid) and np.array_equal(r[o], rr)
    assert np.abs(v[o] - rv).max() <= 1e-13 * np.abs(rv).max()
    assert np.allclose(np.linalg.norm(v, axis=1), np.sqrt(3 * T), rtol=1e-13)
    # momentum removal
    ctx.generate_cubic_grid(n, box, T, seed=7, remove_momentum=True)
    v2 = ctx.download("V")
    # IOHelpers::removeMomentum: the mean velocity is subtracted, then the velocities are scaled back to the target temperature
    vs = v - v.mean(axis=0)
    assert np.abs(v2.sum(axis=0)).max() <= 1e-9 and np.abs(v2 - vs * np.sqrt(T / ((vs ** 2).sum() / (3 * N)))).max() <= 1e-13
    assert abs((v2 ** 2).sum() / (3 * N) - T) <= 1e-13
    # the generated state is a normal state: forces against the oracle
    ctx.update(); m = ctx.traverse_cells()
    b = orc.OracleSim([0, 0, 0], box, rc, precision=prec, geometry="xml", periodic=(1, 1, 1))
    b.set_components(ens); b.upload(ctx.download("ID"), ctx.download("CID"), ctx.download("R"), v2, None, None)
    b.update(); mb = b.traverse_cells()
    assert np.array_equal(ctx.download("CELL"), b.download("CELL"))
    # on the perfect lattice the forces cancel to rounding: the bound is relative to the magnitude of the pair terms (the virial)
    ok, worst = rel_close(ctx.download("F"), b.download("F"), TOL[prec], scale=float(np.abs(b.download("VI")).mean())); assert ok, worst
    ok, worst = rel_close(ctx.download("VI"), b.download("VI"), TOL[prec]); assert ok, worst
    assert abs(m.upot - mb.upot) <= TOL[prec] * abs(mb.upot)
    # blocks of a decomposition: the union is the single domain, every molecule exactly once
    seen = []
    for rank in range(4):
        dd = DomainDecomposition(rank, 4, np.array(box), (2, 2, 1))
        c = capi.Context(list(dd.box_min()), list(dd.box_max()), rc, precision=prec, geometry="xml", periodic=dd.periodic())
        c.set_components(ens); c.decomp_attach(dd)
        nb = c.generate_cubic_grid(n, box, T, seed=7, remove_momentum=False)
        bid, br = c.download("ID"), c.download("R")
        assert len(bid) == nb and (br >= np.array(dd.box_min())).all() and (br < np.array(dd.box_max())).all()
        k = np.searchsorted(rid, bid)
        assert np.array_equal(rid[k], bid) and np.array_equal(br, rr[k]) and np.abs(c.download("V") - rv[k]).max() <= 1e-13 * np.abs(rv).max()
        seen.append(bid)
    assert np.array_equal(np.sort(np.concatenate(seen)), rid)


@pytest.mark.parametrize("fixture", ["generator_1clj_2662.npz", "generator_water_432.npz"])
@pytest.mark.parametrize("prec", ["mixed", "double"])
def test_device_generator_reference_velocities_vs_reference_binary(capi, prec, fixture):
    """ls1b200_generate_cubic_grid with LS1B200_GEN_REFERENCE_VELOCITIES | LS1B200_GEN_REMOVE_MOMENTUM against the UNMODIFIED reference
    binary started from `<generator name="CubicGridGenerator">` (tests/golden/generator_1clj_2662.npz, make_generator_golden.py):
    positions bit-identical to its step-0 checkpoint, velocities (per-cell std::mt19937(0) sequence -- the device evaluates every
    molecule's insertion rank inside its LinkedCells cell in closed form --, momentum removed, rescaled to T) to 1e-13, and the
    generated context then follows the reference binary's trace (U_pot, U_kin, p, beta at 1e-9 / 1e-6). Blocks of a
    decomposition re-seed per cell of THEIR grid, like the ranks of the reference would: checked against the host mirror fed with
    the block's own cells. generator_water_432.npz: a rotating multi-site model whose generated state the reference treats as an
    EXPLOSION -- 100 steps through Domain's heuristics on the device (k_globals, the selective scaling applied on load by
    k_leapfrog_pre / k_scale), see tests/test_oracle_golden.py::test_cubic_grid_generator_mirror_vs_reference_binary."""
    import os
    from helpers import GOLD, TOL, TRACE_TOL, check_generated_state_vs_reference, compare_trace_rows
    from ls1_mardyn_b200 import generators, workloads
    from ls1_mardyn_b200.decomposition import DomainDecomposition
    g = dict(np.load(os.path.join(GOLD, fixture)))
    n, L, T, rc, steps = int(g["n_per_dim"]), float(g["length"]), float(g["temperature"]), float(g["rc"]), int(g["steps"])
    box = [L, L, L]; mixed = prec == "mixed"
    water = str(g["workload"]) == "water"
    ens = workloads.get(str(g["workload"])).ensemble
    ctx = capi.Context([0, 0, 0], box, rc, precision=prec, geometry="xml", periodic=(1, 1, 1))
    ctx.set_components(ens)
    N = ctx.generate_cubic_grid(n, box, T, remove_momentum=True, reference_velocities=True)
    assert N == 2 * n ** 3
    r, v, ids = ctx.download("R"), ctx.download("V"), ctx.download("ID")
    out = check_generated_state_vs_reference(capi.Context, prec, r, v, ids, fixture)      # state == checkpoint; an uploaded copy follows the trace
    # ... and so does the generated context itself (no upload in between)
    ctx.update(); m0 = ctx.traverse_cells()
    term_scale = tol_macro = None
    if water and mixed:      # as in check_generated_state_vs_reference: cancelling U_pot, 40 steps
        term_scale = float(np.abs(ctx.download("VI")).sum()); steps = 40; tol_macro = TOL["mixed"]
    rec = ctx.simulate(steps, float(g["dt"]), T)
    out2 = compare_trace_rows(m0, rec, g[prec + "_trace"], steps, N, L ** 3, TRACE_TOL[prec], ("generated context", prec), term_scale, tol_macro)
    print(fixture, prec, {k: "%.2e" % x for k, x in out.items()}, {k: "%.2e" % x for k, x in out2.items()})
    if water:
        return      # (a box of three cutoff radii is too small for blocks)
    # blocks: a 2 x 2 x 2 decomposition and an uneven 3 x 1 x 2 one
    for nranks, grid in ((8, (2, 2, 2)), (6, (3, 1, 2))):
        total = 0
        for rank in range(nranks):
            dd = DomainDecomposition(rank, nranks, np.array(box), grid)
            c = capi.Context(list(dd.box_min()), list(dd.box_max()), rc, precision=prec, geometry="xml", periodic=dd.periodic())
            c.set_components(ens); c.decomp_attach(dd)
            nb = c.generate_cubic_grid(n, box, T, remove_momentum=False, reference_velocities=True)
            bid, br, bv = c.download("ID"), c.download("R"), c.download("V")
            rid, rr, _, lat = generators.cubic_grid_reference(n, box, dd.box_min(), dd.box_max(), T, 1.0, mixed=mixed, with_lattice_index=True)
            o = np.argsort(bid)
            assert nb == len(rid) and np.array_equal(bid[o], rid) and np.array_equal(br[o], rr)
            # the block's own cell grid, from a plain context over the same box
            k = capi.Context(list(dd.box_min()), list(dd.box_max()), rc, precision=prec, geometry="xml", periodic=(0, 0, 0))
            k.set_components(ens); k.upload(rid, np.zeros(len(rid), dtype=np.uint32), rr, np.zeros_like(rr), None, None); k.update()
            mv = generators.reference_velocities(lat, k.download("CELL"), T, 1.0, mixed=mixed, remove_momentum=False)
            assert np.abs(bv[o] - mv).max() <= 1e-13 * np.abs(mv).max(), (grid, rank)
            total += nb
        assert total == N


@pytest.mark.parametrize("workload,prec,fixture", [("lj", "mixed", "reference_trace_lj16M.npz"), ("lj", "double", "reference_trace_lj16M_double.npz"),
                                                   ("argon", "mixed", "reference_trace_argon_full.npz"), ("water", "mixed", "reference_trace_water_full.npz"),
                                                   ("co2", "mixed", "reference_trace_co2_full.npz")])
def test_benchmark_state_vs_reference_binary(capi, workload, prec, fixture):
    """Parity AT THE BENCHMARK'S FULL SIZE (BASELINE.json configs[1]: 16 000 000 1CLJ molecules, mixed precision and the double build;
    configs[0]: the Argon state point at 101 306 molecules): the device generator builds the state the unmodified reference binary
    builds from its CubicGridGenerator, and ls1b200_simulate (the fused step that bench.py times) is held against the reference
    binary's own ResultWriter rows on that configuration, 100 steps (tests/golden/reference_trace_*.npz, made by
    make_reference_traces_full_size.py on the build container's host cores): U_pot, U_kin (T), p, beta_trans at 1e-6 relative
    (mixed; north_star's bound for T and p over 100 steps) / 1e-9 (double).
    configs[2] and [3] -- 4 000 752 water-like and 8 039 358 2CLJQ molecules, 20 steps: the generated state of a rotating model (all
    molecules aligned, the whole thermal energy in the translation) is an EXPLOSION by the reference's definition, so these two runs go
    through Domain's heuristics (beta_trans = 1, beta_rot^0.0091, per-molecule limits) on the device; their U_pot is what is left of
    pair terms orders of magnitude larger, and is held -- like p -- to north_star's 1e-5 for mixed-precision energy and virial
    sums, relative to the larger of the value and sum |Vi| (strict figure printed); U_kin and the thermostat factors to 1e-6."""
    import os
    from helpers import GOLD, TRACE_TOL, compare_trace_rows
    from ls1_mardyn_b200 import workloads
    fx = os.path.join(GOLD, fixture)
    if not os.path.exists(fx):
        pytest.skip("tests/golden/%s not generated" % fixture)
    g = np.load(fx)
    wl = workloads.get(workload)
    n, L, T, rc, dt, steps = int(g["n_per_dim"]), float(g["length"]), float(g["temperature"]), float(g["rc"]), float(g["dt"]), int(g["steps"])
    assert n == wl.n_per_dim and rc == wl.rc and dt == wl.dt
    ctx = capi.Context([0, 0, 0], [L, L, L], rc, precision=prec, geometry="xml", periodic=(1, 1, 1))
    ctx.set_components(wl.ensemble)
    N = ctx.generate_cubic_grid(n, [L, L, L], T, remove_momentum=True, reference_velocities=True)
    assert N == int(g["molecules"]) == 2 * n ** 3
    ctx.update(); m0 = ctx.traverse_cells()
    term_scale = tol_macro = None
    if workload in ("water", "co2"):
        from helpers import TOL
        term_scale = float(np.abs(ctx.download("VI")).sum()); tol_macro = TOL[prec]
    rec = ctx.simulate(steps, dt, T)
    ctx.close()
    for m in sorted({min(20, steps), min(50, steps), steps}):      # how the deviation grows (printed before anything is asserted)
        o = compare_trace_rows(m0, rec[:m], g["trace"], m, N, L ** 3, 1.0, (workload, prec, m), term_scale)
        print(workload, prec, N, "molecules, first", m, "steps vs the reference binary:", {k: "%.2e" % x for k, x in o.items()})
    compare_trace_rows(m0, rec, g["trace"], steps, N, L ** 3, TRACE_TOL[prec], (workload, prec, "full size"), term_scale, tol_macro)


def test_forces_one_million_molecules_vs_reference_on_the_fly(capi, tmp_path):
    """Parity at a size where the golden fixtures cannot go (they are 250 - 1 458 molecules): a 1 024 000-molecule jittered periodic
    1CLJ liquid (100 x 100 x... BCC lattice 2 * 80^3) is written out, the UNMODIFIED reference (oracle/_ref/mixed/refdump, the dump
    driver linked against the reference's own objects) computes cells, forces, virials and the macroscopic sums on the GPU box's
    host cores, and the CUDA path is held to the same bar as on the fixtures: bit-exact cells, F / Vi / Upot / virial at 1e-5 (mixed).
    Skipped where the reference build is not in the snapshot."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "mixed", "refdump")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mixed/refdump not built (oracle/build_ref.sh)")
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    from make_golden import read_refdump
    from ls1_mardyn_b200 import generators
    from ls1_mardyn_b200.ascii_reader import write_phase_space
    ps = generators.lj_lattice(80, seed=3, jitter=0.1)
    ps.r = ps.r.astype(np.float32).astype(np.float64)
    inp = str(tmp_path / "big.inp"); out = str(tmp_path / "dump.bin")
    write_phase_space(inp, ps)
    subprocess.check_call([exe, inp, repr(2.5), out, "vcp", "periodic", "xml"], stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1)), timeout=1500)
    d = read_refdump(out)
    sel = ~d["mol_halo"].astype(bool)
    o = np.argsort(d["mol_id"][sel])
    a = _mk(capi.Context, ps, 2.5, "mixed"); a.update(); m = a.traverse_cells()
    oa = np.argsort(a.download("ID"))
    assert np.array_equal(a.download("ID")[oa], d["mol_id"][sel][o])
    assert np.array_equal(a.download("CELL")[oa], d["mol_cell"][sel][o]), "cell assignment differs from the reference at 1 M molecules"
    for f, k in (("F", "mol_F"), ("VI", "mol_Vi")):
        ok, worst = rel_close(a.download(f)[oa], d[k][sel][o], TOL["mixed"]); assert ok, (f, worst)
    terms = float(np.abs(d["mol_Vi"][sel]).sum())
    assert abs(m.upot - d["upot"]) <= TOL["mixed"] * max(abs(d["upot"]), terms) and abs(m.virial - d["virial"]) <= TOL["mixed"] * max(abs(d["virial"]), terms)
    print("1 M molecules: strict relative error U_pot %.2e, virial %.2e" % (abs(m.upot - d["upot"]) / abs(d["upot"]), abs(m.virial - d["virial"]) / abs(d["virial"])))
