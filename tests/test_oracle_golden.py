"""Pins the CPU oracle (oracle/oracle.cpp) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py). Mirrors the reference's own VectorizedCellProcessorTest cases
(src/particleContainer/adapter/tests/VectorizedCellProcessorTest.cpp:45-388)."""
import numpy as np
import pytest

from helpers import ABS_SCALE, TOL, check_trace_vs_reference, macro_close, golden_names, load, make_sim, ref_by_id, rel_close
from oracle import oracle as orc

FORCE = golden_names("force_")
PERIODIC = golden_names("periodic_")


def test_known_answers_U0_F0():
    """testForcePotentialCalculationU0 / F0 (VectorizedCellProcessorTest.cpp:45-134): 4 LJ molecules on a square."""
    g = load("force_ForceCalculationTestU0")
    for prec, tol in (("double", 1e-8), ("mixed", 1e-4)):
        s = make_sim(orc.OracleSim, g, prec); s.update(); m = s.traverse_cells()
        F = s.download("F")
        assert abs(m.upot) < tol and abs(m.virial - 96.0) < tol * 96
        # molecule ids 1..4 at (1,1) (2,1) (1,2) (2,2): forces +-24 pointing outwards
        expect = np.array([[-24, -24, 0], [24, -24, 0], [-24, 24, 0], [24, 24, 0]], dtype=float)
        assert np.allclose(F, expect, atol=tol * 24)
    g = load("force_ForceCalculationTestF0")
    for prec, tol in (("double", 1e-8), ("mixed", 1e-4)):
        s = make_sim(orc.OracleSim, g, prec); s.update(); m = s.traverse_cells()
        assert abs(m.upot + 4.0) < tol * 4 and np.abs(s.download("F")).max() < max(tol, 1e-7)


@pytest.mark.parametrize("name", FORCE + PERIODIC)
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_forces_vs_reference(name, prec):
    g = load(name)
    s = make_sim(orc.OracleSim, g, prec)
    s.update()
    m = s.traverse_cells()
    ref = ref_by_id(g, prec)
    tol = TOL[prec]
    # bit-exact binning
    assert np.array_equal(s.download("CELL"), ref["cell"]), "cell assignment differs"
    for k, fld in (("F", "F"), ("M", "M"), ("Vi", "VI")):
        ok, worst = rel_close(s.download(fld), ref[k], tol, ABS_SCALE.get(name))
        assert ok, f"{name} {prec} {k}: worst err/bound = {worst:.3g}"
    assert macro_close(m, g, prec, ref, ABS_SCALE.get(name))


@pytest.mark.parametrize("name", FORCE)
def test_vcp_vs_legacy_reference(name):
    """The reference's own differential test: VectorizedCellProcessor (double) vs LegacyCellProcessor + potforce.h."""
    g = load(name)
    a, b = ref_by_id(g, "double"), ref_by_id(g, "legacy")
    for k in ("F", "M", "Vi"):
        assert np.allclose(a[k], b[k], atol=1e-11, rtol=0)


@pytest.mark.parametrize("name", FORCE + PERIODIC)
def test_lj_tables_and_components(name):
    """Comp2Param / Component: both our host mirror (Ensemble.lj_tables) and the oracle's C++ restatement
    against the parameter streams of the reference."""
    g = load(name)
    ens = g["ensemble"]
    rc = float(g["cutoff"])
    for tabs in (ens.lj_tables(rc), orc.lj_tables(ens, rc)):
        for t, k in zip(tabs, ("eps24", "sig2", "shift6")):
            assert np.allclose(t, g["double_" + k], rtol=1e-15, atol=0), k
    cc, cs, ns = ens.pack()
    ref = g["double_comp_ref"]
    for i in range(len(ens.components)):
        assert np.isclose(cc[i].mass, ref[i, 0], rtol=1e-15)
        assert np.allclose(list(cc[i].inertia), ref[i, 1:4], rtol=1e-15, atol=0)
        assert cc[i].rot_dof == int(ref[i, 4])


@pytest.mark.parametrize("name", PERIODIC)
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_halo_copies_vs_reference(name, prec):
    """DomainDecompBase::populateHaloLayerWithCopies: same multiset of (id, cell, position) halo records."""
    g = load(name)
    s = make_sim(orc.OracleSim, g, prec); s.update()
    owner, r, cell = s.halo()
    ids = g["in_id"][owner]
    mine = sorted(zip(ids.tolist(), cell.tolist(), map(tuple, r.tolist())))
    h = g[f"{prec}_mol_halo"].astype(bool)
    theirs = sorted(zip(g[f"{prec}_mol_id"][h].tolist(), g[f"{prec}_mol_cell"][h].tolist(), map(tuple, g[f"{prec}_mol_r"][h].tolist())))
    assert len(mine) == len(theirs)
    assert mine == theirs


@pytest.mark.parametrize("name", ["1clj_1458", "water_432", "multi_250", "1clj_1458_explosion"])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_trace_vs_reference_binary(name, prec):
    """The integrator and the global values, pinned against the UNMODIFIED reference binary (stock Simulation loop, Leapfrog
    with rigid-body rotation, Domain::calculateGlobalValues, velocity scaling; tests/golden/make_traces.py): 100 steps of
    U_pot, U_kin (sum m v^2 + sum I w^2), p, beta_trans, beta_rot. water_432 and multi_250 rotate (beta_rot != 1)."""
    out = check_trace_vs_reference(orc.OracleSim, name, prec)
    print(name, prec, {k: "%.2e" % v for k, v in out.items()})


@pytest.mark.parametrize("fixture", ["generator_1clj_2662.npz", "generator_water_432.npz"])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_cubic_grid_generator_mirror_vs_reference_binary(prec, fixture):
    """The reference's CubicGridGenerator restated (generators.cubic_grid_reference for the lattice -- coordinates are ONE fused
    multiply-add, as the reference binary contracts them -- and generators.reference_velocities for the per-cell std::mt19937(0)
    sequence, IOHelpers::removeMomentum and its rescaling to T) against the step-0 checkpoint and the trace of the UNMODIFIED
    reference binary started from `<generator name="CubicGridGenerator">` (tests/golden/make_generator_golden.py).
    The cell of a molecule -- the reference re-seeds per LinkedCells cell -- comes from the oracle's bit-exact binning.
    generator_water_432.npz: a rotating multi-site model, whose generated state the reference treats as an EXPLOSION (all thermal
    energy in the translation: beta_trans = 0.5^0.4 < 0.9): 100 steps through Domain's heuristics (Domain.cpp:259-314, :363-368) --
    beta_trans = 1, beta_rot^0.0091, per-molecule energy limits, and the warning / error counters that end the special handling at
    step 80, after which the plain factors apply."""
    from helpers import GOLD, check_generated_state_vs_reference
    from ls1_mardyn_b200 import generators, workloads
    import os
    g = dict(np.load(os.path.join(GOLD, fixture)))
    wl = workloads.get(str(g["workload"]))
    comp = wl.ensemble.components[0]
    mass, rot_dof = float(comp.m()), int(comp.Ipa()[1])
    n, L, T, rc = int(g["n_per_dim"]), float(g["length"]), float(g["temperature"]), float(g["rc"])
    mixed = prec == "mixed"
    ids, r, _, lat = generators.cubic_grid_reference(n, [L, L, L], [0, 0, 0], [L, L, L], T, mass, mixed=mixed, with_lattice_index=True)
    s = orc.OracleSim([0, 0, 0], [L, L, L], rc, precision=prec, geometry="xml", periodic=(1, 1, 1))
    s.set_components(wl.ensemble)
    s.upload(ids, np.zeros(len(ids), dtype=np.uint32), r, np.zeros_like(r), None, None)
    s.update()
    v = generators.reference_velocities(lat, s.download("CELL"), T, mass, rot_dof=rot_dof, mixed=mixed)
    assert len(np.unique(np.round(v, 12), axis=0)) == len(np.unique(np.round(g[prec + "_v"], 12), axis=0))      # the cells share one sequence
    out = check_generated_state_vs_reference(orc.OracleSim, prec, r, v, ids, fixture)
    tr = g[prec + "_trace"]
    if str(g["workload"]) == "water":      # the fixture does go through the heuristics and out of them again
        assert int(g[prec + "_explosions"]) == 81 and (tr[1:80, 8] == 1.0).all() and tr[81, 8] < 0.9 and abs(tr[-1, 8] - 1.0) < 0.05
    print(fixture, prec, {k: "%.2e" % x for k, x in out.items()})
