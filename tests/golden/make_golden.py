#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/{mixed,double}/refdump and MarDyn,
built from /root/reference by oracle/build_ref.sh) on the reference's own test fixtures and on synthetic systems.

Run in the build container only (needs /root/reference); the .npz files are committed and are all the GPU box sees.

  python tests/golden/make_golden.py            # everything
  python tests/golden/make_golden.py forces     # only the force fixtures
"""
from __future__ import annotations

import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ls1_mardyn_b200.ascii_reader import PhaseSpace, read_phase_space, write_phase_space  # noqa: E402
from ls1_mardyn_b200 import generators  # noqa: E402

REF = os.environ.get("LS1_REF", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref")

# fixture, cutoff: exactly the cases of VectorizedCellProcessorTest.cpp:45-134, :136-388
FORCE_FIXTURES = [
    ("ForceCalculationTestU0", 1.1), ("ForceCalculationTestF0", 1.3),
    ("VectorizationLennardJones", 35.0), ("VectorizationLennardJones1CLJ", 35.0),
    ("VectorizationCharge", 35.0), ("VectorizationChargeDipole", 35.0), ("VectorizationChargeQuadrupole", 35.0),
    ("VectorizationDipole", 35.0), ("VectorizationDipoleQuadrupole", 35.0), ("VectorizationQuadrupole", 35.0),
    ("VectorizationWater", 6.16), ("VectorizationMultiComponentMultiPotentials", 35.0),
    ("VectorizationMultiComponentMultiPotentials_50_molecules", 35.0),
]


def read_refdump(path):
    b = open(path, "rb").read()
    assert b[:8] == b"LS1REFD1"
    o = [8]

    def f64(n=1):
        v = np.frombuffer(b, dtype="<f8", count=n, offset=o[0]); o[0] += 8 * n; return v.copy()

    def u64(n=1):
        v = np.frombuffer(b, dtype="<u8", count=n, offset=o[0]); o[0] += 8 * n; return v.copy()

    d = {}
    d["box_min"] = f64(3); d["box_max"] = f64(3); d["halo_width"] = f64(3); d["cell_length"] = f64(3)
    d["cells_per_dim"] = u64(3).astype(np.int64)
    d["cutoff"] = f64()[0]; d["epsilon_rf"] = f64()[0]; d["upot"] = f64()[0]; d["virial"] = f64()[0]
    nc = int(u64()[0]); Cn = int(u64()[0])
    d["eps24"] = f64(Cn * Cn).reshape(Cn, Cn); d["sig2"] = f64(Cn * Cn).reshape(Cn, Cn); d["shift6"] = f64(Cn * Cn).reshape(Cn, Cn)
    comp = []
    for _ in range(nc):
        m, i1, i2, i3 = f64(4); rest = u64(5)
        comp.append([m, i1, i2, i3, *rest.astype(np.float64)])
    d["comp_ref"] = np.array(comp)  # mass, I11,I22,I33, rotdof, nlj,nc,nd,nq
    n = int(u64()[0])
    rec = np.frombuffer(b, dtype=np.dtype([("id", "<u8"), ("cid", "<u8"), ("cell", "<u8"), ("halo", "<u8"), ("r", "<f8", 3), ("v", "<f8", 3),
                                           ("q", "<f8", 4), ("L", "<f8", 3), ("F", "<f8", 3), ("M", "<f8", 3), ("Vi", "<f8", 3)]), count=n, offset=o[0])
    for k in rec.dtype.names:
        d["mol_" + k] = rec[k].copy()
    return d


def run_refdump(prec, inp, cutoff, proc, mode, geom):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "dump.bin")
        subprocess.check_call([os.path.join(REFBIN, prec, "refdump"), inp, repr(cutoff), out, proc, mode, geom],
                              stdout=subprocess.DEVNULL)
        return read_refdump(out)


def save_case(name, ps: PhaseSpace, cutoff, mode, geom, inp):
    """One .npz: inputs (parsed by OUR reader) + reference outputs for mixed/double VCP and double Legacy."""
    out = dict(ps.ensemble.to_arrays())
    out.update(length=ps.length, temperature=ps.temperature, epsilon_rf=ps.epsilon_rf, cutoff=cutoff,
               in_id=ps.id, in_cid=ps.cid, in_r=ps.r, in_v=ps.v, in_q=ps.q, in_L=ps.L,
               mode=np.array(mode), geom=np.array(geom))
    for tag, prec, proc in (("mixed", "mixed", "vcp"), ("double", "double", "vcp"), ("legacy", "double", "legacy")):
        d = run_refdump(prec, inp, cutoff, proc, mode, geom)
        for k, v in d.items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print("wrote", name, "n =", len(ps.id), "ref records =", len(out["mixed_mol_id"]), "upot(double) =", out["double_upot"])


def forces():
    for name, rc in FORCE_FIXTURES:
        inp = os.path.join(REF, "test_input", name + ".inp")
        ps = read_phase_space(inp)
        save_case("force_" + name, ps, rc, "open", "ctor", inp)


def periodic():
    """Periodic-halo force dumps (DomainDecompBase::exchangeMolecules + traverseCells), XML geometry."""
    with tempfile.TemporaryDirectory() as td:
        # (a) reduced-unit 1CLJ liquid on a perturbed BCC lattice, rc = 2.5
        ps = generators.lj_lattice(n_per_dim=9, rho=0.7, temperature=0.85, rc=2.5, seed=1, jitter=0.12)
        inp = os.path.join(td, "lj.inp"); write_phase_space(inp, ps)
        save_case("periodic_1clj_1458", read_phase_space(inp), 2.5, "periodic", "xml", inp)
        # (b) the reference's water-like fixture component on a lattice (1 LJ + 3 charges), random orientations
        water = read_phase_space(os.path.join(REF, "test_input", "VectorizationWater.inp")).ensemble
        ps = generators.lattice(water, n_per_dim=6, rho=250.0 / 37.0 ** 3, temperature=9.437e-4, seed=2, jitter=0.1, rc=14.9)
        inp = os.path.join(td, "w.inp"); write_phase_space(inp, ps)
        save_case("periodic_water_432", read_phase_space(inp), 14.9, "periodic", "xml", inp)
        # (c) all potentials at once: the 2-component multi-potential fixture's components, periodic
        multi = read_phase_space(os.path.join(REF, "test_input", "VectorizationMultiComponentMultiPotentials_50_molecules.inp")).ensemble
        ps = generators.lattice(multi, n_per_dim=5, rho=250.0 / 60.0 ** 3, temperature=8.55e-4, seed=3, jitter=0.1, rc=12.0)
        inp = os.path.join(td, "m.inp"); write_phase_space(inp, ps)
        save_case("periodic_multi_250", read_phase_space(inp), 12.0, "periodic", "xml", inp)
        # (d) the reference's own periodic container fixture
        inp = os.path.join(REF, "test_input", "1clj-regular-12x12x12.inp")
        save_case("periodic_1clj_regular_12", read_phase_space(inp), 1.5, "periodic", "xml", inp)


def border():
    """Pins the cell-border arithmetic (CellBorderAndFlagManager::getBorder, one FMA in the reference binary): molecules
    placed exactly ON, one ulp below and one ulp above every inner border, in each dimension, for borders where
    fma(i, cellLength, haloMin) differs from i*cellLength + haloMin and for some where it does not."""
    import math
    from fractions import Fraction
    from ls1_mardyn_b200.components import lj_fluid
    L = 52.28; rc = 2.5
    nb = int(math.floor(L / float(np.float32(rc)))); clen = L / nb; hmin = 0.0 - clen
    vals = []
    for i in range(2, nb + 1):
        a = i * clen + hmin
        f = float(Fraction(i) * Fraction(clen) + Fraction(hmin))
        for b in {a, f}:
            vals += [b, np.nextafter(b, -1e300), np.nextafter(b, 1e300)]
    vals = np.array(sorted(set(vals)))
    r = []
    for d in range(3):
        for k, x in enumerate(vals):
            p = [0.0, 0.0, 0.0]
            p[d] = x
            p[(d + 1) % 3] = 1.3 + 1.7 * (k % 29)
            p[(d + 2) % 3] = 1.1 + 1.7 * (k // 29) + 6.1 * d
            r.append(p)
    r = np.array(r); n = len(r)
    rng = np.random.default_rng(5)
    ps = PhaseSpace(np.array([L, L, L]), 0.85, 1e10, lj_fluid(rc=rc), np.arange(1, n + 1, dtype=np.uint64), np.zeros(n, dtype=np.uint32),
                    r, rng.normal(size=(n, 3)), np.tile([1.0, 0, 0, 0], (n, 1)), np.zeros((n, 3)))
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "b.inp"); write_phase_space(inp, ps)
        save_case("periodic_border_fma", read_phase_space(inp), rc, "periodic", "xml", inp)


if __name__ == "__main__":
    what = sys.argv[1:] or ["forces", "periodic", "border", "traces"]
    if "forces" in what:
        forces()
    if "periodic" in what:
        periodic()
    if "border" in what:
        border()
    if "traces" in what:
        from make_traces import traces  # noqa: E402
        traces()
