"""Phase-space ``.inp`` reader/writer -- host-side mirror of the reference's ASCIIReader
(reference: src/io/ASCIIReader.cpp:53-260 header, :289-420 molecules; format summary in SURVEY.md A.6).
Only the tokens the hot path needs are interpreted."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .components import Charge, Component, Dipole, Ensemble, LJcenter, Quadrupole


@dataclass
class PhaseSpace:
    length: np.ndarray          # box lengths, box_min is 0 (ASCIIReader "Length")
    temperature: float
    epsilon_rf: float
    ensemble: Ensemble
    id: np.ndarray              # uint64 [n]
    cid: np.ndarray             # uint32 [n], 0-based
    r: np.ndarray               # [n,3]
    v: np.ndarray               # [n,3]
    q: np.ndarray               # [n,4]  (w,x,y,z)
    L: np.ndarray               # [n,3]  angular momentum D


def _tokens(path):
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0]
            for t in line.split():
                yield t


def read_phase_space(path: str) -> PhaseSpace:
    it = _tokens(path)
    nxt = lambda: next(it)
    fl = lambda: float(nxt())
    length = None
    temperature = 0.0
    eps_rf = 1e10
    ens = None
    assert nxt() in ("mardyn", "MDProject")
    nxt(); nxt()  # "trunk" <version>
    fmt = "ICRVQD"
    n = 0
    while True:
        t = nxt()
        if t in ("currentTime", "t"):
            fl()
        elif t in ("Temperature", "T"):
            temperature = fl()
        elif t in ("Length", "L"):
            length = np.array([fl(), fl(), fl()])
        elif t in ("NumberOfComponents", "C"):
            nc = int(nxt())
            comps = []
            for _ in range(nc):
                nl, nch, nd, nq, nt = (int(nxt()) for _ in range(5))
                assert nt == 0, "tersoff no longer supported"
                c = Component()
                for _ in range(nl):
                    x, y, z, m, eps, sig, rc, sh = (fl() for _ in range(8))
                    c.ljcenters.append(LJcenter((x, y, z), m, eps, sig, rc, sh != 0))
                for _ in range(nch):
                    x, y, z, m, q = (fl() for _ in range(5))
                    c.charges.append(Charge((x, y, z), m, q))
                for _ in range(nd):
                    x, y, z, ex, ey, ez, p = (fl() for _ in range(7))
                    c.dipoles.append(Dipole((x, y, z), (ex, ey, ez), p))
                for _ in range(nq):
                    x, y, z, ex, ey, ez, p = (fl() for _ in range(7))
                    c.quadrupoles.append(Quadrupole((x, y, z), (ex, ey, ez), p))
                c.I_override = (fl(), fl(), fl())
                comps.append(c)
            xi = np.ones((nc, nc)); eta = np.ones((nc, nc))
            for i in range(nc - 1):
                for j in range(i + 1, nc):
                    xi[i, j] = fl(); eta[i, j] = fl()
            eps_rf = fl()
            ens = Ensemble(comps, xi, eta)
        elif t in ("NumberOfMolecules", "N"):
            n = int(nxt())
            break
        else:
            raise ValueError(f"unknown token {t!r} in {path}")
    rest = list(it)
    if rest and rest[0] in ("MoleculeFormat", "M"):
        fmt = rest[1]
        rest = rest[2:]
    ncol = {"ICRVQD": 15, "ICRV": 8, "IRV": 7, "ICRVQDV": 18}[fmt]
    a = np.array(rest[: n * ncol], dtype=np.float64).reshape(n, ncol)
    ids = np.array([int(x) for x in np.array(rest[: n * ncol]).reshape(n, ncol)[:, 0]], dtype=np.uint64)
    if fmt == "IRV":
        cid = np.zeros(n, dtype=np.uint32); off = 1
    else:
        cid = (a[:, 1].astype(np.int64) - 1).astype(np.uint32); off = 2
    r = a[:, off:off + 3].copy(); v = a[:, off + 3:off + 6].copy()
    q = np.zeros((n, 4)); q[:, 0] = 1.0
    L = np.zeros((n, 3))
    if fmt in ("ICRVQD", "ICRVQDV"):
        q = a[:, off + 6:off + 10].copy(); L = a[:, off + 10:off + 13].copy()
    return PhaseSpace(length, temperature, eps_rf, ens, ids, cid, r, v, q, L)


def write_phase_space(path: str, ps: PhaseSpace) -> None:
    """Writes an ICRVQD file the reference's ASCIIReader accepts (used to feed synthetic systems to the
    reference build; FullMolecule::write, FullMolecule.cpp:439-446)."""
    with open(path, "w") as f:
        f.write("mardyn trunk 20120726\n currentTime\t0\n")
        f.write(" Length\t%.17g %.17g %.17g\n" % tuple(ps.length))
        f.write(" Temperature\t%.17g\n" % ps.temperature)
        comps = ps.ensemble.components
        f.write(" NumberOfComponents\t%d\n" % len(comps))
        for c in comps:
            f.write("%d\t%d\t%d\t%d\t0\n" % (len(c.ljcenters), len(c.charges), len(c.dipoles), len(c.quadrupoles)))
            for s in c.ljcenters:
                f.write("%.17g %.17g %.17g\t%.17g %.17g %.17g %.17g %d\n" % (*s.r, s.m, s.eps, s.sigma, s.rc, int(s.do_shift)))
            for s in c.charges:
                f.write("%.17g %.17g %.17g\t%.17g %.17g\n" % (*s.r, s.m, s.q))
            for s in c.dipoles:
                f.write("%.17g %.17g %.17g\t%.17g %.17g %.17g %.17g\n" % (*s.r, *s.e, s.abs_my))
            for s in c.quadrupoles:
                f.write("%.17g %.17g %.17g\t%.17g %.17g %.17g %.17g\n" % (*s.r, *s.e, s.abs_q))
            f.write("%.17g %.17g %.17g\n" % tuple(c.I_override))
        nc = len(comps)
        for i in range(nc - 1):
            f.write(" ".join("%.17g %.17g" % (ps.ensemble.xi[i, j], ps.ensemble.eta[i, j]) for j in range(i + 1, nc)) + "\n")
        f.write("%.17g\n" % ps.epsilon_rf)
        f.write(" NumberOfMolecules\t%d\n MoleculeFormat\tICRVQD\n" % len(ps.id))
        for i in range(len(ps.id)):
            f.write("%d\t%d\t%.17g %.17g %.17g\t%.17g %.17g %.17g\t%.17g %.17g %.17g %.17g\t%.17g %.17g %.17g\n" % (
                int(ps.id[i]), int(ps.cid[i]) + 1, *ps.r[i], *ps.v[i], *ps.q[i], *ps.L[i]))


def read_molecules_only(path: str):
    """(id, cid, r, v, q, L) of a phase-space file WITHOUT interpreting its header: what ASCIIReader::readPhaseSpace does in XML
    mode (io/ASCIIReader.cpp:303-318 skips to NumberOfMolecules; the components come from the XML). Reads the checkpoints
    Domain::writeCheckpoint / FullMolecule::write produce (molecules/FullMolecule.cpp:439-446: id cid r v q D, cid 1-based)."""
    toks = open(path).read().split()
    k = max(i for i, t in enumerate(toks) if t in ("NumberOfMolecules", "N"))
    n = int(toks[k + 1]); k += 2
    fmt = "ICRVQD"
    if toks[k] in ("MoleculeFormat", "M"):
        fmt = toks[k + 1]; k += 2
    w = {"ICRVQD": 15, "ICRV": 8, "IRV": 7}[fmt]
    a = np.array(toks[k:k + n * w], dtype=np.float64).reshape(n, w)
    ids = np.array([int(t) for t in toks[k:k + n * w:w]], dtype=np.uint64)
    if fmt == "IRV":
        return ids, np.zeros(n, dtype=np.uint32), a[:, 1:4], a[:, 4:7], np.tile([1.0, 0, 0, 0], (n, 1)), np.zeros((n, 3))
    cid = (a[:, 1] - 1).astype(np.uint32)
    q = a[:, 8:12] if fmt == "ICRVQD" else np.tile([1.0, 0, 0, 0], (n, 1))
    Lm = a[:, 12:15] if fmt == "ICRVQD" else np.zeros((n, 3))
    return ids, cid, a[:, 2:5], a[:, 5:8], q, Lm
