"""Molecule-type definitions and pair-parameter tables -- host-side mirror of ls1-MarDyn's

  * ``Component`` / ``Site``      (reference: src/molecules/Component.{h,cpp}, src/molecules/Site.h)
  * ``Comp2Param`` (LJ part)      (reference: src/molecules/Comp2Param.cpp:11-111)
  * the LJ table layout of ``VectorizedCellProcessor`` (reference: adapter/VectorizedCellProcessor.cpp:39-81)

and their packing into the plain-C structs of ``include/ls1b200.h``. Names follow the reference.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np


# ---- C structs of include/ls1b200.h -------------------------------------------------------------------------
class CComponent(C.Structure):
    _fields_ = [("n_lj", C.c_uint32), ("n_charge", C.c_uint32), ("n_dipole", C.c_uint32), ("n_quadrupole", C.c_uint32),
                ("site_offset", C.c_uint32), ("lj_lookup_id", C.c_uint32), ("rot_dof", C.c_uint32), ("reserved", C.c_uint32),
                ("mass", C.c_double), ("inertia", C.c_double * 3)]


class CSite(C.Structure):
    _fields_ = [("r", C.c_double * 3), ("e", C.c_double * 3), ("param", C.c_double)]


@dataclass
class LJcenter:
    """Site.h: LJcenter(x, y, z, m, eps, sigma, shift6)."""
    r: Sequence[float]
    m: float
    eps: float
    sigma: float
    rc: float = 0.0           # cutoff used for the shift (ASCIIReader column 7)
    do_shift: bool = False


@dataclass
class Charge:
    r: Sequence[float]
    m: float
    q: float


@dataclass
class Dipole:
    r: Sequence[float]
    e: Sequence[float]
    abs_my: float


@dataclass
class Quadrupole:
    r: Sequence[float]
    e: Sequence[float]
    abs_q: float


@dataclass
class Component:
    """One molecule type. Mass / inertia follow Component::updateMassInertia (Component.cpp:151-178): principal-axes
    inertia from the massive sites (LJ centres and charges), rot_dof = number of non-zero principal moments;
    explicitly given moments > 0 override (ASCIIReader.cpp:214-218, Component.cpp:91-95)."""
    ljcenters: List[LJcenter] = field(default_factory=list)
    charges: List[Charge] = field(default_factory=list)
    dipoles: List[Dipole] = field(default_factory=list)
    quadrupoles: List[Quadrupole] = field(default_factory=list)
    I_override: Sequence[float] = (0.0, 0.0, 0.0)

    def m(self) -> float:
        return sum(s.m for s in self.ljcenters) + sum(s.m for s in self.charges)

    def Ipa(self):
        I = [0.0, 0.0, 0.0]
        for s in list(self.ljcenters) + list(self.charges):
            x, y, z = s.r
            I[0] += s.m * (y * y + z * z)
            I[1] += s.m * (x * x + z * z)
            I[2] += s.m * (x * x + y * y)
        rot = sum(1 for d in range(3) if I[d] != 0.0) if (self.ljcenters or self.charges) else 0
        for d in range(3):
            if self.I_override[d] > 0.0:
                I[d] = float(self.I_override[d])   # setI11..33 do not touch _rot_dof
        return I, rot

    def num_sites(self) -> int:
        return len(self.ljcenters) + len(self.charges) + len(self.dipoles) + len(self.quadrupoles)


@dataclass
class Ensemble:
    """Component list + Lorentz-Berthelot mixing coefficients (xi, eta) per component pair i<j
    (ASCIIReader.cpp:229-245; molecules/mixingrules/LorentzBerthelot.h)."""
    components: List[Component]
    xi: np.ndarray = None      # [nc, nc], used for i<j
    eta: np.ndarray = None

    def __post_init__(self):
        nc = len(self.components)
        if self.xi is None:
            self.xi = np.ones((nc, nc))
        if self.eta is None:
            self.eta = np.ones((nc, nc))

    def num_lj_centers(self) -> int:
        return sum(len(c.ljcenters) for c in self.components)

    def lj_tables(self, cutoff_lj: float):
        """eps24[i,j] = 24 eps_ij, sig2[i,j] = sigma_ij^2, shift6[i,j] = 6 * shift, i/j over all LJ centres of all
        components (Comp2Param.cpp:19-111 streamed into VectorizedCellProcessor.cpp:60-81).
        Same component: eps = sqrt(eps_i eps_j), sigma = (s_i+s_j)/2, shift6 of centre i (only if requested).
        Different components: eps = xi sqrt(..), sigma = eta (s_i+s_j)/2, shift ALWAYS from rcLJ."""
        comps = self.components
        start, Cn = [], 0
        for c in comps:
            start.append(Cn)
            Cn += len(c.ljcenters)
        eps24 = np.zeros((Cn, Cn))
        sig2 = np.zeros((Cn, Cn))
        shift6 = np.zeros((Cn, Cn))
        for ci, c in enumerate(comps):
            for a, la in enumerate(c.ljcenters):
                sh_i = 0.0
                if la.do_shift:
                    s2 = la.sigma * la.sigma / (la.rc * la.rc)
                    s6 = s2 * s2 * s2
                    sh_i = 24.0 * la.eps * (s6 - s6 * s6)
                for b, lb in enumerate(c.ljcenters):
                    ia, ib = start[ci] + a, start[ci] + b
                    eps24[ia, ib] = 24.0 * math.sqrt(la.eps * lb.eps)
                    s = 0.5 * (la.sigma + lb.sigma)
                    sig2[ia, ib] = s * s
                    shift6[ia, ib] = sh_i
            for cj in range(ci + 1, len(comps)):
                xi, eta = float(self.xi[ci, cj]), float(self.eta[ci, cj])
                for a, la in enumerate(c.ljcenters):
                    for b, lb in enumerate(comps[cj].ljcenters):
                        ia, jb = start[ci] + a, start[cj] + b
                        e24 = 24.0 * (xi * math.sqrt(la.eps * lb.eps))
                        s = 0.5 * (eta * (la.sigma + lb.sigma))
                        s2 = s * s
                        sp2 = s2 / (cutoff_lj * cutoff_lj)
                        sp6 = sp2 * sp2 * sp2
                        sh = e24 * (sp6 - sp6 * sp6)
                        for (p, q_) in ((ia, jb), (jb, ia)):
                            eps24[p, q_] = e24
                            sig2[p, q_] = s2
                            shift6[p, q_] = sh
        return eps24, sig2, shift6

    def pack(self):
        """-> (CComponent array, CSite array) in the layout of include/ls1b200.h."""
        comps = self.components
        ns = sum(c.num_sites() for c in comps)
        cc = (CComponent * len(comps))()
        cs = (CSite * max(ns, 1))()
        so, lj = 0, 0
        for i, c in enumerate(comps):
            I, rot = c.Ipa()
            cc[i].n_lj, cc[i].n_charge = len(c.ljcenters), len(c.charges)
            cc[i].n_dipole, cc[i].n_quadrupole = len(c.dipoles), len(c.quadrupoles)
            cc[i].site_offset, cc[i].lj_lookup_id, cc[i].rot_dof = so, lj, rot
            cc[i].mass = c.m()
            for d in range(3):
                cc[i].inertia[d] = I[d]
            for s in c.ljcenters:
                for d in range(3):
                    cs[so].r[d] = s.r[d]
                so += 1
            for s in c.charges:
                for d in range(3):
                    cs[so].r[d] = s.r[d]
                cs[so].param = s.q
                so += 1
            for s in c.dipoles:
                for d in range(3):
                    cs[so].r[d] = s.r[d]
                    cs[so].e[d] = s.e[d]
                cs[so].param = s.abs_my
                so += 1
            for s in c.quadrupoles:
                for d in range(3):
                    cs[so].r[d] = s.r[d]
                    cs[so].e[d] = s.e[d]
                cs[so].param = s.abs_q
                so += 1
            lj += len(c.ljcenters)
        return cc, cs, ns

    # ---- (de)serialisation for tests/golden/*.npz ---------------------------------------------------------
    def to_arrays(self) -> dict:
        counts, lj, ch, dp, qp, Io = [], [], [], [], [], []
        for c in self.components:
            counts.append([len(c.ljcenters), len(c.charges), len(c.dipoles), len(c.quadrupoles)])
            for s in c.ljcenters:
                lj.append([*s.r, s.m, s.eps, s.sigma, s.rc, float(s.do_shift)])
            for s in c.charges:
                ch.append([*s.r, s.m, s.q])
            for s in c.dipoles:
                dp.append([*s.r, *s.e, s.abs_my])
            for s in c.quadrupoles:
                qp.append([*s.r, *s.e, s.abs_q])
            Io.append(list(c.I_override))
        return dict(comp_counts=np.array(counts, dtype=np.int64).reshape(-1, 4),
                    comp_lj=np.array(lj, dtype=np.float64).reshape(-1, 8),
                    comp_charge=np.array(ch, dtype=np.float64).reshape(-1, 5),
                    comp_dipole=np.array(dp, dtype=np.float64).reshape(-1, 7),
                    comp_quadrupole=np.array(qp, dtype=np.float64).reshape(-1, 7),
                    comp_I=np.array(Io, dtype=np.float64).reshape(-1, 3),
                    mix_xi=np.asarray(self.xi, dtype=np.float64), mix_eta=np.asarray(self.eta, dtype=np.float64))

    @staticmethod
    def from_arrays(a) -> "Ensemble":
        comps = []
        il = ic = idp = iq = 0
        for k, (nl, nc_, nd, nq) in enumerate(a["comp_counts"]):
            c = Component(I_override=tuple(a["comp_I"][k]))
            for _ in range(nl):
                row = a["comp_lj"][il]; il += 1
                c.ljcenters.append(LJcenter(tuple(row[0:3]), row[3], row[4], row[5], row[6], bool(row[7])))
            for _ in range(nc_):
                row = a["comp_charge"][ic]; ic += 1
                c.charges.append(Charge(tuple(row[0:3]), row[3], row[4]))
            for _ in range(nd):
                row = a["comp_dipole"][idp]; idp += 1
                c.dipoles.append(Dipole(tuple(row[0:3]), tuple(row[3:6]), row[6]))
            for _ in range(nq):
                row = a["comp_quadrupole"][iq]; iq += 1
                c.quadrupoles.append(Quadrupole(tuple(row[0:3]), tuple(row[3:6]), row[6]))
            comps.append(c)
        return Ensemble(comps, np.array(a["mix_xi"]), np.array(a["mix_eta"]))


def lj_fluid(eps: float = 1.0, sigma: float = 1.0, mass: float = 1.0, rc: float = 2.5, shifted: bool = False) -> Ensemble:
    """The 1CLJ model of BASELINE.json configs 1, 2 and 5 (one LJ centre at the origin)."""
    return Ensemble([Component(ljcenters=[LJcenter((0.0, 0.0, 0.0), mass, eps, sigma, rc, shifted)])])
