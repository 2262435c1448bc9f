"""ctypes front end of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The API mirrors ls1_mardyn_b200.capi.Context method for method so a parity test drives both identically."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ls1_mardyn_b200.components import CComponent, CSite, Ensemble

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FIELDS = dict(ID=0, CID=1, R=2, V=3, Q=4, L=5, F=6, M=7, VI=8, CELL=9)
_FIELD_SHAPE = {0: (np.uint64, ()), 1: (np.uint32, ()), 2: (np.float64, (3,)), 3: (np.float64, (3,)), 4: (np.float64, (4,)),
                5: (np.float64, (3,)), 6: (np.float64, (3,)), 7: (np.float64, (3,)), 8: (np.float64, (3,)), 9: (np.uint64, ())}


class Config(C.Structure):
    _fields_ = [("box_min", C.c_double * 3), ("box_max", C.c_double * 3), ("cutoff", C.c_double), ("cutoff_lj", C.c_double),
                ("epsilon_rf", C.c_double), ("precision", C.c_int32), ("geometry", C.c_int32), ("periodic", C.c_int32 * 3),
                ("device", C.c_int32), ("reserved", C.c_int32 * 4)]


class Macro(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upot", "virial", "upot6lj", "upot_xpoles", "virial_raw", "my_rf")]


class Kinetic(C.Structure):
    _fields_ = [("summv2", C.c_double), ("sum_iw2", C.c_double), ("n_molecules", C.c_uint64), ("rot_dof", C.c_uint64)]


class StepRecord(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("upot", "virial", "summv2", "sum_iw2", "temperature", "pressure", "beta_trans", "beta_rot")]


class FlopCount(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "mol_dist", "lj", "charge", "charge_dipole", "dipole", "charge_quadrupole", "dipole_quadrupole", "quadrupole",
        "lj_macro", "charge_macro", "charge_dipole_macro", "dipole_macro", "charge_quadrupole_macro",
        "dipole_quadrupole_macro", "quadrupole_macro", "flops")]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_num_molecules.restype = C.c_uint64
        _LIB.orc_halo.restype = C.c_uint64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class OracleSim:
    def __init__(self, box_min, box_max, cutoff, cutoff_lj=None, epsilon_rf=1e10, precision="mixed", geometry="xml",
                 periodic=(1, 1, 1)):
        cfg = Config()
        for d in range(3):
            cfg.box_min[d] = box_min[d]; cfg.box_max[d] = box_max[d]; cfg.periodic[d] = int(periodic[d])
        cfg.cutoff = cutoff; cfg.cutoff_lj = cutoff if cutoff_lj is None else cutoff_lj; cfg.epsilon_rf = epsilon_rf
        cfg.precision = {"mixed": 0, "double": 1}[precision]; cfg.geometry = {"xml": 0, "ctor": 1}[geometry]
        self.cfg = cfg
        self.h = C.c_void_p(lib().orc_create(C.byref(cfg)))
        self.n = 0

    def close(self):
        if self.h:
            lib().orc_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_components(self, ens: Ensemble, tables=None):
        cc, cs, ns = ens.pack()
        e24, s2, sh6 = tables if tables is not None else ens.lj_tables(self.cfg.cutoff_lj)
        e24 = np.ascontiguousarray(e24, dtype=np.float64); s2 = np.ascontiguousarray(s2, dtype=np.float64); sh6 = np.ascontiguousarray(sh6, dtype=np.float64)
        lib().orc_set_components(self.h, len(ens.components), cc, ns, cs, e24.shape[0], _p(e24, C.c_double), _p(s2, C.c_double), _p(sh6, C.c_double))

    def upload(self, id, cid, r, v, q=None, L=None):
        n = len(id)
        self.n = n
        id = np.ascontiguousarray(id, dtype=np.uint64); cid = np.ascontiguousarray(cid, dtype=np.uint32)
        r = np.ascontiguousarray(r, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
        q = None if q is None else np.ascontiguousarray(q, dtype=np.float64)
        L = None if L is None else np.ascontiguousarray(L, dtype=np.float64)
        lib().orc_upload(self.h, C.c_uint64(n), _p(id, C.c_uint64), _p(cid, C.c_uint32), _p(r, C.c_double), _p(v, C.c_double), _p(q, C.c_double), _p(L, C.c_double))

    def geometry(self):
        cpd = (C.c_int32 * 3)(); cl = (C.c_double * 3)(); hm = (C.c_double * 3)()
        lib().orc_geometry(self.h, cpd, cl, hm)
        return np.array(cpd[:]), np.array(cl[:]), np.array(hm[:])

    def update(self):
        lib().orc_update(self.h)

    def traverse_cells(self):
        m = Macro(); lib().orc_traverse_cells(self.h, C.byref(m)); return m

    def leapfrog_pre(self, dt):
        lib().orc_leapfrog_pre(self.h, C.c_double(dt))

    def leapfrog_post(self, dt):
        k = Kinetic(); lib().orc_leapfrog_post(self.h, C.c_double(dt), C.byref(k)); return k

    def scale_velocities(self, bt, br):
        lib().orc_scale_velocities(self.h, C.c_double(bt), C.c_double(br))

    def download(self, field):
        f = FIELDS[field] if isinstance(field, str) else field
        dt, shp = _FIELD_SHAPE[f]
        out = np.zeros((self.n,) + shp, dtype=dt)
        rc = lib().orc_download(self.h, f, out.ctypes.data_as(C.c_void_p), C.c_uint64(self.n))
        assert rc == 0
        return out

    def halo(self):
        nh = lib().orc_halo(self.h, None, None, None)
        owner = np.zeros(nh, dtype=np.int64); r = np.zeros((nh, 3)); cell = np.zeros(nh, dtype=np.uint64)
        lib().orc_halo(self.h, _p(owner, C.c_int64), _p(r, C.c_double), _p(cell, C.c_uint64))
        return owner, r, cell

    def count_flops(self):
        f = FlopCount(); lib().orc_count_flops(self.h, C.byref(f)); return f

    def simulate(self, nsteps, dt, target_temperature, global_n=0, global_volume=0.0):
        rec = (StepRecord * nsteps)()
        lib().orc_simulate(self.h, C.c_uint64(nsteps), C.c_double(dt), C.c_double(target_temperature), C.c_uint64(global_n), C.c_double(global_volume), rec)
        return np.array([[getattr(x, k) for k, _ in StepRecord._fields_] for x in rec])


def lj_tables(ens: Ensemble, rc_lj: float):
    """Comp2Param restatement in C++ (independent of Ensemble.lj_tables)."""
    nc = len(ens.components)
    nlj = np.array([len(c.ljcenters) for c in ens.components], dtype=np.uint32)
    eps = np.array([s.eps for c in ens.components for s in c.ljcenters], dtype=np.float64)
    sig = np.array([s.sigma for c in ens.components for s in c.ljcenters], dtype=np.float64)
    # the shift of a centre is computed with ITS OWN cutoff column (Component::addLJcenter, Component.cpp:104-116)
    rcs = np.array([s.rc for c in ens.components for s in c.ljcenters], dtype=np.float64)
    dosh = np.array([int(s.do_shift) for c in ens.components for s in c.ljcenters], dtype=np.int32)
    Cn = int(nlj.sum())
    e24 = np.zeros((Cn, Cn)); s2 = np.zeros((Cn, Cn)); sh6 = np.zeros((Cn, Cn))
    xi = np.ascontiguousarray(ens.xi, dtype=np.float64); eta = np.ascontiguousarray(ens.eta, dtype=np.float64)
    lib().orc_lj_tables(nc, _p(nlj, C.c_uint32), _p(eps, C.c_double), _p(sig, C.c_double), _p(dosh, C.c_int32), _p(rcs, C.c_double),
                        _p(xi, C.c_double), _p(eta, C.c_double), C.c_double(rc_lj), _p(e24, C.c_double), _p(s2, C.c_double), _p(sh6, C.c_double))
    return e24, s2, sh6
