"""TEST INFRASTRUCTURE ONLY -- the decomposed (multi-rank) CPU oracle.

One ``OracleRank`` per block of the 3D decomposition: the leaving-molecule / halo-copy selection, the periodic shift and
its clamps are restated here in numpy from the reference's MPI path

    parallel/CommunicationPartner.cpp:435-575   collectMoleculesInRegion: regions, shift, LEAVING / HALO position clamps
    parallel/DomainDecompBase.cpp:317-376       halo width == cutoff, half-open boundary layers [min, min+rc), [max-rc, max)
    parallel/DomainDecomposition.cpp:130-139    block bounds

and the per-block force path is the scalar oracle (oracle.cpp) on the block with its imported halo copies uploaded as
ordinary molecules lying in halo cells (which is what a halo molecule is to the cell processor,
VectorizedCellProcessor.cpp:2744-2819). It is pinned by tests/test_decomposition.py: the union of the blocks reproduces the
single-domain oracle (itself pinned against the unmodified reference). The API mirrors capi.Context in decomposed mode."""
from __future__ import annotations

import numpy as np

from ls1_mardyn_b200.capi import PARTICLE
from . import oracle as orc


def _next_f(x, up, mixed):
    if mixed:
        return float(np.nextafter(np.float32(x), np.float32(np.inf if up else -np.inf)))
    return float(np.nextafter(np.float64(x), np.inf if up else -np.inf))


class OracleRank:
    def __init__(self, dd, cutoff, ens, precision="mixed", epsilon_rf=1e10, geometry="xml", cutoff_lj=None):
        self.dd = dd
        self.bmin, self.bmax = dd.box_min(), dd.box_max()
        self.rc = float(cutoff)
        self.mixed = precision == "mixed"
        self.ens = ens
        self.sim = orc.OracleSim(list(self.bmin), list(self.bmax), cutoff, cutoff_lj=cutoff_lj, epsilon_rf=epsilon_rf, precision=precision,
                                 geometry=geometry, periodic=dd.periodic())
        self.sim.set_components(ens)
        L = dd.global_length
        # clamps against the GLOBAL box in the calc type (CommunicationPartner.cpp:533-562)
        self.wrap_hi = [_next_f(L[d], False, self.mixed) for d in range(3)]
        self.halo_hi = [_next_f(L[d], True, self.mixed) for d in range(3)]
        self.halo_lo = [_next_f(0.0, False, self.mixed) for d in range(3)]
        self.own = np.zeros(0, dtype=PARTICLE)      # molecules this rank owns
        self.ghost = np.zeros(0, dtype=PARTICLE)    # halo copies of this step's exchange
        self.n = 0
        self._uploaded = False

    # ---- state ----
    def upload(self, id, cid, r, v, q=None, L=None):
        n = len(id)
        a = np.zeros(n, dtype=PARTICLE)
        a["id"] = id; a["cid"] = cid; a["r"] = r; a["v"] = v
        a["q"] = q if q is not None else np.tile([1.0, 0, 0, 0], (n, 1))
        if L is not None:
            a["L"] = L
        self.own = a; self.ghost = np.zeros(0, dtype=PARTICLE); self.n = n
        self._uploaded = False

    def num_molecules(self):
        return len(self.own)

    # ---- one side of one stage (== k_export in halo.cu; both restate CommunicationPartner::collectMoleculesInRegion) ----
    def halo_export(self, dim, side):
        dd = self.dd
        glen = float(dd.global_length[dim])
        bmin, bmax, rc = float(self.bmin[dim]), float(self.bmax[dim]), self.rc
        wraps = dd.coords[dim] == 0 if side == 0 else dd.coords[dim] == dd.grid[dim] - 1
        out = []
        x = self.own["r"][:, dim].copy()
        leaving = x < bmin if side == 0 else x >= bmax
        if leaving.any():
            rec = self.own[leaving].copy()
            xn = x[leaving].copy()
            xh = xn.copy()
            if wraps:
                if side == 0:
                    xn = xn + glen; xn[xn >= glen] = self.wrap_hi[dim]            # LEAVING clamp
                    xh = xn - glen; xh[xh >= 0.0] = self.halo_lo[dim]             # the copy that stays behind: HALO clamp
                else:
                    xn = xn - glen; xn[xn < 0.0] = 0.0
                    xh = xn + glen; xh[xh < glen] = self.halo_hi[dim]
            rec["r"][:, dim] = xn; rec["flags"] = 0
            out.append(rec)
            stay = self.own[leaving].copy(); stay["r"][:, dim] = xh; stay["flags"] = 1
            stay["v"] = 0.0; stay["L"] = 0.0
            self.ghost = np.concatenate([self.ghost, stay]); self.own = self.own[~leaving]
            x = x[~leaving]
        # halo copies: owned molecules inside the boundary layer, and ghosts (forwarding to edge / corner neighbours)
        for arr, is_ghost in ((self.own, False), (self.ghost[: len(self.ghost) - (int(leaving.sum()) if leaving.any() else 0)], True)):
            if not len(arr):
                continue
            xa = arr["r"][:, dim]
            sel = xa < bmin + rc if side == 0 else xa >= bmax - rc
            if not is_ghost:
                sel &= (xa >= bmin) & (xa < bmax)
            if not sel.any():
                continue
            rec = arr[sel].copy(); xs = xa[sel].copy(); xn = xs.copy()
            if wraps:
                inside = (xs >= bmin) & (xs < bmax)
                if side == 0:
                    xn = xs + glen; fix = inside & (xn < glen); xn[fix] = self.halo_hi[dim]
                else:
                    xn = xs - glen; fix = inside & (xn >= 0.0); xn[fix] = self.halo_lo[dim]
            rec["r"][:, dim] = xn; rec["flags"] = 1
            out.append(rec)
        return np.concatenate(out) if out else np.zeros(0, dtype=PARTICLE)

    def halo_import(self, records):
        records = np.asarray(records, dtype=PARTICLE)
        if not len(records):
            return
        lv = records["flags"] == 0
        if lv.any():
            self.own = np.concatenate([self.own, records[lv]])
        if (~lv).any():
            g = records[~lv].copy(); g["v"] = 0.0; g["L"] = 0.0
            self.ghost = np.concatenate([self.ghost, g])
        self._uploaded = False

    # ---- phases ----
    def _push(self):
        a = np.concatenate([self.own, self.ghost])
        self.sim.upload(a["id"], a["cid"], a["r"], a["v"], a["q"], a["L"])
        self._uploaded = True

    def _pull(self):
        n = len(self.own)
        for f, k in (("R", "r"), ("V", "v"), ("Q", "q"), ("L", "L")):
            self.own[k] = self.sim.download(f)[:n]

    def update(self):
        self._push()
        self.sim.update()
        self._pull()
        self.n = len(self.own)

    def traverse_cells(self):
        return self.sim.traverse_cells()

    def leapfrog_pre(self, dt):
        assert self._uploaded
        self.sim.leapfrog_pre(dt)
        self._pull()
        self.ghost = np.zeros(0, dtype=PARTICLE)      # the halo of the previous step is gone (deleteOuterParticles)
        self._uploaded = False

    def leapfrog_post(self, dt):
        k = self.sim.leapfrog_post(dt)                # ghosts carry v = L = 0: they add nothing to the sums
        self._pull()
        k.n_molecules = len(self.own)
        rot = np.array([c.Ipa()[1] for c in self.ens.components], dtype=np.int64)
        k.rot_dof = int(rot[self.own["cid"]].sum())
        return k

    def scale_velocities(self, bt, br):
        self.sim.scale_velocities(bt, br)
        self._pull()

    def download(self, field):
        n = len(self.own)
        if field in ("ID", "CID", "R", "V", "Q", "L"):
            return self.own[{"ID": "id", "CID": "cid", "R": "r", "V": "v", "Q": "q", "L": "L"}[field]].copy()
        return self.sim.download(field)[:n]
