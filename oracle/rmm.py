"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of ls1-MarDyn's reduced-memory mode (RMM) for the 1CLJ fluid:
the all-single-precision path of SURVEY.md 8f-2, the next row to be built on the GPU. Nothing in ls1_mardyn_b200/ imports it.

Parity status: PINNED by tests/test_oracle_rmm.py against tests/golden/rmm_1clj_2000.npz, the per-step ResultWriter output of the
reference built in RMM configuration (oracle/build_ref.sh rmm -> oracle/_ref/rmm/MarDyn; the reference compiles there only
from a shadow tree with one `constexpr` removed in two headers, see build_ref.sh), made by tests/golden/make_golden_rmm.py.

What is restated (file:line relative to /root/reference/src):
  * state: position and velocity in FP32, id (molecules/MoleculeRMM.h; CellDataSoARMM.h) -- 32 B per molecule;
  * step (Simulation::simulateOneTimestep, Simulation.cpp:1076-1300, with the RMM integrator):
      1. LeapfrogRMM::computePositions -> PositionCellProcessorRMM: r = fma(dt, v, r) in FP32 (integrators/PositionCellProcessorRMM.h:43-58);
      2. LinkedCells::update + DomainDecompBase::exchangeMolecules: periodic wrap and halo copies with the FP32 clamps
         (parallel/DomainDecompBase.cpp:199-249, :317-376), coordinates stored back as FP32;
      3. VCP1CLJRMM (particleContainer/adapter/VCP1CLJRMM.cpp:195-236, :238-369): every pair inside the cutoff once, FP32 throughout,
         scale = 24 eps / r^2 * (2 lj12 - lj6), and the force goes STRAIGHT into the velocities: v_i += f dt/m, v_j -= f dt/m;
         6 U += 24 eps (lj12 - lj6) + shift6, virial += d . f, both FP32 sums; a pair with a halo copy counts for the macroscopic
         sums iff cell1.index < cell2.index (:137-150), i.e. exactly one of the two mirror images;
      4. LeapfrogRMM::computeVelocities -> VelocityCellProcessorRMM: sum v^2 (NOT mass weighted, :119-127);
      5. Domain::calculateGlobalValues (Domain.cpp:157-252): U = U6/6 + UpotCorr, T = sum v^2 / 3N, beta = (3 N T_target / sum v^2)^0.4,
         p = rho T + rho (virial + VirialCorr) / 3N, with the homogeneous long-range correction (longRange/Homogeneous.cpp:95-99, :128-136);
      6. VelocityScalingThermostat::apply: v *= beta.
    The force of prepare_start (Simulation.cpp:902-938) is applied with dt/2 (no thermostat there): together with step 3 this is the
    leapfrog scheme.
  Pairs are found by brute force over owned molecules and explicit halo copies (the cell structure only selects pairs, it does not
  change them: every pair inside the cutoff lies in neighbouring cells because a cell is at least one cutoff wide); FP32 sums are
  taken in numpy's order, not in the reference's SIMD-lane order, hence the 1e-5 tolerance of the test."""
from __future__ import annotations

import numpy as np

F32 = np.float32


class OracleRMM:
    def __init__(self, length, cutoff, dt, target_temperature, eps=1.0, sigma=1.0, mass=1.0, shifted=False):
        self.L = np.asarray(length, dtype=np.float64)
        self.rc = float(cutoff)
        self.dt = float(dt)
        self.T_target = float(target_temperature)
        # Comp2Param for one component (molecules/Comp2Param.cpp:11-111): eps24, sig2, shift6 -- cast to FP32 (VCP1CLJRMM.cpp:35-40)
        self.eps24 = F32(24.0 * eps)
        self.sig2 = F32(sigma * sigma)
        s6 = (sigma / cutoff) ** 6
        self.shift6 = F32(-24.0 * eps * (s6 * s6 - s6)) if shifted else F32(0.0)      # Component.cpp:118-122
        self.mass = float(mass)
        self.rc2 = F32(cutoff * cutoff)
        self.r = np.zeros((0, 3), dtype=F32)
        self.v = np.zeros((0, 3), dtype=F32)
        self.id = np.zeros(0, dtype=np.uint64)
        self.eps_d, self.sig2_d, self.shifted = float(eps), float(sigma * sigma), bool(shifted)

    # ---- state ----
    def upload(self, id, r, v):
        self.id = np.asarray(id, dtype=np.uint64).copy()
        self.r = np.asarray(r, dtype=np.float64).astype(F32)
        self.v = np.asarray(v, dtype=np.float64).astype(F32)

    @property
    def n(self):
        return len(self.id)

    # ---- long-range correction of the truncated (unshifted) potential: Homogeneous.cpp:95-99, :128-136, :158-181 ----
    def _corrections(self):
        if self.shifted:
            return 0.0, 0.0
        N = float(self.n); rho = N / float(np.prod(self.L)); rc = self.rc

        def ticcu(n):
            return -rc ** (2 * n + 3) / (self.sig2_d ** n * (2 * n + 3))

        def ticcv(n):
            return 2 * n * ticcu(n)
        fac0 = N * N * 24.0 * self.eps_d
        u = fac0 * (ticcu(-6) - ticcu(-3)); w = fac0 * (ticcv(-6) - ticcv(-3))
        fac = np.pi * rho / (3.0 * N)
        return fac * u, -fac * w

    # ---- steps 2: wrap + halo copies ----
    def _wrap(self):
        r = self.r
        for d in range(3):
            L = self.L[d]; x = r[:, d].astype(np.float64)
            lo = x < 0.0
            hi = x >= L
            xl = (x + L).astype(F32)                       # setr stores FP32
            xl = np.where(xl.astype(np.float64) >= L, np.nextafter(F32(L), F32(-np.inf)), xl)      # DomainDecompBase.cpp:236-240
            xh = (x - L).astype(F32)
            xh = np.where(xh.astype(np.float64) < 0.0, F32(0.0), xh)                              # :241-245 (r <= bmin -> bmin)
            r[:, d] = np.where(lo, xl, np.where(hi, xh, r[:, d]))

    def _halo(self):
        """Owned molecules followed by their periodic copies (x, then y incl. x copies, then z): DomainDecompBase.cpp:317-376."""
        pos = self.r.copy(); src = np.arange(self.n)
        for d in range(3):
            L = self.L[d]; x = pos[:, d].astype(np.float64)
            low = (x >= 0.0) & (x < self.rc)               # copy goes above the box
            high = (x >= L - self.rc) & (x < L)            # copy goes below the box
            up = pos[low].copy(); xu = (x[low] + L).astype(F32)
            xu = np.where(xu.astype(np.float64) < L, np.nextafter(F32(L), F32(np.inf)), xu)        # :364-368
            up[:, d] = xu
            dn = pos[high].copy(); xd = (x[high] - L).astype(F32)
            xd = np.where(xd.astype(np.float64) >= 0.0, np.nextafter(F32(0.0), F32(-np.inf)), xd)  # :358-362
            dn[:, d] = xd
            pos = np.concatenate([pos, up, dn]); src = np.concatenate([src, src[low], src[high]])
        return pos, src

    # ---- step 3: forces straight into the velocities ----
    def _forces(self, dt):
        n = self.n
        pos, src = self._halo()
        dt_inv_m = F32(dt / self.mass)
        upot6 = F32(0.0); vir = F32(0.0)
        dv = np.zeros((n, 3), dtype=F32)
        block = 512
        for i0 in range(0, n, block):
            ri = pos[i0:min(i0 + block, n)]
            d = ri[:, None, :] - pos[None, :, :]                                   # FP32
            r2 = d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]
            ii = np.arange(i0, i0 + len(ri))[:, None]; jj = np.arange(len(pos))[None, :]
            mask = (r2 < self.rc2) & (jj != ii)
            r2s = np.where(mask, r2, F32(1.0))
            r2inv = F32(1.0) / r2s
            lj2 = self.sig2 * r2inv; lj4 = lj2 * lj2; lj6 = lj4 * lj2; lj12 = lj6 * lj6; lj12m6 = lj12 - lj6
            scale = np.where(mask, (self.eps24 * r2inv) * (lj12 + lj12m6), F32(0.0))
            f = d * scale[..., None]                                                # force on i
            dv[i0:i0 + len(ri)] += f.sum(axis=1, dtype=F32)
            # macroscopic sums: owned-owned pairs once (j > i), owned-halo pairs with weight 1/2 (one of the two mirror images)
            owned_j = jj < n
            wgt = np.where(owned_j, np.where(jj > ii, F32(1.0), F32(0.0)), F32(0.5)) * mask
            upot6 += ((self.eps24 * lj12m6 + self.shift6) * wgt).sum(dtype=F32)
            vir += ((d[..., 0] * f[..., 0] + d[..., 1] * f[..., 1] + d[..., 2] * f[..., 2]) * wgt).sum(dtype=F32)
        self.v = self.v + dv * dt_inv_m
        return float(upot6), float(vir)

    # ---- the loop ----
    def _globals(self, upot6, vir):
        N = self.n
        uc, wc = self._corrections()
        summv2 = float((self.v.astype(np.float32) ** 2).sum(dtype=F32))            # VelocityCellProcessorRMM: FP32, not mass weighted
        T = summv2 / (3.0 * N)
        beta = (3.0 * N * self.T_target / summv2) ** 0.4 if self.T_target > 0.0 else 1.0
        rho = N / float(np.prod(self.L))
        upot = upot6 / 6.0 + uc
        p = rho * T + rho * ((vir + wc) / N) / 3.0
        return dict(upot=upot, ukin=0.5 * summv2, p=p, beta_trans=beta, T=T, virial=vir + wc)

    def start(self):
        """Simulation::prepare_start (Simulation.cpp:847-1000): wrap, halo, the initial force with dt/2, global values. The
        thermostat is NOT applied here (the factor is only reported): the golden trace pins that."""
        self._wrap()
        u6, w = self._forces(0.5 * self.dt)
        return self._globals(u6, w)

    def step(self):
        self.r = (self.r + F32(self.dt) * self.v).astype(F32)                        # PositionCellProcessorRMM
        self._wrap()
        u6, w = self._forces(self.dt)
        g = self._globals(u6, w)
        self.v = self.v * F32(g["beta_trans"])                                      # VelocityScalingThermostat::apply
        return g

    def simulate(self, nsteps):
        rows = [self.start()]
        for _ in range(nsteps):
            rows.append(self.step())
        return rows
