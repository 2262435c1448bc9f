"""Synthetic input generation: body-centred cubic lattice like the reference's CubicGridGenerator
(reference: src/io/CubicGridGeneratorInternal.cpp:125-187 for molecules per dimension,
src/particleContainer/ParticleCellBase.cpp:51-159 for the two sub-lattices at 1/4 and 3/4 of the spacing and
the equal-magnitude random-direction velocities, src/utils/generator/EqualVelocityAssigner.h).
The random stream is numpy's, not the reference's per-cell mt19937: parity runs feed the SAME generated state to
both sides through a phase-space file, so only the statistical character has to match."""
from __future__ import annotations

import numpy as np

from .ascii_reader import PhaseSpace
from .components import Ensemble, lj_fluid


def lattice(ens: Ensemble, n_per_dim: int, rho: float, temperature: float, seed: int = 0, jitter: float = 0.0,
            rc: float = 0.0, dtype=np.float64) -> PhaseSpace:
    n = int(n_per_dim)
    N = 2 * n ** 3
    L = (N / rho) ** (1.0 / 3.0)
    a = L / n
    rng = np.random.default_rng(seed)
    g = (np.arange(n, dtype=np.float64) + 0.25) * a
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r0 = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    r = np.concatenate([r0, r0 + 0.5 * a], axis=0)
    if jitter > 0.0:
        r += (rng.random(r.shape) - 0.5) * (2.0 * jitter * a)
        r = np.mod(r, L)
        r[r >= L] = 0.0
    nc = len(ens.components)
    cid = (np.arange(N) % nc).astype(np.uint32)
    mass = np.array([c.m() for c in ens.components])[cid]
    inert = np.array([c.Ipa()[0] for c in ens.components])[cid]
    # equal-magnitude velocities in random directions, |v| = sqrt(3 T / m); total momentum removed
    d = rng.normal(size=(N, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v = d * np.sqrt(3.0 * temperature / mass)[:, None]
    v -= (mass[:, None] * v).sum(axis=0) / mass.sum()
    q = np.zeros((N, 4)); q[:, 0] = 1.0
    Lang = np.zeros((N, 3))
    rotating = inert.any(axis=1)
    if rotating.any():
        qq = rng.normal(size=(N, 4))
        qq /= np.linalg.norm(qq, axis=1, keepdims=True)
        q[rotating] = qq[rotating]
        Lang = rng.normal(size=(N, 3)) * np.sqrt(temperature * inert)
    ids = np.arange(1, N + 1, dtype=np.uint64)
    # shuffle so that storage order is not cell order (exercises the binning)
    perm = rng.permutation(N)
    return PhaseSpace(np.array([L, L, L]), float(temperature), 1e10, ens, ids, cid[perm], r[perm], v[perm], q[perm], Lang[perm])


def lj_lattice(n_per_dim: int, rho: float = 0.7, temperature: float = 0.85, rc: float = 2.5, seed: int = 0,
               jitter: float = 0.0) -> PhaseSpace:
    """Reduced-unit 1CLJ liquid (sigma = eps = m = 1): BASELINE.json configs 1c / 2 / 5."""
    return lattice(lj_fluid(rc=rc), n_per_dim, rho, temperature, seed, jitter, rc)


def lattice_block(ens: Ensemble, n_per_dim: int, rho: float, temperature: float, coords, grid, seed: int = 0) -> PhaseSpace:
    """The part of a BCC lattice of (grid * n_per_dim)^3-shaped extent that lies in block `coords` of a `grid` decomposition:
    every rank of a weak-scaling run generates only its own n_per_dim^3 lattice cells (2 n^3 molecules). ``length`` of the
    returned phase space is the GLOBAL box."""
    n = int(n_per_dim)
    a = (2.0 / rho) ** (1.0 / 3.0)
    L = np.array([grid[d] * n * a for d in range(3)])
    rng = np.random.default_rng(seed + 7919 * (1 + (coords[0] * grid[1] + coords[1]) * grid[2] + coords[2]))
    ax = [(np.arange(n, dtype=np.float64) + coords[d] * n + 0.25) * a for d in range(3)]
    X, Y, Z = np.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    r0 = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    r = np.concatenate([r0, r0 + 0.5 * a], axis=0)
    N = len(r)
    nc = len(ens.components)
    cid = (np.arange(N) % nc).astype(np.uint32)
    mass = np.array([c.m() for c in ens.components])[cid]
    d = rng.normal(size=(N, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    v = d * np.sqrt(3.0 * temperature / mass)[:, None]
    v -= (mass[:, None] * v).sum(axis=0) / mass.sum()
    q = np.zeros((N, 4)); q[:, 0] = 1.0
    Lang = np.zeros((N, 3))
    inert = np.array([c.Ipa()[0] for c in ens.components])[cid]
    rotating = inert.any(axis=1)
    if rotating.any():
        qq = rng.normal(size=(N, 4)); qq /= np.linalg.norm(qq, axis=1, keepdims=True)
        q[rotating] = qq[rotating]
        Lang = rng.normal(size=(N, 3)) * np.sqrt(temperature * inert)
    first = ((coords[0] * grid[1] + coords[1]) * grid[2] + coords[2]) * N
    ids = np.arange(first + 1, first + N + 1, dtype=np.uint64)
    perm = rng.permutation(N)
    return PhaseSpace(L, float(temperature), 1e10, ens, ids, cid[perm], r[perm], v[perm], q[perm], Lang[perm])


# ---- host mirror of ls1b200_generate_cubic_grid (csrc/generate.cu), used by the parity test of the device generator ----------------
def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def _fma(a: float, b: float, c: float) -> float:
    """round(a * b + c) with ONE rounding (exact rational arithmetic; Python < 3.13 has no math.fma)."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def cubic_grid_reference(mols_per_dim, sim_box, box_min, box_max, temperature, mass, rot_dof=0, seed=0, mixed=True, with_lattice_index=False):
    """What the device generator produces for the context box [box_min, box_max): (id, r, v before momentum removal), ordered by id.
    Same arithmetic: coordinate = fma(i, spacing, origin) rounded through the calc type (the reference binary contracts the
    expression), membership min <= r < max, velocity direction from splitmix64(seed ^ 2 * index [+ 1]) (reference:
    ParticleCellBase.cpp:51-159, EqualVelocityAssigner.h:16-29). with_lattice_index: also (sub-lattice, i, j, k) per molecule."""
    n = np.broadcast_to(np.asarray(mols_per_dim, dtype=np.int64), (3,))
    sim_box = np.asarray(sim_box, dtype=np.float64)
    ids, rs, lat = [], [], []
    for s, f in enumerate((0.25, 0.25 * 3.0)):
        axes, idxs = [], []
        for d in range(3):
            spacing = sim_box[d] / float(n[d])
            i = np.arange(n[d], dtype=np.int64)
            x = np.array([_fma(float(k), float(spacing), float(spacing * f)) for k in i])
            if mixed:
                x = x.astype(np.float32).astype(np.float64)
            m = (x >= box_min[d]) & (x < box_max[d])
            axes.append(x[m]); idxs.append(i[m])
        X, Y, Z = np.meshgrid(*axes, indexing="ij"); I, J, K = np.meshgrid(*idxs, indexing="ij")
        g = ((s * n[0] + I) * n[1] + J) * n[2] + K
        ids.append(g.ravel()); rs.append(np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1))
        lat.append(np.stack([np.full(g.size, s, dtype=np.int64), I.ravel(), J.ravel(), K.ravel()], axis=1))
    gidx = np.concatenate(ids).astype(np.uint64); r = np.concatenate(rs); lat = np.concatenate(lat)
    with np.errstate(over="ignore"):
        h1 = _splitmix64(np.uint64(seed) ^ (np.uint64(2) * gidx)); h2 = _splitmix64(np.uint64(seed) ^ (np.uint64(2) * gidx + np.uint64(1)))
    u1 = (h1 >> np.uint64(11)).astype(np.float64) / 9007199254740992.0; u2 = (h2 >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    phi = 2.0 * np.pi * u1; theta = np.arccos(2.0 * u2 - 1.0)
    vabs = np.sqrt((3.0 + rot_dof) * temperature / mass)
    v = vabs * np.stack([np.sin(phi), np.cos(phi) * np.sin(theta), np.cos(phi) * np.cos(theta)], axis=1)
    o = np.argsort(gidx)
    if with_lattice_index:
        return (gidx[o] + np.uint64(1)), r[o], v[o], lat[o]
    return (gidx[o] + np.uint64(1)), r[o], v[o]


def reference_velocity_table(count, temperature, mass, rot_dof=0, mixed=True):
    """The velocity sequence every cell of the reference's CubicGridGenerator walks (ParticleCellBase.cpp:66-67 constructs
    EqualVelocityAssigner(T) -- std::mt19937(0) -- per cell; EqualVelocityAssigner.h:16-29): entry k is what the k-th molecule
    inserted into a cell gets. numpy's legacy RandomState(0) is the same Mersenne twister with the same seeding as std::mt19937(0);
    libstdc++'s uniform_real_distribution<double> draws two 32-bit words per number: (w0 + w1 * 2^32) / 2^64. The ensemble's
    temperature passes through a vcp_real_calc variable (float in the mixed build) before it reaches the assigner."""
    w = np.frombuffer(np.random.RandomState(0).bytes(16 * count), dtype="<u4").astype(np.float64).reshape(count, 2, 2)
    u = (w[:, :, 0] + w[:, :, 1] * 4294967296.0) / 18446744073709551616.0
    assert (u < 1.0).all()
    phi = 2 * np.pi * u[:, 0]; theta = np.arccos(2 * u[:, 1] - 1)
    t_calc = float(np.float32(temperature)) if mixed else float(temperature)
    vabs = np.sqrt((3 + rot_dof) * t_calc / mass)
    return np.stack([vabs * np.sin(phi), vabs * np.cos(phi) * np.sin(theta), vabs * np.cos(phi) * np.cos(theta)], axis=1)


def reference_velocities(lattice_index, cell, temperature, mass, rot_dof=0, mixed=True, remove_momentum=True):
    """Velocities of the reference's generator for molecules with lattice index rows (sub-lattice, i, j, k) lying in LinkedCells
    cells `cell` (the caller bins the positions: the cell grid is the container's). Inside a cell the reference inserts in loop
    order i, j, k with sub-lattice 1 before sub-lattice 2 at the same lattice point (ParticleCellBase.cpp:88-136) -- here simply a
    sort, independent of the closed form the device kernel evaluates. Then IOHelpers::removeMomentum (io/IOHelpers.cpp:8-115):
    subtract sum(m v) / sum(m), scale back to the target temperature."""
    lat = np.asarray(lattice_index, dtype=np.int64); cell = np.asarray(cell, dtype=np.int64)
    order = np.lexsort((lat[:, 0], lat[:, 3], lat[:, 2], lat[:, 1], cell))      # by cell, then i, j, k, sub-lattice
    c_sorted = cell[order]
    first = np.r_[0, np.flatnonzero(np.diff(c_sorted)) + 1]
    start = np.repeat(first, np.diff(np.r_[first, len(cell)]))
    rank = np.empty(len(cell), dtype=np.int64)
    rank[order] = np.arange(len(cell)) - start
    v = reference_velocity_table(int(rank.max()) + 1, temperature, mass, rot_dof, mixed)[rank]
    if remove_momentum:
        v = v - v.sum(axis=0) / len(v)
        t_current = mass * (v ** 2).sum() / ((3 + rot_dof) * len(v))
        v = v * np.sqrt(temperature / t_current)
    return v
