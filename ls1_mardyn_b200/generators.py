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
