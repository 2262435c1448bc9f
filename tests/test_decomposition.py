"""The multi-rank path on CPU: the host-side mirror of the reference's DomainDecomposition (process grid, block bounds,
neighbours -- src/parallel/DomainDecomposition.cpp:22-54, :130-139) and the staged leaving-molecule + halo exchange,
checked with the decomposed CPU oracle (oracle/decomposed.py) against the single-domain oracle, in one process and with
two processes over torch.distributed / gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from ls1_mardyn_b200 import generators
from ls1_mardyn_b200.capi import PARTICLE
from ls1_mardyn_b200.components import lj_fluid
from ls1_mardyn_b200.decomposition import DomainDecomposition, balanced_cuts, dims_create, exchange_local
from oracle import oracle as orc
from oracle.decomposed import OracleRank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dims_create_and_bounds():
    assert [dims_create(n) for n in (1, 2, 4, 8, 6, 12)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 1), (3, 2, 2)]
    L = [30.0, 20.0, 10.0]
    for n in (1, 2, 4, 8, 6):
        dds = [DomainDecomposition(r, n, L) for r in range(n)]
        vol = 0.0
        for dd in dds:
            assert dd.rank_of(dd.coords) == dd.rank
            lo, hi = dd.box_min(), dd.box_max()
            vol += float(np.prod(hi - lo))
            for d in range(3):
                # getBoundingBoxMin/Max (DomainDecomposition.cpp:130-139): coords*L/grid, the last block ends exactly at L
                assert lo[d] == dd.coords[d] * L[d] / dd.grid[d]
                if dd.coords[d] + 1 == dd.grid[d]:
                    assert hi[d] == L[d]
                # neighbours are periodic and mutually consistent
                for side in (0, 1):
                    nb = dds[dd.neighbour(d, side)]
                    assert nb.neighbour(d, 1 - side) == dd.rank
                assert dd.periodic()[d] == int(dd.grid[d] == 1)
        assert abs(vol - 6000.0) < 1e-9
    with pytest.raises(ValueError):
        DomainDecomposition(0, 4, L, grid=(3, 1, 1))
    dd = DomainDecomposition(0, 8, L)
    r = np.array([[0.0, 0.0, 0.0], [29.9, 19.9, 9.9], [15.0, 10.0, 5.0], [14.999, 9.999, 4.999]])
    assert list(dd.owner_of(r)) == [0, 7, 7, 0]


def split(ps, n_ranks, grid=None):
    dds = [DomainDecomposition(r, n_ranks, ps.length, grid) for r in range(n_ranks)]
    owner = dds[0].owner_of(ps.r)
    return dds, owner


def make_ranks(ps, dds, owner, prec, rc=2.5):
    ranks = []
    for dd in dds:
        k = OracleRank(dd, rc, ps.ensemble, precision=prec, epsilon_rf=ps.epsilon_rf)
        m = owner == dd.rank
        k.upload(ps.id[m], ps.cid[m], ps.r[m], ps.v[m], ps.q[m], ps.L[m])
        ranks.append(k)
    return ranks


def gather(ranks, field):
    ids = np.concatenate([k.download("ID") for k in ranks])
    val = np.concatenate([k.download(field) for k in ranks])
    o = np.argsort(ids)
    return ids[o], val[o]


def step_all(ranks, dds, dt, single=None):
    for k in ranks:
        k.leapfrog_pre(dt)
    exchange_local(ranks, dds)
    for k in ranks:
        k.update()
    ms = [k.traverse_cells() for k in ranks]
    ks = [k.leapfrog_post(dt) for k in ranks]
    return sum(m.upot for m in ms), sum(m.virial for m in ms), sum(k.summv2 for k in ks), sum(k.n_molecules for k in ks)


@pytest.mark.parametrize("n_ranks,grid", [(2, None), (4, None), (8, None), (3, (1, 3, 1)), (4, (1, 2, 2))])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_decomposed_oracle_equals_single_domain(n_ranks, grid, prec):
    """Union of the blocks == the single periodic domain: forces, energies, and the trajectory over steps with migration."""
    ps = generators.lj_lattice(9, seed=4, jitter=0.15)          # 1458 molecules, L = 12.77: blocks >= 2 rc only for grid 2 -> use rc 2.0
    rc = 2.0
    ps.v *= 6.0                                                 # hot: molecules cross block faces within a few steps
    dds, owner = split(ps, n_ranks, grid)
    ranks = make_ranks(ps, dds, owner, prec, rc)
    one = orc.OracleSim([0, 0, 0], list(ps.length), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=(1, 1, 1))
    one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    # step 0: exchange + forces
    exchange_local(ranks, dds)
    for k in ranks:
        k.update()
    ms = [k.traverse_cells() for k in ranks]
    one.update(); m1 = one.traverse_cells()
    tol = 1e-11 if prec == "double" else 2e-6
    ids, F = gather(ranks, "F")
    assert np.array_equal(ids, np.sort(ps.id))
    o1 = np.argsort(one.download("ID"))
    F1 = one.download("F")[o1]
    scale = np.sqrt((F1 ** 2).mean())
    assert np.abs(F - F1).max() <= tol * max(scale, np.abs(F1).max())
    assert abs(sum(m.upot for m in ms) - m1.upot) <= tol * abs(m1.upot)
    assert abs(sum(m.virial for m in ms) - m1.virial) <= tol * max(abs(m1.virial), np.abs(one.download("VI")).sum())
    # a few steps: ownership migrates, the trajectory stays that of the single domain
    dt = 0.005
    moved = 0
    own0 = {int(i): r for r, k in enumerate(ranks) for i in k.download("ID")}
    for step in range(10):
        up, vir, mv2, n = step_all(ranks, dds, dt)
        one.leapfrog_pre(dt); one.update(); m1 = one.traverse_cells(); k1 = one.leapfrog_post(dt)
        assert n == len(ps.id)
        assert abs(up - m1.upot) <= 10 * tol * abs(m1.upot), (step, up, m1.upot)
        assert abs(mv2 - k1.summv2) <= 10 * tol * k1.summv2
    own1 = {int(i): r for r, k in enumerate(ranks) for i in k.download("ID")}
    moved = sum(1 for i in own0 if own0[i] != own1[i])
    assert moved >= 3, "the test must exercise migration"
    ids, R = gather(ranks, "R")
    R1 = one.download("R")[o1]
    assert np.abs(R - R1).max() <= (1e-9 if prec == "double" else 1e-4)


def test_balanced_cuts_and_inhomogeneous_system():
    """Load-balanced tensor-product decomposition (host logic only; the library side is the same per-rank box): the cuts follow
    the molecule distribution, respect the minimum width, and the union of the blocks still reproduces the single domain."""
    rc = 2.0
    ps = generators.lj_lattice(9, seed=8, jitter=0.1)           # 1458 molecules, L = 12.77
    L = ps.length
    # an inhomogeneous system: keep all molecules of the lower 45 % of x, every fourth one elsewhere (a film with vapour)
    keep = (ps.r[:, 0] < 0.45 * L[0]) | (np.arange(len(ps.id)) % 4 == 0)
    for f in ("id", "cid", "r", "v", "q", "L"):
        setattr(ps, f, getattr(ps, f)[keep])
    n = len(ps.id)
    grid = (2, 1, 1)
    cuts = balanced_cuts(ps.r, L, grid, 2 * rc)
    assert cuts[0][0] == 0.0 and cuts[0][2] == L[0] and len(cuts[1]) == 2
    assert cuts[0][1] < 0.45 * L[0] and cuts[0][1] >= 2 * rc and L[0] - cuts[0][1] >= 2 * rc      # the cut moved into the dense part
    dds = [DomainDecomposition(r, 2, L, grid, cuts=cuts) for r in range(2)]
    owner = dds[0].owner_of(ps.r)
    counts = np.bincount(owner, minlength=2)
    equal = np.bincount(DomainDecomposition(0, 2, L, grid).owner_of(ps.r), minlength=2)
    assert abs(int(counts[0]) - int(counts[1])) < abs(int(equal[0]) - int(equal[1])), (counts, equal)
    with pytest.raises(ValueError):
        balanced_cuts(ps.r, L, (4, 1, 1), 2 * rc)               # four slabs of width 4 do not fit into 12.77
    with pytest.raises(ValueError):
        DomainDecomposition(0, 2, L, grid, cuts=[np.array([0.0, 3.0]), cuts[1], cuts[2]])
    # physics: blocks of unequal width == the single periodic domain, with migration over a few steps
    ps.v *= 5.0
    ranks = make_ranks(ps, dds, owner, "double", rc)
    one = orc.OracleSim([0, 0, 0], list(L), rc, epsilon_rf=ps.epsilon_rf, precision="double", geometry="xml", periodic=(1, 1, 1))
    one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    exchange_local(ranks, dds)
    for k in ranks:
        k.update()
    ms = [k.traverse_cells() for k in ranks]
    one.update(); m1 = one.traverse_cells()
    ids, F = gather(ranks, "F")
    o1 = np.argsort(one.download("ID")); F1 = one.download("F")[o1]
    assert np.abs(F - F1).max() <= 1e-11 * max(np.sqrt((F1 ** 2).mean()), np.abs(F1).max())
    assert abs(sum(m.upot for m in ms) - m1.upot) <= 1e-11 * abs(m1.upot)
    for step in range(6):
        up, vir, mv2, nn = step_all(ranks, dds, 0.005)
        one.leapfrog_pre(0.005); one.update(); m1 = one.traverse_cells(); k1 = one.leapfrog_post(0.005)
        assert nn == n and abs(up - m1.upot) <= 1e-10 * abs(m1.upot) and abs(mv2 - k1.summv2) <= 1e-10 * k1.summv2


def test_two_processes_gloo(tmp_path):
    """world_size 2 over torch.distributed (gloo): the same staged exchange with the messages on the wire."""
    script = os.path.join(ROOT, "tests", "gloo_decomposition_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 2000), PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "OK" in outs[0]
