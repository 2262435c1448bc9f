"""Multi-GPU block decomposition on the device (halo.cu), through the C ABI. Needs a B200: run with -m gpu.

(1) several contexts -- one per block -- on ONE device, messages staged through the host (ls1b200_halo_export_host /
    ls1b200_halo_import_host): every exported record and every force is compared with the decomposed CPU oracle and with the
    single-domain CUDA path; (2) with >= 2 GPUs, the NCCL transport inside the library (torchrun, one process per GPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import TOL, rel_close
from test_decomposition import gather, make_ranks, split

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ls1_mardyn_b200 import capi
    capi.lib()
    return capi


class Recorder:
    """Wraps a rank object and keeps what it exported (to compare CUDA and oracle message by message)."""

    def __init__(self, inner):
        self.inner = inner; self.log = []

    def halo_export(self, dim, side):
        r = self.inner.halo_export(dim, side); self.log.append((dim, side, r)); return r

    def __getattr__(self, k):
        return getattr(self.inner, k)


def make_gpu_ranks(capi, ps, dds, owner, prec, rc):
    ranks = []
    for dd in dds:
        c = capi.Context(list(dd.box_min()), list(dd.box_max()), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=dd.periodic())
        c.set_components(ps.ensemble)
        c.decomp_attach(dd)
        m = owner == dd.rank
        c.upload(ps.id[m], ps.cid[m], ps.r[m], ps.v[m], ps.q[m], ps.L[m])
        ranks.append(c)
    return ranks


def same_records(a, b, exact):
    """Same molecules with the same role; positions bit-identical while the two states are (step 0), afterwards within the
    drift the force tolerance allows."""
    if len(a) != len(b):
        return False
    a = np.sort(a, order=["flags", "id", "r"]); b = np.sort(b, order=["flags", "id", "r"])
    if not (np.array_equal(a["id"], b["id"]) and np.array_equal(a["cid"], b["cid"]) and np.array_equal(a["flags"], b["flags"])):
        return False
    lv = a["flags"] == 0
    if exact:
        return np.array_equal(a["r"], b["r"]) and np.array_equal(a["v"][lv], b["v"][lv]) and np.array_equal(a["q"][lv], b["q"][lv])
    return np.allclose(a["r"], b["r"], rtol=0, atol=1e-4) and np.allclose(a["v"][lv], b["v"][lv], rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("n_ranks,grid", [(2, None), (4, None), (8, None), (3, (1, 3, 1))])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_blocks_on_one_device_vs_oracle(capi, n_ranks, grid, prec):
    from ls1_mardyn_b200 import generators
    from ls1_mardyn_b200.decomposition import exchange_local
    from oracle import oracle as orc
    ps = generators.lj_lattice(9, seed=4, jitter=0.15); ps.v *= 6.0
    rc, dt = 2.0, 0.005
    dds, owner = split(ps, n_ranks, grid)
    G = [Recorder(k) for k in make_gpu_ranks(capi, ps, dds, owner, prec, rc)]
    O = [Recorder(k) for k in make_ranks(ps, dds, owner, prec, rc)]
    one = capi.Context([0, 0, 0], list(ps.length), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=(1, 1, 1))
    one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    one.update(); m1 = one.traverse_cells()
    o1 = np.argsort(one.download("ID"))
    for step in range(8):
        # step 0: both sides see bit-identical inputs -> the north-star tolerance. Later steps compare two trajectories of a hot,
        # jittered (stiff) system whose states have drifted apart by the force tolerance: the bound is widened accordingly.
        tol = TOL[prec] * (1 if step == 0 else 50)
        if step:
            for k in G + O:
                k.leapfrog_pre(dt)
            one.leapfrog_pre(dt); one.update(); m1 = one.traverse_cells()
        exchange_local(G, dds); exchange_local(O, dds)
        # message by message: same molecules, bit-identical shifted positions (integer / exact work: no tolerance)
        for g, o in zip(G, O):
            assert len(g.log) == len(o.log)
            for (d, s, a), (_, _, b) in zip(g.log, o.log):
                assert same_records(a, b, exact=(step == 0)), (step, g.inner.cfg.box_min[:], d, s, len(a), len(b))
            g.log.clear(); o.log.clear()
        for k in G + O:
            k.update()
        mg = [k.traverse_cells() for k in G]; mo = [k.traverse_cells() for k in O]
        for g, o in zip(G, O):
            assert np.array_equal(np.sort(g.download("ID")), np.sort(o.download("ID")))      # same owners
        ids, Fg = gather(G, "F"); _, Fo = gather(O, "F")
        ok, worst = rel_close(Fg, Fo, tol); assert ok, (step, worst)
        ok, worst = rel_close(Fg, one.download("F")[o1], 3 * tol); assert ok, (step, "vs single domain", worst)      # other cells, other summation order
        up = sum(m.upot for m in mg)
        assert abs(up - sum(m.upot for m in mo)) <= tol * abs(up) and abs(up - m1.upot) <= 3 * tol * abs(up)
        vir = sum(m.virial for m in mg)
        assert abs(vir - m1.virial) <= 3 * tol * max(abs(m1.virial), np.abs(one.download("VI")).sum())
        kg = [k.leapfrog_post(dt) for k in G]; ko = [k.leapfrog_post(dt) for k in O]; k1 = one.leapfrog_post(dt)
        assert sum(k.n_molecules for k in kg) == len(ps.id)
        assert abs(sum(k.summv2 for k in kg) - k1.summv2) <= 10 * tol * k1.summv2
    ids, Rg = gather(G, "R")
    assert np.abs(Rg - one.download("R")[o1]).max() <= (1e-9 if prec == "double" else 1e-4)


def _water_lattice(n_per_dim=10, seed=7):
    """Water-like component of the reference's VectorizationWater fixture (1 LJ + 3 charges, rotating) on a lattice whose box
    (L = 74.0 for n = 10, rc = 14.9) is wide enough for two blocks of two cutoff radii per dimension."""
    from helpers import load
    from ls1_mardyn_b200 import generators
    g = load("periodic_water_432")
    ps = generators.lattice(g["ensemble"], n_per_dim, 250.0 / 37.0 ** 3, 9.437e-4, seed=seed, jitter=0.1, rc=14.9)
    return ps, 14.9


@pytest.mark.parametrize("n_ranks", [2, 8])
@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_multisite_blocks(capi, n_ranks, prec):
    """Rotating multi-site molecules through the device-side export / import (120-byte records: orientation and angular
    momentum travel with the molecule), several blocks on one device: every exported record against the decomposed CPU
    oracle, forces and torques against the oracle ranks and against the single-domain CUDA path, over 4 steps."""
    from ls1_mardyn_b200.decomposition import exchange_local
    ps, rc = _water_lattice()
    assert ps.length.min() >= 4 * rc
    dds, owner = split(ps, n_ranks)
    G = [Recorder(k) for k in make_gpu_ranks(capi, ps, dds, owner, prec, rc)]
    O = [Recorder(k) for k in make_ranks(ps, dds, owner, prec, rc)]
    one = capi.Context([0, 0, 0], list(ps.length), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=(1, 1, 1))
    one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    one.update(); one.traverse_cells(); o1 = np.argsort(one.download("ID"))
    dt = 0.0166879
    n_leaving = 0
    for step in range(4):
        tol = TOL[prec] * (1 if step == 0 else 50)      # see test_blocks_on_one_device_vs_oracle
        if step:
            for k in G + O:
                k.leapfrog_pre(dt)
            one.leapfrog_pre(dt); one.update(); one.traverse_cells()
        exchange_local(G, dds); exchange_local(O, dds)
        for g, o in zip(G, O):
            assert len(g.log) == len(o.log)
            for (d, sd, a), (_, _, b) in zip(g.log, o.log):
                assert same_records(a, b, exact=(step == 0)), (step, d, sd, len(a), len(b))
                if step == 0:      # the rotational state of the leaving molecules is bit-identical on both sides
                    a2 = np.sort(a, order=["flags", "id", "r"]); b2 = np.sort(b, order=["flags", "id", "r"]); lv = a2["flags"] == 0
                    assert np.array_equal(a2["q"][lv], b2["q"][lv]) and np.array_equal(a2["L"][lv], b2["L"][lv])
                n_leaving += int((a["flags"] == 0).sum())
            g.log.clear(); o.log.clear()
        for k in G + O:
            k.update()
        for k in G + O:
            k.traverse_cells()
        for g, o in zip(G, O):
            assert np.array_equal(np.sort(g.download("ID")), np.sort(o.download("ID")))
        for f in ("F", "M"):
            _, a = gather(G, f); _, b = gather(O, f)
            ok, worst = rel_close(a, b, tol); assert ok, (step, f, "vs oracle ranks", worst)
            ok, worst = rel_close(a, one.download(f)[o1], 3 * tol); assert ok, (step, f, "vs single domain", worst)
        kg = [k.leapfrog_post(dt) for k in G]; [k.leapfrog_post(dt) for k in O]; k1 = one.leapfrog_post(dt)
        assert sum(k.n_molecules for k in kg) == len(ps.id)
        assert abs(sum(k.sum_iw2 for k in kg) - k1.sum_iw2) <= 10 * tol * max(k1.sum_iw2, 1e-300)
    _, qg = gather(G, "Q")
    assert np.abs(qg - one.download("Q")[o1]).max() <= (1e-9 if prec == "double" else 1e-4)


@pytest.mark.parametrize("n_ranks", [2, 8])
def test_split_step_kernels_on_one_device(capi, n_ranks, monkeypatch):
    """The overlapped multi-GPU step integrates and bins BOUNDARY and INNER slots in separate launches (Region, csrc/common.cuh:
    the exchange runs between them on its own stream inside ls1b200_simulate). With LS1B200_SPLIT_TEST=1 the phase calls of a
    decomposed context run exactly those two launches back to back, so the partition of the slots -- every slot in exactly one
    launch, imports with the boundary -- is checked on ONE device against the decomposed oracle, like the unsplit path."""
    monkeypatch.setenv("LS1B200_SPLIT_TEST", "1")
    test_blocks_on_one_device_vs_oracle(capi, n_ranks, None, "double")
    test_multisite_blocks(capi, n_ranks, "double")


@pytest.mark.parametrize("prec", ["double", "mixed"])
def test_balanced_blocks_on_one_device(capi, prec):
    """Load-balanced decomposition (the tensor-product analogue of parallel/KDDecomposition.cpp): an inhomogeneous
    film-with-vapour system cut where the molecules are, blocks of UNEQUAL width and population as contexts on one device (different
    message capacities per block), against the decomposed CPU oracle record by record and against the single-domain CUDA path."""
    from ls1_mardyn_b200 import generators
    from ls1_mardyn_b200.decomposition import DomainDecomposition, balanced_cuts, exchange_local
    rc = 2.0
    ps = generators.lj_lattice(12, seed=8, jitter=0.1)           # 3456 molecules, L = 17.03
    L = ps.length
    keep = (ps.r[:, 0] < 0.4 * L[0]) | (np.arange(len(ps.id)) % 5 == 0)
    for f in ("id", "cid", "r", "v", "q", "L"):
        setattr(ps, f, getattr(ps, f)[keep])
    grid = (2, 2, 1)
    cuts = balanced_cuts(ps.r, L, grid, 2 * rc)
    dds = [DomainDecomposition(r, 4, L, grid, cuts=cuts) for r in range(4)]
    owner = dds[0].owner_of(ps.r)
    counts = np.bincount(owner, minlength=4)
    equal = np.bincount(DomainDecomposition(0, 4, L, grid).owner_of(ps.r), minlength=4)
    assert counts.max() - counts.min() < equal.max() - equal.min(), (counts, equal)
    ps.v *= 5.0
    G = [Recorder(k) for k in make_gpu_ranks(capi, ps, dds, owner, prec, rc)]
    O = [Recorder(k) for k in make_ranks(ps, dds, owner, prec, rc)]
    one = capi.Context([0, 0, 0], list(L), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=(1, 1, 1))
    one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
    one.update(); m1 = one.traverse_cells(); o1 = np.argsort(one.download("ID"))
    dt = 0.005
    for step in range(6):
        tol = TOL[prec] * (1 if step == 0 else 50)
        if step:
            for k in G + O:
                k.leapfrog_pre(dt)
            one.leapfrog_pre(dt); one.update(); m1 = one.traverse_cells()
        exchange_local(G, dds); exchange_local(O, dds)
        for g, o in zip(G, O):
            assert len(g.log) == len(o.log)
            for (d, sd, a), (_, _, b) in zip(g.log, o.log):
                assert same_records(a, b, exact=(step == 0)), (step, d, sd, len(a), len(b))
            g.log.clear(); o.log.clear()
        for k in G + O:
            k.update()
        mg = [k.traverse_cells() for k in G]
        for k in O:
            k.traverse_cells()
        _, Fg = gather(G, "F"); _, Fo = gather(O, "F")
        ok, worst = rel_close(Fg, Fo, tol); assert ok, (step, worst)
        ok, worst = rel_close(Fg, one.download("F")[o1], 3 * tol); assert ok, (step, "vs single domain", worst)
        up = sum(m.upot for m in mg)
        assert abs(up - m1.upot) <= 3 * tol * abs(m1.upot)
        kg = [k.leapfrog_post(dt) for k in G]; [k.leapfrog_post(dt) for k in O]; one.leapfrog_post(dt)
        assert sum(k.n_molecules for k in kg) == len(ps.id)


@pytest.mark.parametrize("transport", ["peer", "nccl", "peer-overlap2", "nccl-overlap2", "peer-overlap1"])
def test_nccl_two_gpus(capi, tmp_path, transport):
    """The transports inside the library (ls1b200_exchange inside ls1b200_simulate), one process per GPU: records stored
    straight into the neighbour's memory over NVLink (CUDA IPC, ls1b200_peer_attach), and NCCL send / recv. Default: the exchange
    in front of the sort. "-overlap2" (LS1B200_OVERLAP2=1): the exchange on its own stream beside the sort and the force traversal
    of the inner chunks, the imports sorted behind the sorted slots, the boundary chunks reading every cell as two slot ranges
    (1CLJ; the rotating system of the worker takes the in-line path); "-overlap1" (LS1B200_OVERLAP=1): the exchange beside the
    integration / binning of the inner slots. Both overlapped variants are parity-green and measured slower (profiles/r02_multi_gpu.md)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29641",
           os.path.join(ROOT, "tests", "nccl_decomposition_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PYTHONPATH=ROOT, LS1B200_TRANSPORT=transport.split("-")[0], LS1B200_OVERLAP="1" if transport.endswith("overlap1") else "0",
                                LS1B200_OVERLAP2="1" if transport.endswith("overlap2") else "0"))
    assert p.returncode == 0 and "NCCL-DECOMP-OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
