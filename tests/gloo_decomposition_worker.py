"""Worker of tests/test_decomposition.py::test_two_processes_gloo: rank r of a 2-rank decomposed oracle run whose halo /
leaving-molecule messages travel over torch.distributed (gloo). Rank 0 checks the result against the single domain."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ls1_mardyn_b200 import generators  # noqa: E402
from ls1_mardyn_b200.capi import PARTICLE  # noqa: E402
from ls1_mardyn_b200.decomposition import DomainDecomposition, exchange_dist  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle.decomposed import OracleRank  # noqa: E402


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ps = generators.lj_lattice(9, seed=7, jitter=0.15)
    ps.v *= 6.0
    rc, dt, prec = 2.0, 0.004, "double"
    dd = DomainDecomposition(rank, world, ps.length)
    m = dd.owner_of(ps.r) == rank
    k = OracleRank(dd, rc, ps.ensemble, precision=prec, epsilon_rf=ps.epsilon_rf)
    k.upload(ps.id[m], ps.cid[m], ps.r[m], ps.v[m], ps.q[m], ps.L[m])
    exchange_dist(k, dd, dist, PARTICLE); k.update(); k.traverse_cells()
    import torch
    trace = []
    for step in range(5):
        k.leapfrog_pre(dt)
        exchange_dist(k, dd, dist, PARTICLE)
        k.update(); mac = k.traverse_cells(); kin = k.leapfrog_post(dt)
        t = torch.tensor([mac.upot, mac.virial, kin.summv2, float(kin.n_molecules)], dtype=torch.float64)
        dist.all_reduce(t)                       # the collective of Domain::calculateGlobalValues
        trace.append(t.numpy().copy())
    if rank == 0:
        one = orc.OracleSim([0, 0, 0], list(ps.length), rc, epsilon_rf=ps.epsilon_rf, precision=prec, geometry="xml", periodic=(1, 1, 1))
        one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v, ps.q, ps.L)
        one.update(); one.traverse_cells()
        for step in range(5):
            one.leapfrog_pre(dt); one.update(); m1 = one.traverse_cells(); k1 = one.leapfrog_post(dt)
            t = trace[step]
            assert int(t[3]) == len(ps.id)
            assert abs(t[0] - m1.upot) <= 1e-10 * abs(m1.upot), (step, t[0], m1.upot)
            assert abs(t[2] - k1.summv2) <= 1e-10 * k1.summv2
        print("OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
