"""torchrun worker: N ranks, one GPU each; the fused ls1b200_simulate loop with the NCCL halo exchange inside the library
against the same system on a single GPU (computed by rank 0). Prints NCCL-DECOMP-OK on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ls1_mardyn_b200 import capi, generators  # noqa: E402
from ls1_mardyn_b200.decomposition import DomainDecomposition  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    n = int(os.environ.get("LS1_TEST_N", "24"))
    ps = generators.lj_lattice(n, seed=3, jitter=0.1); ps.v *= 3.0
    N = len(ps.id); V = float(np.prod(ps.length))
    rc, dt, T, steps = 2.5, 0.002, 0.85, 40
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    for prec, tol in (("double", 1e-9), ("mixed", 1e-6)):
        dd = DomainDecomposition(rank, world, ps.length)
        c = capi.Context(list(dd.box_min()), list(dd.box_max()), rc, precision=prec, periodic=dd.periodic(), device=local)
        c.set_components(ps.ensemble)
        if prec == "mixed":
            uid = [capi.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
        c.decomp_attach(dd, uid[0])
        m = dd.owner_of(ps.r) == rank
        c.upload(ps.id[m], ps.cid[m], ps.r[m], ps.v[m])
        if os.environ.get("LS1B200_TRANSPORT", "peer") != "nccl":      # NVLink stores into the neighbours' buffers instead of ncclSend / Recv
            c.peer_attach_collective(dist)
        c.exchange(); c.update(); c.traverse_cells()
        rec = c.simulate(steps, dt, T, global_n=N, global_volume=V)
        cnt = torch.tensor([c.num_molecules()]); dist.all_reduce(cnt)
        assert int(cnt) == N, (int(cnt), N)
        if rank == 0:
            one = capi.Context([0, 0, 0], list(ps.length), rc, precision=prec, device=local)
            one.set_components(ps.ensemble); one.upload(ps.id, ps.cid, ps.r, ps.v)
            one.update(); one.traverse_cells()
            ref = one.simulate(steps, dt, T)
            for k, i in dict(upot=0, summv2=2, temperature=4, beta_trans=6).items():
                err = np.abs(rec[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max()
                assert err <= tol, (prec, k, err)
            rho = N / V
            err = np.abs(rec[:, 5] - ref[:, 5]).max() / max(np.abs(ref[:, 5]).max(), rho * np.abs(ref[:, 4]).max())
            assert err <= tol, (prec, "pressure", err)
            print(f"{prec} ({os.environ.get('LS1B200_TRANSPORT', 'peer')} transport): {world} ranks vs 1 GPU over {steps} steps ok (N={N}, owned by rank 0 now: {c.num_molecules()})")
        c.peer_detach_collective(dist)      # (a barrier inside) nobody frees receive buffers a neighbour still has mapped
        c.close()
        dist.barrier()
    # ---- rotating multi-site molecules (water-like: 1 LJ + 3 charges): the 120-byte records with orientation and angular momentum
    from ls1_mardyn_b200 import workloads
    wl = workloads.get("water")
    pw = generators.lattice(wl.ensemble, 10, wl.rho, wl.temperature, seed=7, jitter=0.05, rc=wl.rc)
    Nw = len(pw.id); Vw = float(np.prod(pw.length))
    uid = [capi.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dd = DomainDecomposition(rank, world, pw.length)
    c = capi.Context(list(dd.box_min()), list(dd.box_max()), wl.rc, epsilon_rf=pw.epsilon_rf, precision="double", periodic=dd.periodic(), device=local)
    c.set_components(pw.ensemble)
    c.decomp_attach(dd, uid[0])
    m = dd.owner_of(pw.r) == rank
    c.upload(pw.id[m], pw.cid[m], pw.r[m], pw.v[m], pw.q[m], pw.L[m])
    if os.environ.get("LS1B200_TRANSPORT", "peer") != "nccl":
        c.peer_attach_collective(dist)
    c.exchange(); c.update(); c.traverse_cells()
    rec = c.simulate(20, wl.dt, wl.temperature, global_n=Nw, global_volume=Vw)
    if rank == 0:
        one = capi.Context([0, 0, 0], list(pw.length), wl.rc, epsilon_rf=pw.epsilon_rf, precision="double", device=local)
        one.set_components(pw.ensemble); one.upload(pw.id, pw.cid, pw.r, pw.v, pw.q, pw.L)
        one.update(); one.traverse_cells()
        ref = one.simulate(20, wl.dt, wl.temperature)
        for k, i in dict(upot=0, summv2=2, sum_iw2=3, temperature=4).items():
            err = np.abs(rec[:, i] - ref[:, i]).max() / np.abs(ref[:, i]).max()
            assert err <= 1e-9, ("water", k, err)
        print(f"water-like, double: {world} ranks vs 1 GPU over 20 steps ok (N={Nw})")
    c.peer_detach_collective(dist)
    c.close()
    dist.barrier()
    if rank == 0:
        print("NCCL-DECOMP-OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
