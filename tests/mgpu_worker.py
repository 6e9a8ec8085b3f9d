"""Worker for the multi-rank tests, started once per rank by torch.distributed.run.

  --mode cpu : gloo backend, host logic only (slab generation, ghost-last localisation, halo lists,
               unique-id plumbing) checked against the localisation of the global matrix
  --mode gpu : one GPU per rank, NCCL inside libopmb200 (halo exchange + all-reduce), block-Jacobi
               DILU / ILU0, checked against the oracle emulating the same ranks in one process
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from opm_simulators_b200 import generators, partition  # noqa: E402


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--prec", default="dilu")
    ap.add_argument("--b", type=int, default=3)
    ap.add_argument("--collectives", default="nccl", choices=["nccl", "p2p"])
    ap.add_argument("--schedule", default="auto", choices=["levels", "tiles", "auto"])
    ap.add_argument("--halo-overlap", type=int, default=1)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.mode == "gpu":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group("gloo")

    nx, ny, nzl, b = 7, 6, 4, args.b
    nz = nzl * world
    kw = dict(nx=nx, ny=ny, nz=nz, b=b, seed=321, sigma=1.0, kz_mult=0.5)
    # every rank generates ONLY its own slab ...
    slab = generators.blackoil_system(z_range=(rank * nzl, (rank + 1) * nzl), with_rhs=False, **kw)
    n_global = nx * ny * nz
    bounds = partition.partition_bounds(n_global, world)
    owner_of = lambda g: np.searchsorted(bounds, g, side="right") - 1  # noqa: E731
    ls = partition.localize_rows(slab["row0"], slab["rowptr"], slab["gcol"], slab["val"], owner_of, rank)
    # ... which must equal the localisation of the global matrix (integer maps bit-exact)
    full = generators.blackoil_system(**kw)
    A = full["A"]
    part = partition.partition_simple(n_global, world)
    assert np.array_equal(part, owner_of(np.arange(n_global)))
    ref = partition.localize(A, part, rank)
    assert np.array_equal(ref.l2g, ls.l2g) and np.array_equal(ref.A.rowptr, ls.A.rowptr)
    assert np.array_equal(ref.A.col, ls.A.col) and np.array_equal(ref.A.val, ls.A.val)
    for k in ref.halo:
        assert np.array_equal(ref.halo[k], ls.halo[k]), k
    # halo lists agree pairwise: what I send to o is what o expects from me (global ids, same order)
    mine = {int(o): ls.l2g[ls.halo["send_rows"][ls.halo["send_ptr"][k]:ls.halo["send_ptr"][k + 1]]].tolist()
            for k, o in enumerate(ls.halo["neighbors"])}
    want = {int(o): ls.l2g[ls.halo["recv_rows"][ls.halo["recv_ptr"][k]:ls.halo["recv_ptr"][k + 1]]].tolist()
            for k, o in enumerate(ls.halo["neighbors"])}
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, want))
    for o, (m_o, w_o) in enumerate(gathered):
        if o != rank and rank in w_o:
            assert w_o[rank] == mine[o], "halo mismatch"

    if args.mode == "gpu":
        from opm_simulators_b200.flexible_solver import Communication, FlexibleSolver, MatrixAdapter
        from oracle import oracle as orc
        from opm_simulators_b200 import _lib

        _lib.check(_lib.lib().opmb200_set_device(local_rank))
        ids = [Communication.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = Communication(rank, world, ids[0])
        tol = 1e-8
        fs = FlexibleSolver(MatrixAdapter(ls.A, ls.n_interior, comm, ls.halo),
                            {"tol": tol, "maxiter": 200, "preconditioner": {"type": args.prec, "relaxation": 0.9},
                             "b200": {"schedule": args.schedule, "halo_overlap": args.halo_overlap}})
        if args.collectives == "p2p":  # collectives inside the library's own kernels over peer memory
            def allgather(blob):
                out = [None] * world
                dist.all_gather_object(out, blob)
                return out
            fs.enable_p2p(allgather)
        # oracle: all ranks emulated in this process
        locs = [partition.localize(A, part, r) for r in range(world)]
        ps = orc.ParSystem([dict(rowptr=l.A.rowptr, col=l.A.col, val=l.A.val, interior=l.n_interior, l2g=l.l2g)
                            for l in locs], nglobal=n_global)
        ps.prec_update(args.prec, 0.9)
        rng = np.random.default_rng(17)
        dg = rng.standard_normal(n_global * b)
        d_loc = [l.scatter_global(dg) for l in locs]
        # SpMV: owner rows only, ghost rows zero (WellOperators.hpp:432-468)
        y = np.full(ls.n * b, np.nan)
        fs.op.apply(d_loc[rank], y)
        yo = orc.spmv(ls.A.rowptr, ls.A.col, ls.A.val, d_loc[rank], interior=ls.n_interior)
        assert rel_err(y, yo) < 1e-10 and not y[ls.n_interior * b:].any()
        # dot: owner rows, summed over ranks
        dot = fs.dot(d_loc[rank], d_loc[rank])
        assert abs(dot - float(dg @ dg)) < 1e-10 * float(dg @ dg)
        # block-Jacobi preconditioner + copyOwnerToAll
        for l in d_loc:
            l.reshape(-1, b)[0:0] = 0
        vo = ps.prec_apply(d_loc)
        for _ in range(3):  # back to back: exercises the flow control of the halo exchange
            v = np.zeros(ls.n * b)
            fs.preconditioner().apply(v, d_loc[rank])
            assert rel_err(v, vo[rank]) < 1e-10, rel_err(v, vo[rank])
        # whole solve
        rhs = [l.scatter_global(full["rhs2"]) for l in locs]
        for l, r_ in zip(locs, rhs):
            r_.reshape(-1, b)[l.n_interior:] = 0.0
        x, r = np.zeros(ls.n * b), rhs[rank].copy()
        res = fs.apply(x, r)
        xo, ro, reso, hist = ps.bicgstab(rhs, tol=tol, maxiter=200)
        assert res.converged and abs(res.iterations - reso["iterations"]) <= 1, (res.iterations, reso["iterations"])
        assert rel_err(x, xo[rank]) < 1e-6, rel_err(x, xo[rank])
        # and the assembled global solution solves the global system
        xs = [None] * world
        dist.all_gather_object(xs, ls.owner_part(x))
        xg = np.concatenate(xs).reshape(-1)
        true = np.linalg.norm(full["rhs2"] - A.to_scipy() @ xg) / np.linalg.norm(full["rhs2"])
        assert true < 10 * tol, true
        if rank == 0:
            print(f"mgpu {args.prec} b={b} {args.collectives}: ranks={world} iterations={res.iterations} (oracle {reso['iterations']}) "
                  f"true reduction={true:.2e}")
        fs.close()
        comm.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
