"""Sweep times of the LOCAL system of one rank of a z-slab partition, on one GPU (no communicator: the sweeps of a
block-Jacobi preconditioner are local).  python scripts/rank_slab.py CONFIG NZ_PER_RANK WORLD [ranks...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, args = sys.argv[:1], sys.argv[1:]
import bench  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

config, nzr, world = args[0], int(args[1]), int(args[2])
ranks = [int(r) for r in args[3:]] or list(range(world))
sched = os.environ.get("SCHED", "auto")
extra = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv)}
for rank in ranks:
    w = bench.workload(config, rank, world, nz_per_rank=nzr)
    for prec in os.environ.get("PREC", "dilu").split(","):
        fs = FlexibleSolver(MatrixAdapter(w["A"], w["n_interior"]), {"preconditioner": {"type": prec}, "b200": {"schedule": sched, **extra}})
        i = fs.info()
        t = {name: fs.time_kernel(what, 3, 10)[0] for what, name in ((4, "lower"), (5, "upper"), (0, "spmv"), (2, "update"))}
        print(f"{config} nz/rank {nzr} rank {rank}/{world} {prec}: rows {i['n_rows']} interior {w['n_interior']} schedule {i['schedule']} "
              f"levels/steps {i['n_levels']} chunks {i['n_chunks']} chunk_rows {i['chunk_rows']} padded {i['padded_blocks'] / i['nnzb']:.3f} | "
              + " ".join(f"{k} {v * 1e3:.1f} us" for k, v in t.items()), flush=True)
        fs.close()
