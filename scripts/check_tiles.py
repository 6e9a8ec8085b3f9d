"""Tile walkers against the level schedule, bit for bit: repeated preconditioner applications and solves with
random right-hand sides.  python scripts/check_tiles.py cfg scale prec chunk_rows [poll_warps] [rhs_warps] [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
cfg, scale, prec, cr = sys.argv[1], float(sys.argv[2]), sys.argv[3], int(sys.argv[4])
pw = int(sys.argv[5]) if len(sys.argv) > 5 else 4
rw = int(sys.argv[6]) if len(sys.argv) > 6 else 2
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
s = generators.config(cfg, scale=scale)
A = s["A"]
ref = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-4, "maxiter": 200, "preconditioner": {"type": prec}, "b200": {"schedule": "levels"}})
tw = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-4, "maxiter": 200, "preconditioner": {"type": prec},
                                       "b200": {"schedule": "tiles", "chunk_rows": cr, "poll_warps": pw, "rhs_warps": rw}})
print("schedule", tw.info()["schedule"], "chunks", tw.info()["n_chunks"], "chunk_rows", tw.info()["chunk_rows"], "n", A.n, "b", A.b)
rng = np.random.default_rng(1)
bad = 0
for k in range(reps):
    d = rng.standard_normal(A.n * A.b)
    v0, v1 = np.zeros_like(d), np.zeros_like(d)
    ref.preconditioner().apply(v0, d)
    tw.preconditioner().apply(v1, d)
    nd = int(np.count_nonzero(v0 != v1))
    if nd:
        bad += 1
        idx = np.flatnonzero(v0 != v1)
        print(f"apply {k}: {nd} entries differ, first rows {np.unique(idx // A.b)[:8]}, max rel {np.max(np.abs(v0 - v1)) / np.max(np.abs(v0)):.2e}")
for name in ("rhs", "rhs2"):
    x0, r0 = np.zeros(A.n * A.b), s[name].copy()
    x1, r1 = np.zeros(A.n * A.b), s[name].copy()
    a = ref.apply(x0, r0); b = tw.apply(x1, r1)
    same = np.array_equal(ref.history(), tw.history())
    print(name, "iterations", a.iterations, b.iterations, "history identical", same)
    pass  # (the two layouts sum their dot products in different orders: histories agree to rounding only)
print("BAD" if bad else "OK", bad)
