"""Small target for ncu captures: create a solver and run a few preconditioner applications / solves.
   ncu --set full --clock-control none --import-source on -k regex:tw_sweep -s 2 -c 2 -o gpurun_out/x \
       python scripts/ncu_target.py C3 tiles dilu 3"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
sched = sys.argv[2] if len(sys.argv) > 2 else "tiles"
prec = sys.argv[3] if len(sys.argv) > 3 else "dilu"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
what = sys.argv[5] if len(sys.argv) > 5 else "apply"
scale = float(os.environ.get("SCALE", "1.0"))
dims = os.environ.get("DIMS")  # e.g. 1000x8x4: one free-running tile
s = generators.blackoil_system(*[int(v) for v in dims.split("x")], b=3, seed=5) if dims else generators.config(cfg, scale=scale)
A = s["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "preconditioner": {"type": prec}, "b200": {"schedule": sched, "chunk_rows": int(os.environ.get("CHUNK_ROWS", "0"))}})
d = s["rhs2"]
if what == "apply":
    for _ in range(reps):
        fs.time_kernel(1, 0, 1)
elif what == "update":
    for _ in range(reps):
        fs.update()
else:
    for _ in range(reps):
        x, r = np.zeros_like(d), d.copy()
        fs.apply(x, r)
fs.close()
