"""Timing experiment: the upper tile sweep alone, the SpMV alone, and both side by side on two streams
(opmb200_time_kernel what=6; results of that run are meaningless, only the time counts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
prec = sys.argv[2] if len(sys.argv) > 2 else "dilu"
A = generators.config(cfg, scale=1.0, with_rhs=False)["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": "tiles"}})
for what, name in ((5, "upper sweep alone"), (0, "spmv alone"), (6, "upper sweep || spmv"), (5, "upper sweep alone")):
    ms, _ = fs.time_kernel(what, 3, 20)
    print(f"{name:24s} {ms*1e3:8.1f} us", flush=True)
fs.close()
