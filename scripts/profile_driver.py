"""small driver for ncu captures: one C3 prepare + solve through the C ABI"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
prec = sys.argv[2] if len(sys.argv) > 2 else "dilu"
s = generators.config(cfg)
A = s["A"]
sched = sys.argv[3] if len(sys.argv) > 3 else "levels"
fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": prec}, "b200": {"schedule": sched}})
x, r = np.zeros(A.n * A.b), s["rhs2"].copy()
print(fs.apply(x, r), fs.info()["kernel_launches"])
