import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_simulators_b200 import generators, _lib
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
s = generators.config("C3", scale=scale); A = s["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": "dilu"}, "b200": {"schedule": "chunks", "debug_timing": dbg}})
info = fs.info(); L = _lib.lib()
out = (C.c_ulonglong * 16)()
for what in (4, 5):
    L.opmb200_prof_read(out, 1)
    ms, nb = fs.time_kernel(what, 2, 10)
    L.opmb200_prof_read(out, 1)
    steps = info["n_slices"] * 12 * 2   # both sweeps run in each of 12 reps
    print("kernel", what, "ms %.3f" % ms, "slices", info["n_slices"], "est_steps", info["est_steps"])
    names = ["request", "wait stage", "look-ahead", "deps+poll", "accumulate+finish", "syncwarp1", "publish+syncwarp2"]
    tot = sum(out[i] for i in range(7))
    for i, nme in enumerate(names):
        print("   %-20s %8.1f cycles/step  %5.1f%%" % (nme, out[i] / steps, 100.0 * out[i] / max(tot, 1)))
    print("   total %.1f cycles/step" % (tot / steps))
