"""In-kernel phase profile of the chunk sweeps (build with OPMB200_PROFILE=1):
   OPMB200_PROFILE=1 python opm_simulators_b200/build.py && python scripts/prof_chunk.py [cfg] [scale] [prec] [prefetch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_simulators_b200 import generators, _lib
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
prec = sys.argv[3] if len(sys.argv) > 3 else "dilu"
pf = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dims = [int(v) for v in sys.argv[5].split("x")] if len(sys.argv) > 5 else None
s = generators.config(cfg, nx=dims[0], ny=dims[1], nz=dims[2]) if dims else generators.config(cfg, scale=scale); A = s["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": "chunks", "prefetch_slices": pf}})
lib = _lib.lib()
out = (ctypes.c_ulonglong * 16)()
names = ["late stage wait", "look-ahead issue", "ext validate/poll", "compute+stores", "syncwarp+release", "loop overhead", "prologue", "wait for chunk start"]
for what, name in ((4, "lower"), (5, "upper")):
    lib.opmb200_prof_read(out, 1)
    ms, nb = fs.time_kernel(what, 0, 10)
    lib.opmb200_prof_read(out, 1)
    steps, warps = out[8], out[9]
    print(f"{name}: {ms:.3f} ms/launch, {steps} steps over {warps} chunk-walks; cycles per step:")
    for i in range(8):
        print(f"   {names[i]:20s} {out[i] / max(steps, 1):9.1f}")
    print(f"   {'total':20s} {sum(out[i] for i in range(8)) / max(steps, 1):9.1f}")
    print(f"   poll-loop iterations per step (t>0): {out[10] / max(steps, 1):.3f}; at chunk start, per chunk: {out[11] / max(warps, 1):.1f}")
fs.close()
