"""In-kernel profile of the tile walkers.  Build a profiling copy of the library and point OPMB200_LIB at it:
   OPMB200_LIB=$PWD/opm_simulators_b200/libopmb200_prof.so OPMB200_PROFILE=1 python opm_simulators_b200/build.py
   OPMB200_LIB=$PWD/opm_simulators_b200/libopmb200_prof.so python scripts/prof_tiles.py [cfg] [scale] [prec] [prefetch] [chunk_rows]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import _lib, generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
prec = sys.argv[3] if len(sys.argv) > 3 else "dilu"
pf = int(sys.argv[4]) if len(sys.argv) > 4 else 12
cr = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dims = os.environ.get("DIMS")
if dims:
    A = generators.blackoil_system(*[int(x) for x in dims.split("x")], b=3, seed=5, with_rhs=False)["A"]
else:
    A = generators.config(cfg, scale=scale, with_rhs=False)["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec},
                                      "b200": {"schedule": "tiles", "prefetch_steps": pf, "chunk_rows": cr,
                                               "poll_warps": int(os.environ.get("POLL", "4")), "rhs_warps": int(os.environ.get("RHS", "2")),
                                               "debug_timing": int(os.environ.get("DBG", "0"))}})
info = fs.info()
lib = _lib.lib()
out = (C.c_ulonglong * 64)()
names = {0: "compute: wait for the externals, steps >= 8", 6: "compute: wait for the externals, steps < 8", 1: "compute: the step's block",
         2: "compute: named barrier + release", 7: "compute: first step's loads",
         8: "loader: prefetch", 9: "loader: wait for a free stage", 10: "loader: expect_tx + TMA",
         16: "poll w0: wait for the step's record", 17: "poll w0: poll loop + park", 18: "poll w0: signal",
         32: "rhs w0: lists, next rhs loads", 33: "rhs w0: wait for a free stage + next record", 34: "rhs w0: park + signal",
         28: "publisher: wait for the step", 29: "publisher: stores + release"}
for what, name in ((4, "lower"), (5, "upper")):
    fs.time_kernel(what, 2, 3)
    lib.opmb200_prof_read(out, 1)
    reps = 5
    ms, _ = fs.time_kernel(what, 0, reps)
    lib.opmb200_prof_read(out, 1)
    # time_kernel runs lower+upper per repetition: both sweeps accumulate into the same counters
    steps, chunks = out[24], out[25]
    print(f"{name}: {ms:.3f} ms/launch; chunks {info['n_chunks']} chunk_rows {info['chunk_rows']}; counters cover lower+upper "
          f"of {reps} repetitions: {steps} steps, {chunks} chunk walks, poll loads (warp 0) {out[26] / max(steps, 1):.2f} per step")
    print(f"   steps >= 8 that waited: for the right-hand side {out[36] / max(steps, 1):.3f} of all steps, {out[38] / max(out[36], 1):.0f} cycles each; "
          f"for the externals {out[37] / max(steps, 1):.3f}, {out[39] / max(out[37], 1):.0f} cycles each")
    print(f"   hop anatomy: release -> publisher's store {out[44] / max(out[43], 1):.0f} cycles; store -> valid sample (waited-for records only) "
          f"{out[41] / max(out[40], 1):.0f} ns avg, {out[42]} ns max, {out[40] / max(steps, 1):.2f} per step; last arrival -> restart {out[46] / max(out[45], 1):.0f} cycles")
    for i, nm in names.items():
        print(f"   {nm:42s} {out[i] / max(steps, 1):10.1f} cycles per step")
fs.close()
