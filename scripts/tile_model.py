"""Separates the step time and the hop cost of the tile walkers with box grids:
one tile with a long line (steady-state step time), tiles side by side in j, tiles stacked in k, the full C3 box.
Timing experiments (b200.debug_timing, OPMB200_TWDBG builds) switch parts of the kernel off.
   python scripts/tile_model.py [variants]     variants: comma list of  name:poll_warps:prefetch:debug[:load_warps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

variants = [v.split(":") for v in (sys.argv[1] if len(sys.argv) > 1 else "base:3:12:0").split(",")]
grids = [(600, 10, 4), (60, 220, 4), (60, 10, 85), (60, 220, 85)]
if len(sys.argv) > 2:
    grids = [tuple(int(x) for x in g.split("x")) for g in sys.argv[2].split(",")]
b = int(os.environ.get("B", "3"))
for dims in grids:
    A = generators.blackoil_system(*dims, b=b, seed=5, with_rhs=False)["A"]
    for name, pw, pf, dbg, *rest in variants:
        t0 = time.time()
        fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": os.environ.get("PREC", "dilu")},
                                              "b200": {"schedule": "tiles", "poll_warps": int(pw), "prefetch_steps": int(pf),
                                                       "debug_timing": int(dbg), "ctas_per_sm": int(os.environ.get("CPS", "1")),
                                                       "chunk_rows": int(os.environ.get("TILE", "0"))}})
        info = fs.info()
        t1 = time.time()
        lo, _ = fs.time_kernel(4, 3, 10)
        up, _ = fs.time_kernel(5, 3, 10)
        # critical path in steps: line length + skew across the whole grid
        path = dims[0] + dims[1] + dims[2] - 2
        print(f"{dims} {name:10s} chunks {info['n_chunks']:4d} tile {info['chunk_rows']} path {path:4d} steps: "
              f"lower {lo*1e3:8.1f} us ({lo*1e6*1.965/path:6.0f} cycles/path step)  upper {up*1e3:8.1f} us ({up*1e6*1.965/path:6.0f})"
              f"  [create {t1-t0:.1f}s timing {time.time()-t1:.1f}s]", flush=True)
        fs.close()
