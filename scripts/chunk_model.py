"""Separates the step time, the k-hop and the j-hop of the chunk sweeps with four box grids:
one chunk, chunks stacked in k, chunks side by side in j, and the full C3 box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
prec = sys.argv[1] if len(sys.argv) > 1 else "dilu"
pf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cases = [(60, 32, 1), (60, 32, 85), (60, 220, 1), (60, 220, 85)] if len(sys.argv) <= 3 else [tuple(int(v) for v in sys.argv[3].split("x"))]
for dims in cases:
    s = generators.config("C3", nx=dims[0], ny=dims[1], nz=dims[2]); A = s["A"]
    for sched in ("levels", "chunks"):
        fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": sched, "prefetch_slices": pf, "debug_timing": int(os.environ.get("DBG", "0"))}})
        info = fs.info()
        lo, _ = fs.time_kernel(4, 3, 20); up, _ = fs.time_kernel(5, 3, 20)
        print(f"{dims} {sched:7s} levels/groups {info['n_levels']:5d} slices {info['n_slices']:6d} chunks {info['n_chunks']:4d}  lower {lo*1e3:8.1f} us  upper {up*1e3:8.1f} us", flush=True)
        fs.close()
