"""Step cost and hop cost of the tile walkers from three degenerate box grids (no in-kernel instrumentation):
   one tile (nx x 8 x 4: nx + 10 free-running steps), a chain of tiles along j (60 x 8m x 4) and along k (60 x 8 x 4m).
   T = steps_on_path * c + hops * H.    python scripts/tile_chain.py [dbg,dbg,...]   (DBG bits need an OPMB200_TWDBG build)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

dbgs = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
prec = os.environ.get("PREC", "dilu")
MHZ = float(os.environ.get("MHZ", "1920"))
extra = {k: int(v) for k, v in (kv.split("=") for kv in os.environ.get("OPTS", "").split(",") if kv)}
cases = [("one tile", 4000, 8, 4), ("one tile short", 1000, 8, 4), ("j chain", 60, 216, 4), ("k chain", 60, 8, 84), ("C3", 60, 220, 85)]
only = os.environ.get("CASES")
for name, nx, ny, nz in cases:
    if only and not any(name.startswith(o) for o in only.split(",")):
        continue
    A = generators.blackoil_system(nx, ny, nz, b=3, seed=5, with_rhs=False)["A"]
    for dbg in dbgs:
        fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec},
                                              "b200": {"schedule": "tiles", "chunk_rows": -804, "debug_timing": dbg, **extra}})
        out = []
        for what in (4, 5):
            ms, _ = fs.time_kernel(what, 3, 10)
            out.append(ms)
        ntj, ntk = -(-ny // 8), -(-nz // 4)
        steps = nx + ny + nz - 2
        hops = ntj - 1 + ntk - 1
        cyc = [ms * 1e-3 * MHZ * 1e6 for ms in out]
        print(f"{name:15s} {nx}x{ny}x{nz} dbg {dbg:3d}: lower {out[0]*1e3:8.1f} us upper {out[1]*1e3:8.1f} us; path {steps} steps + {hops} hops; "
              f"cycles/step if hops were free: {cyc[0]/steps:6.0f} / {cyc[1]/steps:6.0f}", flush=True)
        fs.close()
