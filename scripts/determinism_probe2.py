"""Bisecting the rare wrong preconditioner application of the tile walkers (scripts/determinism_probe.py found 2 of 80
applications on C3 differing from the first by up to 3e-3): per variant (runtime options; OPMB200_LIB selects a
differently compiled library), N applications of the same right-hand side compared bitwise with the level schedule's
result; for the wrong ones: how many entries, which grid cells.   python scripts/determinism_probe2.py N variant..."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
VARIANTS = {
    "default": {},
    "prefetch0": {"prefetch_steps": 0},
    "poll1": {"poll_warps": 1},
    "rhs1": {"rhs_warps": 1},
    "poll1rhs1": {"poll_warps": 1, "rhs_warps": 1},
    "cps2": {"ctas_per_sm": 2},
    "tile10x4": {"chunk_rows": -1004},
    "tile4x8": {"chunk_rows": -408},
}
names = sys.argv[2:] or ["default"]
cfg = generators.CONFIGS["C3"]
s = generators.config("C3", scale=1.0, with_rhs=False)
A = s["A"]
nx, ny = cfg["nx"], cfg["ny"]
d = torch.from_numpy(np.random.default_rng(3).standard_normal(A.n * A.b)).cuda()
for prec in os.environ.get("PROBE_PRECS", "dilu").split(","):
    ref = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": "levels"}})
    v_ref = torch.empty_like(d)
    ref.preconditioner().apply(v_ref, d)
    ref.close()
    for name in names:
        fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": dict(VARIANTS[name], schedule="tiles")})
        bad = []
        v = torch.empty_like(d)
        for k in range(N):
            v.fill_(float("nan"))
            fs.preconditioner().apply(v, d)
            if not torch.equal(v, v_ref):
                diff = (v != v_ref) | torch.isnan(v)
                rows = torch.unique(torch.nonzero(diff).flatten() // A.b).cpu().numpy()
                cells = [(int(r % nx), int(r // nx % ny), int(r // (nx * ny))) for r in rows[:6]]
                bad.append({"apply": k, "rows_wrong": int(len(rows)), "first_cells_ijk": cells,
                            "last_cell_ijk": (int(rows[-1] % nx), int(rows[-1] // nx % ny), int(rows[-1] // (nx * ny))),
                            "max_rel": float(((v - v_ref).abs().max() / v_ref.abs().max()).cpu())})
        print(json.dumps({"lib": os.environ.get("OPMB200_LIB", "default"), "prec": prec, "variant": name, "applies": N,
                          "wrong": len(bad), "events": bad[:6]}), flush=True)
        fs.close()
