"""Is every stage run-to-run bit-reproducible at full size?  Repeats the preconditioner application, the SpMV and the
scalar product on fixed inputs and long solves (tol 1e-6, ~50 iterations), per sweep schedule, and reports how many
repetitions differ bitwise from the first.   python scripts/determinism_probe.py [C3] [dilu]"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
prec = sys.argv[2] if len(sys.argv) > 2 else "dilu"
s = generators.config(cfg, scale=1.0)
A = s["A"]
rhs_d = torch.from_numpy(s["rhs2"]).cuda()
d = torch.from_numpy(np.random.default_rng(3).standard_normal(A.n * A.b)).cuda()
for schedule in ("levels", "tiles"):
    for graph in (1, 0):
        fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-6, "maxiter": 200, "preconditioner": {"type": prec},
                                              "b200": {"schedule": schedule, "cuda_graph": graph}})
        out = {"schedule": schedule, "cuda_graph": graph}
        if graph:
            v0 = torch.empty_like(d)
            fs.preconditioner().apply(v0, d)
            bad, worst = 0, 0.0
            for _ in range(80):
                v = torch.empty_like(d)
                fs.preconditioner().apply(v, d)
                if not torch.equal(v, v0):
                    bad += 1
                    worst = max(worst, float((v - v0).abs().max() / v0.abs().max()))
            out["prec_apply_80"] = {"differ": bad, "worst_rel": worst}
            y0 = torch.empty_like(d)
            fs.op.apply(d, y0)
            bad = 0
            for _ in range(40):
                y = torch.empty_like(d)
                fs.op.apply(d, y)
                bad += int(not torch.equal(y, y0))
            out["spmv_40_differ"] = bad
            out["dot_40_distinct"] = len({fs.dot(d, v0) for _ in range(40)})
        sol = []
        for _ in range(5):
            x, r = torch.zeros_like(rhs_d), rhs_d.clone()
            torch.cuda.synchronize()
            res = fs.apply(x, r)
            h = fs.history()
            sol.append((res.iterations, len(h), hashlib.sha1(x.cpu().numpy().tobytes()).hexdigest()[:10],
                        hashlib.sha1(h[:21].tobytes()).hexdigest()[:8], hashlib.sha1(h[:41].tobytes()).hexdigest()[:8]))
        out["solves"] = sol
        print(json.dumps(out), flush=True)
        fs.close()
