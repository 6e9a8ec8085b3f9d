"""What does one BiCGSTAB iteration cost, and what is fixed per solve?  Solves of different length (tolerances) on both
right-hand sides of a configuration, device-resident vectors, every solve's device time listed (not only the mean).
   python scripts/iter_cost.py [C3] [dilu]"""
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
fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": prec}})
vals_d = torch.from_numpy(A.val).cuda()
out = {"config": cfg, "prec": prec, "runs": []}
for rhs_name in ("rhs2", "rhs"):
    rhs_d = torch.from_numpy(s[rhs_name]).cuda()
    x_d, r_d = torch.zeros_like(rhs_d), torch.empty_like(rhs_d)
    for tol in (1e-2, 1e-4, 1e-6):
        for with_update in (0, 1):
            times, its, launches = [], [], []
            for rep in range(6):
                if with_update:
                    fs.update(vals_d)
                x_d.zero_()
                r_d.copy_(rhs_d)
                torch.cuda.synchronize()
                l0 = fs.info()["kernel_launches"]
                res = fs.apply(x_d, r_d, tol)
                i_ = fs.info()
                times.append(round(i_["t_solve_ms"], 3))
                its.append(res.iterations)
                launches.append(i_["kernel_launches"] - l0)
            out["runs"].append({"rhs": rhs_name, "tol": tol, "update_before_each_solve": with_update, "iterations": its[-1],
                                "solve_ms": times, "launches": launches[-1], "history_len": len(fs.history())})
            print(out["runs"][-1], flush=True)
fs.close()
print(json.dumps(out))
