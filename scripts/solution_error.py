"""Achieved parity of the SOLUTION per BASELINE config: relative error of x and of the
iteration count against the oracle at several reductions, for the library named by OPMB200_LIB
(default build, or one built with OPMB200_EXTRA_FLAGS="-fmad=false").  The oracle is built with
-ffp-contract=off, nvcc contracts a*b+c into an FMA by default; the dot products are tree sums here and
running sums there.  Prints one JSON line per case.

  python scripts/solution_error.py                     # default library
  OPMB200_LIB=/root/repo/opm_simulators_b200/libopmb200_nofma.so OPMB200_EXTRA_FLAGS=-fmad=false \
      python opm_simulators_b200/build.py && OPMB200_LIB=... python scripts/solution_error.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [("C2", 1.0), ("C3", 0.25), ("C3", 1.0), ("C5", 0.12)]
if len(sys.argv) > 1:
    CASES = [(a.split(":")[0], float(a.split(":")[1])) for a in sys.argv[1:]]
tag = os.path.basename(os.environ.get("OPMB200_LIB", "libopmb200.so"))
for cfg, scale in CASES:
    s = generators.config(cfg, scale=scale)
    A = s["A"]
    for prec in ("dilu", "ilu0"):
        ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)
        ps.prec_update(prec)
        for tol in (1e-2, 1e-6, 1e-10):
            fs = FlexibleSolver(MatrixAdapter(A), {"tol": tol, "maxiter": 400, "preconditioner": {"type": prec}})
            d = np.random.default_rng(3).standard_normal(A.n * A.b)
            v = np.zeros_like(d)
            fs.preconditioner().apply(v, d)
            e_prec = float(np.linalg.norm(v - ps.prec_apply([d])[0]) / np.linalg.norm(v))
            x, r = np.zeros(A.n * A.b), s["rhs2"].copy()
            res = fs.apply(x, r)
            xo, _, ro, ho = ps.bicgstab([s["rhs2"]], tol=tol, maxiter=400)
            h = fs.history()
            m = min(len(h), len(ho))
            print(json.dumps({"lib": tag, "config": cfg, "scale": scale, "n": A.n, "b": A.b, "prec": prec, "tol": tol,
                              "it_gpu": res.iterations, "it_oracle": ro["iterations"],
                              "half_steps": [len(h) - 1, len(ho) - 1],
                              "x_rel_err": float(np.linalg.norm(x - xo[0]) / np.linalg.norm(xo[0])),
                              "prec_apply_rel_err": e_prec,
                              "history_max_rel_dev": float(np.max(np.abs(h[:m] - ho[:m]) / ho[:m]))}), flush=True)
            fs.close()
