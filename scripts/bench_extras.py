"""Timing of the operators either side of the smoother on one B200, live CUDA
events on the library's stream through opmb200_time_kernel:

  * the operator with standard wells kept outside the matrix, y = (A - C^T D^-1 B) x (well kernel + SpMV with the
    perforated rows corrected in place), beside the plain SpMV;
  * the CPR transfer pieces: quasi-IMPES weights, coarse (pressure) matrix entries, restriction + prolongation;
  * one BiCGSTAB+DILU solve with and without the wells, checked against its own true residual.

   python scripts/bench_extras.py [C3] [scale]  ->  one JSON line (kept under profiles/)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, WellModelMatrixAdapter  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
s = generators.config(cfg, scale=scale)
A = s["A"]
op = WellModelMatrixAdapter(A, None)
fs = FlexibleSolver(op, {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": "dilu"}})
out = {"config": cfg, "scale": scale, "rows": A.n, "nnzb": A.nnzb, "b": A.b, "peak_GBps": peak, "kernels": {}}


def timed(name, what):
    ms, nbytes = fs.time_kernel(what, 3, 20)
    out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1),
                            "frac_of_peak": round(nbytes / ms / 1e6 / peak, 3)}


def solve(tag):
    x, r = np.zeros(A.n * A.b), s["rhs"].copy()
    res = fs.apply(x, r)
    y = np.zeros_like(x)
    op.apply(x, y)  # the operator the solve used (with or without the wells)
    out[tag] = {"iterations": res.iterations, "converged": bool(res.converged), "device_ms": round(fs.info()["t_solve_ms"], 3),
                "true_reduction": float(np.linalg.norm(s["rhs"] - y) / np.linalg.norm(s["rhs"]))}


timed("spmv (no wells)", 0)
solve("solve_without_wells")
# marginal cost of one BiCGSTAB iteration: two solves of different length, so that the fixed part of a solve (staging,
# initial defect, read-back) cancels -- bench.py's iteration_ms divides the whole solve by its iteration count
marg = []
for tol in (1e-2, 1e-5):
    x, r = np.zeros(A.n * A.b), s["rhs2"].copy()
    fs.apply(x, r, tol)
    marg.append(((len(fs.history()) - 1) / 2, fs.info()["t_solve_ms"]))  # executed iterations = half steps / 2
(i0, t0), (i1, t1) = marg
b_ = A.b
b_iter = 2 * (A.nnzb * (8 * b_ * b_ + 4) + 4 * (A.n + 1) + 16 * b_ * A.n) \
    + 2 * ((A.nnzb - A.n) * (8 * b_ * b_ + 4) + 16 * b_ * b_ * A.n + 40 * b_ * A.n + 16 * (A.n + 1) + 8 * A.n) + 19 * 8 * b_ * A.n
if i1 > i0:
    m = (t1 - t0) / (i1 - i0)
    out["marginal_iteration"] = {"solves": marg, "ms": round(m, 4), "fixed_ms_per_solve": round(t0 - i0 * m, 4),
                                 "algorithmic_MB": round(b_iter / 1e6, 1), "roofline_frac": round(b_iter / 1e6 / peak / m, 4)}
# Norne has 36 wells, a full-field model a few hundred: 200 wells x 40 perforations, 4 well equations (black oil)
wells = generators.standard_wells(A, n_wells=200, perfs=40, seed=5, strength=0.05)
op.set_wells(wells)
out["wells"] = {"n_wells": len(wells["ptr"]) - 1, "n_perforations": int(len(wells["cells"])), "dim_wells": int(wells["Dinv"].shape[-1])}
timed("well kernel + spmv (A - C^T D^-1 B)", 0)
solve("solve_with_wells")
op.set_wells(None)
timed("cpr quasi-IMPES weights", 7)
timed("cpr coarse entries", 8)
timed("cpr restrict + prolongate", 9)
fs.close()
print(json.dumps(out))
