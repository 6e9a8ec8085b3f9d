"""Replays a linear system dumped by a Flow build through libopmb200:

  python scripts/replay_system.py --matrix J.mm --rhs r.mm [--block 3]          # Dune storeMatrixMarket dumps
  python scripts/replay_system.py --export-dir export/ --block 3                 # exportSystem.hpp raw binaries
  ... [--options opts.json | --prec dilu|ilu0 --tol 1e-2 --maxiter 200] [--schedule auto|levels|tiles] [--check] [--reps 3]

--check solves the same system with the CPU oracle (the restated Dune path) and reports the differences."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_simulators_b200 import matrixmarket
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter

ap = argparse.ArgumentParser()
ap.add_argument("--matrix"); ap.add_argument("--rhs"); ap.add_argument("--export-dir")
ap.add_argument("--block", type=int, default=None)
ap.add_argument("--options"); ap.add_argument("--prec", default="dilu"); ap.add_argument("--tol", type=float, default=1e-2)
ap.add_argument("--maxiter", type=int, default=200); ap.add_argument("--schedule", default="auto")
ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
if a.export_dir:
    A, rhs = matrixmarket.import_system(a.export_dir, a.block or 3)
else:
    A = matrixmarket.read_matrix(a.matrix, a.block)
    rhs = matrixmarket.read_vector(a.rhs)
opts = json.load(open(a.options)) if a.options else {"solver": "bicgstab", "tol": a.tol, "maxiter": a.maxiter,
                                                      "preconditioner": {"type": a.prec, "relaxation": 1.0}}
opts.setdefault("b200", {})["schedule"] = a.schedule
print(f"system: {A.n} block rows, {A.nnzb} blocks of {A.b}x{A.b}, |rhs| = {np.linalg.norm(rhs):.6e}")
t = time.perf_counter(); fs = FlexibleSolver(MatrixAdapter(A), opts); info = fs.info()
print(f"create + first factorisation {time.perf_counter() - t:.2f} s (analysis {info['t_analysis_s']:.2f} s), "
      f"{info['n_levels']} level sets, structurally symmetric: {bool(info['structurally_symmetric'])}")
for rep in range(a.reps):
    fs.update(A.val)
    x, r = np.zeros_like(rhs), rhs.copy()
    res = fs.apply(x, r)
    i = fs.info()
    print(f"solve {rep}: iterations {res.iterations} reduction {res.reduction:.3e} converged {res.converged} | "
          f"update {i['t_update_ms']:.3f} ms solve {i['t_solve_ms']:.3f} ms (device, host buffers)")
y = np.zeros_like(rhs); fs.op.apply(x, y)
print(f"true relative residual |b - A x| / |b| = {np.linalg.norm(rhs - y) / np.linalg.norm(rhs):.3e}")
if a.check:
    from oracle import oracle as orc
    po = opts.get("preconditioner", {})
    xo, ro, _ = orc.solve_serial(A.rowptr, A.col, A.val, rhs, prec=str(po.get("type", "ilu0")).lower().replace("paroverilu0", "ilu0"),
                                 tol=float(opts.get("tol", 1e-2)), maxiter=int(opts.get("maxiter", 200)),
                                 relaxation=float(po.get("relaxation", 1.0)))
    err = np.linalg.norm(x - xo) / max(np.linalg.norm(xo), 1e-300)
    print(f"oracle: iterations {ro['iterations']} reduction {ro['reduction']:.3e}; |x - x_oracle| / |x_oracle| = {err:.3e}")
    print(json.dumps({"replay_check": {"n": A.n, "b": A.b, "iterations": res.iterations, "oracle_iterations": ro["iterations"],
                                       "x_rel_err": float(err), "padded_blocks_ratio": info["padded_blocks"] / info["nnzb"],
                                       "schedule": info["schedule"]}}))
fs.close()
