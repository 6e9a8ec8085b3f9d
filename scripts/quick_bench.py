import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sys, time, json
import numpy as np
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t=time.time(); s = generators.config(cfg, scale=scale); A = s["A"]; print("gen", time.time()-t, A.n, A.nnzb, flush=True)
sched = sys.argv[3] if len(sys.argv) > 3 else "levels"
variants = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
for prec in (sys.argv[5].split(",") if len(sys.argv) > 5 else ("dilu", "ilu0")):
    for thr in variants:
        t=time.time()
        fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": prec}, "b200": ({"schedule": "levels", "throttle_levels": thr} if sched == "levels" else {"schedule": sched, "chunk_rows": thr, "ctas_per_sm": int(os.environ.get("CPS", "1")), "poll_warps": int(os.environ.get("POLL", "4"))})})
        info = fs.info(); 
        print(prec, sched, "variant", thr, "schedule built", info["schedule"], "create+update %.2fs" % (time.time()-t), "analysis %.2fs" % info["t_analysis_s"], "update_ms %.3f" % info["t_update_ms"], "levels", info["n_levels"], "slices", info["n_slices"], "padded", info["padded_blocks"]/info["nnzb"], "chunks", info["n_chunks"], "chunk_rows", info["chunk_rows"], "est_steps", info["est_steps"], flush=True)
        for what, name in ((0,"spmv"),(1,"prec_apply"),(4,"lower"),(5,"upper"),(2,"prec_update"),(3,"vec3")):
            ms, nb = fs.time_kernel(what, 3, 20)
            print("   %-12s %8.3f ms  %8.1f GB/s (algorithmic %.1f MB)" % (name, ms, nb/ms/1e6, nb/1e6), flush=True)
        for tol in (1e-2, 1e-6):
            x, r = np.zeros(A.n*A.b), s["rhs2"].copy()
            t=time.time()
            try:
                res = fs.apply(x, r, tol)
            except Exception as e:
                res = repr(e)[:60]
            w=time.time()-t
            print("   solve tol", tol, res, "device ms %.3f" % fs.info()["t_solve_ms"], "wall %.3f" % w, flush=True)
        fs.close()
