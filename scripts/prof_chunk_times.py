"""Per-chunk timeline of one lower sweep (OPMB200_PROFILE build): when each chunk became resident, when
its first step's dependencies had arrived, when it finished.  Prints the lags between neighbouring tiles."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_simulators_b200 import generators, _lib
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
dims = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "60x220x85").split("x")]
what = int(sys.argv[2]) if len(sys.argv) > 2 else 4
s = generators.config("C3", nx=dims[0], ny=dims[1], nz=dims[2]); A = s["A"]
fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": "dilu"}, "b200": {"schedule": "chunks"}})
info = fs.info(); nc = info["n_chunks"]; cr = info["chunk_rows"]
lib = _lib.lib()
ms, _ = fs.time_kernel(what, 2, 1)   # the LAST launch of the timed kind leaves its stamps (lower then upper run in pairs)
buf = (ctypes.c_ulonglong * (3 * nc))()
lib.opmb200_prof_chunks(buf, nc)
t = np.array(buf, dtype=np.int64).reshape(nc, 3).astype(np.float64)
t0 = t[:, 0].min(); t = (t - t0) / 1e3  # us
print(f"{dims} chunks {nc} chunk_rows {cr}; kernel {ms*1e3:.1f} us; span of stamps {t.max():.1f} us")
print("chunk  resident   ready    done   duration(ready->done)")
for c in list(range(0, min(nc, 40))) + list(range(max(40, nc - 8), nc)):
    print(f"{c:5d} {t[c,0]:9.1f} {t[c,1]:8.1f} {t[c,2]:8.1f} {t[c,2]-t[c,1]:8.1f}")
if cr < 0:
    TJ, TK = (-cr) // 100, (-cr) % 100
    ntj = (dims[1] + TJ - 1) // TJ; ntk = (dims[2] + TK - 1) // TK
    if ntj * ntk == nc:
        R = t[:, 1].reshape(ntk, ntj)
        print("mean lag of ready time to the left tile  (j):", np.diff(R, axis=1).mean(), "us; median", np.median(np.diff(R, axis=1)))
        print("mean lag of ready time to the lower tile (k):", np.diff(R, axis=0).mean(), "us; median", np.median(np.diff(R, axis=0)))
        D = (t[:, 2] - t[:, 1]).reshape(ntk, ntj)
        print("chunk duration: mean", D.mean(), "min", D.min(), "max", D.max(), " first chunk", D[0, 0])
fs.close()
