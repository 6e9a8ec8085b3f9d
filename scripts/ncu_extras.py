"""ncu target for the kernels of the well operator and the CPR pieces: one launch of each on C3.
   ncu --set full --clock-control none --import-source on -k regex:"well_z|spmv_kernel|cpr_" -c 8 -o gpurun_out/r02_ncu_extras \
       python scripts/ncu_extras.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from opm_simulators_b200 import generators  # noqa: E402
from opm_simulators_b200.flexible_solver import FlexibleSolver, PressureTransferPolicy, WellModelMatrixAdapter  # noqa: E402

A = generators.config(sys.argv[1] if len(sys.argv) > 1 else "C3", scale=1.0, with_rhs=False)["A"]
wells = generators.standard_wells(A, n_wells=200, perfs=40, seed=5, strength=0.05)
op = WellModelMatrixAdapter(A, wells)
fs = FlexibleSolver(op, {"preconditioner": {"type": "dilu"}})
x = np.random.default_rng(0).standard_normal(A.n * A.b)
y = np.zeros_like(x)
op.apply(x, y)  # well_z_kernel + spmv_kernel<.., WELLS>
pol = PressureTransferPolicy(fs, 0, False)
pol.quasi_impes_weights()
pol.calculateCoarseEntries()
c = pol.moveToCoarseLevel(x)
pol.moveToFineLevel(c, y)
fs.close()
print("ncu_extras done")
