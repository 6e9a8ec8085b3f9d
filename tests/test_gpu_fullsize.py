"""Full-size (BASELINE.json configs[2], 1.12M cells) checks through size-independent properties:
the oracle is far too slow to run here inside a test, so use linearity, a scipy cross-check of the
SpMV, consistency of the preconditioner with its own factors, and the true residual of the solve."""
import numpy as np
import pytest

from conftest import rel_err
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return generators.config("C3")


@pytest.mark.parametrize("prec", ["dilu", "ilu0"])
def test_c3_full_size_properties(c3, prec):
    A = c3["A"]
    assert A.n == 1122000 and A.nnzb == 7780000
    fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": prec}})
    info = fs.info()
    assert info["n_levels"] == 60 + 220 + 85 - 2
    rng = np.random.default_rng(0)
    S = A.to_scipy()
    x1, x2 = rng.standard_normal(A.n * 3), rng.standard_normal(A.n * 3)
    y1, y2, y3 = np.zeros_like(x1), np.zeros_like(x1), np.zeros_like(x1)
    fs.op.apply(x1, y1)
    assert rel_err(y1, S @ x1) < 1e-12
    # linearity of the preconditioner: M^-1 (2 d1 - 3 d2) == 2 M^-1 d1 - 3 M^-1 d2
    fs.preconditioner().apply(y1, x1)
    fs.preconditioner().apply(y2, x2)
    fs.preconditioner().apply(y3, 2 * x1 - 3 * x2)
    assert rel_err(y3, 2 * y1 - 3 * y2) < 1e-10
    # M^-1 is a good approximate inverse: |A M^-1 d - d| well below |d|
    fs.op.apply(y1, y2)
    assert rel_err(y2, x1) < 0.9
    # the solve reaches the requested reduction on the TRUE residual
    rhs = c3["rhs2"]
    x, r = np.zeros(A.n * 3), rhs.copy()
    res = fs.apply(x, r)
    assert res.converged and 0 < res.iterations <= 200
    true = np.linalg.norm(rhs - S @ x) / np.linalg.norm(rhs)
    assert true < 1.05e-2 and abs(true - res.reduction) < 1e-6
    assert rel_err(r, rhs - S @ x) < 1e-6
    # deterministic: a second solve reproduces the first bit for bit
    x2_, r2_ = np.zeros(A.n * 3), rhs.copy()
    res2 = fs.apply(x2_, r2_)
    assert res2.iterations == res.iterations and np.array_equal(x2_, x)
