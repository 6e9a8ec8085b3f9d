"""Full-size (BASELINE.json configs[2], 1.12M cells) parity against the oracle.  The oracle needs
seconds here (SpMV 0.1 s, a factorisation 0.3 s, an apply 0.2 s, the 10-iteration solve 8 s), so the
oversubscribed regime of the real workload (thousands of CTAs on 148 SMs, ticket wrap-around, CTA
retirement order) is compared value by value -- SpMV, Dinv, one DILU and one ILU0 apply to 1e-10, the
solve to iterations +-1 with its defect history -- and not only through size-independent properties.
Dune's BiCGSTAB counting conventions (iterations=(int)it, breakdown thresholds) are restated, not
pinned: "+-1 iteration" is against the restatement."""
import os

import numpy as np
import pytest

from conftest import rel_err
from opm_simulators_b200 import generators
from opm_simulators_b200.flexible_solver import FlexibleSolver, MatrixAdapter
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10
SCHEDULES = [s for s in os.environ.get("OPMB200_FULLSIZE_SCHEDULES", "levels,tiles").split(",") if s]


@pytest.fixture(scope="module")
def c3():
    return generators.config("C3")


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("prec", ["dilu", "ilu0"])
def test_c3_full_size_against_the_oracle(c3, prec, schedule):
    A = c3["A"]
    assert A.n == 1122000 and A.nnzb == 7780000
    fs = FlexibleSolver(MatrixAdapter(A), {"tol": 1e-2, "maxiter": 200, "preconditioner": {"type": prec},
                                          "b200": {"schedule": schedule}})
    info = fs.info()
    assert info["n_levels"] == 60 + 220 + 85 - 2
    rng = np.random.default_rng(0)
    x1 = rng.standard_normal(A.n * 3)
    # SpMV
    y1 = np.zeros_like(x1)
    fs.op.apply(x1, y1)
    assert rel_err(y1, orc.spmv(A.rowptr, A.col, A.val, x1)) < TOL
    # factorisation + one preconditioner application
    v = np.zeros_like(x1)
    fs.preconditioner().apply(v, x1)
    if prec == "dilu":
        Dinv = orc.dilu_update(A.rowptr, A.col, A.val)
        assert rel_err(fs.dinv(), Dinv) < TOL
        vo = orc.dilu_apply(A.rowptr, A.col, A.val, Dinv, x1)
    else:
        lu = orc.ilu0_decompose(A.rowptr, A.col, A.val)
        assert rel_err(fs.ilu0(), lu) < TOL
        vo = orc.ilu0_apply(A.rowptr, A.col, lu, x1)
    assert rel_err(v, vo) < TOL
    # the solve: same stopping step +-1, same defect history, solution at the bar of the small cases
    rhs = c3["rhs2"]
    x, r = np.zeros(A.n * 3), rhs.copy()
    res = fs.apply(x, r)
    xo, ro, ho = orc.solve_serial(A.rowptr, A.col, A.val, rhs, prec=prec, tol=1e-2, maxiter=200)
    assert res.converged and abs(res.iterations - ro["iterations"]) <= 1
    h = fs.history()
    m = min(len(h), len(ho))
    assert m >= 2 and np.allclose(h[:m], ho[:m], rtol=1e-6)
    if res.iterations == ro["iterations"] and len(h) == len(ho):
        assert rel_err(x, xo) < 1e-8
    true = np.linalg.norm(rhs - orc.spmv(A.rowptr, A.col, A.val, x)) / np.linalg.norm(rhs)
    assert true < 1.05e-2 and abs(true - res.reduction) < 1e-6
    # deterministic: a second solve reproduces the first bit for bit
    x2_, r2_ = np.zeros(A.n * 3), rhs.copy()
    res2 = fs.apply(x2_, r2_)
    assert res2.iterations == res.iterations and np.array_equal(x2_, x)
    fs.close()


@pytest.mark.parametrize("prec,applies", [("dilu", 400), ("ilu0", 200)])
def test_tile_walkers_equal_the_level_schedule_on_every_application(c3, prec, applies):
    """Regression test of a rare hand-over race: the right-hand-side warps' parity wait for the NEXT step's record was
    not anchored on the phase before it at the kStages-th step of a tile, and once in ~1e5 tiles the compute warps
    decoded a stale record -- about 1 % of the preconditioner applications on C3 were wrong by 1e-6..3e-3
    (profiles/r02_determinism.md).  Every application must be bit-identical to the level schedule's."""
    torch = pytest.importorskip("torch")
    A = c3["A"]
    d = torch.from_numpy(np.random.default_rng(3).standard_normal(A.n * 3)).cuda()
    ref = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": "levels"}})
    v_ref = torch.empty_like(d)
    ref.preconditioner().apply(v_ref, d)
    ref.close()
    fs = FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": prec}, "b200": {"schedule": "tiles"}})
    assert fs.info()["schedule"] == 1
    v = torch.empty_like(d)
    wrong = 0
    for _ in range(applies):
        v.fill_(float("nan"))
        fs.preconditioner().apply(v, d)
        wrong += int(not torch.equal(v, v_ref))
    fs.close()
    assert wrong == 0, f"{wrong} of {applies} applications differ from the level schedule"
