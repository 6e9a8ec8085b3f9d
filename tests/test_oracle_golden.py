"""Pins the CPU oracle (oracle/opm_oracle.c) against every golden vector / known-answer test the
reference's own tests hold for this path, and against the reference's own
dependency-free C solver compiled unmodified into oracle/_ref (opm/simulators/linalg/mixed/*.c).
CPU only."""
import numpy as np
import pytest

from conftest import coo_to_bcsr, pattern_to_bcsr, rel_err
from opm_simulators_b200 import generators
from opm_simulators_b200.bcsr import BCSR
from oracle import oracle as orc


# ---- integer golden vectors ------------------------------------------------------------------
@pytest.mark.parametrize("case", range(4))
def test_graphcoloring_golden(golden, case):
    """tests/test_graphcoloring.cpp:94-397 -- SYMMETRIC / UPPER / LOWER level sets, bit-exact"""
    g = golden["coloring"][case]
    A = pattern_to_bcsr(g["rows"], 1)
    for name, kind in (("SYMMETRIC", orc.COLOR_SYMMETRIC), ("UPPER", orc.COLOR_UPPER), ("LOWER", orc.COLOR_LOWER)):
        assert orc.level_sets(A.rowptr, A.col, kind) == g[name], (g["name"], name)


def test_partition_simple_golden(golden):
    """tests/test_partitionCells.cpp:116-131"""
    for g in golden["partition_simple"]:
        assert orc.partition_simple(g["num_cells"], g["num_domains"]).tolist() == g["part"]


def test_reorder_maps_are_inverse():
    A = generators.laplace_like(6, 1, np.random.default_rng(1), dims=3)
    _, rows, _ = orc.row_coloring(A.rowptr, A.col, orc.COLOR_LOWER)
    r2n, n2r = orc.reorder_maps(rows)
    assert np.array_equal(r2n[n2r], np.arange(A.n)) and np.array_equal(r2n, rows)


def test_levels_of_box_grid_are_hyperplanes():
    """natural ordering of an nx*ny*nz 7-point grid has nx+ny+nz-2 levels = hyperplanes i+j+k"""
    s = generators.blackoil_system(5, 4, 3, b=1, with_rhs=False)
    color, rows, ptr = orc.row_coloring(s["A"].rowptr, s["A"].col, orc.COLOR_LOWER)
    assert len(ptr) - 1 == 5 + 4 + 3 - 2
    i, j, k = np.meshgrid(np.arange(5), np.arange(4), np.arange(3), indexing="ij")
    expect = (i + j + k).transpose(2, 1, 0).reshape(-1)
    assert np.array_equal(color, expect)


# ---- block inverse ----------------------------------------------------------------------------
@pytest.mark.parametrize("b", [1, 2, 3, 4, 5])
def test_invert_block(b):
    rng = np.random.default_rng(b)
    for _ in range(20):
        M = rng.uniform(-1, 1, (b, b)) + 2 * np.eye(b)
        assert rel_err(orc.invert_block(M), np.linalg.inv(M)) < 1e-12


def test_invert_block4_lu_fallback_and_singular():
    """matrixblock.hh:192-224: |det| < 1e-40 -> pivoted LU; exactly singular -> MatrixBlockError"""
    M = np.diag([1e-12, 1e-12, 1e-12, 1e-12]) + 1e-14 * np.arange(16).reshape(4, 4)
    assert abs(np.linalg.det(M)) < 1e-40
    assert rel_err(orc.invert_block(M), np.linalg.inv(M)) < 1e-10
    S = np.ones((4, 4))
    with pytest.raises(orc.OracleError) as e:
        orc.invert_block(S)
    assert e.value.code == 2


# ---- DILU known answers (tests/test_dilu.cpp) -------------------------------------------------
def _dilu_2x2(a01=True, a10=True):
    rows = [[0] + ([1] if a01 else []), ([0] if a10 else []) + [1]]
    A = pattern_to_bcsr(rows, 2)
    A.val[:] = 0
    d = A.diag_index()
    A.val[d[0]] = [[3, 1], [2, 1]]
    A.val[d[1]] = [[-1, 0], [0, -1]]
    if a01:
        A.val[d[0] + 1] = np.eye(2)
    if a10:
        A.val[d[1] - 1] = 2 * np.eye(2)
    return A


@pytest.mark.parametrize("a01,a10", [(True, True), (True, False), (False, True), (False, False)])
def test_dilu_2x2_known_answers(a01, a10):
    """tests/test_dilu.cpp:29-160 (Dinv), :165-470 (apply), :781-850 (block diagonal == ILU)"""
    A = _dilu_2x2(a01, a10)
    D = A.to_dense()
    D00, D11 = D[:2, :2], D[2:, 2:]
    dinv = orc.dilu_update(A.rowptr, A.col, A.val)
    D11_expected = D11 - D[2:, :2] @ np.linalg.inv(D00) @ D[:2, 2:] if (a01 and a10) else D11
    assert rel_err(np.linalg.inv(dinv[0]), D00) < 1e-9
    assert rel_err(np.linalg.inv(dinv[1]), D11_expected) < 1e-9
    # DILU of a 2x2 block matrix is exact (M == A): apply(b) == A^-1 b  (the reference test's
    # new_x = x + M^-1 (b - A x) collapses to this)
    b = np.array([2.0, 1.0, 3.0, 4.0])
    v = orc.dilu_apply(A.rowptr, A.col, A.val, dinv, b)
    assert rel_err(v, np.linalg.solve(D, b)) < 1e-9
    # ILU0 is exact as well
    lu = orc.ilu0_decompose(A.rowptr, A.col, A.val)
    assert rel_err(orc.ilu0_apply(A.rowptr, A.col, lu, b), np.linalg.solve(D, b)) < 1e-9


def _dilu_3x3():
    """tests/test_dilu.cpp:476-595, 599-777"""
    A = pattern_to_bcsr([[0], [1, 2], [1, 2]], 3)
    blocks = {
        (0, 0): [[3, 1, 2], [2, 3, 1], [2, 1, 0]],
        (1, 1): [[1, 0, 1], [4, 1, 0], [3, 1, 3]],
        (1, 2): [[1, 0, 2], [0, 1, 1], [0, 1, 3]],
        (2, 1): [[1, 0, 2], [0, 1, 4], [5, 1, 1]],
        (2, 2): [[1, 3, 2], [2, 1, 3], [3, 1, 2]],
    }
    rows = A.row_of_entry()
    for k in range(A.nnzb):
        A.val[k] = blocks[(int(rows[k]), int(A.col[k]))]
    return A


def test_dilu_3x3_known_answers():
    A = _dilu_3x3()
    D = A.to_dense()
    dinv = orc.dilu_update(A.rowptr, A.col, A.val)
    D11 = D[3:6, 3:6]
    D22 = D[6:, 6:] - D[6:, 3:6] @ np.linalg.inv(D11) @ D[3:6, 6:]
    assert rel_err(np.linalg.inv(dinv[0]), D[:3, :3]) < 1e-9
    assert rel_err(np.linalg.inv(dinv[1]), D11) < 1e-9
    assert rel_err(np.linalg.inv(dinv[2]), D22) < 1e-9
    b = np.array([2.0, 1, 2, 2, 3, 2, 0, 2, 1])
    assert rel_err(orc.dilu_apply(A.rowptr, A.col, A.val, dinv, b), np.linalg.solve(D, b)) < 1e-9


def test_dilu_asymmetric_structure_skips_missing_transpose():
    """DILU.hpp:196-201: the A_ij Dinv_j A_ji term exists only when A_ji is stored"""
    A = pattern_to_bcsr([[0], [0, 1]], 2)  # A_10 present, A_01 absent
    dinv = orc.dilu_update(A.rowptr, A.col, A.val)
    d = A.diag_index()
    assert rel_err(dinv[1], np.linalg.inv(A.val[d[1]])) < 1e-13


@pytest.mark.parametrize("b", [1, 2, 3, 4])
def test_dilu_and_ilu0_against_dense_formulas(b):
    """generic check on a 3-D 7-point pattern against straightforward dense numpy restatements"""
    rng = np.random.default_rng(10 + b)
    A = generators.laplace_like(4, b, rng, dims=3, asym=0.3)
    n, D = A.n, A.to_dense()
    blk = lambda i, j: D[i * b:(i + 1) * b, j * b:(j + 1) * b]  # noqa: E731
    pat = {(int(i), int(j)) for i, j in zip(A.row_of_entry(), A.col)}
    # DILU
    dinv = orc.dilu_update(A.rowptr, A.col, A.val)
    ref = []
    for i in range(n):
        acc = blk(i, i).copy()
        for j in range(i):
            if (i, j) in pat and (j, i) in pat:
                acc -= blk(i, j) @ ref[j] @ blk(j, i)
        ref.append(np.linalg.inv(acc))
    assert rel_err(dinv, np.array(ref)) < 1e-12
    d = rng.standard_normal(n * b)
    Dinv = np.zeros_like(D)
    for i in range(n):
        Dinv[i * b:(i + 1) * b, i * b:(i + 1) * b] = ref[i]
    Dd = np.linalg.inv(Dinv)
    L = np.tril(D, -1)
    U = np.triu(D, 1)
    for i in range(n):  # strict block triangles
        L[i * b:(i + 1) * b, i * b:(i + 1) * b] = 0
        U[i * b:(i + 1) * b, i * b:(i + 1) * b] = 0
    Mdilu = (Dd + L) @ Dinv @ (Dd + U)
    assert rel_err(orc.dilu_apply(A.rowptr, A.col, A.val, dinv, d), np.linalg.solve(Mdilu, d)) < 1e-11
    # ILU0: L U agrees with A on the pattern
    lu = orc.ilu0_decompose(A.rowptr, A.col, A.val)
    Lf, Uf = np.eye(n * b), np.zeros((n * b, n * b))
    rows = A.row_of_entry()
    for k in range(A.nnzb):
        i, j = int(rows[k]), int(A.col[k])
        if j < i:
            Lf[i * b:(i + 1) * b, j * b:(j + 1) * b] = lu[k]
        elif j == i:
            Uf[i * b:(i + 1) * b, j * b:(j + 1) * b] = np.linalg.inv(lu[k])
        else:
            Uf[i * b:(i + 1) * b, j * b:(j + 1) * b] = lu[k]
    P = Lf @ Uf
    for (i, j) in pat:
        assert np.allclose(P[i * b:(i + 1) * b, j * b:(j + 1) * b], blk(i, j), rtol=1e-10, atol=1e-12)
    v = orc.ilu0_apply(A.rowptr, A.col, lu, d)
    assert rel_err(v, np.linalg.solve(P, d)) < 1e-11
    assert rel_err(orc.ilu0_apply(A.rowptr, A.col, lu, d, relaxation=0.9), 0.9 * v) < 1e-15


# ---- whole-solve golden vectors ---------------------------------------------------------------
@pytest.mark.parametrize("b", [1, 3])
@pytest.mark.parametrize("prec", ["ilu0", "dilu"])
def test_matr33_golden_solution(golden, b, prec):
    """tests/test_flexiblesolver.cpp:83-130, tests/test_preconditionerfactory.cpp:116-149"""
    A = coo_to_bcsr(golden["matr33"], b)
    opts = golden["options_flexiblesolver_1x1"]
    x, res, _ = orc.solve_serial(A.rowptr, A.col, A.val, golden["rhs3"], prec=prec,
                                 tol=float(opts["tol"]), maxiter=int(opts["maxiter"]))
    tol = golden["matr33_solution_tol_percent"] / 100
    if b == 3 or prec == "ilu0":
        # block-tridiagonal: ILU0 (and block DILU) are exact -> converged at the first half step
        assert res["it"] == 0.5 and res["iterations"] == 0 and res["converged"]
        assert np.allclose(x, golden["matr33_solution"], rtol=tol)
    else:
        assert res["converged"]


def test_matr33rep_unpreconditioned_bicgstab_golden(golden):
    """tests/test_preconditionerfactory.cpp:231-376: plain Dune::BiCGSTABSolver on A*A, the one
    stored result that exercises several BiCGSTAB iterations"""
    opts = golden["options_flexiblesolver_simple"]
    for b in (1, 3):
        A = coo_to_bcsr(golden["matr33rep"], b)
        x, res, hist = orc.solve_serial(A.rowptr, A.col, A.val, golden["rhs3rep"], prec="nothing",
                                        tol=float(opts["tol"]), maxiter=int(opts["maxiter"]),
                                        op_repeats=golden["matr33rep_repeats"])
        assert res["converged"]
        assert np.allclose(x, golden["matr33rep_solution"], rtol=golden["matr33rep_solution_tol_percent"] / 100)


def test_solver_adapter_tridiagonal():
    """tests/gpuistl/test_solver_adapter.cpp:88-117: 10 block rows of (1,-2,1)*I3, b = A*1,
    x0_i = 0.1*i -> x = 1"""
    n, b = 10, 3
    rows = [[j for j in (i - 1, i, i + 1) if 0 <= j < n] for i in range(n)]
    A = pattern_to_bcsr(rows, b)
    r = A.row_of_entry()
    A.val[:] = np.where((r == A.col)[:, None, None], -2.0, 1.0) * np.eye(b)
    rhs = orc.spmv(A.rowptr, A.col, A.val, np.ones(n * b))
    x0 = np.repeat(0.1 * np.arange(n), b)
    x, res, _ = orc.solve_serial(A.rowptr, A.col, A.val, rhs, prec="ilu0", tol=1e-12, maxiter=200, x0=x0)
    assert np.allclose(x, 1.0, rtol=1e-11)


# ---- against the reference's own compiled C solver -------------------------------------------
needs_ref = pytest.mark.skipif(not orc.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("use_dilu", [True, False])
def test_against_reference_mixed_solver(use_dilu):
    """opm/simulators/linalg/mixed/{bsr,prec,bslv}.c compiled unmodified: same SpMV, same
    M^-1 d (DILU and ILU0 factor + apply), same converged solution."""
    s = generators.blackoil_system(7, 6, 5, b=3, seed=77, sigma=1.0)
    A, rhs = s["A"], s["rhs"]
    ref = orc.RefMixedSolver(A.rowptr, A.col, A.val, tol=1e-10, maxiter=200, use_dilu=use_dilu)
    x = np.random.default_rng(3).standard_normal(A.n * 3)
    assert rel_err(orc.spmv(A.rowptr, A.col, A.val, x), ref.spmv(x)) < 1e-13
    ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)
    ps.prec_update("dilu" if use_dilu else "ilu0")
    assert rel_err(ps.prec_apply([x])[0], ref.factor_apply(x, use_dilu)) < 1e-10
    xr, count, red = ref.solve(rhs)
    xo, res, _ = orc.solve_serial(A.rowptr, A.col, A.val, rhs, prec="dilu" if use_dilu else "ilu0",
                                  tol=1e-10, maxiter=200)
    assert res["converged"] and red < 1e-10
    assert rel_err(xo, xr) < 1e-7 and rel_err(xo, s["xstar"]) < 1e-7
    assert abs(res["iterations"] - count) <= 3  # different (left/right) recurrences, same method


# ---- parallel semantics -------------------------------------------------------------------------
def test_make_overlap_rows_invalid():
    A = generators.laplace_like(4, 2, np.random.default_rng(5), dims=2)
    val = orc.make_overlap_rows_invalid(A.rowptr, A.col, A.val, 10)
    rows = A.row_of_entry()
    for k in range(A.nnzb):
        if rows[k] >= 10:
            assert np.array_equal(val[k], np.eye(2) if A.col[k] == rows[k] else np.zeros((2, 2)))
        else:
            assert np.array_equal(val[k], A.val[k])


def test_ghost_last_spmv():
    """WellOperators.hpp:432-468: interior rows only, ghost rows of y zeroed"""
    A = generators.laplace_like(4, 3, np.random.default_rng(6), dims=2)
    x = np.random.default_rng(7).standard_normal(A.n * 3)
    y = orc.spmv(A.rowptr, A.col, A.val, x, interior=11)
    full = A.to_dense() @ x
    assert rel_err(y[:33], full[:33]) < 1e-14 and np.all(y[33:] == 0)
    y2 = orc.spmv_scaleadd(A.rowptr, A.col, A.val, -1.0, x, full, interior=11)
    assert np.abs(y2[:33]).max() < 1e-12 and np.all(y2[33:] == 0)
