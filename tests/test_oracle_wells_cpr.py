"""CPU checks of the oracle's well operator (WellOperators.hpp:84-109,144-164; StandardWellEquations.cpp:132-148) and of
the numpy restatements of the CPR transfer pieces (getQuasiImpesWeights.hpp:64-111, PressureTransferPolicy.hpp,
gpuistl/detail/cpr_amg_operations.cu:35-178) against dense linear algebra.  The reference holds no stored numbers for
these (tests/gpuistl/test_GpuPressureTransferPolicy.cpp compares two live implementations), so the dense formulas are
the anchor: "parity unpinned" beyond them."""
import numpy as np
import pytest

from conftest import rel_err
from opm_simulators_b200 import generators
from opm_simulators_b200.bcsr import BCSR
from oracle import oracle as orc


def dense_well_matrix(wells, n, b):
    """sum_w C_w^T D_w^-1 B_w as a dense (n*b) x (n*b) matrix"""
    W = np.zeros((n * b, n * b))
    ptr, cells = wells["ptr"], wells["cells"]
    for w in range(len(ptr) - 1):
        dw = wells["Dinv"].shape[-1]
        Bw = np.zeros((dw, n * b))
        Cw = np.zeros((dw, n * b))
        for p in range(ptr[w], ptr[w + 1]):
            c = cells[p]
            Bw[:, c * b:(c + 1) * b] += wells["B"][p]
            Cw[:, c * b:(c + 1) * b] += wells["C"][p]
        W += Cw.T @ wells["Dinv"][w] @ Bw
    return W


@pytest.mark.parametrize("b,dw,shared", [(3, 4, 0), (3, 3, 2), (4, 5, 1), (2, 2, 0), (1, 2, 0)])
def test_well_apply_equals_dense_formula(b, dw, shared):
    A = generators.blackoil_system(6, 5, 4, b=b, seed=3, with_rhs=False)["A"] if b >= 3 else \
        generators.laplace_like(8, b, np.random.default_rng(2), dims=2)
    wells = generators.standard_wells(A, n_wells=5, perfs=6, dim_wells=dw, seed=11, shared_cells=shared)
    rng = np.random.default_rng(0)
    x, y0 = rng.standard_normal(A.n * b), rng.standard_normal(A.n * b)
    y = orc.well_apply(wells, x, y0, b)
    assert rel_err(y, y0 - dense_well_matrix(wells, A.n, b) @ x) < 1e-13
    if shared:  # two wells perforate the same cell
        assert len(np.unique(wells["cells"])) < len(wells["cells"])


def test_bicgstab_with_wells_solves_the_combined_operator():
    """WellModelMatrixAdapter: the Krylov operator is A - C^T D^-1 B, the preconditioner is A's"""
    A = generators.blackoil_system(6, 5, 4, b=3, seed=5, with_rhs=False)["A"]
    wells = generators.standard_wells(A, n_wells=4, perfs=5, seed=1)
    M = A.to_dense() - dense_well_matrix(wells, A.n, 3)
    rng = np.random.default_rng(1)
    xstar = rng.uniform(-1, 1, A.n * 3)
    rhs = M @ xstar
    for prec in ("dilu", "ilu0"):
        ps = orc.ParSystem.serial(A.rowptr, A.col, A.val)
        ps.set_wells(wells)
        ps.prec_update(prec)
        x, r, res, hist = ps.bicgstab([rhs], tol=1e-12, maxiter=200)
        assert res["converged"] and rel_err(x[0], xstar) < 1e-9
        # b is overwritten with the (recursively updated) residual of the combined operator
        assert np.linalg.norm(r[0] - (rhs - M @ x[0])) < 1e-9 * np.linalg.norm(rhs)
        ps0 = orc.ParSystem.serial(A.rowptr, A.col, A.val)  # without wells the same rhs gives another solution
        ps0.prec_update(prec)
        x0 = ps0.bicgstab([rhs], tol=1e-12, maxiter=200)[0]
        assert rel_err(x0[0], xstar) > 1e-3


def _tridiag_fixture(b, seed=0):
    """tests/gpuistl/test_GpuPressureTransferPolicy.cpp:46-120: N = 10 block-tridiagonal, entries U(-10, 10)"""
    n = 10
    rng = np.random.default_rng(seed)
    r, c = [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n:
                r.append(i)
                c.append(j)
    return BCSR.from_block_coo(n, np.array(r), np.array(c), rng.uniform(-10, 10, (len(r), b, b)))


@pytest.mark.parametrize("b", [1, 2, 3, 4])
@pytest.mark.parametrize("transpose", [False, True])
def test_cpr_pieces_against_dense_definitions(b, transpose):
    A = _tridiag_fixture(b)
    for p in range(b):
        w = orc.quasi_impes_weights(A.rowptr, A.col, A.val, p, transpose)
        assert np.allclose(np.abs(w).max(axis=1), 1.0)
        # definition: D^T w (transpose: D w) is a multiple of e_p
        for i in range(A.n):
            D = A.val[A.rowptr[i] + list(A.col[A.rowptr[i]:A.rowptr[i + 1]]).index(i)]
            t = (D if transpose else D.T) @ w[i]
            off = np.delete(t, p)
            assert np.all(np.abs(off) < 1e-10 * max(1.0, abs(t[p])))
        coarse = orc.cpr_coarse_entries(A.rowptr, A.col, A.val, w, p, transpose)
        # dense definition: non-transposed  Ac = R A P with R = blockdiag(w_i^T), P = e_p per block;
        #                   transposed      Ac = R A P with R = e_p^T per block, P = blockdiag(w_i)
        Ad = A.to_dense()
        Rw = np.zeros((A.n, A.n * b))
        Pe = np.zeros((A.n * b, A.n))
        for i in range(A.n):
            Rw[i, i * b:(i + 1) * b] = w[i]
            Pe[i * b + p, i] = 1.0
        Ac = (Pe.T @ Ad @ Rw.T) if transpose else (Rw @ Ad @ Pe)
        rows = np.repeat(np.arange(A.n), np.diff(A.rowptr))
        assert rel_err(coarse, Ac[rows, A.col]) < 1e-13
        rng = np.random.default_rng(3)
        fine = rng.standard_normal(A.n * b)
        rc = orc.cpr_restrict(fine, w, p, transpose)
        assert rel_err(rc, (Pe.T if transpose else Rw) @ fine) < 1e-13
        lhs = rng.standard_normal(A.n)
        back = orc.cpr_prolongate(lhs, fine, w, p, transpose)
        if transpose:
            assert rel_err(back, Rw.T @ lhs) < 1e-13
        else:
            expect = fine.copy()
            expect[p::b] = lhs
            assert np.array_equal(back, expect)
