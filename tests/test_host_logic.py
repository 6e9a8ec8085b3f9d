"""CPU-side tests: the C-ABI library loads and exports every symbol of include/opmb200.h, the
integer analysis done by the library is bit-exact against the reference's golden vectors and the
oracle, option-tree / factory error contracts, file formats, partition bookkeeping.
No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, coo_to_bcsr, pattern_to_bcsr
from opm_simulators_b200 import _lib, generators, matrixmarket, partition
from opm_simulators_b200.bcsr import BCSR
from opm_simulators_b200.flexible_solver import (FlexibleSolver, InvalidArgument, MatrixAdapter, PreconditionerFactory,
                                                 PropertyTree, setup_property_tree)
from oracle import oracle as orc


def _coloring(A, kind):
    n = A.n
    color, rows, ptr = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n + 1, np.int32)
    nl = C.c_int32()
    _lib.check(_lib.lib().opmb200_row_coloring(n, A.rowptr, A.col, kind, color, rows, ptr, C.byref(nl)))
    return color, rows, ptr[: nl.value + 1]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "opmb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(opmb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/opmb200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert L.opmb200_version() == 100


@pytest.mark.parametrize("case", range(4))
def test_row_coloring_golden(golden, case):
    """tests/test_graphcoloring.cpp:94-397 through the C ABI (bit-exact)"""
    g = golden["coloring"][case]
    A = pattern_to_bcsr(g["rows"], 1)
    for name, kind in (("SYMMETRIC", 0), ("LOWER", 1), ("UPPER", 2)):
        _, rows, ptr = _coloring(A, kind)
        assert [rows[ptr[i]:ptr[i + 1]].tolist() for i in range(len(ptr) - 1)] == g[name]


@pytest.mark.parametrize("seed", range(5))
def test_row_coloring_matches_oracle_on_irregular_patterns(seed):
    rng = np.random.default_rng(seed)
    n = 300
    dense = rng.random((n, n)) < 0.02
    if seed % 2 == 0:
        dense |= dense.T
    dense |= np.eye(n, dtype=bool)
    A = BCSR.from_dense_pattern(dense, 1, fill=1.0)
    for kind in (0, 1, 2):
        c1, r1, p1 = _coloring(A, kind)
        c2, r2, p2 = orc.row_coloring(A.rowptr, A.col, kind)
        assert np.array_equal(c1, c2) and np.array_equal(r1, r2) and np.array_equal(p1, p2)


def test_row_coloring_c2_like_grid_matches_oracle():
    s = generators.config("C2", scale=0.35, with_rhs=False)
    A = s["A"]
    assert A.is_structurally_symmetric()
    c1, r1, p1 = _coloring(A, 1)
    c2, r2, p2 = orc.row_coloring(A.rowptr, A.col, 1)
    assert np.array_equal(r1, r2) and np.array_equal(p1, p2)


def test_missing_diagonal_is_reported():
    A = pattern_to_bcsr([[0, 1], [0]], 1)
    with pytest.raises(_lib.MatrixBlockError):
        _coloring(A, 1)


def test_partition_simple_golden(golden):
    """tests/test_partitionCells.cpp:116-131"""
    for g in golden["partition_simple"]:
        assert partition.partition_simple(g["num_cells"], g["num_domains"]).tolist() == g["part"]
        b = partition.partition_bounds(g["num_cells"], g["num_domains"])
        assert np.array_equal(np.searchsorted(b, np.arange(g["num_cells"]), side="right") - 1, g["part"])


def test_localize_ghost_last_structure():
    s = generators.blackoil_system(6, 5, 8, b=2, seed=5)
    A = s["A"]
    part = partition.partition_simple(A.n, 3)
    covered = np.zeros(A.n, int)
    for r in range(3):
        ls = partition.localize(A, part, r)
        K = ls.n_interior
        assert np.all(part[ls.l2g[:K]] == r) and np.all(part[ls.l2g[K:]] != r)
        assert np.all(np.diff(ls.l2g[:K]) > 0) and np.all(np.diff(ls.l2g[K:]) > 0)
        covered[ls.l2g[:K]] += 1
        # owner rows are complete and carry the global values; ghost rows are identity rows
        rows = ls.A.row_of_entry()
        for i in range(K):
            g = ls.l2g[i]
            sl = slice(ls.A.rowptr[i], ls.A.rowptr[i + 1])
            gl = slice(A.rowptr[g], A.rowptr[g + 1])
            assert sorted(ls.l2g[ls.A.col[sl]].tolist()) == A.col[gl].tolist()
            assert np.all(np.diff(ls.A.col[sl]) > 0)
        for i in range(K, ls.n):
            sl = slice(ls.A.rowptr[i], ls.A.rowptr[i + 1])
            assert ls.A.col[sl].tolist() == [i] and np.array_equal(ls.A.val[sl][0], np.eye(2))
        # the slab path (rank generates its own rows) builds the identical system
        z0, z1 = np.searchsorted(partition.partition_bounds(A.n, 3), [0])[0], None
        lo, hi = partition.partition_bounds(A.n, 3)[r: r + 2]
        ls2 = partition.localize_rows(int(lo), A.rowptr[lo:hi + 1] - A.rowptr[lo], A.col[A.rowptr[lo]:A.rowptr[hi]],
                                      A.val[A.rowptr[lo]:A.rowptr[hi]], lambda g: part[g], r)
        assert np.array_equal(ls2.l2g, ls.l2g) and np.array_equal(ls2.A.rowptr, ls.A.rowptr)
        assert np.array_equal(ls2.A.col, ls.A.col) and np.array_equal(ls2.A.val, ls.A.val)
        for k in ("neighbors", "send_ptr", "send_rows", "recv_ptr", "recv_rows"):
            assert np.array_equal(ls.halo[k], ls2.halo[k])
    assert np.all(covered == 1)


def test_halo_lists_are_consistent_between_ranks():
    s = generators.blackoil_system(5, 4, 9, b=1, seed=2)
    A, P = s["A"], 4
    part = partition.partition_simple(A.n, P)
    ls = [partition.localize(A, part, r) for r in range(P)]
    for r in range(P):
        h = ls[r].halo
        for k, o in enumerate(h["neighbors"]):
            recv_g = ls[r].l2g[h["recv_rows"][h["recv_ptr"][k]:h["recv_ptr"][k + 1]]]
            ho = ls[o].halo
            ko = list(ho["neighbors"]).index(r)
            send_g = ls[o].l2g[ho["send_rows"][ho["send_ptr"][ko]:ho["send_ptr"][ko + 1]]]
            assert np.array_equal(recv_g, send_g)


def test_slab_generation_matches_whole_grid():
    full = generators.blackoil_system(5, 4, 6, b=3, seed=9, with_rhs=False)["A"]
    slab = generators.blackoil_system(5, 4, 6, b=3, seed=9, z_range=(2, 4))
    lo, hi = 2 * 20, 4 * 20
    assert np.array_equal(slab["gcol"], full.col[full.rowptr[lo]:full.rowptr[hi]])
    assert np.array_equal(slab["val"], full.val[full.rowptr[lo]:full.rowptr[hi]])


def test_generated_systems_have_the_reference_sparsity_rule():
    s = generators.config("C2", scale=0.3)
    A = s["A"]
    rows = A.row_of_entry()
    assert A.is_structurally_symmetric()
    assert np.all(A.diag_index() >= 0)
    for i in range(0, A.n, 97):
        c = A.col[A.rowptr[i]:A.rowptr[i + 1]]
        assert np.all(np.diff(c) > 0)
    assert len(s["rhs"]) == A.n * 3 and np.isfinite(s["rhs"]).all()
    assert 4.0 < A.nnzb / A.n <= 7.2 and rows[-1] == A.n - 1


# ---- option tree / factory error contracts ------------------------------------------------------
def test_property_tree_semantics(golden):
    prm = PropertyTree(golden["options_flexiblesolver_1x1"])
    assert prm.get("tol", 1e-2) == 0.5 and prm.get("maxiter", 200) == 20
    assert prm.get("preconditioner.type", "x") == "ilu0"
    assert prm.get("preconditioner.relaxation", 1.0) == 1.0
    assert prm.get_child("preconditioner").get("type", "") == "ilu0"
    assert prm.get_child_optional("nope") is None
    with pytest.raises(InvalidArgument):
        prm.get("nope")
    prm.put("preconditioner.relaxation", 0.9)
    assert PropertyTree(prm.to_json()).get("preconditioner.relaxation", 1.0) == 0.9
    p2 = setup_property_tree("ilu0")
    assert p2.get("preconditioner.type", "") == "paroverilu0" and p2.get("preconditioner.relaxation", 1.0) == 0.9
    assert setup_property_tree("dilu").get("preconditioner.type", "") == "dilu"
    with pytest.raises(InvalidArgument):
        setup_property_tree("cprw")


def test_unknown_preconditioner_and_solver_are_invalid_argument():
    """tests/test_preconditionerfactory.cpp:183-228 error contract; FlexibleSolver_impl.hpp:326-329"""
    A = pattern_to_bcsr([[0, 1], [0, 1]], 3)
    with pytest.raises(InvalidArgument, match="not registered in the factory"):
        FlexibleSolver(MatrixAdapter(A), {"preconditioner": {"type": "nothing_registered"}})
    # straight through the C ABI as well (the option check runs before any device work)
    h = C.c_void_p()
    L = _lib.lib()
    for js, frag in ((b'{"preconditioner": {"type": "amg"}}', "not registered in the factory"),
                     (b'{"solver": "gmres", "preconditioner": {"type": "dilu"}}', "Solver gmres not known"),
                     (b'{"solver": ', "JSON parse error")):
        rc = L.opmb200_create(js, 3, A.n, A.nnzb, A.rowptr, A.col, A.n, None, None, C.byref(h))
        assert rc == _lib.BAD_OPTIONS and frag in L.opmb200_last_error().decode()
    rc = L.opmb200_create(None, 9, A.n, A.nnzb, A.rowptr, A.col, A.n, None, None, C.byref(h))
    assert rc == _lib.INVALID_ARGUMENT
    # the tuning keys of the reference's GPU preconditioners (StandardPreconditioners_gpu_serial.hpp:77-80, 92-96) are
    # accepted and type-checked; mixed precision storage is refused
    rc = L.opmb200_create(b'{"preconditioner": {"type": "dilu", "split_matrix": "maybe"}}', 3, A.n, A.nnzb, A.rowptr, A.col,
                          A.n, None, None, C.byref(h))
    assert rc == _lib.BAD_OPTIONS and "not a bool" in L.opmb200_last_error().decode()
    rc = L.opmb200_create(b'{"preconditioner": {"type": "dilu", "mixed_precision_scheme": 1}}', 3, A.n, A.nnzb, A.rowptr,
                          A.col, A.n, None, None, C.byref(h))
    assert rc == _lib.BAD_OPTIONS and "mixed_precision_scheme" in L.opmb200_last_error().decode()
    rc = L.opmb200_create(b'{"preconditioner": {"type": "gpudilu", "split_matrix": "true", "tune_gpu_kernels": "false", '
                          b'"reorder": "true"}}', 3, A.n, A.nnzb, A.rowptr, A.col, A.n, None, None, C.byref(h))
    assert rc in (_lib.SUCCESS, _lib.CUDA_ERROR), L.opmb200_last_error().decode()  # past the option check (no device here)
    if rc == _lib.SUCCESS:
        L.opmb200_destroy(h)


def test_add_creator_plugin_hook():
    """PreconditionerFactory::addCreator (tests/test_preconditionerfactory.cpp:200-217)"""
    made = []

    def creator(op, prm):
        made.append(prm.get("type", ""))
        raise InvalidArgument(_lib.BAD_OPTIONS, "creator was called")

    PreconditionerFactory.addCreator("MyPlugin", creator)
    A = pattern_to_bcsr([[0]], 1)
    with pytest.raises(InvalidArgument, match="creator was called"):
        PreconditionerFactory.create(MatrixAdapter(A), PropertyTree({"type": "myplugin"}))
    assert made == ["myplugin"]


def test_matrix_without_diagonal_is_rejected_at_create():
    A = pattern_to_bcsr([[0, 1], [0]], 2)
    h = C.c_void_p()
    rc = _lib.lib().opmb200_create(b'{"preconditioner": {"type": "dilu"}}', 2, A.n, A.nnzb, A.rowptr, A.col, A.n,
                                   None, None, C.byref(h))
    assert rc == _lib.DIAGONAL_MISSING


# ---- file formats -----------------------------------------------------------------------------
def test_matrixmarket_roundtrip_and_block_reinterpretation(golden, tmp_path):
    A3 = coo_to_bcsr(golden["matr33"], 3)
    p = tmp_path / "m.txt"
    matrixmarket.write_matrix(str(p), A3)
    B3 = matrixmarket.read_matrix(str(p))
    assert B3.b == 3 and np.array_equal(B3.col, A3.col) and np.allclose(B3.val, A3.val, rtol=1e-15)
    B1 = matrixmarket.read_matrix(str(p), block_size=1)
    assert B1.n == 9 and B1.nnzb == 63 and np.allclose(B1.to_dense(), A3.to_dense())
    v = np.array(golden["rhs3"])
    matrixmarket.write_vector(str(tmp_path / "v.txt"), v, 3)
    assert np.array_equal(matrixmarket.read_vector(str(tmp_path / "v.txt")), v)
    matrixmarket.export_system(str(tmp_path / "sys"), A3, v)
    A4, r4 = matrixmarket.import_system(str(tmp_path / "sys"), 3)
    assert np.array_equal(A4.val, A3.val) and np.array_equal(r4, v)


# ---- sweep schedules (host-only planner, opmb200_plan_schedule) -------------------------------------
def _plan(A, schedule, chunk_rows=0, n_interior=None):
    import ctypes as C
    n = A.n
    ns, nc, cr, est = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    r2n = np.zeros(n, np.int32)
    sf = np.zeros(n + 1, np.int32)
    cf = np.zeros(n + 2, np.int32)
    _lib.check(_lib.lib().opmb200_plan_schedule(A.b, n, A.nnzb, A.rowptr, A.col, n if n_interior is None else n_interior,
                                                schedule, chunk_rows, C.byref(ns), C.byref(nc), C.byref(cr), C.byref(est),
                                                r2n.ctypes.data, sf.ctypes.data, cf.ctypes.data))
    return dict(n_slices=ns.value, n_chunks=nc.value, chunk_rows=cr.value, est=est.value, r2n=r2n,
                slice_first=sf[:ns.value + 1], chunk_first=cf[:nc.value + 1])


def _check_schedule(A, P, chunks):
    n = A.n
    assert sorted(P["r2n"].tolist()) == list(range(n))
    pos = np.empty(n, np.int64)
    pos[P["r2n"]] = np.arange(n)
    sl = np.searchsorted(P["slice_first"], pos, side="right") - 1          # slice of every row
    assert (np.diff(P["slice_first"]) <= 32).all() and (np.diff(P["slice_first"]) > 0).all()
    rows = np.repeat(np.arange(n), np.diff(A.rowptr))
    lo = A.col < rows
    # a row depends only on rows of earlier slices (lower sweep; the upper sweep is the mirror image)
    assert (sl[A.col[lo]] < sl[rows[lo]]).all()
    if chunks:
        ch = np.searchsorted(P["chunk_first"], sl, side="right") - 1
        assert (ch[A.col[lo]] <= ch[rows[lo]]).all()                       # in-order chunk start cannot deadlock
        hi = A.col > rows
        assert (ch[A.col[hi]] >= ch[rows[hi]]).all()


@pytest.mark.parametrize("dims", [(12, 20, 9), (7, 33, 5), (16, 8, 1), (5, 1, 1)])
def test_chunk_schedule_box_grids_use_line_tiles(dims):
    s = generators.blackoil_system(*dims, b=2, seed=5, with_rhs=False)
    A = s["A"]
    for sched, cr in ((0, 0), (1, 0), (1, 64)):
        P = _plan(A, sched, cr)
        _check_schedule(A, P, sched == 1)


def test_chunk_schedule_tiles_on_a_box_grid():
    A = generators.blackoil_system(60, 64, 16, b=1, seed=5, with_rhs=False)["A"]
    P = _plan(A, 1, 0)
    assert P["chunk_rows"] < 0, "a clean box grid is cut into tiles of 32 grid lines"
    _check_schedule(A, P, True)
    # fewer, fuller steps than contiguous chunks of 32 lines
    assert P["n_slices"] < _plan(A, 1, 32 * 60)["n_slices"]


def test_chunk_schedule_falls_back_when_tiles_would_deadlock():
    # non-neighbour connections pointing against the tile order make the tile numbering invalid
    s = generators.blackoil_system(10, 24, 8, b=1, seed=6, nnc=40, with_rhs=False)
    A = s["A"]
    P = _plan(A, 1, 0)
    _check_schedule(A, P, True)


def test_schedules_with_ghost_rows_and_degenerate_patterns():
    """ghost-last local systems (identity ghost rows behind the owners) and tiny patterns through every schedule"""
    rng = np.random.default_rng(9)
    full = generators.blackoil_system(10, 12, 8, b=2, seed=8, with_rhs=False)["A"]
    part = partition.partition_simple(full.n, 2)
    for rank in (0, 1):
        ls = partition.localize(full, part, rank)
        for sched, cr in ((0, 0), (1, 0), (1, 128), (1, -804)):
            P = _plan(ls.A, sched, cr, n_interior=ls.n_interior)
            assert sorted(P["r2n"].tolist()) == list(range(ls.A.n))
            # owner rows: every lower neighbour sits in an earlier slice; ghost rows depend on nothing
            pos = np.empty(ls.A.n, np.int64)
            pos[P["r2n"]] = np.arange(ls.A.n)
            sl = np.searchsorted(P["slice_first"], pos, side="right") - 1
            rows = np.repeat(np.arange(ls.A.n), np.diff(ls.A.rowptr))
            own_lo = (ls.A.col < rows) & (rows < ls.n_interior)
            assert (sl[ls.A.col[own_lo]] < sl[rows[own_lo]]).all()
            if sched == 1:
                ch = np.searchsorted(P["chunk_first"], sl, side="right") - 1
                own = rows < ls.n_interior
                lo, hi = own & (ls.A.col < rows), own & (ls.A.col > rows)
                assert (ch[ls.A.col[lo]] <= ch[rows[lo]]).all() and (ch[ls.A.col[hi]] >= ch[rows[hi]]).all()
    one = pattern_to_bcsr([[0]], 3, rng)
    for sched in (0, 1):
        P = _plan(one, sched)
        assert P["n_slices"] == 1 and P["r2n"].tolist() == [0]
