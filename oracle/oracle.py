"""ctypes front-end of the CPU oracle (oracle/opm_oracle.c) and of the reference-owned
``mixed`` C solver (oracle/_ref/libopm_mixed_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py -- never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(HERE, "_build", "liboracle.so")
_REF = os.path.join(HERE, "_ref", "libopm_mixed_ref.so")

COLOR_SYMMETRIC, COLOR_LOWER, COLOR_UPPER = 0, 1, 2
PREC_NONE, PREC_DILU, PREC_ILU0 = 0, 1, 2
PREC_KIND = {"nothing": PREC_NONE, "none": PREC_NONE, "dilu": PREC_DILU, "ilu0": PREC_ILU0,
             "paroverilu0": PREC_ILU0}
ERR_NAMES = {0: "ok", 1: "diagonal entry missing", 2: "singular matrix block",
             3: "BiCGSTAB breakdown", 4: "defect is NaN/Inf", 5: "bad argument"}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}: {ERR_NAMES.get(code, '?')}")
        self.code = code


def build(force: bool = False) -> None:
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(HERE, "opm_oracle.c")):
        subprocess.run(["make", "-C", HERE, "all"], check=True, capture_output=True)
    elif not os.path.exists(_REF) and os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("reduction", C.c_double), ("converged", C.c_int),
                ("conv_rate", C.c_double), ("it", C.c_double), ("norm0", C.c_double), ("norm", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_row_coloring.argtypes = [C.c_int, _i32p, _i32p, C.c_int, _i32p, _i32p, _i32p]
        L.orc_row_coloring.restype = C.c_int
        L.orc_reorder_maps.argtypes = [C.c_int, _i32p, _i32p, _i32p]
        L.orc_partition_simple.argtypes = [C.c_int, C.c_int, _i32p]
        L.orc_invert_block.argtypes = [C.c_int, _f64p]
        L.orc_invert_block.restype = C.c_int
        L.orc_spmv.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p]
        L.orc_spmv_scaleadd.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_double, _f64p, _f64p]
        L.orc_make_overlap_rows_invalid.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int]
        L.orc_dilu_update.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p]
        L.orc_dilu_update.restype = C.c_int
        L.orc_dilu_apply.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p]
        L.orc_ilu0_decompose.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int]
        L.orc_ilu0_decompose.restype = C.c_int
        L.orc_ilu0_apply.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, _f64p, _f64p]
        L.orc_par_create.argtypes = [C.c_int, C.c_int, C.c_long]
        L.orc_par_create.restype = C.c_void_p
        L.orc_par_destroy.argtypes = [C.c_void_p]
        L.orc_par_set_sub.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_void_p]
        L.orc_par_set_sub.restype = C.c_int
        L.orc_par_prec_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_par_prec_update.restype = C.c_int
        PP = C.POINTER(C.POINTER(C.c_double))
        L.orc_par_prec_apply.argtypes = [C.c_void_p, PP, PP]
        L.orc_par_prec_apply.restype = C.c_int
        L.orc_par_copy_owner_to_all.argtypes = [C.c_void_p, PP]
        L.orc_par_dot.argtypes = [C.c_void_p, PP, PP]
        L.orc_par_dot.restype = C.c_double
        L.orc_par_bicgstab.argtypes = [C.c_void_p, PP, PP, C.c_double, C.c_int, C.c_int, C.POINTER(Result),
                                       C.c_void_p, C.POINTER(C.c_int)]
        L.orc_par_bicgstab.restype = C.c_int
        L.orc_par_dinv.argtypes = [C.c_void_p, C.c_int]
        L.orc_par_dinv.restype = C.POINTER(C.c_double)
        L.orc_par_lu.argtypes = [C.c_void_p, C.c_int]
        L.orc_par_lu.restype = C.POINTER(C.c_double)
        L.orc_well_apply.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.orc_par_set_wells.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_par_set_wells.restype = C.c_int
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _aligned_zeros(n, align=64):
    """zero-filled float64 vector whose data pointer is `align`-byte aligned (the reference's
    vec_inner2 promises 64-byte alignment to the compiler, mixed/bslv.c:88-91)"""
    raw = np.zeros(n + align // 8)
    off = (-raw.ctypes.data % align) // 8
    return raw[off: off + n]


# --------------------------------------------------------------------------------------------
# integer artefacts
# --------------------------------------------------------------------------------------------
def row_coloring(rowptr, col, kind=COLOR_LOWER):
    """-> (color[n], level_rows[n], level_ptr[nlevels+1])  (GraphColoring.hpp:246-307)"""
    rowptr, col = _i32(rowptr), _i32(col)
    n = len(rowptr) - 1
    color = np.zeros(n, np.int32)
    rows = np.zeros(n, np.int32)
    ptr = np.zeros(n + 1, np.int32)
    nl = lib().orc_row_coloring(n, rowptr, col, kind, color, rows, ptr)
    if nl < 0:
        raise OracleError(-nl)
    return color, rows, ptr[: nl + 1].copy()


def level_sets(rowptr, col, kind=COLOR_LOWER):
    _, rows, ptr = row_coloring(rowptr, col, kind)
    return [rows[ptr[i]: ptr[i + 1]].tolist() for i in range(len(ptr) - 1)]


def reorder_maps(level_rows):
    level_rows = _i32(level_rows)
    n = len(level_rows)
    r2n = np.zeros(n, np.int32)
    n2r = np.zeros(n, np.int32)
    lib().orc_reorder_maps(n, level_rows, r2n, n2r)
    return r2n, n2r


def partition_simple(num_cells, num_domains):
    part = np.zeros(num_cells, np.int32)
    lib().orc_partition_simple(num_cells, num_domains, part)
    return part


def invert_block(blk):
    a = _f64(blk).copy()
    b = a.shape[0]
    rc = lib().orc_invert_block(b, a.reshape(-1))
    if rc:
        raise OracleError(rc)
    return a


# --------------------------------------------------------------------------------------------
# serial kernels on one BCSR (rowptr, col, val[nnzb,b,b])
# --------------------------------------------------------------------------------------------
def spmv(rowptr, col, val, x, interior=None):
    rowptr, col, val, x = _i32(rowptr), _i32(col), _f64(val), _f64(x)
    n, b = len(rowptr) - 1, val.shape[-1]
    y = np.zeros(n * b)
    lib().orc_spmv(n, b, rowptr, col, val.reshape(-1), n if interior is None else interior, x.reshape(-1), y)
    return y


def spmv_scaleadd(rowptr, col, val, alpha, x, y, interior=None):
    rowptr, col, val, x = _i32(rowptr), _i32(col), _f64(val), _f64(x)
    n, b = len(rowptr) - 1, val.shape[-1]
    y = _f64(y).reshape(-1).copy()
    lib().orc_spmv_scaleadd(n, b, rowptr, col, val.reshape(-1), n if interior is None else interior,
                            alpha, x.reshape(-1), y)
    return y


def well_apply(wells, x, y, b):
    """y -= C^T (Dinv (B x)) well after well (WellOperators.hpp:84-91, StandardWellEquations.cpp:132-148); returns y"""
    wp, wc = _i32(wells["ptr"]), _i32(wells["cells"])
    B, Cm, Di = _f64(wells["B"]), _f64(wells["C"]), _f64(wells["Dinv"])
    y = _f64(y).reshape(-1).copy()
    lib().orc_well_apply(len(wp) - 1, int(Di.shape[-1]), b, wp, wc, B.reshape(-1), Cm.reshape(-1), Di.reshape(-1),
                         _f64(x).reshape(-1), y)
    return y


def make_overlap_rows_invalid(rowptr, col, val, interior):
    rowptr, col = _i32(rowptr), _i32(col)
    val = _f64(val).copy()
    n, b = len(rowptr) - 1, val.shape[-1]
    lib().orc_make_overlap_rows_invalid(n, b, rowptr, col, val.reshape(-1), interior)
    return val


def dilu_update(rowptr, col, val):
    rowptr, col, val = _i32(rowptr), _i32(col), _f64(val)
    n, b = len(rowptr) - 1, val.shape[-1]
    dinv = np.zeros((n, b, b))
    rc = lib().orc_dilu_update(n, b, rowptr, col, val.reshape(-1), dinv.reshape(-1))
    if rc:
        raise OracleError(rc)
    return dinv


def dilu_apply(rowptr, col, val, dinv, d):
    rowptr, col, val = _i32(rowptr), _i32(col), _f64(val)
    n, b = len(rowptr) - 1, val.shape[-1]
    v = np.zeros(n * b)
    lib().orc_dilu_apply(n, b, rowptr, col, val.reshape(-1), _f64(dinv).reshape(-1), _f64(d).reshape(-1), v)
    return v


def ilu0_decompose(rowptr, col, val, interior=None):
    rowptr, col = _i32(rowptr), _i32(col)
    lu = _f64(val).copy()
    n, b = len(rowptr) - 1, lu.shape[-1]
    rc = lib().orc_ilu0_decompose(n, b, rowptr, col, lu.reshape(-1), n if interior is None else interior)
    if rc:
        raise OracleError(rc)
    return lu


def ilu0_apply(rowptr, col, lu, d, interior=None, relaxation=1.0, v0=None):
    rowptr, col, lu = _i32(rowptr), _i32(col), _f64(lu)
    n, b = len(rowptr) - 1, lu.shape[-1]
    v = np.zeros(n * b) if v0 is None else _f64(v0).reshape(-1).copy()
    lib().orc_ilu0_apply(n, b, rowptr, col, lu.reshape(-1), n if interior is None else interior,
                         _f64(d).reshape(-1), v)
    if abs(relaxation - 1.0) > 1e-15:
        v *= relaxation
    return v


# --------------------------------------------------------------------------------------------
# (multi-subdomain) solver
# --------------------------------------------------------------------------------------------
class ParSystem:
    """P subdomains emulating P MPI ranks of Flow in one process.

    ``subs`` is a list of dicts with keys rowptr, col, val[nnzb,b,b], interior, l2g (None if serial).
    """

    def __init__(self, subs, nglobal=None):
        L = lib()
        self.subs = []
        self.b = int(np.asarray(subs[0]["val"]).shape[-1])
        self.nsub = len(subs)
        if nglobal is None:
            nglobal = sum(int(len(s["rowptr"]) - 1 if s.get("interior") is None else s["interior"]) for s in subs)
        self.h = L.orc_par_create(self.nsub, self.b, nglobal)
        for p, s in enumerate(subs):
            rp, cl, vl = _i32(s["rowptr"]), _i32(s["col"]), _f64(s["val"])
            n = len(rp) - 1
            interior = int(s.get("interior", n) if s.get("interior") is not None else n)
            l2g = None if s.get("l2g") is None else _i32(s["l2g"])
            self.subs.append((rp, cl, vl, l2g, n, interior))  # keep alive
            rc = L.orc_par_set_sub(self.h, p, n, interior, rp, cl, vl.reshape(-1),
                                   None if l2g is None else l2g.ctypes.data_as(C.c_void_p))
            if rc:
                raise OracleError(rc)
        self.kind = PREC_NONE

    @classmethod
    def serial(cls, rowptr, col, val):
        return cls([dict(rowptr=rowptr, col=col, val=val, interior=None, l2g=None)])

    def __del__(self):
        try:
            lib().orc_par_destroy(self.h)
        except Exception:
            pass

    def _pp(self, vecs):
        arr = (C.POINTER(C.c_double) * self.nsub)()
        for p, v in enumerate(vecs):
            assert v.dtype == np.float64 and v.flags.c_contiguous
            arr[p] = v.ctypes.data_as(C.POINTER(C.c_double))
        return arr

    def set_wells(self, wells, p=0):
        """wells: dict(ptr[nw+1], cells[nperf], B[nperf,dw,b], C[nperf,dw,b], Dinv[nw,dw,dw]) kept OUTSIDE the matrix
        (WellModelMatrixAdapter): every operator application of bicgstab() adds y -= C^T Dinv B x"""
        wp, wc = _i32(wells["ptr"]), _i32(wells["cells"])
        B, Cm, Di = _f64(wells["B"]), _f64(wells["C"]), _f64(wells["Dinv"])
        self._wells = getattr(self, "_wells", {})
        self._wells[p] = (wp, wc, B, Cm, Di)  # keep alive
        rc = lib().orc_par_set_wells(self.h, p, len(wp) - 1, int(Di.shape[-1]), wp, wc, B.reshape(-1), Cm.reshape(-1),
                                     Di.reshape(-1))
        if rc:
            raise OracleError(rc)

    def prec_update(self, kind, relaxation=1.0):
        if isinstance(kind, str):
            kind = PREC_KIND[kind.lower()]
        self.kind = kind
        rc = lib().orc_par_prec_update(self.h, kind, relaxation)
        if rc:
            raise OracleError(rc)

    def dinv(self, p=0):
        n, b = self.subs[p][4], self.b
        return np.ctypeslib.as_array(lib().orc_par_dinv(self.h, p), shape=(n, b, b)).copy()

    def lu(self, p=0):
        nnzb, b = self.subs[p][2].shape[0], self.b
        return np.ctypeslib.as_array(lib().orc_par_lu(self.h, p), shape=(nnzb, b, b)).copy()

    def prec_apply(self, d, v0=None):
        d = [_f64(x).reshape(-1) for x in d]
        v = [np.zeros_like(x) for x in d] if v0 is None else [_f64(x).reshape(-1).copy() for x in v0]
        lib().orc_par_prec_apply(self.h, self._pp(v), self._pp(d))
        return v

    def copy_owner_to_all(self, v):
        v = [_f64(x).reshape(-1).copy() for x in v]
        lib().orc_par_copy_owner_to_all(self.h, self._pp(v))
        return v

    def dot(self, x, y):
        x = [_f64(a).reshape(-1) for a in x]
        y = [_f64(a).reshape(-1) for a in y]
        return lib().orc_par_dot(self.h, self._pp(x), self._pp(y))

    def bicgstab(self, b, x0=None, tol=1e-2, maxiter=200, op_repeats=1):
        """-> (x list, residual list (b overwritten as Dune does), result dict, history)"""
        r = [_f64(a).reshape(-1).copy() for a in b]
        x = [np.zeros_like(a) for a in r] if x0 is None else [_f64(a).reshape(-1).copy() for a in x0]
        res = Result()
        hist = np.zeros(2 * maxiter + 2)
        nh = C.c_int(0)
        rc = lib().orc_par_bicgstab(self.h, self._pp(x), self._pp(r), tol, maxiter, op_repeats, C.byref(res),
                                    hist.ctypes.data_as(C.c_void_p), C.byref(nh))
        if rc:
            raise OracleError(rc)
        return x, r, res.as_dict(), hist[: nh.value].copy()


def solve_serial(rowptr, col, val, b, prec="dilu", tol=1e-2, maxiter=200, relaxation=1.0, x0=None, op_repeats=1):
    ps = ParSystem.serial(rowptr, col, val)
    ps.prec_update(prec, relaxation)
    x, r, res, hist = ps.bicgstab([b], None if x0 is None else [x0], tol, maxiter, op_repeats)
    return x[0], res, hist


# --------------------------------------------------------------------------------------------
# the reference's own C solver (opm/simulators/linalg/mixed/*.c), compiled unmodified
# --------------------------------------------------------------------------------------------
class _BsrMatrix(C.Structure):
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("nnz", C.c_int), ("b", C.c_int),
                ("rowptr", C.POINTER(C.c_int)), ("colidx", C.POINTER(C.c_int)),
                ("dbl", C.POINTER(C.c_double)), ("flt", C.POINTER(C.c_float))]


class _Prec(C.Structure):
    _fields_ = [("L", C.POINTER(_BsrMatrix)), ("D", C.POINTER(_BsrMatrix)), ("U", C.POINTER(_BsrMatrix)),
                ("noffsets", C.c_int), ("offsets", C.c_void_p)]


class _BslvMemory(C.Structure):
    _fields_ = [("use_dilu", C.c_bool), ("tol", C.c_double), ("max_iter", C.c_int),
                ("e", C.POINTER(C.c_double)), ("n", C.c_int), ("dtmp", C.POINTER(C.POINTER(C.c_double))),
                ("P", C.POINTER(_Prec))]


_ref = None


def ref_available() -> bool:
    return os.path.exists(_REF)


def ref_lib():
    global _ref
    if _ref is None:
        R = C.CDLL(_REF)
        R.bsr_alloc.restype = C.POINTER(_BsrMatrix)
        R.bsr_init.argtypes = [C.POINTER(_BsrMatrix), C.c_int, C.c_int, C.c_int]
        R.bsr_free.argtypes = [C.POINTER(_BsrMatrix)]
        R.bsr_vdspmv3.argtypes = [C.POINTER(_BsrMatrix), C.c_void_p, C.c_void_p]
        R.bslv_alloc.restype = C.POINTER(_BslvMemory)
        R.bslv_init.argtypes = [C.POINTER(_BslvMemory), C.c_double, C.c_int, C.POINTER(_BsrMatrix), C.c_bool]
        R.bslv_free.argtypes = [C.POINTER(_BslvMemory)]
        R.bslv_pbicgstab3d.argtypes = [C.POINTER(_BslvMemory), C.POINTER(_BsrMatrix), C.c_void_p, C.c_void_p]
        R.bslv_pbicgstab3d.restype = C.c_int
        R.prec_dilu_factorize.argtypes = [C.POINTER(_Prec), C.POINTER(_BsrMatrix)]
        R.prec_ilu0_factorize.argtypes = [C.POINTER(_Prec), C.POINTER(_BsrMatrix)]
        R.prec_dapply3c.argtypes = [C.POINTER(_Prec), C.c_void_p]
        _ref = R
    return _ref


class RefMixedSolver:
    """Drives ``bslv_pbicgstab3d`` the way mixed/wrapper.hpp:33-95 does (blocks transposed to
    column-major, 3x3 only, structurally symmetric patterns only, serial)."""

    def __init__(self, rowptr, col, val, tol=1e-2, maxiter=200, use_dilu=True):
        R = ref_lib()
        rowptr, col, val = _i32(rowptr), _i32(col), _f64(val)
        assert val.shape[-1] == 3, "the reference mixed solver is 3x3 only"
        self.n = len(rowptr) - 1
        nnz = len(col)
        self.A = R.bsr_alloc()
        R.bsr_init(self.A, self.n, nnz, 3)
        C.memmove(self.A.contents.rowptr, rowptr.ctypes.data, rowptr.nbytes)
        C.memmove(self.A.contents.colidx, col.ctypes.data, col.nbytes)
        self.mem = R.bslv_alloc()
        # e[] is indexed up to max_iter inclusive by the solver; allocate one spare iteration
        R.bslv_init(self.mem, tol, maxiter + 1, self.A, use_dilu)
        self.mem.contents.max_iter = maxiter
        self.set_values(val)

    def set_values(self, val):
        cm = np.ascontiguousarray(np.transpose(_f64(val), (0, 2, 1)))  # row-major -> column-major blocks
        C.memmove(self.A.contents.dbl, cm.ctypes.data, cm.nbytes)

    def solve(self, b):
        n3 = 3 * self.n
        bb = _aligned_zeros(n3 + 8)
        bb[:n3] = _f64(b).reshape(-1)
        x = _aligned_zeros(n3 + 8)
        count = ref_lib().bslv_pbicgstab3d(self.mem, self.A, bb.ctypes.data, x.ctypes.data)
        red = self.mem.contents.e[count] if count > 0 else float("nan")
        return x[:n3].copy(), count, red

    def spmv(self, x):
        n3 = 3 * self.n
        xx = np.zeros(n3 + 8)
        xx[:n3] = _f64(x).reshape(-1)
        y = np.zeros(n3 + 8)
        ref_lib().bsr_vdspmv3(self.A, xx.ctypes.data, y.ctypes.data)
        return y[:n3].copy()

    def factor_apply(self, d, use_dilu):
        """factorise and apply the preconditioner once: returns M^-1 d"""
        R = ref_lib()
        P = self.mem.contents.P
        (R.prec_dilu_factorize if use_dilu else R.prec_ilu0_factorize)(P, self.A)
        n3 = 3 * self.n
        x = np.zeros(n3 + 8)
        x[:n3] = _f64(d).reshape(-1)
        R.prec_dapply3c(P, x.ctypes.data)
        return x[:n3].copy()

    def __del__(self):
        try:
            R = ref_lib()
            R.bslv_free(self.mem)
            R.bsr_free(self.A)
        except Exception:
            pass


# --------------------------------------------------------------------------------------------
# CPR pieces around the smoother (numpy restatements; test infrastructure like everything here)
# --------------------------------------------------------------------------------------------
def _diag_index(rowptr, col):
    n = len(rowptr) - 1
    d = np.empty(n, np.int64)
    for i in range(n):
        k = rowptr[i] + np.searchsorted(col[rowptr[i]:rowptr[i + 1]], i)
        assert col[k] == i
        d[i] = k
    return d


def quasi_impes_weights(rowptr, col, val, pressure_index, transpose=False):
    """Amg::getQuasiImpesWeights (getQuasiImpesWeights.hpp:64-111): per row solve D^T w = e_p (transpose: D w = e_p)
    with the diagonal block D, then w /= max|w|.  -> [n, b]"""
    val = _f64(val)
    b = val.shape[-1]
    D = val[_diag_index(rowptr, col)]
    rhs = np.zeros(b)
    rhs[pressure_index] = 1.0
    M = D if transpose else np.transpose(D, (0, 2, 1))
    w = np.linalg.solve(M, np.broadcast_to(rhs, (len(D), b))[..., None])[..., 0]
    return w / np.abs(w).max(axis=1, keepdims=True)


def cpr_coarse_entries(rowptr, col, val, weights, pressure_index, transpose=False):
    """PressureTransferPolicy::calculateCoarseEntries (gpuistl/detail/cpr_amg_operations.cu:79-123):
    transpose False: sum_j A_k[j][p] w_row[j];  True: sum_j A_k[p][j] w_col[j].  -> [nnzb]"""
    val = _f64(val)
    w = _f64(weights).reshape(len(rowptr) - 1, -1)
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    if transpose:
        return np.einsum("kj,kj->k", val[:, pressure_index, :], w[np.asarray(col)])
    return np.einsum("kj,kj->k", val[:, :, pressure_index], w[rows])


def cpr_restrict(fine, weights, pressure_index, transpose=False):
    """PressureTransferPolicy::moveToCoarseLevel (cpr_amg_operations.cu:126-151)"""
    w = _f64(weights)
    if w.ndim != 2:
        raise ValueError("weights must be [n, b]")
    f = _f64(fine).reshape(-1, w.shape[-1])
    return f[:, pressure_index].copy() if transpose else np.einsum("ik,ik->i", f, w)


def cpr_prolongate(coarse, fine, weights, pressure_index, transpose=False):
    """PressureTransferPolicy::moveToFineLevel (PressureTransferPolicy.hpp:148-162): returns the updated fine vector"""
    w = _f64(weights)
    b = w.shape[-1]
    f = _f64(fine).reshape(-1, b).copy()
    c = _f64(coarse).reshape(-1)
    if transpose:
        f[:] = c[:, None] * w
    else:
        f[:, pressure_index] = c
    return f.reshape(-1)
