"""On-disk system formats of the reference, host side (numpy only).

* MatrixMarket with the ``% ISTL_STRUCT blocked r c`` header line -- what Dune's
  ``storeMatrixMarket``/``readMatrixMarket`` produce and what Flow dumps when ``verbosity > 10``
  (opm/simulators/linalg/WriteSystemMatrixHelper.hpp:63-94, ISTLSolver.hpp:433-440); the reference
  fixtures tests/matr33.txt, tests/rhs3.txt are in this format.
* the raw binaries of opm/simulators/linalg/exportSystem.hpp:40-139
  (``rows.i32 / cols.i32 / data.f64 / r.f64``).
"""
from __future__ import annotations

import os

import numpy as np

from .bcsr import BCSR


def _header(lines):
    if not lines or not lines[0].startswith("%%MatrixMarket"):
        raise ValueError("not a MatrixMarket file")
    fmt = lines[0].split()
    block = None
    k = 1
    while k < len(lines) and lines[k].startswith("%"):
        tok = lines[k][1:].split()
        if len(tok) >= 4 and tok[0] == "ISTL_STRUCT" and tok[1] == "blocked":
            block = (int(tok[2]), int(tok[3]))
        k += 1
    return fmt, block, k


def read_matrix(path: str, block_size: int | None = None) -> BCSR:
    """Read a (blocked) coordinate MatrixMarket file into BCSR.

    ``block_size`` overrides the ISTL_STRUCT header: the reference's tests read the very same
    file as a 1x1-block and as a 3x3-block matrix (tests/test_flexiblesolver.cpp:83-130)."""
    with open(path) as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    fmt, block, k = _header(lines)
    if fmt[2] != "coordinate":
        raise ValueError("matrix file must be in coordinate format")
    nr, nc, nnz = (int(t) for t in lines[k].split()[:3])
    data = np.array([ln.split() for ln in lines[k + 1: k + 1 + nnz]], dtype=np.float64)
    if len(data) != nnz:
        raise ValueError("truncated MatrixMarket file")
    b = block_size if block_size is not None else (block[0] if block else 1)
    if nr % b or nc % b:
        raise ValueError("matrix size is not a multiple of the block size")
    r = data[:, 0].astype(np.int64) - 1
    c = data[:, 1].astype(np.int64) - 1
    v = data[:, 2]
    return BCSR.from_scalar_coo(nr // b, b, r, c, v)


def read_vector(path: str) -> np.ndarray:
    with open(path) as f:
        lines = [ln.strip() for ln in f if ln.strip()]
    fmt, _, k = _header(lines)
    if fmt[2] != "array":
        raise ValueError("vector file must be in array format")
    nr, nc = (int(t) for t in lines[k].split()[:2])
    vals = np.array([float(ln.split()[0]) for ln in lines[k + 1: k + 1 + nr * nc]])
    if len(vals) != nr * nc:
        raise ValueError("truncated MatrixMarket file")
    return vals


def write_matrix(path: str, A: BCSR) -> None:
    """Dune::storeMatrixMarket layout: block after block, every scalar entry of a block listed."""
    b = A.b
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"% ISTL_STRUCT blocked {b} {b}\n")
        f.write(f"{A.n * b} {A.n * b} {A.nnzb * b * b}\n")
        for i in range(A.n):
            for k in range(A.rowptr[i], A.rowptr[i + 1]):
                j = A.col[k]
                for r in range(b):
                    for c in range(b):
                        f.write(f"{i * b + r + 1} {j * b + c + 1} {A.val[k, r, c]:.17g}\n")


def write_vector(path: str, v: np.ndarray, b: int) -> None:
    v = np.asarray(v, dtype=np.float64).reshape(-1)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write(f"% ISTL_STRUCT blocked {b} 1\n")
        f.write(f"{len(v)} 1\n")
        for x in v:
            f.write(f"{x:.17g}\n")


# ---- exportSystem.hpp raw binaries ----------------------------------------------------------
def export_system(dirname: str, A: BCSR, rhs: np.ndarray) -> None:
    os.makedirs(dirname, exist_ok=True)
    A.rowptr.astype(np.int32).tofile(os.path.join(dirname, "rows.i32"))
    A.col.astype(np.int32).tofile(os.path.join(dirname, "cols.i32"))
    A.val.astype(np.float64).tofile(os.path.join(dirname, "data.f64"))
    np.asarray(rhs, np.float64).tofile(os.path.join(dirname, "r.f64"))


def import_system(dirname: str, b: int):
    rowptr = np.fromfile(os.path.join(dirname, "rows.i32"), np.int32)
    col = np.fromfile(os.path.join(dirname, "cols.i32"), np.int32)
    val = np.fromfile(os.path.join(dirname, "data.f64"), np.float64).reshape(-1, b, b)
    rhs = np.fromfile(os.path.join(dirname, "r.f64"), np.float64)
    return BCSR(rowptr, col, val), rhs
