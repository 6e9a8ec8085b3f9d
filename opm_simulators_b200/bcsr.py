"""Host-side block-CSR container with the exact memory layout of
``Dune::BCRSMatrix<Opm::MatrixBlock<double,b,b>>`` (contiguous row-major b x b blocks in row
order, opm/simulators/linalg/gpuistl/GpuSparseMatrix.cpp:164-167): ``rowptr[n+1]`` int32,
``col[nnzb]`` int32 ascending per row, ``val[nnzb, b, b]`` float64.
"""
from __future__ import annotations

import hashlib

import numpy as np


class BCSR:
    def __init__(self, rowptr, col, val):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        if self.val.ndim != 3 or self.val.shape[1] != self.val.shape[2]:
            raise ValueError("val must have shape [nnzb, b, b]")
        if self.rowptr[-1] != len(self.col) or len(self.col) != self.val.shape[0]:
            raise ValueError("inconsistent BCSR arrays")

    @property
    def n(self) -> int:
        return len(self.rowptr) - 1

    @property
    def b(self) -> int:
        return self.val.shape[1]

    @property
    def nnzb(self) -> int:
        return len(self.col)

    # ---- construction -----------------------------------------------------------------------
    @classmethod
    def from_block_coo(cls, n, rows, cols, blocks):
        """rows/cols: block indices (duplicates are summed); blocks[k, b, b]."""
        rows = np.asarray(rows, np.int64)
        cols = np.asarray(cols, np.int64)
        blocks = np.asarray(blocks, np.float64)
        key = rows * n + cols
        order = np.argsort(key, kind="stable")
        key = key[order]
        uniq, first = np.unique(key, return_index=True)
        val = np.add.reduceat(blocks[order], first, axis=0) if len(key) else blocks[:0]
        r = (uniq // n).astype(np.int64)
        c = (uniq % n).astype(np.int32)
        rowptr = np.zeros(n + 1, np.int64)
        np.add.at(rowptr, r + 1, 1)
        return cls(np.cumsum(rowptr).astype(np.int32), c, val)

    @classmethod
    def from_scalar_coo(cls, n, b, r, c, v):
        """scalar coordinates -> b x b blocks; a block exists iff any of its scalars is listed."""
        r = np.asarray(r, np.int64)
        c = np.asarray(c, np.int64)
        bi, bj = r // b, c // b
        key = bi * n + bj
        uniq, inv = np.unique(key, return_inverse=True)
        val = np.zeros((len(uniq), b, b))
        np.add.at(val, (inv, r % b, c % b), np.asarray(v, np.float64))
        rowptr = np.zeros(n + 1, np.int64)
        np.add.at(rowptr, (uniq // n) + 1, 1)
        return cls(np.cumsum(rowptr).astype(np.int32), (uniq % n).astype(np.int32), val)

    @classmethod
    def from_dense_pattern(cls, pattern, b, fill=None, rng=None):
        """pattern: n x n 0/1 array; values random (rng) or ``fill``."""
        pattern = np.asarray(pattern)
        n = pattern.shape[0]
        rows, cols = np.nonzero(pattern)
        if rng is not None:
            blocks = rng.uniform(-1, 1, size=(len(rows), b, b))
        else:
            blocks = np.full((len(rows), b, b), 0.0 if fill is None else fill)
        return cls.from_block_coo(n, rows, cols, blocks)

    # ---- queries ----------------------------------------------------------------------------
    def row_of_entry(self) -> np.ndarray:
        return np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rowptr))

    def diag_index(self) -> np.ndarray:
        rows = self.row_of_entry()
        idx = np.full(self.n, -1, np.int64)
        k = np.nonzero(self.col == rows)[0]
        idx[rows[k]] = k
        return idx

    def to_dense(self) -> np.ndarray:
        n, b = self.n, self.b
        D = np.zeros((n * b, n * b))
        rows = self.row_of_entry()
        for k in range(self.nnzb):
            i, j = rows[k], self.col[k]
            D[i * b:(i + 1) * b, j * b:(j + 1) * b] = self.val[k]
        return D

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.bsr_matrix((self.val, self.col, self.rowptr), shape=(self.n * self.b, self.n * self.b))

    def is_structurally_symmetric(self) -> bool:
        rows = self.row_of_entry().astype(np.int64)
        k1 = np.sort(rows * self.n + self.col)
        k2 = np.sort(self.col.astype(np.int64) * self.n + rows)
        return bool(np.array_equal(k1, k2))

    def sha256(self) -> str:
        h = hashlib.sha256()
        for a in (self.rowptr, self.col, self.val):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()

    def copy(self) -> "BCSR":
        return BCSR(self.rowptr.copy(), self.col.copy(), self.val.copy())
