"""Domain decomposition bookkeeping on the host: contiguous partition (Opm::partitionCellsSimple,
opm/simulators/flow/partitionCells.cpp:734-751), ghost-last local systems with one overlap layer
(FlowGenericVanguard.hpp:79, ISTLSolver.hpp:299-306, findOverlapRowsAndColumns.hpp:92-135) and the
owner/copy index lists the halo exchange needs (gpuistl/GpuAwareMPISender.hpp:164-222).
The integer work is done by the library (opmb200_partition_simple / opmb200_localize); the numpy
slab variant builds the same structure from rows a rank generated on its own.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .bcsr import BCSR


def partition_simple(num_cells: int, num_domains: int) -> np.ndarray:
    part = np.zeros(num_cells, np.int32)
    _lib.check(_lib.lib().opmb200_partition_simple(num_cells, num_domains, part))
    return part


def partition_bounds(num_cells: int, num_domains: int) -> np.ndarray:
    """first cell of every domain of partition_simple (length num_domains+1)"""
    q, r = divmod(num_cells, num_domains)
    sizes = np.full(num_domains, q, np.int64)
    sizes[:r] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


class LocalSystem:
    """one rank's ghost-last system: owner rows [0, n_interior), ghost rows behind them"""

    def __init__(self, A: BCSR, n_interior: int, l2g: np.ndarray, src: np.ndarray | None = None):
        self.A = A
        self.n_interior = int(n_interior)
        self.l2g = np.ascontiguousarray(l2g, np.int32)
        self.src = src
        self.halo = None  # dict(neighbors, send_ptr, send_rows, recv_ptr, recv_rows)

    @property
    def n(self):
        return self.A.n

    def scatter_global(self, xg: np.ndarray) -> np.ndarray:
        b = self.A.b
        return np.ascontiguousarray(xg.reshape(-1, b)[self.l2g].reshape(-1))

    def owner_part(self, xl: np.ndarray) -> np.ndarray:
        b = self.A.b
        return xl.reshape(-1, b)[: self.n_interior]


def localize(A: BCSR, part: np.ndarray, rank: int) -> LocalSystem:
    """ghost-last local system of `rank` out of the global matrix (library call)"""
    L = _lib.lib()
    part = np.ascontiguousarray(part, np.int32)
    nl, ni, nz = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(L.opmb200_localize(A.n, A.rowptr, A.col, part, rank, C.byref(nl), C.byref(ni), C.byref(nz),
                                  None, None, None, None))
    l2g = np.zeros(nl.value, np.int32)
    rowptr = np.zeros(nl.value + 1, np.int32)
    col = np.zeros(nz.value, np.int32)
    src = np.zeros(nz.value, np.int64)
    _lib.check(L.opmb200_localize(A.n, A.rowptr, A.col, part, rank, C.byref(nl), C.byref(ni), C.byref(nz),
                                  l2g.ctypes.data, rowptr.ctypes.data, col.ctypes.data, src.ctypes.data))
    b = A.b
    val = np.where((src >= 0)[:, None, None], A.val[np.maximum(src, 0)], np.eye(b)[None])
    ls = LocalSystem(BCSR(rowptr, col, val), ni.value, l2g, src)
    ls.halo = build_halo(ls, lambda g: part[g], rank)
    return ls


def localize_rows(row0: int, rowptr, gcol, val, owner_of, rank: int) -> LocalSystem:
    """same structure from the owned rows [row0, row0+n) with GLOBAL column ids (numpy; used when
    every rank generates only its own slab)"""
    rowptr = np.asarray(rowptr, np.int64)
    gcol = np.asarray(gcol, np.int64)
    n_own = len(rowptr) - 1
    b = val.shape[-1]
    own = (gcol >= row0) & (gcol < row0 + n_own)
    ghosts = np.unique(gcol[~own])
    lcol = np.where(own, gcol - row0, n_own + np.searchsorted(ghosts, gcol))
    rows = np.repeat(np.arange(n_own, dtype=np.int64), np.diff(rowptr))
    order = np.lexsort((lcol, rows))
    ng = len(ghosts)
    n = n_own + ng
    rp = np.concatenate([rowptr, rowptr[-1] + 1 + np.arange(ng)]).astype(np.int32)
    cl = np.concatenate([lcol[order], n_own + np.arange(ng)]).astype(np.int32)
    vl = np.concatenate([val[order], np.broadcast_to(np.eye(b), (ng, b, b))])
    l2g = np.concatenate([row0 + np.arange(n_own), ghosts]).astype(np.int32)
    ls = LocalSystem(BCSR(rp, cl, vl), n_own, l2g)
    ls.halo = build_halo(ls, owner_of, rank)
    return ls


def build_halo(ls: LocalSystem, owner_of, rank: int):
    """owner/copy lists per neighbour, both sides ordered by ascending global index.  The send
    lists rely on the structural symmetry of Flow's Jacobians (an owner row is a copy on rank o
    iff it has a column owned by o)."""
    A, K = ls.A, ls.n_interior
    ghost_g = ls.l2g[K:].astype(np.int64)
    ghost_owner = np.asarray(owner_of(ghost_g), np.int64) if len(ghost_g) else np.zeros(0, np.int64)
    rows = A.row_of_entry()
    m = (rows < K) & (A.col >= K)
    pair_row = rows[m].astype(np.int64)
    pair_own = ghost_owner[A.col[m] - K] if m.any() else np.zeros(0, np.int64)
    neighbors = np.unique(ghost_owner)
    send_ptr, recv_ptr, send_rows, recv_rows = [0], [0], [], []
    for o in neighbors:
        r = K + np.nonzero(ghost_owner == o)[0]          # ghosts are stored in ascending global order
        s = np.unique(pair_row[pair_own == o])            # owners are stored in ascending global order
        recv_rows.append(r)
        send_rows.append(s)
        recv_ptr.append(recv_ptr[-1] + len(r))
        send_ptr.append(send_ptr[-1] + len(s))
    cat = lambda xs: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), np.int32)  # noqa: E731
    return dict(neighbors=np.ascontiguousarray(neighbors, np.int32), send_ptr=np.array(send_ptr, np.int32),
                send_rows=cat(send_rows), recv_ptr=np.array(recv_ptr, np.int32), recv_rows=cat(recv_rows))
