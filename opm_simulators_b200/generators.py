"""Synthetic black-oil-shaped Jacobians.

Sparsity rule of the reference's TPFA linearizer (opm/models/discretization/common/
tpfalinearizer.hh:513-528, 597-608): one block row per active cell in natural order (i fastest,
then j, then k), columns = self + face neighbours + NNC neighbours, sorted, diagonal present,
structurally symmetric.

Everything random is drawn from per-z-plane ``numpy.random.Philox`` streams keyed on
(seed, purpose, plane), so any z-slab of a grid can be generated on its own (one slab per GPU)
and is bit-identical to the same rows of the whole-grid matrix.
"""
from __future__ import annotations

import numpy as np

from .bcsr import BCSR

# named configurations of BASELINE.json (C1 is the tests/matr33.txt fixture)
CONFIGS = {
    "C2": dict(nx=46, ny=112, nz=22, b=3, seed=1002, sigma=2.0, kz_mult=0.1, n_active=44431, nnc=500),
    "C3": dict(nx=60, ny=220, nz=85, b=3, seed=1003, sigma=3.0, kz_mult=0.01),
    "C4": dict(nx=400, ny=400, nz=400, b=3, seed=1004, sigma=2.0, kz_mult=0.1),
    "C4slab": dict(nx=400, ny=400, nz=50, b=3, seed=1004, sigma=2.0, kz_mult=0.1),
    "C5": dict(nx=200, ny=200, nz=200, b=4, seed=1005, sigma=2.0, kz_mult=0.1),
    "C5slab": dict(nx=200, ny=200, nz=25, b=4, seed=1005, sigma=2.0, kz_mult=0.1),
}

_P_FACE, _P_PERT, _P_PORV, _P_RHS, _P_MASK, _P_NNC = range(6)


def _rng(seed, purpose, plane):
    return np.random.Generator(np.random.Philox(key=[(seed << 8) | purpose, plane & 0xFFFFFFFF]))


def mobility_template(b: int):
    """(M_up, M_dn): coupling blocks seen from the upstream / downstream cell of a face.
    Column 0 is pressure (strong, symmetric part); the remaining columns are saturation-like
    unknowns that only couple through the upstream cell (scaled by 0.3 on the other side)."""
    base = np.array([[1.00, 0.20, -0.05, 0.02],
                     [0.80, -0.15, 0.10, 0.03],
                     [0.60, 0.05, 0.25, -0.04],
                     [0.40, 0.02, -0.03, 0.30]])
    M_up = base[:b, :b].copy()
    M_dn = M_up.copy()
    M_dn[:, 1:] *= 0.3
    return M_up, M_dn


def blackoil_system(nx, ny, nz, b=3, seed=1000, sigma=2.0, kz_mult=1.0, z_range=None,
                    n_active=None, nnc=0, acc=(0.005, 1.0, 1.0, 1.0), pert=0.05, with_rhs=True):
    """Build rows of the cells in planes z_range=[z0,z1) (default: all) of an nx x ny x nz grid.

    Returns dict(A=BCSR with GLOBAL column indices restricted... see below, ...):
      * whole grid (z_range None): ``A`` is square over the active cells, ``rhs``, ``xstar``,
        ``rhs2`` are given, ``cell_index`` maps active cell -> natural grid index.
      * slab: ``rowptr, gcol, val`` hold the owned rows with global column ids, to be
        localised by :func:`opm_simulators_b200.partition.localize`.
    """
    z0, z1 = (0, nz) if z_range is None else z_range
    if (n_active is not None or nnc) and z_range is not None:
        raise ValueError("active masks / NNCs are only supported for whole-grid generation")
    npl = nx * ny
    M_up, M_dn = mobility_template(b)
    accv = np.asarray(acc[:b], dtype=np.float64)

    # ---- per-plane random fields (planes z0-1 .. z1 needed for the z faces) ----------------
    def face_fields(k):
        g = _rng(seed, _P_FACE, k).standard_normal((3, ny, nx))
        T = np.exp(sigma * g)
        T[2] *= kz_mult  # z face between plane k and k+1
        return T

    def pert_fields(k):
        # six directed connections per cell: -z -y -x +x +y +z
        return _rng(seed, _P_PERT, k).uniform(-1.0, 1.0, (6, ny, nx, b, b))

    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")  # [ny, nx]
    rows_l, cols_l, blks_l = [], [], []
    diag = np.zeros((z1 - z0, ny, nx, b, b))
    tsum = np.zeros((z1 - z0, ny, nx))
    Tprev = face_fields(z0 - 1) if z0 > 0 else None
    for k in range(z0, z1):
        T = face_fields(k)
        R = pert_fields(k)
        cell = (k * npl + jj * nx + ii).astype(np.int64)  # natural index [ny, nx]
        # (direction id, neighbour offset, transmissibility array, validity mask, upstream?)
        # a face's upstream side is its lower-index cell (fixed "flow direction" -> asymmetry)
        conns = []
        if k > 0:
            conns.append((0, -npl, Tprev[2], np.ones((ny, nx), bool), False))
        conns.append((1, -nx, np.roll(T[1], 1, axis=0), jj > 0, False))
        conns.append((2, -1, np.roll(T[0], 1, axis=1), ii > 0, False))
        conns.append((3, +1, T[0], ii < nx - 1, True))
        conns.append((4, +nx, T[1], jj < ny - 1, True))
        if k < nz - 1:
            conns.append((5, +npl, T[2], np.ones((ny, nx), bool), True))
        for d, off, Tf, ok, upstream in conns:
            Tf = np.where(ok, Tf, 0.0)
            Mo = M_dn if upstream else M_up   # block multiplying the NEIGHBOUR's unknowns
            Md = M_up if upstream else M_dn   # this cell's own contribution to its diagonal
            offblk = -Tf[..., None, None] * (Mo + pert * R[d])
            diag[k - z0] += Tf[..., None, None] * (Md + pert * R[d])
            tsum[k - z0] += Tf
            sel = ok
            rows_l.append(cell[sel])
            cols_l.append(cell[sel] + off)
            blks_l.append(offblk[sel])
        Tprev = T
        porv = _rng(seed, _P_PORV, k).uniform(0.5, 1.5, (ny, nx))
        # accumulation: compressibility-like on pressure, pore-volume-like on saturations
        diag[k - z0] += (tsum[k - z0] * porv)[..., None, None] * np.diag(accv)

    cells = (np.arange(z0, z1)[:, None, None] * npl + (jj * nx + ii)[None]).reshape(-1).astype(np.int64)
    rows = np.concatenate(rows_l + [cells])
    cols = np.concatenate(cols_l + [cells])
    blks = np.concatenate(blks_l + [diag.reshape(-1, b, b)])
    ncell = nx * ny * nz

    if z_range is not None:
        order = np.argsort(rows * ncell + cols, kind="stable")
        rows, cols, blks = rows[order], cols[order], blks[order]
        rowptr = np.zeros(len(cells) + 1, np.int64)
        np.add.at(rowptr, rows - z0 * npl + 1, 1)
        return dict(rowptr=np.cumsum(rowptr), gcol=cols, val=blks, row0=z0 * npl, nrows=len(cells),
                    nglobal=ncell, b=b)

    # ---- whole grid: optional active mask (column-wise pinch-out) and NNCs ------------------
    cell_index = np.arange(ncell, dtype=np.int64)
    if n_active is not None and n_active < ncell:
        # every (i,j) column keeps a contiguous k-range; trim columns until n_active cells remain
        rm = _rng(seed, _P_MASK, 0)
        keep = np.ones((nz, ny, nx), bool)
        top = rm.integers(nz // 2, nz + 1, (ny, nx))
        order_cols = rm.permutation(npl)
        excess = ncell - n_active
        for c in order_cols:
            if excess <= 0:
                break
            j, i = divmod(c, nx)
            cut = min(int(top[j, i]), excess)
            keep[nz - cut:, j, i] = False
            excess -= cut
        active = keep.reshape(-1)
        new_id = np.full(ncell, -1, np.int64)
        new_id[active] = np.arange(active.sum())
        selc = active[rows] & active[cols]
        # drop couplings to inactive cells; their share of the diagonal stays (boundary-like)
        rows, cols, blks = new_id[rows[selc]], new_id[cols[selc]], blks[selc]
        cell_index = np.nonzero(active)[0]
        ncell = int(active.sum())
    if nnc:
        rn = _rng(seed, _P_NNC, 0)
        a = rn.integers(0, ncell, 4 * nnc)
        c = rn.integers(0, ncell, 4 * nnc)
        ok = np.abs(a - c) > 1
        pairs = np.unique(np.stack([np.minimum(a, c)[ok], np.maximum(a, c)[ok]], 1), axis=0)[:nnc]
        Tn = np.exp(sigma * rn.standard_normal(len(pairs))) * 0.1
        Rn = rn.uniform(-1, 1, (2, len(pairs), b, b))
        lo, hi = pairs[:, 0], pairs[:, 1]
        rows = np.concatenate([rows, lo, hi, lo, hi])
        cols = np.concatenate([cols, hi, lo, lo, hi])
        blks = np.concatenate([blks,
                               -Tn[:, None, None] * (M_dn + pert * Rn[0]),
                               -Tn[:, None, None] * (M_up + pert * Rn[1]),
                               Tn[:, None, None] * (M_up + pert * Rn[0]),
                               Tn[:, None, None] * (M_dn + pert * Rn[1])])
    A = BCSR.from_block_coo(ncell, rows, cols, blks)
    out = dict(A=A, cell_index=cell_index, b=b, dims=(nx, ny, nz))
    if with_rhs:
        rr = _rng(seed, _P_RHS, 0)
        xstar = rr.uniform(-1.0, 1.0, ncell * b)
        out["xstar"] = xstar
        out["rhs"] = A.to_scipy() @ xstar
        out["rhs2"] = rr.standard_normal(ncell * b)
    return out


def config(name: str, scale: float = 1.0, **over):
    """Named BASELINE.json configuration; ``scale`` < 1 shrinks every grid dimension (tests)."""
    p = dict(CONFIGS[name])
    p.update(over)
    if scale != 1.0:
        for k in ("nx", "ny", "nz"):
            p[k] = max(2, int(round(p[k] * scale)))
        if p.get("n_active"):
            p["n_active"] = int(p["nx"] * p["ny"] * p["nz"] * 0.39)
        if p.get("nnc"):
            p["nnc"] = max(1, int(p["nnc"] * scale ** 3))
    return blackoil_system(**p)


def laplace_like(n_side: int, b: int, rng, dims: int = 2, asym: float = 0.0) -> BCSR:
    """small random block matrices on a 5-/7-point pattern for unit tests (diagonally dominant)."""
    shape = (n_side,) * dims
    n = n_side ** dims
    idx = np.arange(n).reshape(shape)
    rows, cols = [np.arange(n)], [np.arange(n)]
    for ax in range(dims):
        lo = np.take(idx, np.arange(n_side - 1), axis=ax).reshape(-1)
        hi = np.take(idx, np.arange(1, n_side), axis=ax).reshape(-1)
        rows += [lo, hi]
        cols += [hi, lo]
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    blks = rng.uniform(-1, 1, (len(rows), b, b))
    isdiag = rows == cols
    blks[isdiag] += (2.0 * dims + 1.0) * np.eye(b)
    if asym:
        blks[~isdiag & (rows < cols)] *= (1.0 + asym)
    return BCSR.from_block_coo(n, rows, cols, blks)


def standard_wells(A: BCSR, n_wells: int = 8, perfs: int = 12, dim_wells: int | None = None, seed: int = 7,
                   strength: float = 0.2, shared_cells: int = 0):
    """Synthetic StandardWell equations for the operator A - sum_w C_w^T D_w^-1 B_w
    (wells/StandardWellEquations.hpp: duneB_, duneC_ with one dim_wells x b block per perforation, invDuneD_).

    Every well perforates `perfs` cells of a vertical-ish line of rows (stride = a random step), wells may share
    `shared_cells` cells (two wells perforating one cell, which the reference allows).  The blocks are scaled by
    `strength` relative to the perforated cell's diagonal block, and C = -(B + 20 % noise) with D close to 2 I, so that
    -C^T D^-1 B is close to a positive semi-definite addition: A - C^T D^-1 B stays as well conditioned as A (a random
    sign pattern makes the combined operator indefinite and BiCGSTAB chaotic: 1e-15 on the rhs moved x by 7 %).
    Returns dict(ptr, cells, B, C, Dinv) in the layout of include/opmb200.h (opmb200_set_wells)."""
    rng = np.random.default_rng(seed)
    b = A.b
    dw = b + 1 if dim_wells is None else dim_wells  # numWellEq = numEq + 1 for the black-oil StandardWell
    n = A.n
    ptr, cells = [0], []
    for w in range(n_wells):
        start = int(rng.integers(0, max(1, n - 1)))
        step = int(rng.integers(1, max(2, n // (4 * perfs) + 2)))
        c = [(start + k * step) % n for k in range(perfs if w % 3 else max(1, perfs // 2))]
        c = list(dict.fromkeys(c))  # a well perforates a cell once
        if shared_cells and w > 0 and cells:
            for extra in cells[:shared_cells]:
                if extra not in c:
                    c.append(extra)
        cells += c
        ptr.append(len(cells))
    cells = np.asarray(cells, np.int32)
    n_perf = len(cells)
    diag = np.empty((n_perf,))
    for k, cell in enumerate(cells):
        row = slice(A.rowptr[cell], A.rowptr[cell + 1])
        d = row.start + int(np.searchsorted(A.col[row], cell))
        diag[k] = np.abs(A.val[d]).max()
    scale = np.sqrt(strength * diag)[:, None, None]
    Bm = rng.uniform(-1.0, 1.0, (n_perf, dw, b)) * scale
    Cm = -(Bm + 0.2 * rng.uniform(-1.0, 1.0, (n_perf, dw, b)) * scale)
    D = rng.uniform(-0.3, 0.3, (n_wells, dw, dw)) + np.eye(dw) * 2.0
    Dinv = np.linalg.inv(D)
    return dict(ptr=np.asarray(ptr, np.int32), cells=cells, B=np.ascontiguousarray(Bm), C=np.ascontiguousarray(Cm),
                Dinv=np.ascontiguousarray(Dinv))
