"""Host-only replay of the tile walkers' step tables (opmb200_plan_tiles): a scalar Python interpreter walks
the chunks in ticket order, step by step, with a ring of `ring` positions per chunk and a global record
array for the dependencies the ring does not serve -- exactly the data flow of tw_sweep_kernel
(opm_simulators_b200/csrc/tile_kernels.cuh) -- and must reproduce the oracle's DILU apply.  This pins the
integer side of the schedule (codes, ring aliasing, external lists, chunk order) without a GPU."""
import ctypes as C

import numpy as np
import pytest

from conftest import pattern_to_bcsr, rel_err
from opm_simulators_b200 import _lib, generators, partition
from oracle import oracle as orc

RING, EXT = 1 << 30, 1 << 29


def plan_tiles(A, schedule=1, chunk_rows=0, n_interior=None, direction=0):
    n = A.n
    ni = n if n_interior is None else n_interior
    info = np.zeros(8, np.int32)
    r2n = np.zeros(max(n, 1), np.int32)
    f = _lib.lib().opmb200_plan_tiles
    _lib.check(f(A.b, n, A.nnzb, A.rowptr, A.col, ni, schedule, chunk_rows, info.ctypes.data, r2n.ctypes.data,
                 None, None, None, direction, None, None, None))
    P = dict(schedule=int(info[0]), R=int(info[1]), ring=int(info[2]), S=(int(info[3]), int(info[4])),
             n_steps=int(info[5]), n_chunks=int(info[6]), chunk_rows=int(info[7]), r2n=r2n[:n])
    if P["schedule"] != 1:
        return P
    S, RP = P["S"][direction], (P["R"] + 3) & ~3
    step_first = np.zeros(P["n_steps"] + 1, np.int32)
    chunk_first = np.zeros(P["n_chunks"] + 1, np.int32)
    flags = np.zeros(P["n_steps"], np.int32)
    codes = np.zeros(P["n_steps"] * S * RP, np.int32)
    ext = np.zeros(P["n_steps"] * 64, np.int32)
    n_ext = np.zeros(P["n_steps"], np.int32)
    _lib.check(f(A.b, n, A.nnzb, A.rowptr, A.col, ni, schedule, chunk_rows, info.ctypes.data, r2n.ctypes.data,
                 step_first.ctypes.data, chunk_first.ctypes.data, flags.ctypes.data, direction, codes.ctypes.data,
                 ext.ctypes.data, n_ext.ctypes.data))
    P.update(step_first=step_first, chunk_first=chunk_first, flags=flags, codes=codes.reshape(P["n_steps"], S, RP),
             ext=ext.reshape(P["n_steps"], 64), n_ext=n_ext, RP=RP)
    return P


def replay_dilu(A, Dinv, d, n_interior=None, **kw):
    """DILU apply (DILU.hpp:253-304) through the step tables -> v in natural order"""
    n, b = A.n, A.b
    ni = n if n_interior is None else n_interior
    diag = A.diag_index()
    d = d.reshape(n, b)
    y = np.full((n, b), np.nan)  # by position: "global records"
    v = np.full((n, b), np.nan)
    for direction in (0, 1):
        P = plan_tiles(A, n_interior=n_interior, direction=direction, **kw)
        assert P["schedule"] == 1
        r2n, S, ring_n = P["r2n"], P["S"][direction], P["ring"]
        out = y if direction == 0 else v
        chunks = range(P["n_chunks"]) if direction == 0 else range(P["n_chunks"] - 1, -1, -1)
        for c in chunks:
            ring = np.full((ring_n, b), np.nan)
            steps = range(P["chunk_first"][c], P["chunk_first"][c + 1])
            for st in (steps if direction == 0 else reversed(steps)):
                q0, q1 = P["step_first"][st], P["step_first"][st + 1]
                assert 0 < q1 - q0 <= P["R"] and P["n_ext"][st] <= 64
                extv = np.array([out[p] for p in P["ext"][st][: P["n_ext"][st]]]).reshape(-1, b)
                assert not np.isnan(extv).any(), "an external dependency is produced later in ticket order"
                res = np.zeros((q1 - q0, b))
                for rho, q in enumerate(range(q0, q1)):
                    i = r2n[q]
                    ghost = i >= ni
                    assert bool(P["flags"][st] & 1) == ghost
                    if direction == 0:
                        ents = [] if ghost else list(range(A.rowptr[i], diag[i]))
                    else:
                        ents = [] if ghost else list(range(A.rowptr[i + 1] - 1, diag[i], -1))
                    assert len(ents) <= S
                    acc = np.zeros(b) if direction else d[i].copy()
                    for k in range(S):
                        code = P["codes"][st, k, rho]
                        if k >= len(ents):
                            assert code == RING | ring_n  # "no dependency": the zero record behind the ring
                            continue
                        if code & RING:
                            x = ring[code & (ring_n - 1)]
                        else:
                            assert code & EXT
                            x = extv[code & 63]
                        assert not np.isnan(x).any()
                        if direction:
                            acc += A.val[ents[k]] @ x
                        else:
                            acc -= A.val[ents[k]] @ x
                    Di = np.eye(b) if ghost else Dinv[i]
                    res[rho] = Di @ acc if direction == 0 else y[q] - Di @ acc
                for rho, q in enumerate(range(q0, q1)):  # the step's stores happen after all its loads
                    ring[q & (ring_n - 1)] = res[rho]
                    out[q] = res[rho]
    vn = np.zeros((n, b))
    vn[P["r2n"]] = v
    return vn.reshape(-1)


CASES = [
    ("box b3", lambda: generators.blackoil_system(9, 13, 11, b=3, seed=3, with_rhs=False)["A"], {}),
    ("box b4", lambda: generators.blackoil_system(7, 9, 10, b=4, seed=4, with_rhs=False)["A"], {}),
    ("box b2", lambda: generators.laplace_like(9, 2, np.random.default_rng(1), dims=3, asym=0.2), {}),
    ("box b1", lambda: generators.laplace_like(12, 1, np.random.default_rng(2), dims=3, asym=0.2), {}),
    ("box b3 contiguous 64", lambda: generators.blackoil_system(9, 8, 7, b=3, seed=11, with_rhs=False)["A"],
     {"chunk_rows": 64}),
    ("box b3 strips", lambda: generators.blackoil_system(6, 21, 5, b=3, seed=5, with_rhs=False)["A"],
     {"chunk_rows": -2002}),
    ("c2-like irregular", lambda: generators.config("C2", scale=0.3, with_rhs=False)["A"], {"chunk_rows": 256}),
    ("2d b3", lambda: generators.laplace_like(23, 3, np.random.default_rng(3), dims=2, asym=0.3), {}),
]


@pytest.mark.parametrize("name,make,kw", CASES, ids=[c[0] for c in CASES])
def test_step_tables_replay_the_dilu_apply(name, make, kw):
    A = make()
    if max(np.diff(A.rowptr)) > 9:
        pytest.skip("pattern too wide for the tile walkers")
    Dinv = orc.dilu_update(A.rowptr, A.col, A.val)
    d = np.random.default_rng(7).standard_normal(A.n * A.b)
    P = plan_tiles(A, **kw)
    if P["schedule"] != 1:
        pytest.skip("the analysis kept the level schedule (rows wider than the tile walkers' slots)")
    v = replay_dilu(A, Dinv, d, **kw)
    assert rel_err(v, orc.dilu_apply(A.rowptr, A.col, A.val, Dinv, d)) < 1e-13


def test_auto_schedule_picks_tiles_on_box_grids_only():
    box = generators.blackoil_system(12, 20, 8, b=3, seed=3, with_rhs=False)["A"]
    P = plan_tiles(box, schedule=2)
    assert P["schedule"] == 1 and P["chunk_rows"] < 0 and P["R"] == 32 and P["S"] == (3, 3)
    irregular = generators.config("C2", scale=0.3, with_rhs=False)["A"]
    assert plan_tiles(irregular, schedule=2)["schedule"] == 0
    rng = np.random.default_rng(8)
    dense = np.eye(60, dtype=bool)
    for i in range(60):
        for j in rng.choice(60, 6, replace=False):
            dense[i, j] = dense[j, i] = True
    from opm_simulators_b200.bcsr import BCSR
    wide = BCSR.from_dense_pattern(dense, 3, rng=rng)
    assert plan_tiles(wide, schedule=1)["schedule"] == 0  # rows wider than 4 slots: level schedule


def test_step_tables_on_a_ghost_last_local_system():
    """a rank's slab: ghost rows behind the owners, the plane below becomes a 4th UPPER entry"""
    s = generators.blackoil_system(6, 7, 9, b=3, seed=9, with_rhs=False)
    A = s["A"]
    part = partition.partition_simple(A.n, 3)
    loc = partition.localize(A, part, 1)
    B = loc.A
    vals = orc.make_overlap_rows_invalid(B.rowptr, B.col, B.val, loc.n_interior)
    P = plan_tiles(B, n_interior=loc.n_interior, direction=1)
    assert P["schedule"] == 1 and P["S"] == (3, 4)
    from opm_simulators_b200.bcsr import BCSR
    Bi = BCSR(B.rowptr, B.col, vals)
    # the oracle's serial DILU on the local matrix with identity ghost rows == the rank's local apply
    Dinv = orc.dilu_update(Bi.rowptr, Bi.col, Bi.val)
    d = np.random.default_rng(5).standard_normal(B.n * 3)
    v = replay_dilu(Bi, Dinv, d, n_interior=loc.n_interior)
    assert rel_err(v, orc.dilu_apply(Bi.rowptr, Bi.col, Bi.val, Dinv, d)) < 1e-13
