"""include/opmb200/dune_adapter.hpp (the Dune-shaped C++ layer over the C ABI) compiled against the
stand-in headers of tests/cpp/stubs and run on the reference's matr33 fixture."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, coo_to_bcsr
from opm_simulators_b200 import matrixmarket

EXE = os.path.join(ROOT, "tests", "cpp", "test_dune_adapter")


EXE_PAR = os.path.join(ROOT, "tests", "cpp", "test_dune_adapter_par")


def _build(exe, source):
    src = os.path.join(ROOT, "tests", "cpp", source)
    deps = [src, os.path.join(ROOT, "tests", "cpp", "stubs", "dune_stubs.hpp"),
            os.path.join(ROOT, "include", "opmb200", "dune_adapter.hpp"), os.path.join(ROOT, "include", "opmb200.h")]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe, src,
                    "-L" + os.path.join(ROOT, "opm_simulators_b200"), "-lopmb200", "-pthread",
                    "-Wl,-rpath," + os.path.join(ROOT, "opm_simulators_b200")], check=True)


def build_adapter_test():
    _build(EXE, "test_dune_adapter.cpp")
    _build(EXE_PAR, "test_dune_adapter_par.cpp")


def _write_rank_files(tmp_path, A, part, world, rhs_global=None):
    """every rank's ghost-last local system + what a Dune::OwnerOverlapCopyCommunication of that rank knows:
    attribute per local index (owner 1 / copy 3) and, per other rank, the (global index, attribute there) pairs"""
    from opm_simulators_b200 import partition

    locs = [partition.localize(A, part, r) for r in range(world)]
    for r, ls in enumerate(locs):
        pre = str(tmp_path / f"r{r}_")
        ls.A.rowptr.astype(np.int32).tofile(pre + "rowptr.i32")
        ls.A.col.astype(np.int32).tofile(pre + "col.i32")
        ls.A.val.astype(np.float64).tofile(pre + "val.f64")
        ls.l2g.astype(np.int32).tofile(pre + "l2g.i32")
        attr = np.where(np.arange(ls.n) < ls.n_interior, 1, 3).astype(np.int32)
        attr.tofile(pre + "attr.i32")
        peers = [world - 1]
        for o, lo in enumerate(locs):
            if o == r:
                continue
            order = np.argsort(lo.l2g)
            ga = np.stack([lo.l2g[order], np.where(order < lo.n_interior, 1, 3)], 1).astype(np.int32)
            peers += [o, len(ga)] + ga.reshape(-1).tolist()
        np.array(peers, np.int32).tofile(pre + "peers.i32")
        if rhs_global is not None:
            b = ls.scatter_global(rhs_global)
            b.reshape(-1, A.b)[ls.n_interior:] = 0.0
            b.tofile(pre + "rhs.f64")
    return locs


def test_adapter_compiles_against_dune_shaped_headers():
    build_adapter_test()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_adapter_runs_reference_flexiblesolver_test(golden, tmp_path):
    build_adapter_test()
    A = coo_to_bcsr(golden["matr33"], 3)
    matrixmarket.write_matrix(str(tmp_path / "matr33.txt"), A)
    matrixmarket.write_vector(str(tmp_path / "rhs3.txt"), np.array(golden["rhs3"]), 3)
    import json
    with open(tmp_path / "options.json", "w") as f:
        json.dump(golden["options_flexiblesolver_1x1"], f)
    r = subprocess.run([EXE, str(tmp_path / "matr33.txt"), str(tmp_path / "rhs3.txt"), str(tmp_path / "options.json")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


def test_flatten_halo_equals_the_partition_halo_lists(tmp_path):
    """Opm::b200::flattenHalo on a Dune-shaped communication object == the owner/copy lists of partition.build_halo
    (what the Python tests hand to opmb200_create); host only"""
    from opm_simulators_b200 import generators, partition

    build_adapter_test()
    A = generators.blackoil_system(5, 4, 9, b=3, seed=3, with_rhs=False)["A"]
    world = 3
    part = partition.partition_simple(A.n, world)
    locs = _write_rank_files(tmp_path, A, part, world)
    for r, ls in enumerate(locs):
        res = subprocess.run([EXE_PAR, str(tmp_path), str(r), str(world), "flatten"], capture_output=True, text=True, timeout=60)
        assert res.returncode == 0, res.stdout + res.stderr
        got = {}
        for line in open(tmp_path / f"r{r}_halo.txt"):
            k, *v = line.split()
            got[k] = [int(x) for x in v]
        assert got["interior"] == [ls.n_interior]
        for k in ("neighbors", "send_ptr", "send_rows", "recv_ptr", "recv_rows"):
            assert got[k] == ls.halo[k].tolist(), k


@pytest.mark.gpu
def test_two_ranks_through_the_cpp_adapter(tmp_path):
    """Solver(op, comm, nccl, json) of dune_adapter.hpp on two GPUs (one process per rank, NCCL id handed round through
    a file where Flow would MPI_Bcast): iterations and solution against the oracle emulating the same two ranks"""
    import json

    import torch

    from opm_simulators_b200 import generators, partition
    from oracle import oracle as orc

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    build_adapter_test()
    s = generators.blackoil_system(7, 6, 8, b=3, seed=321, sigma=1.0, kz_mult=0.5)
    A, world = s["A"], 2
    part = partition.partition_simple(A.n, world)
    locs = _write_rank_files(tmp_path, A, part, world, s["rhs2"])
    with open(tmp_path / "options.json", "w") as f:
        json.dump({"solver": "bicgstab", "tol": "1e-8", "maxiter": "200", "verbosity": "0",
                   "preconditioner": {"type": "dilu"}}, f)
    procs = [subprocess.Popen([EXE_PAR, str(tmp_path), str(r), str(world), "solve"], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    ps = orc.ParSystem([dict(rowptr=l.A.rowptr, col=l.A.col, val=l.A.val, interior=l.n_interior, l2g=l.l2g) for l in locs],
                       nglobal=A.n)
    ps.prec_update("dilu")
    bs = [l.scatter_global(s["rhs2"]) for l in locs]
    for l, bb in zip(locs, bs):
        bb.reshape(-1, 3)[l.n_interior:] = 0.0
    xo, _, ro, _ = ps.bicgstab(bs, tol=1e-8, maxiter=200)
    for r, l in enumerate(locs):
        x = np.fromfile(tmp_path / f"r{r}_x.f64")
        it, conv = [int(v) for v in open(tmp_path / f"r{r}_result.txt").read().split()]
        assert conv == 1 and abs(it - ro["iterations"]) <= 1
        ref = xo[r].reshape(-1, 3)[: l.n_interior].reshape(-1)
        assert np.linalg.norm(x - ref) <= 1e-6 * np.linalg.norm(ref)
