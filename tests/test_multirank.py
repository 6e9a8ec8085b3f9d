"""world_size-2 tests: `gloo` on CPU for the host logic of the N>1 path, NCCL on >= 2 GPUs for the
halo exchange / all-reduce / block-Jacobi path (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "mgpu_worker.py")


def launch(nproc, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + extra
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)


def test_two_rank_host_logic_gloo():
    r = launch(2, ["--mode", "cpu"], 29611)
    assert r.returncode == 0 and "MGPU_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("collectives", ["nccl", "p2p"])
@pytest.mark.parametrize("prec,b", [("dilu", 3), ("ilu0", 3), ("dilu", 4)])
def test_two_rank_block_jacobi(prec, b, collectives):
    """halo exchange + all-reduce over NCCL, and over the library's own peer-memory kernels"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = launch(2, ["--mode", "gpu", "--prec", prec, "--b", str(b), "--collectives", collectives],
               29620 + b + (10 if collectives == "p2p" else 0))
    assert r.returncode == 0 and "MGPU_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", ["tiles", "levels"])
@pytest.mark.parametrize("prec", ["dilu", "ilu0"])
def test_two_rank_block_jacobi_both_schedules(prec, schedule):
    """the tile walkers (the default on box grids) and the level schedule on local systems with ghost rows (4th upper
    entry of the first plane, ghost rows ParallelOverlappingILU0 leaves alone)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = launch(2, ["--mode", "gpu", "--prec", prec, "--b", "3", "--collectives", "p2p", "--schedule", schedule],
               29660 + (1 if prec == "ilu0" else 0) + (2 if schedule == "levels" else 0))
    assert r.returncode == 0 and "MGPU_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("collectives", ["nccl", "p2p"])
def test_two_rank_block_jacobi_serialised_halo(collectives):
    """b200.halo_overlap = 0: the halo copy and the SpMV on one stream (the default runs the copy beside the SpMV of the
    rows that read no ghost value; every other two-rank test covers that)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    r = launch(2, ["--mode", "gpu", "--prec", "ilu0", "--b", "3", "--collectives", collectives, "--halo-overlap", "0"],
               29680 + (1 if collectives == "p2p" else 0))
    assert r.returncode == 0 and "MGPU_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
