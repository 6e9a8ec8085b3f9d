import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        return json.load(f)


def coo_to_bcsr(coo, b):
    from opm_simulators_b200.bcsr import BCSR

    return BCSR.from_scalar_coo(coo["nrows"] // b, b, np.array(coo["i"]) - 1, np.array(coo["j"]) - 1,
                                np.array(coo["v"]))


def pattern_to_bcsr(rows, b, rng=None, dominant=True):
    """rows: list of column lists (one per block row) -> BCSR with random, diagonally dominant blocks"""
    from opm_simulators_b200.bcsr import BCSR

    rng = rng or np.random.default_rng(0)
    r = np.concatenate([[i] * len(c) for i, c in enumerate(rows)]).astype(np.int64)
    c = np.concatenate([sorted(cc) for cc in rows]).astype(np.int64)
    blk = rng.uniform(-1, 1, (len(r), b, b))
    if dominant:
        blk[r == c] += 4.0 * np.eye(b)
    return BCSR.from_block_coo(len(rows), r, c, blk)


def rel_err(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
