"""include/opmb200/dune_adapter.hpp (the Dune-shaped C++ layer over the C ABI) compiled against the
stand-in headers of tests/cpp/stubs and run on the reference's matr33 fixture."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, coo_to_bcsr
from opm_simulators_b200 import matrixmarket

EXE = os.path.join(ROOT, "tests", "cpp", "test_dune_adapter")


def build_adapter_test():
    src = os.path.join(ROOT, "tests", "cpp", "test_dune_adapter.cpp")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "stubs", "dune_stubs.hpp"),
            os.path.join(ROOT, "include", "opmb200", "dune_adapter.hpp"), os.path.join(ROOT, "include", "opmb200.h")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", EXE, src,
                    "-L" + os.path.join(ROOT, "opm_simulators_b200"), "-lopmb200",
                    "-Wl,-rpath," + os.path.join(ROOT, "opm_simulators_b200")], check=True)


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
