"""bench.py's reference arm on the host (no GPU): the JSON line carries the contract's keys and names the same workload
as the GPU arm (the driver compares the two `config.workload` strings)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, ROOT)
    import bench
    from opm_simulators_b200 import generators

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    c = generators.CONFIGS["C3"]
    n = c["nx"] * c["ny"] * c["nz"]
    assert line["config"]["workload"] == bench.workload_string("C3", (c["nx"], c["ny"], c["nz"]), n, "dilu", 0.01)
