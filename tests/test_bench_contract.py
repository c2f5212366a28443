"""The benchmark contract on the CPU side: `bench.py --impl reference` (the reference arm: the oracle port timed on the
host cores) prints ONE JSON line with the agreed keys, for the default workload and for the FFA workload."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("workload", [None, "c2", "ffa", "zn1", "orb"])
def test_reference_arm_json_line(workload):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"]
    if workload:
        cmd += ["--workload", workload]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    if workload is None:  # the default job is the north-star configuration, strong scaling, and both arms build
        # their `config` object with the same function from the same arguments
        c = d["config"]
        assert d["scaling"] == "strong" and c["scaling"] == "strong" and c["nharm"] == 4 and c["nbin"] == 128
        assert c["n_fdot"] == 64 and c["n_events"] == 9_999_999 and c["n_freq_total"] == 16384
        sys.path.insert(0, ROOT)
        import bench

        assert c == bench.fold_config(bench.WORKLOADS["c3"][0], 128, 4, 9_999_999, 16384, 64, 1, "strong")
