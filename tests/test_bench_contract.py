"""The parts of bench.py's contract that need no GPU: the reference arm (the unmodified reference on
the host cores, or the restatement when it is not built) prints one JSON line with the keys the
driver reads, and ranks other than 0 of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys

from conftest import ROOT

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def run_bench(*args, env=None):
    full_env = dict(os.environ)
    full_env.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=full_env, timeout=600)


def test_reference_arm_prints_one_json_line():
    proc = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "400000")
    assert proc.returncode == 0, proc.stderr
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in REQUIRED:
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "suffixes/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["vs_baseline"] is None and line["dtype"] == "u8"
    assert line["config"]["workload"].startswith("genome3g")
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1 and base["value"] == line["value"]
    assert "400000" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "suffixes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    proc = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                     env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert proc.returncode == 0, proc.stderr
    assert proc.stdout.strip() == ""
