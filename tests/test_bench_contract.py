"""CPU: bench.py's reference arm prints exactly one JSON line on stdout with the contract's keys (the GPU arm shares the
printing code; it is exercised on the B200 by the driver)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


import pytest


@pytest.mark.parametrize("arm", ["reference", "port"])
def test_reference_arm_prints_one_json_line(arm):
    """kind "reference": the unmodified tree staged under baseline/_ref (tools/stage_reference.sh) runs through its own
    BaseVideoModel(cfg)(task_dict); kind "port": the oracle restatement, the fallback when the tree is not staged."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    if arm == "reference":
        if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "models", "base")):
            pytest.skip("reference tree not staged (tools/stage_reference.sh)")
    else:
        env["FSAR_REF_ROOT"] = os.path.join(ROOT, "baseline", "_absent")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d" % len(lines)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "episodes/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("episodes/sec (5-way 1-shot")
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == arm and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
