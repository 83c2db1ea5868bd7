"""CPU: the reference arm of bench.py prints ONE JSON line with the keys the driver reads (no GPU needed)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--cpu-rows", "2000"],
                         check=True, capture_output=True, text=True, timeout=300).stdout.strip().splitlines()
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["n_gpus"] == 1
    assert line["metric"].startswith("ELBO+grad steps/sec") and "steps/s" in line["unit"]
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] >= 3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None and line["data"] == "synthetic"


def test_reference_arm_on_other_ranks_exits_quietly():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
