"""bench.py's reference arm runs on CPU (the oracle port on the host cores): the JSON contract of that line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                                   "cfg1", "--steps", "1", "--warmup", "1"], cwd=ROOT, timeout=600).decode()
    line = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "wgan_gp_train_samples_per_sec"
    assert line["unit"] == "samples/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "cfg1" in line["config"]["workload"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                                   "--workload", "cfg1"], cwd=ROOT, env=env, timeout=120).decode()
    assert out.strip() == ""
