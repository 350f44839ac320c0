"""bench.py's output contract, checked on the CPU-runnable arm: `--impl reference` (the oracle port on the host cores) must
print exactly ONE JSON line on stdout with the keys the driver reads, and the algorithmic constants bench.py derives
 must match the survey's figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--config", "cfg1", "--steps", "2", "--warmup", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("cfg1")


def test_reference_arm_runs_on_rank_zero_only(monkeypatch):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--config", "cfg1", "--gpus", "2"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_flops_and_bytes_match_the_survey():
    # cfg2 F = 83.15 GFLOP per sample, 6 * sum(A) = 132.1 MB per sample; cfg1 4F = 0.064 GF; cfg3 4F = 0.51 GF
    assert abs(bench.rb_flops_per_sample(bench.CONFIGS["cfg2"]) / 1e9 - 83.15) < 0.01
    assert abs(bench.elementwise_bytes_per_sample(bench.CONFIGS["cfg2"]) / 1e6 - 132.1) < 0.1
    assert abs(4 * bench.rb_flops_per_sample(bench.CONFIGS["cfg1"]) / 1e9 - 0.064) < 0.002
    assert abs(4 * bench.rb_flops_per_sample(bench.CONFIGS["cfg3"]) / 1e9 - 0.51) < 0.01
    assert abs(4 * bench.rb_flops_per_sample(bench.CONFIGS["cfg5"]) / 1e9 - 8.76) < 0.01
