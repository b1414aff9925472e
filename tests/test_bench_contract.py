"""The reference arm of bench.py (`--impl reference`) is the first thing the driver runs at round end and needs no GPU: run it
here on a tiny budget (the bounded-sample path) and check the JSON line against the contract; ranks > 0 of a torchrun launch
must exit 0 without output."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", env.get("WORLD_SIZE", "1"),
                           "--steps", "1", "--warmup", "0", "--ref-budget-seconds", "1"], capture_output=True, text=True, timeout=900,
                          env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "frames/sec SegNet(T)+ORB" and line["unit"] == "frames/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 0
    assert abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    assert line["config"]["workload"].startswith("Bayesian SegNet Basic T=6 + ORB(2000) x2")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "T=2 of 6 samples" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout[-500:], r.stderr[-500:])
