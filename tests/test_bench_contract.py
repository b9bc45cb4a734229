"""CPU-only: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints one JSON line with the
contract's keys; with RANK != 0 it prints nothing and exits 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ); env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                          capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_reference_arm_prints_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "x25519_shared_key_ops_per_sec" and d["unit"] == "ops/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    v = d["ed25519_verify"]                     # the second headline of BASELINE.json's metric, same arm
    assert v["metric"] == "ed25519_verify_ops_per_sec" and v["value"] > 0 and v["cpu_baseline"]["cores"] >= 1


def test_reference_arm_verify_metric():
    env = dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--metric", "ed25519_verify", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["metric"] == "ed25519_verify_ops_per_sec" and d["value"] > 0 and d["x25519_shared"]["value"] > 0


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
