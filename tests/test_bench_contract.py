"""bench.py's driver contract, checked on CPU through the reference arm (`--impl reference`): one JSON line with the
keys the driver reads, the same-config / same-steps CPU run of the reference's file, and silence from non-zero ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl")


def _run(env_extra=None, args=()):
    env = dict(os.environ, **(env_extra or {}))
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--no-train", "--shape", "4,8,16,16", *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f32" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    from oracle import build_ref
    assert cb["kind"] == ("reference" if build_ref.available() else "port")      # oracle/_ref: the reference's own file
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["steps"] == 1 and d["warmup"] == 1                                  # --steps / --warmup are honoured
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
