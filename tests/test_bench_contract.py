"""bench.py contract on the CPU: the reference arm prints ONE JSON line with the keys the driver
reads (metric / value / unit / n_gpus / steps / warmup / ms_per_step / higher_is_better / scaling /
config.workload / impl / cpu_baseline / e2e).  The sample is shrunk to one lead time on one process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, NBH_BENCH_REF_PROCS="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["scaling"] == "strong"
    assert rec["metric"] == "neighbourhood gridpoint-evals/s" and rec["unit"] == "gridpoint-evals/s"
    assert rec["value"] > 0 and rec["ms_per_step"] > 0 and rec["steps"] == 1 and rec["warmup"] == 0
    assert "workload" in rec["config"] and rec["config"]["shape"] == [18, 36, 970, 1042]
    cb = rec["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] == 1 and cb["value"] == rec["value"] and cb["sample"]
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0,
                          "d2h_bytes_per_step": 0}
    assert rec["gpu_launches"] == 0
