"""CPU: the reference arm of bench.py (the oracle port timed on host cores) prints one JSON line with the contract's
keys; run on BASELINE config 1 (1-D, 128-point grid), the reference's own CPU-runnable case."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "synthetic_1d_g128", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    z = json.loads(lines[0])
    assert z["impl"] == "reference" and z["metric"] == "wiski_streaming_updates_per_sec" and z["unit"] == "updates/s"
    assert z["higher_is_better"] is True and z["value"] > 0 and z["steps"] == 2
    assert z["config"]["workload"] == "synthetic_1d_g128" and z["config"]["m"] == 128
    assert z["cpu_baseline"]["kind"] == "port" and z["cpu_baseline"]["cores"] >= 1 and z["cpu_baseline"]["value"] == z["value"]
    assert z["e2e"] == {"value": z["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert z["gpu_launches"] == 0


def test_gpu_arm_refuses_to_run_without_cuda():
    """No CPU fallback: without a CUDA device the product arm must fail loudly, not time the oracle."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
