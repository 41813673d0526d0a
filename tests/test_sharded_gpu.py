"""GPU parity of the row-sharded multi-GPU path: the sharded model on N ranks (NCCL, peer-memory exchange, CUDA-graph
replay) streams the same points as the single-GPU model and must agree step by step (tools/sharded_parity.py, fp32
1e-2 / fp64 1e-4).  Runs under ``pytest -m gpu`` whenever the box shows at least 2 GPUs (N = 2, 4, 8 as available;
default and dual-layout exchange); skipped on a 1-GPU box.  ``gpurun --gpus 2 -- 'python -m pytest tests -m gpu -k sharded'``."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        return torch.cuda.device_count()
    except Exception:       # noqa: BLE001
        return 0


CASES = [
    # (world, extra args): 32^4 fp32 (fused tensor-core pair kernels + chunked layouts) and a generic 3-D grid in fp64
    (2, ["--n0", "48", "--steps", "5", "--graphs"]),
    (2, ["--n0", "48", "--steps", "5", "--graphs", "--dual"]),
    (2, ["--dims", "3", "--grid", "16", "--n0", "40", "--steps", "4", "--dtype", "f64"]),
    (2, ["--n0", "40", "--steps", "3", "--q", "3"]),
    (4, ["--n0", "48", "--steps", "5", "--graphs"]),
    (4, ["--n0", "48", "--steps", "4", "--graphs", "--dual"]),
    (8, ["--n0", "48", "--steps", "5", "--graphs"]),
    (8, ["--n0", "48", "--steps", "4", "--graphs", "--dual"]),
]


@pytest.mark.parametrize("world,extra", CASES, ids=[f"n{w}-" + "_".join(a.strip("-") for a in e if a.startswith("--")) for w, e in CASES])
def test_sharded_stream_matches_single_gpu(world, extra):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs, {_ngpu()} visible")
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "sharded_parity.py"), "--watchdog", "150",
           *extra]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0 and lines, (out.stdout[-1500:], out.stderr[-3000:])
    z = json.loads(lines[-1])
    assert z["parity"] is True and z["world"] == world, z
    if "--graphs" in extra:
        assert z["cuda_graphs"] is True and z["graph_replays"] > 0, z
