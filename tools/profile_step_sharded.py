"""Torch-profiler view of the row-sharded streaming step on rank 0 (run under torchrun, one rank per GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
from online_gp_b200 import settings as S
from online_gp_b200.parallel import Comm, ShardedOnlineSKIRegression

rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
dist.init_process_group("nccl", device_id=dev)
d, g, q, n_init, _lr, _ = bench.WORKLOADS["powerplant_4d_g32"]
x, y = bench.synth_stream(d, n_init + 64)
with S.max_root_decomposition_size(512), S.max_cholesky_size(2048), S.cg_tolerance(1e-2), S.sharded_dual_layout("--single" not in sys.argv):
    model = ShardedOnlineSKIRegression(x[:n_init].to(dev), y[:n_init].to(dev), lr=5e-3, grid_size=g, grid_bound=1.0, comm=Comm())
    xd, yd = x[n_init:].to(dev), y[n_init:].to(dev)
    def step(t):
        model.evaluate(xd[t:t + 1], yd[t:t + 1]); model.update(xd[t:t + 1], yd[t:t + 1])
    for t in range(4):
        step(t)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(4, 12):
        step(t)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("eager ms/step (no profiler): %.2f" % (e0.elapsed_time(e1) / 8))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for t in range(12, 16):
            step(t)
        torch.cuda.synchronize()
if rank == 0:
    ev = prof.key_averages()
    kern = [e for e in ev if e.self_device_time_total > 0 and e.self_cpu_time_total == 0]
    print("kernel-only device time per step: %.2f ms" % (sum(e.self_device_time_total for e in kern) / 4 / 1e3))
    for e in sorted(kern, key=lambda e: -e.self_device_time_total)[:24]:
        print("%-80s n/step=%6.1f  ms/step=%7.3f" % (e.key[:80], e.count / 4, e.self_device_time_total / 4 / 1e3))
    for e in sorted(ev, key=lambda e: -e.self_cpu_time_total)[:10]:
        print("CPU %-66s n/step=%6.1f  cpu ms/step=%7.3f" % (e.key[:66], e.count / 4, e.self_cpu_time_total / 4 / 1e3))
dist.barrier()
sys.stdout.flush()
os._exit(0)
