"""Torch-profiler view of one streaming step on the bench workload: kernel time vs wall time, op counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

args = type("A", (), {})()
d, g, q, n_init, _lr, _ = bench.WORKLOADS["powerplant_4d_g32"]
dev = torch.device("cuda:0")
from online_gp_b200 import settings as S
model, xs, ys = bench.build_model(d, g, n_init, _lr, torch.float32, dev)
xd, yd = xs.to(dev), ys.to(dev)
with S.max_root_decomposition_size(512), S.max_cholesky_size(2048), S.cg_tolerance(1e-2):
    for t in range(4):
        bench.one_step(model, xd[t:t + 1], yd[t:t + 1])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for t in range(4, 8):
            bench.one_step(model, xd[t:t + 1], yd[t:t + 1])
        torch.cuda.synchronize()
ev = prof.key_averages()
tot_cuda = sum(e.self_device_time_total for e in ev)
print("steps=4 total kernel time per step: %.2f ms" % (tot_cuda / 4 / 1e3))
rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:22]
for e in rows:
    print("%-70s n/step=%6.1f  cuda ms/step=%7.3f" % (e.key[:70], e.count / 4, e.self_device_time_total / 4 / 1e3))
ncu = sum(e.count for e in ev if e.self_device_time_total > 0)
print("device-side events per step:", ncu / 4)
cpu_rows = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:12]
for e in cpu_rows:
    print("CPU %-66s n/step=%6.1f  cpu ms/step=%7.3f" % (e.key[:66], e.count / 4, e.self_cpu_time_total / 4 / 1e3))
