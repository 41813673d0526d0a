"""Kernel timeline of ONE graph-replayed streaming step (1 GPU, bench workload): start / duration / stream of every kernel,
from the torch profiler (CUPTI).  Shows which kernels overlap (settings.overlap_root_update) and where the device idles.

  python tools/timeline_step.py [--workload powerplant_4d_g32] > gpurun_out/timeline.txt
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="powerplant_4d_g32")
    a = ap.parse_args()
    d, g, q, n_init, lr, desc = bench.WORKLOADS[a.workload]
    from online_gp_b200 import settings as S
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:           # torchrun: the row-sharded engine behind the same class (rank 0 prints its own timeline)
        import torch.distributed as dist
        from online_gp_b200.models import OnlineSKIRegression
        from online_gp_b200.models.stems import Identity
        from online_gp_b200.parallel import Comm
        dist.init_process_group("nccl", device_id=dev)
        x, y = bench.synth_stream(d, n_init + bench.STREAM_EXTRA)
        with S.max_root_decomposition_size(bench.MAX_ROOT), S.max_cholesky_size(bench.MAX_CHOL):
            model = OnlineSKIRegression(Identity(d), x[:n_init].to(dev), y[:n_init].to(dev), lr=lr, grid_size=g, grid_bound=1.0,
                                        comm=Comm())
        model.set_lr(lr)
        xs, ys = x[n_init:], y[n_init:]
    else:
        model, xs, ys = bench.build_model(d, g, n_init, lr, torch.float32, dev)
    xd, yd = xs.to(dev), ys.to(dev)
    with S.max_root_decomposition_size(bench.MAX_ROOT), S.max_cholesky_size(bench.MAX_CHOL), S.cg_tolerance(bench.CG_TOL):
        model.enable_cuda_graphs(True, warmup_calls=1)
        t = 0
        for _ in range(6):
            bench.one_step(model, xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q]); t += 1
        torch.cuda.synchronize()
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(2):
                bench.one_step(model, xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q]); t += 1
            torch.cuda.synchronize()
    if rank != 0:
        torch.distributed.barrier()
        os._exit(0)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print(f"# {len(evs)} device events over 2 steps; columns: start_us  dur_us  end_us  stream  name")
    last_end = 0.0
    for e in evs:
        st, en = e.time_range.start - t0, e.time_range.end - t0
        flag = " <-- overlaps previous" if st < last_end - 1.0 else ""
        if en - st >= 20.0 or flag:
            print(f"{st:10.1f} {en - st:9.1f} {en:10.1f}  {getattr(e, 'stream', '?')!s:>4}  {e.name[:90]}{flag}")
        last_end = max(last_end, en)


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        sys.stdout.flush()
        torch.distributed.barrier()
        os._exit(0)
