"""Row-sharded path on real GPUs vs the single-GPU path (run under torchrun, one rank per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/sharded_parity.py [--dims 4 --grid 32 --n0 48 --steps 4 --q 1]

Every rank streams the same synthetic points through ``ShardedOnlineSKIRegression``; rank 0 also runs the same
stream through the single-GPU ``OnlineSKIRegression`` and compares RMSE / NLL / loss / hyper-parameters step by
step (fp32 tolerance 1e-2 relative as in BASELINE.json's north_star; indices are shared by construction).
Exit status 0 = parity, 1 = mismatch.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", dest="d", type=int, default=4)
    ap.add_argument("--grid", dest="g", type=int, default=32)
    ap.add_argument("--n0", type=int, default=48)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--q", type=int, default=1)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--watchdog", type=float, default=0.0, help="dump all Python stacks and exit after this many seconds")
    ap.add_argument("--dual", action="store_true", help="settings.sharded_dual_layout")
    ap.add_argument("--graphs", action="store_true", help="replay the sharded step as CUDA graphs (eager single-GPU reference)")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
        dist.init_process_group("nccl", device_id=dev)
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import Identity
    from online_gp_b200.parallel import Comm, ShardedOnlineSKIRegression

    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    torch.set_default_dtype(dtype)
    gen = torch.Generator().manual_seed(11)
    n = args.n0 + args.steps * args.q
    X = (torch.rand(n, args.d, generator=gen) * 2 - 1).to(dtype).to(dev)
    y = torch.sin(3 * X.sum(-1, keepdim=True)) + 0.1 * torch.randn(n, 1, generator=gen).to(dtype).to(dev)
    rows = []
    with S.max_root_decomposition_size(512), S.max_cholesky_size(2048), S.sharded_dual_layout(args.dual):
        sh = ShardedOnlineSKIRegression(X[:args.n0], y[:args.n0], lr=1e-2, grid_size=args.g, grid_bound=1.0, comm=Comm())
        if args.graphs:
            sh.enable_cuda_graphs(True, warmup_calls=1)
        for t in range(args.steps):
            s = slice(args.n0 + t * args.q, args.n0 + (t + 1) * args.q)
            rmse, nll = sh.evaluate(X[s], y[s])
            _, loss = sh.update(X[s], y[s])
            rows.append((rmse, nll, loss, float(sh._noise())))
        ls_sh = sh.covar_module.base_kernel.base_kernel.lengthscale.detach().reshape(-1).tolist()
        ok, worst, ref_rows = True, 0.0, []
        if rank == 0:
            one = OnlineSKIRegression(Identity(args.d), X[:args.n0], y[:args.n0], lr=1e-2, grid_size=args.g, grid_bound=1.0)
            one.set_lr(1e-2)
            for t in range(args.steps):
                s = slice(args.n0 + t * args.q, args.n0 + (t + 1) * args.q)
                with S.detach_interp_coeff(True):
                    rmse, nll = one.evaluate(X[s], y[s])
                _, loss = one.update(X[s], y[s], update_stem=True)
                ref_rows.append((rmse, nll, loss, float(one.gp.likelihood.second_noise.reshape(-1)[0])))
            ls_one = one.gp.covar_module.base_kernel.base_kernel.lengthscale.detach().reshape(-1).tolist()
            tol = 1e-2 if dtype == torch.float32 else 1e-4
            for a, b in zip(rows, ref_rows):
                for u, v in zip(a, b):
                    err = abs(u - v) / max(1.0, abs(v))
                    worst = max(worst, err)
                    ok = ok and err <= tol
            for u, v in zip(ls_sh, ls_one):
                worst = max(worst, abs(u - v))
                ok = ok and abs(u - v) <= tol
            print(json.dumps({"world": world, "d": args.d, "g": args.g, "n0": args.n0, "q": args.q, "steps": args.steps,
                              "dtype": args.dtype, "cuda_graphs": bool(args.graphs and sh._graphs is not None and not sh._graphs.failed
                                                                          and sh._graphs.upd is not None),
                              "graph_replays": 0 if sh._graphs is None else sh._graphs.replays, "rank_root": int(sh.L_loc.shape[1]), "rows_per_rank": int(sh.L_loc.shape[0]),
                              "worst_rel_err": worst, "parity": bool(ok), "sharded": rows, "single_gpu": ref_rows}))
    if world > 1:
        flag = torch.tensor([0 if ok else 1], device=dev)
        dist.broadcast(flag, 0)
        ok = int(flag.item()) == 0
        sh.enable_cuda_graphs(False)
        torch.cuda.synchronize()
        dist.barrier()
    sys.stdout.flush()
    os._exit(0 if ok else 1)       # no process-group teardown under live CUDA-graph state (it can hang)


if __name__ == "__main__":
    main()
