"""Online (rmse, nll, gp_loss, noise) trace of the bench stream (powerplant-shaped 32^4, fp32), GPU model or CPU port:
answers whether the streamed model keeps learning although its root rank stays at n_init (SURVEY F9 / VERDICT weak #12).

  python tools/online_trace.py --device cuda --steps 400 --out profiles/r02_online_trace_gpu.json      (GPU box)
  python tools/online_trace.py --device cpu  --steps 200 --out profiles/r02_online_trace_cpu.json      (host cores, oracle port)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--workload", default="powerplant_4d_g32")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    d, g, q, n_init, lr, desc = bench.WORKLOADS[a.workload]
    rows = []
    t0 = time.time()
    if a.device == "cpu":
        args = type("A", (), {"dtype": "f32"})()
        res = bench.cpu_steps(args, a.workload, max_steps=a.steps, budget_s=1e9, warm=0)
        rows = [{"step": t, "rmse": r, "nll": n, "gp_loss": l} for t, (r, n, l) in enumerate(res["trace"])]
        extra = {"cores": res["cores"], "kron_modes": res["kron_modes"], "s_per_step": sum(res["times"]) / len(res["times"])}
    else:
        from online_gp_b200 import settings as S
        dev = torch.device("cuda:0")
        model, xs, ys = bench.build_model(d, g, n_init, lr, torch.float32, dev)
        xd, yd = xs.to(dev), ys.to(dev)
        with S.max_root_decomposition_size(bench.MAX_ROOT), S.max_cholesky_size(bench.MAX_CHOL), S.cg_tolerance(bench.CG_TOL):
            model.enable_cuda_graphs(True, warmup_calls=1)
            for t in range(a.steps):
                rmse, nll, loss = bench.one_step(model, xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q])
                rows.append({"step": t, "rmse": rmse, "nll": nll, "gp_loss": loss, "noise": float(model.noise.mean())})
        extra = {"cuda_graphs": model._graphs is not None and not model._graphs.failed}
    out = {"workload": a.workload, "device": a.device, "steps": len(rows), "wall_s": time.time() - t0, **extra, "trace": rows}
    # running means over windows of 50 steps: the learning signal
    w = 50
    out["window_mean"] = [{"steps": f"{i}-{i + w - 1}", "rmse": sum(r["rmse"] for r in rows[i:i + w]) / len(rows[i:i + w]),
                           "nll": sum(r["nll"] for r in rows[i:i + w]) / len(rows[i:i + w])} for i in range(0, len(rows), w)]
    s = json.dumps(out)
    if a.out:
        open(a.out, "w").write(s)
    print(json.dumps({k: v for k, v in out.items() if k != "trace"}))


if __name__ == "__main__":
    main()
