#!/usr/bin/env python
"""bench.py — WISKI streaming updates/sec on B200 (BASELINE.json metric), with roofline and CPU baseline.

A "step" is one pass of the reference's streaming loop body (experiments/regression.py:49-54) on one batch of q
synthetic points:  OnlineSKIRegression.evaluate(x_t, y_t)  (posterior mean + variance at the new points, caches
rebuilt)  +  OnlineSKIRegression.update(x_t, y_t)  (one Adam step on the Woodbury MLL over lengthscales /
outputscale / noise  +  condition_on_observations in place).  Default workload = BASELINE.json configs[1]:
powerplant-shaped 4-D stream, 32^4 inducing grid, batch_size 1, fp32 (reference default dtype), n_init = 430.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = updates/s with inputs resident in HBM; `e2e` = the same loop fed from
pinned host memory (H2D copy of every batch and the D2H metric reads inside the timed region).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
    # name: (d, g, q, n_init, description)
    "powerplant_4d_g32": (4, 32, 1, 430, "powerplant-shaped 4-D stream, 32^4 inducing grid, batch_size=1"),
    "road3d_3d_g128": (3, 128, 8, 19569, "3droad-shaped 3-D stream, 128^3 grid, batch_size=8"),
    "malaria_2d_g256": (2, 256, 6, 10, "malaria-shaped 2-D stream, 256^2 grid, batch_size=6"),
    "synthetic_1d_g128": (1, 128, 1, 25, "1-D synthetic regression, 128-point grid, batch_size=1"),
    "target_2d_g1024": (2, 1024, 1, 430, "2-D stream, 1024^2 grid, batch_size=1"),
}
MAX_ROOT, MAX_CHOL, CG_TOL = 512, 2048, 1e-2      # config/regression.yaml:24-27


def synth_stream(d, n, seed=0):
    """x ~ U(-1,1)^d, y = sin(3 sum x) + 0.1 eps, z-scored; torch.Generator().manual_seed(0)."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=gen) * 2 - 1
    y = torch.sin(3 * x.sum(-1)) + 0.1 * torch.randn(n, generator=gen)
    y = (y - y.mean()) / y.std()
    return x, y.unsqueeze(-1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        return z["hbm_gbs"], z.get("bf16_tflops_sustained", z["bf16_tflops"]), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], None, set()
        for line in self.f.read().splitlines():
            parts = [t.strip() for t in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def build_model(d, g, n_init, dtype, device):
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import Identity
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        x, y = synth_stream(d, n_init + 4096)
        x, y = x.to(dtype), y.to(dtype)
        with S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL):
            model = OnlineSKIRegression(Identity(d), x[:n_init].to(device), y[:n_init].to(device), lr=5e-3,
                                        grid_size=g, grid_bound=1.0)
            model.set_lr(5e-3)       # base_lr / 10 for powerplant (experiments/regression.py:138)
    finally:
        torch.set_default_dtype(prev)
    return model, x[n_init:], y[n_init:]


def one_step(model, xb, yb):
    from online_gp_b200 import settings as S
    with S.detach_interp_coeff(True):
        rmse, nll = model.evaluate(xb, yb)
    stem_loss, gp_loss = model.update(xb, yb, update_stem=True)
    return rmse, nll, gp_loss


def run_gpu(args):
    import torch.distributed as dist
    from online_gp_b200 import _lib, ops
    from online_gp_b200 import settings as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        # the row <-> column exchanges are large point-to-point transfers: let NCCL spread them over enough channels
        # to fill NVLink 5 (its default of a few p2p channels per peer reaches ~200 GB/s of the 900 GB/s per direction)
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
        dist.init_process_group("nccl", device_id=device)
    d, g, q, n_init, desc = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    lib = _lib.load()

    if world > 1:
        return run_gpu_sharded(args, rank, world, device, dtype)

    model, xs, ys = build_model(d, g, n_init, dtype, device)
    m = g ** d
    model.gp._kernel_cache["WtW"].root_decomposition()      # Cholesky regime: the root is built lazily
    r = model.gp._kernel_cache["WtW"].root.shape[-1]
    ctx = (S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL))
    for c in ctx:
        c.__enter__()
    K, W = args.steps, args.warmup
    need = (2 * K + W + 8) * q
    assert xs.shape[0] >= need, "stream too short"
    xd, yd = xs.to(device), ys.to(device)
    xh, yh = xs.pin_memory(), ys.pin_memory()

    def batch_dev(t):
        return xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q]

    def batch_host(t):
        return (xh[t * q:(t + 1) * q].to(device, non_blocking=True), yh[t * q:(t + 1) * q].to(device, non_blocking=True))

    t = 0
    for _ in range(W):
        one_step(model, *batch_dev(t))
        t += 1
    prof_names = {"wiski_kron_toeplitz_mm", "wiski_kron_toeplitz_bwd_cols", "wiski_gram", "wiski_panel_rmul",
                  "wiski_panel_lowrank_update", "wiski_gather", "wiski_scatter_add", "wiski_interp_fwd",
                  "wiski_kron_fused_pair_apply", "wiski_kron_fused_pair_grad", "wiski_kron_axis_apply",
                  "wiski_kron_axis_contract"}
    use_graphs = not args.no_graphs
    KP = K
    if use_graphs:
        # graph replays run no Python, so the per-kernel CUDA-event timings behind `roofline` / `per_op_ms_per_step`
        # come from an eager pass of the same steps on the same state, right before the timed region
        KP = min(K, 5)
        ops.PROFILE = {"names": prof_names, "events": {}}
        phases = _PhaseTimer(model)
        torch.cuda.synchronize()
        for _ in range(KP):
            one_step(model, *batch_dev(t))
            t += 1
        torch.cuda.synchronize()
        phase_ms = phases.stop(KP)
        prof, ops.PROFILE = ops.PROFILE, None
        model.enable_cuda_graphs(True, warmup_calls=1)
        for _ in range(3):                       # 1 eager, 1 capture + first replay, 1 replay: all untimed
            one_step(model, *batch_dev(t))
            t += 1
        use_graphs = model._graphs is not None and not model._graphs.failed and model._graphs.upd is not None
    else:
        ops.PROFILE = {"names": prof_names, "events": {}}
        phase_ms = None
    # ---- device-resident timing (value)
    clocks = ClockSampler(local_rank)
    launches0 = lib.wiski_launch_count() + model.graph_launches
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        one_step(model, *batch_dev(t))
        t += 1
    e1.record()
    torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1)
    launches = lib.wiski_launch_count() + model.graph_launches - launches0
    if ops.PROFILE is not None:
        prof, ops.PROFILE = ops.PROFILE, None
    # ---- end-to-end timing from pinned host memory
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        one_step(model, *batch_host(t))
        t += 1
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    clk = clocks.stop()

    # ---- per-op device time -> roofline of the dominant op
    b = 4 if dtype == torch.float32 else 8
    per_op = {}
    for name, evs in prof["events"].items():
        tot = sum(a.elapsed_time(bb) for a, bb, _ in evs)
        per_op[name] = {"calls_per_step": len(evs) / KP, "ms_per_step": tot / KP, "ms_per_call": tot / len(evs)}
    step_ms = ms_dev / K
    dom = max(per_op, key=lambda n: per_op[n]["ms_per_step"])
    hbm_peak, tf_peak, peak_src = peaks()
    # algorithmic bytes / flops per launch (DESIGN.md "Kernels and rooflines")
    big = [(a, bb) for a, bb, ar in prof["events"].get("wiski_kron_toeplitz_mm", [])]
    alg = {
        "wiski_kron_toeplitz_mm": ("hbm", 2.0 * m * r * b),                  # m x r panel, ideal single pass
        "wiski_kron_toeplitz_bwd_cols": ("hbm", 2.0 * m * r * b),            # read Z and X once
        "wiski_kron_fused_pair_apply": ("hbm", 2.0 * m * r * b),             # read + write the panel (two axes per pass)
        "wiski_kron_fused_pair_grad": ("hbm", 2.5 * m * r * b),              # read Z and P (+ write Z' on one of the two passes)
        "wiski_panel_lowrank_update": ("hbm", 2.0 * m * r * b),              # read + write the panel
        "wiski_gram": ("tensor", 2.0 * m * r * r),                           # flops
        "wiski_panel_rmul": ("tensor", 2.0 * m * r * r),
    }
    # DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed `ncu --set full`
    # capture profiles/r01_ncu_kernels_v2.md, bytes; the two launches per step of the pair kernels are averaged
    ncu_traffic = {"wiski_kron_fused_pair_grad": 0.5 * (5.41e9 + 3.63e9), "wiski_kron_fused_pair_apply": 0.5 * (4.08e9 + 3.58e9),
                   "wiski_gram": 4.41e9, "wiski_panel_rmul": 3.58e9, "wiski_panel_lowrank_update": 3.56e9}
    roof = None
    if dom in alg:
        bound, work = alg[dom]
        # use the slowest-size calls only (panel-sized calls dominate; m x 1 calls of the same op are excluded)
        evs = prof["events"][dom]
        times = sorted(a.elapsed_time(bb) for a, bb, _ in evs)
        big_t = [x for x in times if x >= 0.5 * times[-1]]
        avg_ms = sum(big_t) / len(big_t)
        if bound == "hbm":
            ach = work / (avg_ms * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": ncu_traffic.get(dom), "peak_source": peak_src,
                    "ms_per_launch": avg_ms, "algorithmic_bytes_per_launch": work,
                    "note": "m x r fp32 panel pass; at g = 32 the two-axes Kronecker passes are FP32-FMA / "
                            "constant-operand issue limited, not DRAM limited (ncu: FMA pipe ~50 %, DRAM ~20 %)"}
        else:
            ach = work / (avg_ms * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": ach / tf_peak, "traffic": ncu_traffic.get(dom), "peak_source": peak_src + " (bf16 dense cuBLAS)",
                    "ms_per_launch": avg_ms, "algorithmic_flops_per_launch": work,
                    "note": "fp32 result via 3xTF32: the kernel issues 3x these flops on the tensor pipe"}
    out = {
        "metric": "wiski_streaming_updates_per_sec", "value": K / (ms_dev * 1e-3), "unit": "updates/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "d": d, "grid": g, "m": m, "q": q, "n_init": n_init,
                   "root_rank": r, "stencil": 4 ** d, "l2": "panels (m*r*%d B = %.2f GB each) are far larger than L2" % (b, m * r * b / 1e9),
                   "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)",
                   "root_update_mode": S.root_update_mode.value(),
                   "cuda_graphs": bool(use_graphs),
                   "per_op_timing": ("eager pass of %d steps before the timed region (graph replays run no host code)" % KP)
                   if not args.no_graphs else "CUDA events inside the timed region"},
        "clocks": clk,
        "e2e": {"value": K / (ms_e2e * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": q * (d + 1) * b,
                "d2h_bytes_per_step": 3 * b + 4},
        "gpu_launches": int(launches),
        "roofline": roof,
        "per_op_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_op.items())},
    }
    if phase_ms:
        # the reference scripts' sub-timings (wiski_regression.py:125-148): mll_time = MLL forward + backward + Adam,
        # fantasy_time = condition_on_observations; device time (CUDA events) of the eager pass, ms per step
        out["phase_ms_per_step_eager"] = phase_ms
    try:
        out["cg_mvm"] = cg_mvm_roofline(model, m, r, b, hbm_peak, peak_src)
    except Exception as err:                      # noqa: BLE001 - a secondary metric must not lose the bench line
        out["cg_mvm"] = {"error": f"{type(err).__name__}: {err}"}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, budget_s=args.cpu_budget)
    for c in ctx:
        c.__exit__(None, None, None)
    print(json.dumps(out))


class _PhaseTimer:
    """CUDA-event timing of the three phases of a streaming step on the eager path (never fatal for the bench)."""

    def __init__(self, model):
        self.model, self.events, self.saved = model, {"evaluate": [], "mll_time": [], "fantasy_time": []}, []
        try:
            self._wrap(model, "evaluate", "evaluate")
            self._wrap(model, "_update_gp_tensor", "mll_time")
            self._wrap(model.gp, "condition_on_observations", "fantasy_time")
        except Exception:                         # noqa: BLE001
            self.stop(1)

    def _wrap(self, obj, name, key):
        fn = getattr(obj, name)
        events = self.events[key]

        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn(*a, **k)
            e1.record()
            events.append((e0, e1))
            return res

        self.saved.append((obj, name))
        setattr(obj, name, timed)

    def stop(self, steps):
        for obj, name in self.saved:
            try:
                delattr(obj, name)                # drop the instance attribute: the class method is visible again
            except Exception:                     # noqa: BLE001
                pass
        self.saved = []
        try:
            return {k: round(sum(a.elapsed_time(b) for a, b in v) / steps, 4) for k, v in self.events.items() if v}
        except Exception:                         # noqa: BLE001
            return None


def cg_mvm_roofline(model, m, r, b, hbm_peak, peak_src, reps=20):
    """BASELINE.json's second metric: achieved GB/s of the CG matrix-vector product  w = (I + L^T K L) v  — the closure
    GPyTorch's linear_cg calls — as one fused pass over the two m x r panels (wiski_q_matvec), and
    of the panel Kronecker-Toeplitz MVM  K L  that feeds it.  CUDA events around `reps` back-to-back launches on the
    model's own panels (3.6 GB per launch: far larger than L2)."""
    from online_gp_b200 import ops
    with torch.no_grad():
        L = model.gp._root_panels()[0]
        K = model.gp.Kuu.items[0].detach()
        KL = K._matmul(L)
        v = torch.randn(r, 1, dtype=L.dtype, device=L.device)
        res = {}
        for name, fn, nbytes in (("q_matvec", lambda: ops.q_matvec(L, KL, v), 2.0 * m * r * b),
                                 ("kron_toeplitz_mm_panel", lambda: K._matmul(L), 2.0 * m * r * b)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            gbs = nbytes / (ms * 1e-3) / 1e9
            res[name] = {"ms_per_call": ms, "algorithmic_bytes_per_call": nbytes, "achieved": gbs, "unit": "GB/s",
                         "peak": hbm_peak, "frac": gbs / hbm_peak, "peak_source": peak_src}
        model.gp._dump_caches()
    return res


def run_gpu_sharded(args, rank, world, device, dtype):
    """N > 1: ONE model, inducing-grid rows sharded across ranks (strong scaling), online_gp_b200/parallel.py."""
    import torch.distributed as dist
    from online_gp_b200 import _lib
    from online_gp_b200 import settings as S
    from online_gp_b200.parallel import Comm, ShardedOnlineSKIRegression

    d, g, q, n_init, desc = WORKLOADS[args.workload]
    lib = _lib.load()
    K, W = args.steps, args.warmup
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    x, y = synth_stream(d, n_init + 4096)
    x, y = x.to(dtype), y.to(dtype)
    ctx = (S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL),
           S.sharded_dual_layout(bool(args.dual_layout)))
    for c in ctx:
        c.__enter__()
    model = ShardedOnlineSKIRegression(x[:n_init].to(device), y[:n_init].to(device), lr=5e-3, grid_size=g,
                                       grid_bound=1.0, comm=Comm())
    torch.set_default_dtype(prev)
    xs, ys = x[n_init:], y[n_init:]
    xd, yd = xs.to(device), ys.to(device)
    xh, yh = xs.pin_memory(), ys.pin_memory()
    m, r = g ** d, model.L_loc.shape[1]
    b = 4 if dtype == torch.float32 else 8

    def step(xb, yb):
        rmse, nll = model.evaluate(xb, yb)
        _, loss = model.update(xb, yb)
        return rmse, nll, loss

    t = 0
    for _ in range(W):
        step(xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q])
        t += 1
    use_graphs = not args.no_graphs
    if use_graphs:
        model.enable_cuda_graphs(True, warmup_calls=1)
        for _ in range(3):                       # 1 eager, 1 capture + first replay, 1 replay: all untimed
            step(xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q])
            t += 1
        ok = torch.tensor([int(model._graphs is not None and not model._graphs.failed and model._graphs.upd is not None)],
                          device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        use_graphs = bool(ok.item())
        if not use_graphs:
            model.enable_cuda_graphs(False)
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))

    def timed(from_host):
        nonlocal t
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            if from_host:
                xb = xh[t * q:(t + 1) * q].to(device, non_blocking=True)
                yb = yh[t * q:(t + 1) * q].to(device, non_blocking=True)
            else:
                xb, yb = xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q]
            step(xb, yb)
            t += 1
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    l0 = lib.wiski_launch_count() + model.graph_launches
    ms_dev = timed(False)
    launches = lib.wiski_launch_count() + model.graph_launches - l0
    ms_e2e = timed(True)
    clk = clocks.stop()
    for c in ctx:
        c.__exit__(None, None, None)
    if rank == 0:
        out = {
            "metric": "wiski_streaming_updates_per_sec", "value": K / (ms_dev * 1e-3), "unit": "updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "d": d, "grid": g, "m": m, "q": q,
                       "n_init": n_init, "root_rank": r, "stencil": 4 ** d,
                       "parallelism": f"inducing-grid rows sharded over {world} GPUs (grid axis 0), r x r algebra replicated",
                       "rows_per_gpu": m // world,
                       "l2": "per-GPU panel slab (%.2f GB) larger than L2" % (m // world * r * b / 1e9),
                       "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)",
                       "cuda_graphs": bool(use_graphs), "dual_layout": bool(args.dual_layout),
                       "exchange": "peer memory" if model.comm.xbuf is not None else "nccl all_to_all"},
            "clocks": clk,
            "e2e": {"value": K / (ms_e2e * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": q * (d + 1) * b,
                    "d2h_bytes_per_step": 3 * b},
            "gpu_launches": int(launches),
            "roofline": None,
        }
        print(json.dumps(out), flush=True)
    # captured graphs hold NCCL work: release them (and everything queued) before the communicator is torn down
    model.enable_cuda_graphs(False)
    del model
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)          # skip interpreter teardown: destroying the process group under live CUDA-graph state can hang


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_steps(args, max_steps, budget_s, warm=1):
    """The reference's algorithm for the same step on host cores: oracle/wiski_matfree.py (literal SVD root update,
    torch CPU autograd for the hyper gradient), all host threads."""
    from oracle.gridkernel import Hypers
    from oracle.interp import create_grid
    from oracle.wiski_matfree import WiskiMatFree

    d, g, q, n_init, desc = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    x, y = synth_stream(d, n_init + 4096)
    grid = create_grid([g] * d, [(-1.1, 1.1)] * d)
    hyp = Hypers(d, learn_noise=True, dtype=dtype)
    t0 = time.time()
    model = WiskiMatFree(grid, hyp, x[:n_init].to(dtype), y[:n_init, 0].to(dtype), torch.ones(n_init, dtype=dtype),
                         max_cholesky_size=MAX_CHOL, max_root=MAX_ROOT, dtype=dtype, update_mode="svd")
    init_s = time.time() - t0
    opt = torch.optim.Adam(hyp.params(), lr=5e-3)
    xs, ys = x[n_init:].to(dtype), y[n_init:, 0].to(dtype)
    times = []
    t = 0
    start = time.time()
    while len(times) < max_steps + warm:
        s0 = time.time()
        xb, yb = xs[t * q:(t + 1) * q], ys[t * q:(t + 1) * q]
        pieces = model.pieces()
        with torch.no_grad():
            mean, cov = model.predict(xb, pieces=pieces)                       # evaluate()
            var = cov.diagonal() + hyp.noise
            _ = float((mean - yb).pow(2).mean().sqrt())
            _ = float(-torch.distributions.Normal(mean, var.sqrt()).log_prob(yb).mean())
        opt.zero_grad()
        loss = -model.mll(pieces=pieces)                                       # _update_gp()
        loss.backward()
        opt.step()
        with torch.no_grad():
            model.condition_on_observations(xb, yb, torch.ones(q, dtype=dtype))   # condition, literal SVD update
        times.append(time.time() - s0)
        t += 1
        if len(times) > warm and time.time() - start > budget_s:
            break
    timed = times[warm:] if len(times) > warm else times
    return timed, cores, init_s, model.L.shape[1]


def cpu_baseline(args, budget_s):
    timed, cores, init_s, r = cpu_steps(args, max_steps=3, budget_s=budget_s, warm=1)
    per = sum(timed) / len(timed)
    return {"value": 1.0 / per, "unit": "updates/s", "cores": cores, "kind": "port",
            "sample": f"{len(timed)} timed step(s) after 1 warm-up of the same workload (m={WORKLOADS[args.workload][1] ** WORKLOADS[args.workload][0]}, r={r}), "
                      f"oracle/wiski_matfree.py on torch CPU, {per:.2f} s/step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, g, q, n_init, desc = WORKLOADS[args.workload]
    timed, cores, init_s, r = cpu_steps(args, max_steps=args.steps, budget_s=args.cpu_budget * 4, warm=min(args.warmup, 1))
    per = sum(timed) / len(timed)
    val = 1.0 / per
    out = {
        "impl": "reference", "metric": "wiski_streaming_updates_per_sec", "value": val, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "d": d, "grid": g, "m": g ** d, "q": q,
                   "n_init": n_init, "root_rank": r,
                   "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)"},
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": cores, "kind": "port",
                         "sample": f"{len(timed)} timed step(s) (time-boxed) of the same workload on host cores; the "
                                   f"reference package itself needs GPyTorch/BoTorch, which cannot be installed "
                                   f"offline, so this is the oracle port oracle/wiski_matfree.py"},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="powerplant_4d_g32", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dual-layout", action="store_true",
                    help="N > 1: settings.sharded_dual_layout (two row <-> column exchanges per step instead of four)")
    ap.add_argument("--no-graphs", action="store_true", help="run the timed steps eagerly instead of replaying CUDA graphs")
    ap.add_argument("--cpu-budget", type=float, default=30.0, help="seconds of CPU work for the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
