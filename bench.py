#!/usr/bin/env python
"""bench.py — WISKI streaming updates/sec on B200 (BASELINE.json metric), with roofline, parity digest and CPU baseline.

A "step" is one pass of the reference's streaming loop body (experiments/regression.py:49-54) on one batch of q
synthetic points:  OnlineSKIRegression.evaluate(x_t, y_t)  (posterior mean + variance at the new points, caches
rebuilt)  +  OnlineSKIRegression.update(x_t, y_t)  (one Adam step on the Woodbury MLL over lengthscales /
outputscale / noise  +  condition_on_observations in place).  Default workload = BASELINE.json configs[1]:
powerplant-shaped 4-D stream, 32^4 inducing grid, batch_size 1, fp32 (reference default dtype), n_init = 430.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Prints ONE JSON line (rank 0).  `value` = updates/s with inputs resident in HBM; `e2e` = the same loop fed from
pinned host memory (H2D copy of every batch and the D2H metric reads inside the timed region); `roofline` = the
dominant kernel against its measured roof; `parity` = the streamed (rmse, nll, loss) of the first steps next to the
CPU port's on the same stream, the last timed step and max |B^T L - I|; `cpu_baseline` = the CPU port timed on the
host cores; `secondary` (N = 1, default workload) = the north-star target grid (1024^2) measured the same way.
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
    # name: (d, g, q, n_init, lr, description)     lr = base_lr / 10 (experiments/regression.py:138)
    "powerplant_4d_g32": (4, 32, 1, 430, 5e-3, "powerplant-shaped 4-D stream, 32^4 inducing grid, batch_size=1"),
    "road3d_3d_g128": (3, 128, 8, 19569, 1e-3, "3droad-shaped 3-D stream, 128^3 grid, batch_size=8"),
    "malaria_2d_g256": (2, 256, 6, 10, 5e-3, "malaria-shaped 2-D stream, 256^2 grid, batch_size=6"),
    "synthetic_1d_g128": (1, 128, 1, 25, 5e-3, "1-D synthetic regression, 128-point grid, batch_size=1"),
    "target_2d_g1024": (2, 1024, 1, 430, 5e-3, "2-D stream, 1024^2 grid, batch_size=1 (north-star target grid)"),
}
MAX_ROOT, MAX_CHOL, CG_TOL = 512, 2048, 1e-2      # config/regression.yaml:24-27
STREAM_EXTRA = 8192                               # stream points generated beyond n_init


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


# ------------------------------------------------------------------------------------------------ shared pieces
PROF_NAMES = {"wiski_kron_toeplitz_mm", "wiski_kron_toeplitz_bwd_cols", "wiski_gram", "wiski_panel_rmul",
              "wiski_panel_lowrank_update", "wiski_panel_lowrank_update2", "wiski_gather", "wiski_scatter_add",
              "wiski_interp_fwd", "wiski_kron_fused_pair_apply", "wiski_kron_fused_pair_grad", "wiski_kron_axis_apply",
              "wiski_kron_axis_contract"}


def algorithmic_work(m, r, b, d, g):
    """Algorithmic bytes / flops per launch of the panel-sized ops, m = rows on this GPU."""
    return {
        "wiski_kron_toeplitz_mm": ("hbm", 2.0 * m * r * b),                  # m x r panel, ideal single pass
        "wiski_kron_toeplitz_bwd_cols": ("hbm", 2.0 * m * r * b),            # read Z and X once
        "wiski_kron_fused_pair_apply": ("hbm", 2.0 * m * r * b),             # read + write the panel (two axes per pass)
        "wiski_kron_fused_pair_grad": ("hbm", 2.5 * m * r * b),              # read Z and P (+ write Z' on one of the two passes)
        "wiski_panel_lowrank_update": ("hbm", 2.0 * m * r * b),              # read + write one panel
        "wiski_panel_lowrank_update2": ("hbm", 4.0 * m * r * b),             # read + write both panels (L and B)
        "wiski_gram": ("tensor", 2.0 * m * r * r),                           # flops
        "wiski_panel_rmul": ("tensor", 2.0 * m * r * r),
        "wiski_kron_axis_apply": ("tensor", 2.0 * g * m * r),                # one axis as a batched GEMM (g >= 64)
        "wiski_kron_axis_contract": ("tensor", 2.0 * g * m * r),
    }


# dram__bytes_read.sum + dram__bytes_write.sum per launch from ONE `ncu --set full` capture of the step's kernels on
# 1 x B200 (profiles/r02_ncu_step_kernels.md; BASELINE config 2: m = 2^20, r = 432, fp32).  Ops with two launches per
# step (the two pair passes) carry the mean of the two.
NCU_TRAFFIC_C2 = {"wiski_panel_rmul": 3.583e9, "wiski_gram": 4.059e9, "wiski_kron_fused_pair_apply": 3.704e9,
                  "wiski_kron_fused_pair_grad": 4.522e9, "wiski_panel_lowrank_update2": 7.840e9}


def ncu_traffic(dom, m, r, b, d, g):
    if (m, r, b, d, g) == (1 << 20, 432, 4, 4, 32) and dom in NCU_TRAFFIC_C2:
        return NCU_TRAFFIC_C2[dom], "profiles/r02_ncu_step_kernels.md (ncu --set full, same workload, 1 x B200)"
    return None, "no ncu capture for this shape (see profiles/)"


def roofline_of(prof_events, KP, m, r, b, d, g):
    """per-op device time (CUDA events around every library call of an eager pass) -> per-op table + roofline of the
    dominant op (frac against the measured peak of MEASURED_PEAKS.json)."""
    per_op = {}
    for name, evs in prof_events.items():
        tot = sum(a.elapsed_time(bb) for a, bb, _ in evs)
        per_op[name] = {"calls_per_step": len(evs) / KP, "ms_per_step": tot / KP, "ms_per_call": tot / len(evs)}
    if not per_op:
        return per_op, None
    hbm_peak, tf_peak, peak_src = peaks()
    alg = algorithmic_work(m, r, b, d, g)
    dom = max(per_op, key=lambda n: per_op[n]["ms_per_step"])
    roof = None
    if dom in alg:
        bound, work = alg[dom]
        # panel-sized calls only (m x 1 calls of the same op are excluded)
        times = sorted(a.elapsed_time(bb) for a, bb, _ in prof_events[dom])
        big_t = [x for x in times if x >= 0.5 * times[-1]]
        avg_ms = sum(big_t) / len(big_t)
        if bound == "hbm":
            ach = work / (avg_ms * 1e-3) / 1e9
            traffic, tsrc = ncu_traffic(dom, m, r, b, d, g)
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "ms_per_launch": avg_ms,
                    "algorithmic_bytes_per_launch": work}
        else:
            ach = work / (avg_ms * 1e-3) / 1e12
            traffic, tsrc = ncu_traffic(dom, m, r, b, d, g)
            # The fp32 GEMMs run as 3xTF32 on the tensor pipe: B200 has no fp32 tensor path, kind::tf32 runs at half the
            # bf16 rate, and an fp32-accurate product costs three tf32 MMAs.  The roof of an fp32-accurate GEMM is therefore
            # the measured dense bf16 rate / 6; `frac` is quoted against it, `frac_of_bf16_peak` against the raw figure.
            fp32_peak = tf_peak / 6.0
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": ach / fp32_peak, "frac_of_bf16_peak": ach / tf_peak, "bf16_peak": tf_peak,
                    "traffic": traffic, "traffic_source": tsrc,
                    "peak_source": peak_src + " (bf16 dense cuBLAS) / 6: kind::tf32 = half the bf16 rate, 3 tf32 MMAs per "
                                              "fp32-accurate product (3xTF32)",
                    "ms_per_launch": avg_ms, "algorithmic_flops_per_launch": work,
                    "hbm_view": {"algorithmic_bytes_per_launch": 2.0 * m * r * b, "achieved_GBps": 2.0 * m * r * b / (avg_ms * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": 2.0 * m * r * b / (avg_ms * 1e-3) / 1e9 / hbm_peak},
                    "note": "algorithmic flops = 2 m r r2 (fp32 GEMM); the kernel issues 3x that many tf32 flops"}
    return per_op, roof


def parity_block(gpu_trace, cpu_trace, btl_err, last):
    """gpu_trace / cpu_trace: [(rmse, nll, loss)] per stream step from step 0 (cpu_trace may be shorter or None)."""
    out = {"last_timed_step": {"rmse": last[0], "nll": last[1], "gp_loss": last[2]},
           "max_abs_BtL_minus_I": btl_err,
           "stream_head": [{"step": t, "rmse": r, "nll": n, "gp_loss": l} for t, (r, n, l) in enumerate(gpu_trace[:8])]}
    if cpu_trace:
        n = min(len(cpu_trace), len(gpu_trace))
        rel = lambda a, b: abs(a - b) / max(1.0, abs(b))
        out["vs_cpu_port"] = {
            "steps_compared": n,
            "max_rel_diff": {"rmse": max(rel(gpu_trace[t][0], cpu_trace[t][0]) for t in range(n)),
                             "nll": max(rel(gpu_trace[t][1], cpu_trace[t][1]) for t in range(n)),
                             "gp_loss": max(rel(gpu_trace[t][2], cpu_trace[t][2]) for t in range(n))},
            "cpu_stream_head": [{"step": t, "rmse": r, "nll": n_, "gp_loss": l} for t, (r, n_, l) in enumerate(cpu_trace[:8])],
            "tolerance": "north_star: 1e-2 relative in fp32 (|a - b| / max(1, |b|))"}
    return out


# ------------------------------------------------------------------------------------------------ GPU arm, N = 1
def build_model(d, g, n_init, lr, dtype, device):
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import Identity
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        x, y = synth_stream(d, n_init + STREAM_EXTRA)
        x, y = x.to(dtype), y.to(dtype)
        with S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL):
            model = OnlineSKIRegression(Identity(d), x[:n_init].to(device), y[:n_init].to(device), lr=lr,
                                        grid_size=g, grid_bound=1.0)
            model.set_lr(lr)
    finally:
        torch.set_default_dtype(prev)
    return model, x[n_init:], y[n_init:]


def one_step(model, xb, yb):
    from online_gp_b200 import settings as S
    with S.detach_interp_coeff(True):
        rmse, nll = model.evaluate(xb, yb)
    stem_loss, gp_loss = model.update(xb, yb, update_stem=True)
    return rmse, nll, gp_loss


def gpu_single(args, workload, K, W, device, cpu_budget, with_cg=True):
    """One model on one GPU: returns the bench record (dict) of `workload`."""
    from online_gp_b200 import _lib, ops
    from online_gp_b200 import settings as S

    d, g, q, n_init, lr, desc = WORKLOADS[workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    lib = _lib.load()
    model, xs, ys = build_model(d, g, n_init, lr, dtype, device)
    m = g ** d
    model.gp._kernel_cache["WtW"].root_decomposition()      # Cholesky regime: the root is built lazily
    r = model.gp._kernel_cache["WtW"].root.shape[-1]
    ctx = (S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL))
    for c in ctx:
        c.__enter__()
    try:
        need = (2 * K + W + 16) * q
        assert xs.shape[0] >= need, "stream too short"
        xd, yd = xs.to(device), ys.to(device)
        xh, yh = xs.pin_memory(), ys.pin_memory()
        trace = []

        def step_dev(t):
            trace.append(one_step(model, xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q]))

        def step_host(t):
            trace.append(one_step(model, xh[t * q:(t + 1) * q].to(device, non_blocking=True),
                                  yh[t * q:(t + 1) * q].to(device, non_blocking=True)))

        t = 0
        for _ in range(W):
            step_dev(t)
            t += 1
        use_graphs = not args.no_graphs
        # per-kernel CUDA-event timings behind `roofline` / `per_op_ms_per_step`: an eager pass of the same steps on
        # the same state right before the timed region (graph replays run no host code)
        KP = min(K, 5)
        ops.PROFILE = {"names": PROF_NAMES, "events": {}}
        phases = _PhaseTimer(model)
        torch.cuda.synchronize()
        # (serial schedule for this pass: a kernel timed while a background launch shares the SMs says nothing about the kernel)
        with S.overlap_root_update(False):
            for _ in range(KP):
                step_dev(t)
                t += 1
        torch.cuda.synchronize()
        phase_ms = phases.stop(KP)
        prof, ops.PROFILE = ops.PROFILE, None
        if use_graphs:
            model.enable_cuda_graphs(True, warmup_calls=1)
            for _ in range(3):                       # 1 eager, 1 capture + first replay, 1 replay: all untimed
                step_dev(t)
                t += 1
            use_graphs = model._graphs is not None and not model._graphs.failed and model._graphs.upd is not None
        # ---- device-resident timing (value)
        clocks = ClockSampler(device.index or 0)
        launches0 = lib.wiski_launch_count() + model.graph_launches
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_dev(t)
            t += 1
        e1.record()
        torch.cuda.synchronize()
        ms_dev = e0.elapsed_time(e1)
        launches = lib.wiski_launch_count() + model.graph_launches - launches0
        last = trace[-1]
        # ---- end-to-end timing from pinned host memory
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            step_host(t)
            t += 1
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1)
        clk = clocks.stop()

        b = 4 if dtype == torch.float32 else 8
        per_op, roof = roofline_of(prof["events"], KP, m, r, b, d, g)
        with torch.no_grad():
            wtw = model.gp._kernel_cache["WtW"]
            Lp, Bp = wtw._panels(wtw.root)[0], wtw._panels(wtw.inv_root)[0]
            G = ops.gram(Bp, Lp)
            # B^T L = I on the kept directions (zero-padded columns stay zero in both panels)
            live = (torch.linalg.vector_norm(Lp, dim=0) > 0).to(G.dtype)
            btl = float((G - torch.diag(live)).abs().max())
        out = {
            "metric": "wiski_streaming_updates_per_sec", "value": K / (ms_dev * 1e-3), "unit": "updates/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "description": desc, "d": d, "grid": g, "m": m, "q": q, "n_init": n_init,
                       "root_rank": r, "stencil": 4 ** d, "lr": lr,
                       "l2": "panels (m*r*%d B = %.2f GB each) are far larger than L2" % (b, m * r * b / 1e9),
                       "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)",
                       "root_update_mode": S.root_update_mode.value(),
                       "kron_directional_grad": bool(S.kron_directional_grad.on()),
                       "kron_tensor_core_pairs": bool(lib.wiski_kron_tc_enable(1)) or True,
                       "cuda_graphs": bool(use_graphs),
                       "per_op_timing": "eager pass of %d steps before the timed region, serial schedule (settings.overlap_root_update off: "
                                        "single-kernel times; the timed region overlaps the background passes)" % KP,
                       "overlap_root_update": True},
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
        if with_cg:
            try:
                out["cg_mvm"] = cg_mvm_roofline(model, m, r, b)
            except Exception as err:                      # noqa: BLE001 - a secondary metric must not lose the bench line
                out["cg_mvm"] = {"error": f"{type(err).__name__}: {err}"}
        cpu_trace = None
        if not args.no_cpu_baseline:
            del model
            torch.cuda.empty_cache()
            base = cpu_baseline(args, workload, budget_s=cpu_budget)
            cpu_trace = base.pop("trace")
            out["cpu_baseline"] = base
        out["parity"] = parity_block(trace, cpu_trace, btl, last)
    finally:
        for c in ctx:
            c.__exit__(None, None, None)
    return out


def run_gpu(args):
    import torch.distributed as dist

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
        return run_gpu_sharded(args, rank, world, device)

    out = gpu_single(args, args.workload, args.steps, args.warmup, device, args.cpu_budget)
    if args.workload == "powerplant_4d_g32" and not args.no_secondary:
        # the north-star target grid (1024^2, >= 50x the CPU path on 1 x B200) measured the same way in the same run;
        # a failure here must not lose the primary line
        try:
            sec = gpu_single(args, "target_2d_g1024", min(args.steps, 8), 3, device, min(args.cpu_budget, 30.0),
                             with_cg=False)
            cb = sec.get("cpu_baseline")
            if cb:
                sec["speedup_vs_cpu_port_e2e"] = sec["e2e"]["value"] / cb["value"]
            out["secondary"] = [sec]
        except Exception as err:                          # noqa: BLE001
            out["secondary"] = [{"config": {"workload": "target_2d_g1024"}, "error": f"{type(err).__name__}: {err}"}]
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


def cg_mvm_roofline(model, m, r, b, reps=20):
    """BASELINE.json's second metric: achieved GB/s of the CG matrix-vector product  w = (I + L^T K L) v  — the closure
    GPyTorch's linear_cg calls — as one fused pass over the two m x r panels (wiski_q_matvec), and
    of the panel Kronecker-Toeplitz MVM  K L  that feeds it.  CUDA events around `reps` back-to-back launches on the
    model's own panels (3.6 GB per launch: far larger than L2)."""
    from online_gp_b200 import ops
    hbm_peak, _, peak_src = peaks()
    with torch.no_grad():
        L = model.gp._root_panels()[0]
        K = model.gp.Kuu.items[0].detach()
        KL = K._matmul(L)
        v = torch.randn(r, 1, dtype=L.dtype, device=L.device)
        res = {}
        npass = max(1, len(K.sizes) // 2) if all(s == 32 for s in K.sizes) and len(K.sizes) % 2 == 0 else len(K.sizes)
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
        res["kron_toeplitz_mm_panel"]["passes"] = npass
        res["kron_toeplitz_mm_panel"]["frac_per_pass"] = res["kron_toeplitz_mm_panel"]["frac"] * npass
        res["kron_toeplitz_mm_panel"]["note"] = ("quoted against ONE ideal read + write of the panel; the MVM is %d pair/axis "
                                                 "passes, each moving that much" % npass)
        model.gp._dump_caches()
    return res


# ------------------------------------------------------------------------------------------------ GPU arm, N > 1
def run_gpu_sharded(args, rank, world, device):
    """N > 1: ONE model, inducing-grid rows sharded across ranks (strong scaling), online_gp_b200/parallel.py."""
    import torch.distributed as dist
    from online_gp_b200 import _lib, ops
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import Identity
    from online_gp_b200.parallel import Comm

    d, g, q, n_init, lr, desc = WORKLOADS[args.workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    lib = _lib.load()
    K, W = args.steps, args.warmup
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    x, y = synth_stream(d, n_init + STREAM_EXTRA)
    x, y = x.to(dtype), y.to(dtype)
    ctx = (S.max_root_decomposition_size(MAX_ROOT), S.max_cholesky_size(MAX_CHOL), S.cg_tolerance(CG_TOL),
           S.sharded_dual_layout(not args.single_layout))
    for c in ctx:
        c.__enter__()
    # the same class as N = 1, with the communicator: it drives the row-sharded engine (online_gp_b200/parallel.py)
    wrapper = OnlineSKIRegression(Identity(d), x[:n_init].to(device), y[:n_init].to(device), lr=lr, grid_size=g,
                                  grid_bound=1.0, comm=Comm())
    wrapper.set_lr(lr)
    model = wrapper.engine
    torch.set_default_dtype(prev)
    xs, ys = x[n_init:], y[n_init:]
    xd, yd = xs.to(device), ys.to(device)
    xh, yh = xs.pin_memory(), ys.pin_memory()
    m, r = g ** d, model.L_loc.shape[1]
    b = 4 if dtype == torch.float32 else 8
    trace = []

    def step(xb, yb):
        rmse, nll = wrapper.evaluate(xb, yb)
        _, loss = wrapper.update(xb, yb)
        trace.append((rmse, nll, loss))

    t = 0
    for _ in range(W):
        step(xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q])
        t += 1
    # per-op timings of this rank's kernels (eager pass, like N = 1) -> roofline of the dominant kernel on rank 0
    KP = min(K, 5)
    ops.PROFILE = {"names": PROF_NAMES, "events": {}}
    torch.cuda.synchronize()
    with S.overlap_root_update(False):       # serial schedule while single kernels are being timed
        for _ in range(KP):
            step(xd[t * q:(t + 1) * q], yd[t * q:(t + 1) * q])
            t += 1
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    use_graphs = not args.no_graphs
    if use_graphs:
        wrapper.enable_cuda_graphs(True, warmup_calls=1)
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
    last = trace[-1]
    ms_e2e = timed(True)
    clk = clocks.stop()
    with torch.no_grad():
        G = ops.gram(model.B_loc, model.L_loc)
        nz = (torch.linalg.vector_norm(model.L_loc, dim=0) > 0).to(G.dtype)
        dist.all_reduce(G)
        dist.all_reduce(nz, op=dist.ReduceOp.MAX)
        btl = float((G - torch.diag(nz)).abs().max())
    per_op, roof = roofline_of(prof["events"], KP, m // world, r, b, d, g)
    for c in ctx:
        c.__exit__(None, None, None)
    if rank == 0:
        out = {
            "metric": "wiski_streaming_updates_per_sec", "value": K / (ms_dev * 1e-3), "unit": "updates/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "d": d, "grid": g, "m": m, "q": q,
                       "n_init": n_init, "root_rank": r, "stencil": 4 ** d, "lr": lr,
                       "parallelism": f"inducing-grid rows sharded over {world} GPUs (grid axis 0), r x r algebra replicated",
                       "model_class": type(wrapper).__name__ + "(comm=Comm())",
                       "rows_per_gpu": m // world,
                       "l2": "per-GPU panel slab (%.2f GB) larger than L2" % (m // world * r * b / 1e9),
                       "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)",
                       "cuda_graphs": bool(use_graphs), "dual_layout": not args.single_layout,
                       "kron_directional_grad": bool(S.kron_directional_grad.on()),
                       "exchange": "peer memory" if model.comm.xbuf is not None else "nccl all_to_all",
                       "per_op_timing": "rank 0, eager pass of %d steps before the timed region, serial schedule" % KP,
                       "overlap_root_update": True},
            "clocks": clk,
            "e2e": {"value": K / (ms_e2e * 1e-3), "unit": "updates/s", "h2d_bytes_per_step": q * (d + 1) * b,
                    "d2h_bytes_per_step": 3 * b},
            "gpu_launches": int(launches),
            "roofline": roof,
            "per_op_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_op.items())},
            "parity": parity_block(trace, None, btl, last),
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
def cpu_steps(args, workload, max_steps, budget_s, warm=1):
    """The reference's algorithm for the same step on host cores: oracle/wiski_matfree.py (literal SVD root update,
    torch CPU autograd for the hyper gradient), all host threads.  Returns per-step times and the (rmse, nll, loss)
    trace from stream step 0 (the same stream, initial set, learning rate and protocol as the GPU arm)."""
    from oracle import wiski_matfree as wm
    from oracle.gridkernel import Hypers
    from oracle.interp import create_grid

    d, g, q, n_init, lr, desc = WORKLOADS[workload]
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    x, y = synth_stream(d, n_init + STREAM_EXTRA)
    grid = create_grid([g] * d, [(-1.1, 1.1)] * d)
    hyp = Hypers(d, learn_noise=True, dtype=dtype)
    kron_modes = wm.pick_kron_mode([g] * d, MAX_ROOT, dtype)
    t0 = time.time()
    model = wm.WiskiMatFree(grid, hyp, x[:n_init].to(dtype), y[:n_init, 0].to(dtype), torch.ones(n_init, dtype=dtype),
                            max_cholesky_size=MAX_CHOL, max_root=MAX_ROOT, dtype=dtype, update_mode="svd", fold="batched")
    init_s = time.time() - t0
    opt = torch.optim.Adam(hyp.params(), lr=lr)
    xs, ys = x[n_init:].to(dtype), y[n_init:, 0].to(dtype)
    times, trace = [], []
    t = 0
    start = time.time()
    while len(times) < max_steps + warm:
        s0 = time.time()
        xb, yb = xs[t * q:(t + 1) * q], ys[t * q:(t + 1) * q]
        pieces = model.pieces()
        with torch.no_grad():
            mean, cov = model.predict(xb, pieces=pieces)                       # evaluate()
            var = cov.diagonal() + hyp.noise
            rmse = float((mean - yb).pow(2).mean().sqrt())
            nll = float(-torch.distributions.Normal(mean, var.sqrt()).log_prob(yb).mean())
        opt.zero_grad()
        loss = -model.mll(pieces=pieces, skip_logdet_forward=True)             # _update_gp() (:137 skip_logdet_forward)
        loss.backward()
        opt.step()
        with torch.no_grad():
            model.condition_on_observations(xb, yb, torch.ones(q, dtype=dtype))   # condition, literal SVD update
        times.append(time.time() - s0)
        trace.append((rmse, nll, float(loss.detach())))
        t += 1
        if len(times) > warm and time.time() - start > budget_s:
            break
    timed = times[warm:] if len(times) > warm else times
    return {"times": timed, "cores": cores, "init_s": init_s, "r": model.L.shape[1], "trace": trace,
            "kron_modes": kron_modes}


def cpu_baseline(args, workload, budget_s):
    res = cpu_steps(args, workload, max_steps=10, budget_s=budget_s, warm=1)
    timed = res["times"]
    per = sum(timed) / len(timed)
    d, g = WORKLOADS[workload][0], WORKLOADS[workload][1]
    return {"value": 1.0 / per, "unit": "updates/s", "cores": res["cores"], "kind": "port", "trace": res["trace"],
            "sample": f"{len(timed)} timed step(s) after 1 warm-up of the same workload and stream (m={g ** d}, r={res['r']}), "
                      f"oracle/wiski_matfree.py on torch CPU, {per:.2f} s/step (min {min(timed):.2f}, max {max(timed):.2f}); "
                      f"per-axis Toeplitz product = faster of dense GEMM / GPyTorch's FFT form on this host: {res['kron_modes']}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, g, q, n_init, lr, desc = WORKLOADS[args.workload]
    res = cpu_steps(args, args.workload, max_steps=args.steps, budget_s=args.cpu_budget * 4, warm=min(args.warmup, 1))
    timed = res["times"]
    per = sum(timed) / len(timed)
    val = 1.0 / per
    out = {
        "impl": "reference", "metric": "wiski_streaming_updates_per_sec", "value": val, "unit": "updates/s",
        "n_gpus": args.gpus, "steps": len(timed), "warmup": min(args.warmup, 1), "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "d": d, "grid": g, "m": g ** d, "q": q,
                   "n_init": n_init, "root_rank": res["r"], "lr": lr,
                   "step": "evaluate + update (Adam step on Woodbury MLL + condition_on_observations)"},
        "cpu_baseline": {"value": val, "unit": "updates/s", "cores": res["cores"], "kind": "port",
                         "sample": f"{len(timed)} timed step(s) (time-boxed) of the same workload on host cores; the "
                                   f"reference package itself needs GPyTorch/BoTorch, which cannot be installed "
                                   f"offline, so this is the oracle port oracle/wiski_matfree.py "
                                   f"(Toeplitz products: {res['kron_modes']})"},
        "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "parity": {"stream_head": [{"step": t, "rmse": a, "nll": b_, "gp_loss": c} for t, (a, b_, c) in enumerate(res["trace"][:8])]},
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
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip the north-star target grid record")
    ap.add_argument("--dual-layout", action="store_true", help="(default since round 2; kept for old command lines)")
    ap.add_argument("--single-layout", action="store_true",
                    help="N > 1: turn settings.sharded_dual_layout off (four row <-> column exchanges per step instead of two)")
    ap.add_argument("--no-graphs", action="store_true", help="run the timed steps eagerly instead of replaying CUDA graphs")
    ap.add_argument("--cpu-budget", type=float, default=45.0, help="seconds of CPU work for the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
