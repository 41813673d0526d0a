"""Diagnostic (GPU): stream of tests/model_cases.py::test_long_fp32_stream... under several kernel settings; prints the first
step whose (rmse, nll) leaves the 1e-2 band of the fp64 oracle."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle.gridkernel import Hypers
from oracle.interp import create_grid
from oracle.wiski_matfree import WiskiMatFree
import online_gp_b200.settings as S
from online_gp_b200 import ops, _lib
from online_gp_b200.models import OnlineSKIRegression
from online_gp_b200.models.stems import Identity

def run(d, g, n0, steps, tag, **kw):
    dev = "cuda:0"
    gen = torch.Generator().manual_seed(21)
    X = torch.rand(n0 + steps, d, generator=gen, dtype=torch.float64) * 2 - 1
    side = int(round(n0 ** 0.5)) + 1
    lat = torch.stack(torch.meshgrid(*[torch.linspace(-0.9, 0.9, side, dtype=torch.float64)] * d, indexing="ij"), -1).reshape(-1, d)
    X[:n0] = lat[:n0] + 0.01 * (torch.rand(n0, d, generator=gen, dtype=torch.float64) - 0.5)
    if kw.get("sites"):
        X[n0:] = X[torch.randint(0, n0, (steps,), generator=gen)]
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps, generator=gen, dtype=torch.float64)).unsqueeze(-1)
    torch.set_default_dtype(torch.float32)
    ctxs = [S.backward_gemm_tf32_passes(kw.get("passes", 3)), S.kron_directional_grad(kw.get("directional", True))]
    for c in ctxs: c.__enter__()
    orig_gram = ops._gram
    if kw.get("nosym"):
        ops._gram = lambda A, B, symmetric=False: orig_gram(A, B, False)
    try:
        with warnings.catch_warnings(), S.max_cholesky_size(0), S.max_root_decomposition_size(128):
            warnings.simplefilter("ignore")
            reg = OnlineSKIRegression(Identity(d), X[:n0].float().to(dev), y[:n0].float().to(dev), lr=5e-3, grid_size=g, grid_bound=1.0)
            reg.set_lr(5e-3)
            hyp = Hypers(d, learn_noise=True)
            orc = WiskiMatFree(create_grid([g] * d, [(-1.1, 1.1)] * d), hyp, X[:n0], y[:n0, 0], torch.ones(n0, dtype=torch.float64),
                               max_cholesky_size=0, max_root=128, update_mode="svd", root_tol=1e-5)
        opt = torch.optim.Adam(hyp.params(), lr=5e-3)
        first, worst = None, 0.0
        with warnings.catch_warnings(), S.max_cholesky_size(2048), S.max_root_decomposition_size(128):
            warnings.simplefilter("ignore")
            for t in range(steps):
                xt, yt = X[n0 + t:n0 + t + 1], y[n0 + t:n0 + t + 1]
                with S.detach_interp_coeff(True):
                    rmse, nll = reg.evaluate(xt.float().to(dev), yt.float().to(dev))
                _, loss = reg.update(xt.float().to(dev), yt.float().to(dev))
                pieces = orc.pieces()
                with torch.no_grad():
                    mo, co = orc.predict(xt, pieces=pieces)
                    var_o = co.diagonal() + hyp.noise
                    rmse_o = float((mo - yt[:, 0]).pow(2).mean().sqrt())
                    nll_o = float(-torch.distributions.Normal(mo, var_o.sqrt()).log_prob(yt[:, 0]).mean())
                opt.zero_grad()
                lo = -orc.mll(pieces=pieces, skip_logdet_forward=True)
                lo.backward()
                opt.step()
                with torch.no_grad():
                    orc.condition_on_observations(xt, yt[:, 0], torch.ones(1, dtype=torch.float64))
                e = max(abs(rmse - rmse_o) / max(1, abs(rmse_o)), abs(nll - nll_o) / max(1, abs(nll_o)))
                worst = max(worst, e)
                if t < 3 or (first is None and e > 1e-2) or t % 50 == 49:
                    print(f"   [{tag}] step {t}: rmse {rmse:.6f}/{rmse_o:.6f} nll {nll:.6f}/{nll_o:.6f} loss {loss:.6f}/{float(lo):.6f} "
                          f"noise {float(reg.noise.mean()):.6f}/{float(hyp.noise):.6f}")
                if first is None and e > 1e-2:
                    first = t
        print(f"[{tag}] d={d} g={g} n0={n0} steps={steps}: worst {worst:.3e}, first step beyond 1e-2: {first}")
    finally:
        ops._gram = orig_gram
        for c in ctxs: c.__exit__(None, None, None)
        torch.set_default_dtype(torch.float32)

def time_rmul():
    P = torch.randn(1 << 20, 432, device="cuda:0")
    M = torch.randn(432, 432, device="cuda:0")
    ref = None
    for terms in (3, 2, 1):
        for _ in range(3):
            Z = ops._rmul(P, M, terms=terms)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            Z = ops._rmul(P, M, terms=terms)
        e1.record()
        torch.cuda.synchronize()
        if ref is None:
            ref = Z.double()
        err = (Z.double() - ref)
        print(f"[rmul m=2^20 r=432 terms={terms}] {e0.elapsed_time(e1) / 10:.3f} ms  max|err| vs 3-pass {float(err.abs().max()):.3e}  "
              f"mean err {float(err.mean()):.3e}  (max|Z| {float(ref.abs().max()):.1f})")


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    time_rmul()
    run(2, 48, 128, steps, "sites passes3 (default)", sites=True)
    run(2, 48, 128, steps, "sites passes2", sites=True, passes=2)
    run(2, 48, 128, steps, "fresh passes3 (default)")
    run(2, 48, 128, steps, "fresh passes2", passes=2)
    if steps <= 100:
        run(2, 48, 128, steps, "nosym", nosym=True)
        run(2, 40, 128, steps, "g40 (SIMT gemm)")
        run(2, 64, 128, steps, "g64 (TC axes)")
