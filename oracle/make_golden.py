"""Generates tests/golden/*.npz — frozen known-answer vectors for the WISKI hot path.

Inputs are the fixtures of the reference's own tests; outputs come from the dense exact-GP
oracle (``oracle.exact_gp``), i.e. from the identity the reference's live test asserts, NOT from the WISKI
restatements they are used to check.  Interpolation vectors (G0) come from ``oracle.interp`` (the GPyTorch
restatement; unpinned at the GPyTorch-version level, see oracle/__init__.py).

Run from the repo root:  python -m oracle.make_golden
"""
import os
import numpy as np
import torch

from .interp import create_grid, interpolate
from .gridkernel import Hypers
from .exact_gp import DenseExactSKIGP

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def g0_interp():
    """Interpolation index/value vectors, fp32 and fp64 targets, incl. both boundary branches."""
    out = {}
    gen = torch.Generator().manual_seed(0)
    for name, sizes, bounds in [("d1", [20], [(-4.0, 14.0)]), ("d2", [5, 7], [(0.0, 1.0), (-1.0, 2.0)]),
                                ("d4", [32] * 4, [(-1.1, 1.1)] * 4), ("d3", [10, 12, 9], [(-1.1, 1.1)] * 3)]:
        grid = create_grid(sizes, bounds)
        d = len(sizes)
        lo = torch.tensor([b[0] for b in bounds], dtype=torch.float64)
        hi = torch.tensor([b[1] for b in bounds], dtype=torch.float64)
        x = lo + (hi - lo) * torch.rand(64, d, generator=gen, dtype=torch.float64)
        # boundary rows: outside the data bounds but inside the extended grid (one-hot branches of A.1)
        glo = torch.stack([g.min() for g in grid]).double()
        ghi = torch.stack([g.max() for g in grid]).double()
        x[0] = glo + 1e-3 * (ghi - glo)
        x[1] = ghi - 1e-3 * (ghi - glo)
        x[2] = glo
        x[3] = ghi
        x[4] = lo
        x[5] = hi
        for dt, tag in [(torch.float64, "f64"), (torch.float32, "f32")]:
            xt = x.to(dt)
            idx, val = interpolate(grid, xt)
            out[f"{name}_{tag}_x"] = _np(xt)
            out[f"{name}_{tag}_idx"] = _np(idx)
            out[f"{name}_{tag}_val"] = _np(val)
        out[f"{name}_sizes"] = np.array(sizes)
        out[f"{name}_bounds"] = np.array(bounds)
    np.savez(os.path.join(OUT, "g0_interp.npz"), **out)


def g1_mll():
    """tests/mlls/test_batched_woodbury_marginal_log_likelihood.py:19-43 (seed 10, n=10, d=2, g=5, noise .1)."""
    torch.set_default_dtype(torch.float64)
    out = {}
    for batched in (False, True):
        torch.random.manual_seed(10)
        train_x = torch.rand(10, 2)
        train_y = torch.sin(2 * train_x[:, 0] + 3 * train_x[:, 1]).unsqueeze(-1)
        train_y_var = 0.1 * torch.ones_like(train_y)
        if batched:
            train_y = torch.cat((train_y, train_y + 0.3 * torch.randn_like(train_y),
                                 train_y + 0.3 * torch.randn_like(train_y)), dim=1)
            train_y_var = train_y_var.repeat(1, 3)
        grid = create_grid([5, 5], [(0.0, 1.0), (0.0, 1.0)])
        tag = "t3" if batched else "t1"
        out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_yvar"] = _np(train_x), _np(train_y), _np(train_y_var)
        for learn in (False, True):
            mlls, grads = [], []
            for o in range(train_y.shape[1]):
                hyp = Hypers(2, kind="rbf", has_scale=True, learn_noise=learn)
                gp = DenseExactSKIGP(grid, hyp, train_x, train_y[:, o], train_y_var[:, o])
                v = gp.mll()
                v.backward()
                mlls.append(v.item())
                grads.append(np.concatenate([_np(p.grad).reshape(-1) for p in hyp.params()]))
            lt = "learn" if learn else "fixed"
            out[f"{tag}_{lt}_mll"] = np.array(mlls)
            out[f"{tag}_{lt}_grad"] = np.stack(grads)     # [t, (raw_ls0, raw_ls1, raw_outputscale[, raw_noise])]
        xs = torch.rand(6, 2)
        hyp = Hypers(2, learn_noise=False)
        mean, cov = DenseExactSKIGP(grid, hyp, train_x, train_y[:, 0], train_y_var[:, 0]).posterior(xs)
        out[f"{tag}_xs"], out[f"{tag}_mean"], out[f"{tag}_cov"] = _np(xs), _np(mean), _np(cov)
    np.savez(os.path.join(OUT, "g1_mll.npz"), **out)


def g2_sequence():
    """tests/models/test_woodbury_gp_model.py:63-101 (stale file; fixture inputs + intended checks only):
    1-D, g=20 on (-4,14), RBF (no ScaleKernel) lengthscale 10, homoskedastic noise 0.01, update sequence
    5+2+1+1+1, test points [5, 8]: posterior after every update == exact GP refit on all data."""
    torch.set_default_dtype(torch.float64)
    xs = torch.tensor([2.0, 3.0, 4.0, 1.0, 7.0])
    labels = torch.sin(xs) + torch.tensor([0.1, 0.2, -0.1, -0.2, -0.2])
    new_points = torch.tensor([2.4, 4.7])
    new_targets = torch.sin(new_points) + torch.tensor([0.1, -0.15])
    pts = [xs, new_points, torch.tensor([2.3]), torch.tensor([4.1]), torch.tensor([4.3])]
    tgs = [labels, new_targets, torch.sin(torch.tensor([2.3])), torch.sin(torch.tensor([4.1])) + 1,
           torch.sin(torch.tensor([4.3]))]
    test_points = torch.tensor([5.0, 8.0]).unsqueeze(-1)
    grid = create_grid([20], [(-4.0, 14.0)])
    hyp = Hypers(1, kind="rbf", has_scale=False, learn_noise=True)
    hyp.set(lengthscale=10.0, noise=0.01)
    out = {"test_points": _np(test_points), "lengthscale": 10.0, "noise": 0.01}
    X, Y = torch.zeros(0, 1), torch.zeros(0)
    for k, (p, t) in enumerate(zip(pts, tgs)):
        X, Y = torch.cat([X, p.unsqueeze(-1)]), torch.cat([Y, t])
        gp = DenseExactSKIGP(grid, hyp, X, Y, torch.ones_like(Y))
        mean, cov = gp.posterior(test_points)
        out[f"x{k}"], out[f"y{k}"] = _np(p.unsqueeze(-1)), _np(t)
        out[f"mean{k}"], out[f"cov{k}"], out[f"mll{k}"] = _np(mean), _np(cov), gp.mll().item()
    np.savez(os.path.join(OUT, "g2_sequence.npz"), **out)


def g3_strategy():
    """tests/models/test_woodbury_prediction_strategy.py:19-47,111-116 (stale; fixture inputs):
    xs=[.2,.3,.4,.1,.7], g in {4,10} on (-.4,1.4), RBF default lengthscale, noise .1, new points [.5,.8],
    fantasy x=.45, y = sin(x)+.12: posterior before / after the fantasy update."""
    torch.set_default_dtype(torch.float64)
    xs = torch.tensor([0.20, 0.30, 0.40, 0.10, 0.70])
    labels = torch.sin(xs) + torch.tensor([0.1, 0.2, -0.1, -0.2, -0.2])
    new_points = torch.tensor([0.5, 0.8]).unsqueeze(-1)
    fant_x = torch.tensor([0.45])
    fant_y = torch.sin(fant_x) + 0.12
    out = {"xs": _np(xs.unsqueeze(-1)), "labels": _np(labels), "new_points": _np(new_points),
           "fant_x": _np(fant_x.unsqueeze(-1)), "fant_y": _np(fant_y), "noise": 0.1}
    for g in (4, 10):
        grid = create_grid([g], [(-0.4, 1.4)])
        hyp = Hypers(1, kind="rbf", has_scale=False, learn_noise=True)
        hyp.set(noise=0.1)
        gp = DenseExactSKIGP(grid, hyp, xs.unsqueeze(-1), labels, torch.ones_like(labels))
        mean, cov = gp.posterior(new_points)
        gp2 = DenseExactSKIGP(grid, hyp, torch.cat([xs, fant_x]).unsqueeze(-1), torch.cat([labels, fant_y]),
                              torch.ones(6))
        mean2, cov2 = gp2.posterior(new_points)
        out[f"g{g}_mean"], out[f"g{g}_cov"], out[f"g{g}_mll"] = _np(mean), _np(cov), gp.mll().item()
        out[f"g{g}_fant_mean"], out[f"g{g}_fant_cov"], out[f"g{g}_fant_mll"] = _np(mean2), _np(cov2), gp2.mll().item()
    np.savez(os.path.join(OUT, "g3_strategy.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g0_interp()
    g1_mll()
    g2_sequence()
    g3_strategy()
    print("wrote", sorted(os.listdir(OUT)))
