"""CPU: the oracle restatements against the frozen golden vectors and against each other."""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle.exact_gp import DenseExactSKIGP
from oracle.gridkernel import Hypers, kron_dense, kron_toeplitz_matmul, kuu_columns
from oracle.interp import create_grid, dense_wt, interpolate, left_interp
from oracle.wiski_matfree import WiskiMatFree
from oracle.wiski_ref import WiskiRef

T64 = torch.float64


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_create_grid_spacing():
    # bounds (-4,14), g=20 -> h=1.0 but actual spacing g*h/(g-1) = 1.0526
    g = create_grid([20], [(-4.0, 14.0)])[0]
    assert g.dtype == torch.float32
    assert abs(float(g[1] - g[0]) - 20.0 / 19.0) < 1e-6
    assert abs(float(g[0]) + 5.0) < 1e-6 and abs(float(g[-1]) - 15.0) < 1e-6


@pytest.mark.parametrize("name", ["d1", "d2", "d3", "d4"])
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_interp_golden(golden_dir, name, tag):
    z = _load(golden_dir, "g0_interp.npz")
    grid = create_grid(z[f"{name}_sizes"].tolist(), [tuple(b) for b in z[f"{name}_bounds"].tolist()])
    x = torch.from_numpy(z[f"{name}_{tag}_x"])
    idx, val = interpolate(grid, x)
    assert np.array_equal(idx.numpy(), z[f"{name}_{tag}_idx"])      # bit-exact indices
    assert np.array_equal(val.numpy(), z[f"{name}_{tag}_val"])
    # partition of unity and index range
    assert torch.allclose(val.sum(-1).double(), torch.ones(x.shape[0], dtype=T64), atol=1e-5)
    m = int(np.prod(z[f"{name}_sizes"]))
    assert idx.min() >= 0 and idx.max() < m


def test_interp_out_of_bounds_raises():
    grid = create_grid([10], [(0.0, 1.0)])
    with pytest.raises(RuntimeError, match="out of bounds"):
        interpolate(grid, torch.tensor([[1.5]]))


def test_left_interp_matches_dense_w():
    grid = create_grid([6, 7], [(0.0, 1.0)] * 2)
    x = torch.rand(9, 2, dtype=T64)
    idx, val = interpolate(grid, x)
    rhs = torch.randn(42, 3, dtype=T64)
    assert torch.allclose(left_interp(idx, val, rhs), dense_wt(idx, val, 42).t() @ rhs, atol=1e-12)


def test_kron_matmul_matches_dense():
    grid = create_grid([5, 6, 4], [(0.0, 1.0)] * 3)
    hyp = Hypers(3)
    cols = kuu_columns(grid, hyp)
    X = torch.randn(120, 4, dtype=T64)
    assert torch.allclose(kron_toeplitz_matmul(cols, X), kron_dense(cols) @ X, atol=1e-12)


def test_g1_mll_and_grads(golden_dir):
    """Literal WISKI restatement vs the dense exact GP frozen in g1 (the reference's live test, rtol 1e-5)."""
    z = _load(golden_dir, "g1_mll.npz")
    grid = create_grid([5, 5], [(0.0, 1.0), (0.0, 1.0)])
    for tag in ("t1", "t3"):
        x, y, yvar = (torch.from_numpy(z[f"{tag}_{k}"]) for k in ("x", "y", "yvar"))
        for lt, learn in (("fixed", False), ("learn", True)):
            for o in range(y.shape[1]):
                hyp = Hypers(2, learn_noise=learn)
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    model = WiskiRef(grid, hyp, x, y[:, o], yvar[:, o])
                    v = model.mll()
                v.backward()
                g = np.concatenate([p.grad.numpy().reshape(-1) for p in hyp.params()])
                assert np.allclose(v.item(), z[f"{tag}_{lt}_mll"][o], rtol=1e-5, atol=1e-8)
                assert np.allclose(g, z[f"{tag}_{lt}_grad"][o], rtol=1e-4, atol=1e-7)


def test_g2_update_sequence(golden_dir):
    z = _load(golden_dir, "g2_sequence.npz")
    grid = create_grid([20], [(-4.0, 14.0)])
    tp = torch.from_numpy(z["test_points"])
    for cls, kw in ((WiskiRef, {}), (WiskiMatFree, {}), (WiskiMatFree, {"update_mode": "sym"})):
        hyp = Hypers(1, has_scale=False, learn_noise=True)
        hyp.set(lengthscale=10.0, noise=0.01)
        model = None
        for k in range(5):
            xk, yk = torch.from_numpy(z[f"x{k}"]), torch.from_numpy(z[f"y{k}"])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                if model is None:
                    model = cls(grid, hyp, xk, yk, torch.ones_like(yk), **kw)
                else:
                    model.condition_on_observations(xk, yk, torch.ones_like(yk))
            mean, cov = model.predict(tp)
            # jitter 1e-8 in the Cholesky of W^T W perturbs the ill-conditioned l=10 problem slightly
            assert np.allclose(mean.detach().numpy(), z[f"mean{k}"], rtol=1e-4, atol=1e-6)
            assert np.allclose(cov.detach().numpy(), z[f"cov{k}"], rtol=1e-3, atol=1e-7)
            assert np.allclose(model.mll().item(), z[f"mll{k}"], rtol=1e-4)


def test_g3_fantasy(golden_dir):
    z = _load(golden_dir, "g3_strategy.npz")
    for g in (4, 10):
        grid = create_grid([g], [(-0.4, 1.4)])
        hyp = Hypers(1, has_scale=False, learn_noise=True)
        hyp.set(noise=0.1)
        xs, labels = torch.from_numpy(z["xs"]), torch.from_numpy(z["labels"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = WiskiRef(grid, hyp, xs, labels, torch.ones_like(labels))
            npts = torch.from_numpy(z["new_points"])
            mean, cov = model.predict(npts)
            assert np.allclose(mean.detach().numpy(), z[f"g{g}_mean"], rtol=1e-5, atol=1e-6)
            assert np.allclose(cov.detach().numpy(), z[f"g{g}_cov"], rtol=1e-5, atol=1e-6)
            model.condition_on_observations(torch.from_numpy(z["fant_x"]), torch.from_numpy(z["fant_y"]), torch.ones(1, dtype=T64))
            mean, cov = model.predict(npts)
        assert np.allclose(mean.detach().numpy(), z[f"g{g}_fant_mean"], rtol=1e-5, atol=1e-6)
        assert np.allclose(cov.detach().numpy(), z[f"g{g}_fant_cov"], rtol=1e-5, atol=1e-6)
        assert np.allclose(model.mll().item(), z[f"g{g}_fant_mll"], rtol=1e-5)


def test_matfree_lowrank_is_projected_update():
    """F9: with r < m the root update keeps only the component of v in span(L): L'L'^T = LL^T + (Pv)(Pv)^T."""
    torch.manual_seed(1)
    grid = create_grid([8, 8], [(-1.1, 1.1)] * 2)
    X = torch.rand(12, 2, dtype=T64) * 2 - 1
    y = torch.sin(3 * X.sum(-1))
    for mode in ("svd", "sym"):
        mf = WiskiMatFree(grid, Hypers(2, learn_noise=True), X[:8], y[:8], torch.ones(8, dtype=T64),
                          max_cholesky_size=0, update_mode=mode)
        L0, B0 = mf.L.clone(), mf.B.clone()
        assert torch.allclose(B0.t() @ L0, torch.eye(L0.shape[1], dtype=T64), atol=1e-9)
        mf.condition_on_observations(X[8:], y[8:], torch.ones(4, dtype=T64))
        idx, val = interpolate(grid, X[8:])
        V = dense_wt(idx, val, 64)
        PV = L0 @ (B0.t() @ V)
        assert torch.allclose(mf.L @ mf.L.t(), L0 @ L0.t() + PV @ PV.t(), atol=1e-9)
        assert torch.allclose(mf.B.t() @ mf.L, torch.eye(L0.shape[1], dtype=T64), atol=1e-8)


def test_numpy_index_oracle_agrees_with_torch_oracle_and_golden(golden_dir):
    """Second, torch-free restatement of the interpolation stencils (oracle/interp_np.py): indices identical to the
    torch oracle and to the frozen golden vectors, values equal to rounding (the tap weights multiply in a different
    association order)."""
    import os
    import numpy as np
    import torch
    from oracle.interp import create_grid, interpolate
    from oracle.interp_np import interpolate_np
    z = np.load(os.path.join(golden_dir, "g0_interp.npz"))
    for name in ("d1", "d2", "d3", "d4"):
        grid = create_grid(z[f"{name}_sizes"].tolist(), [tuple(b) for b in z[f"{name}_bounds"].tolist()])
        grids_np = [g.numpy() for g in grid]
        for tag in ("f32", "f64"):
            x = z[f"{name}_{tag}_x"][:6]
            idx_np, val_np = interpolate_np(grids_np, x)
            assert np.array_equal(idx_np, z[f"{name}_{tag}_idx"][:6]), (name, tag)
            idx_t, val_t = interpolate(grid, torch.from_numpy(x))
            assert np.array_equal(idx_np, idx_t.numpy())
            eps = np.finfo(x.dtype).eps
            assert np.allclose(val_np, z[f"{name}_{tag}_val"][:6], rtol=16 * eps, atol=16 * eps)
    # boundary stencils (one-hot on the nearest of the first / last four grid points)
    grid = create_grid([10], [(0.0, 1.0)])
    x = np.array([[-0.12], [1.12], [0.5]], dtype=np.float64)
    idx_np, val_np = interpolate_np([grid[0].numpy()], x)
    idx_t, val_t = interpolate(grid, torch.from_numpy(x))
    assert np.array_equal(idx_np, idx_t.numpy()) and np.allclose(val_np, val_t.numpy(), atol=1e-12)
