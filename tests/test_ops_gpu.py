"""GPU parity of every C-ABI kernel against the CPU oracle (run with -m gpu on the B200 box).

Tolerances: interpolation indices bit-exact; fp64 kernels rtol 1e-10 (pure reorderings of fp64 sums);
fp32 kernels rtol 1e-4 against the fp64 oracle evaluated on the same fp32 inputs.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.gridkernel import Hypers, kron_toeplitz_matmul as o_kron, kuu_columns
from oracle.interp import create_grid, interpolate as o_interp, left_interp as o_left_interp

DEV = "cuda:0"
DT = [torch.float64, torch.float32]


def _tol(dt):
    return dict(rtol=1e-10, atol=1e-11) if dt == torch.float64 else dict(rtol=2e-4, atol=2e-5)


def _ops():
    from online_gp_b200 import ops
    return ops


def _pad_cols(cols, dt):
    gmax = max(c.numel() for c in cols)
    out = torch.zeros(len(cols), gmax, dtype=dt)
    for i, c in enumerate(cols):
        out[i, : c.numel()] = c.to(dt)
    return out


@pytest.mark.parametrize("name", ["d1", "d2", "d3", "d4"])
@pytest.mark.parametrize("tag,dt", [("f32", torch.float32), ("f64", torch.float64)])
def test_interp_golden(golden_dir, name, tag, dt):
    ops = _ops()
    z = np.load(os.path.join(golden_dir, "g0_interp.npz"))
    grid = create_grid(z[f"{name}_sizes"].tolist(), [tuple(b) for b in z[f"{name}_bounds"].tolist()])
    spec = ops.GridSpec(grid)
    x = torch.from_numpy(z[f"{name}_{tag}_x"]).to(DEV)
    idx, val = ops.interpolate(x, spec)
    assert idx.dtype == torch.int64 and val.dtype == dt
    assert np.array_equal(idx.cpu().numpy(), z[f"{name}_{tag}_idx"]), "interpolation indices must be bit-exact"
    ref = z[f"{name}_{tag}_val"]
    # values: same op order with explicit rn intrinsics -> expected bit-exact; allow 2 ulp
    eps = np.finfo(ref.dtype).eps
    assert np.allclose(val.cpu().numpy(), ref, rtol=4 * eps, atol=4 * eps)


def test_interp_out_of_bounds():
    ops = _ops()
    spec = ops.GridSpec(create_grid([10, 10], [(0.0, 1.0)] * 2))
    with pytest.raises(RuntimeError, match="out of bounds"):
        ops.interpolate(torch.tensor([[0.5, 1.7]], device=DEV), spec)
    idx, val = ops.interpolate(torch.zeros(0, 2, device=DEV), spec)     # empty input
    assert idx.shape == (0, 16) and val.shape == (0, 16)


@pytest.mark.parametrize("dt", DT)
def test_interp_backward(dt):
    ops = _ops()
    grid = create_grid([9, 11], [(-1.0, 1.0)] * 2)
    spec = ops.GridSpec(grid)
    x = (torch.rand(7, 2, dtype=torch.float64) * 1.8 - 0.9).to(dt)
    gv = torch.randn(7, 16, dtype=dt)
    xo = x.clone().requires_grad_(True)
    _, vo = o_interp(grid, xo)
    (vo * gv).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    _, vg = ops.interpolate(xg, spec)
    (vg * gv.to(DEV)).sum().backward()
    tol = dict(rtol=1e-9, atol=1e-9) if dt == torch.float64 else dict(rtol=1e-3, atol=1e-3)
    assert torch.allclose(xg.grad.cpu(), xo.grad, **tol)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("c", [1, 3, 40, 200])
def test_gather_scatter(dt, c):
    ops = _ops()
    grid = create_grid([7, 8, 6], [(-1.0, 1.0)] * 3)
    spec = ops.GridSpec(grid)
    x = (torch.rand(11, 3, dtype=torch.float64) * 2 - 1).to(dt)
    idx, val = o_interp(grid, x)
    src = torch.randn(spec.m, c, dtype=dt)
    ref = o_left_interp(idx, val.double(), src.double())
    out = ops.left_interp(idx.to(DEV), val.to(DEV), src.to(DEV))
    assert torch.allclose(out.cpu().double(), ref, **_tol(dt))
    rhs = torch.randn(11, c, dtype=dt)
    ref_t = torch.zeros(spec.m, c, dtype=torch.float64)
    ref_t.index_add_(0, idx.reshape(-1), (val.double().unsqueeze(-1) * rhs.double().unsqueeze(1)).reshape(-1, c))
    out_t = ops.left_t_interp(idx.to(DEV), val.to(DEV), rhs.to(DEV), spec.m)
    assert torch.allclose(out_t.cpu().double(), ref_t, **_tol(dt))


KRON_CASES = [([32, 32], 16), ([32, 32, 32, 32], 32), ([32, 32], 48), ([128], 3), ([12, 9], 5), ([32, 32], 33), ([10, 10, 10], 7), ([8, 8, 8, 8], 16), ([5, 40, 6], 4),
              ([300], 2), ([64, 16], 1), ([16, 16, 16, 16], 1)]


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("sizes,c", KRON_CASES)
def test_kron_toeplitz_mm(dt, sizes, c):
    ops = _ops()
    d = len(sizes)
    grid = create_grid(sizes, [(-1.1, 1.1)] * d)
    hyp = Hypers(d, kind="rbf")
    with torch.no_grad():
        hyp.raw_lengthscale.copy_(torch.linspace(-1.0, 0.5, d))
    cols = [cc.detach() for cc in kuu_columns(grid, hyp)]
    m = int(np.prod(sizes))
    X = torch.randn(m, c, dtype=dt, generator=torch.Generator().manual_seed(m + c))
    ref = o_kron([cc.double() for cc in cols], X.double())
    out = ops.kron_toeplitz_matmul(_pad_cols(cols, dt).to(DEV), sizes, X.to(DEV))
    tol = _tol(dt)
    tol["atol"] = tol["atol"] * max(1.0, float(ref.abs().max()))      # error relative to the output scale
    assert torch.allclose(out.cpu().double(), ref, **tol)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("sizes,c", [([20], 3), ([12, 9], 5), ([6, 7, 5], 4), ([8, 8, 8, 8], 6), ([40, 6], 3),
                                      ([32, 32], 17), ([32, 32], 32), ([32, 32, 32, 32], 16)])
def test_kron_toeplitz_backward(dt, sizes, c):
    ops = _ops()
    d = len(sizes)
    m = int(np.prod(sizes))
    gen = torch.Generator().manual_seed(3)
    cols = [torch.rand(g, generator=gen, dtype=torch.float64) for g in sizes]
    X = torch.randn(m, c, generator=gen, dtype=torch.float64)
    Z = torch.randn(m, c, generator=gen, dtype=torch.float64)
    cols_o = [cc.clone().requires_grad_(True) for cc in cols]
    Xo = X.clone().requires_grad_(True)
    (o_kron(cols_o, Xo) * Z).sum().backward()
    cg = _pad_cols(cols, dt).to(DEV).requires_grad_(True)
    Xg = X.to(dt).to(DEV).requires_grad_(True)
    (ops.kron_toeplitz_matmul(cg, sizes, Xg) * Z.to(dt).to(DEV)).sum().backward()
    tol = dict(rtol=1e-9, atol=1e-9) if dt == torch.float64 else dict(rtol=1e-3, atol=1e-2 * max(1.0, (m * c) ** 0.5 / 100))
    for i, g in enumerate(sizes):
        assert torch.allclose(cg.grad[i, :g].cpu().double(), cols_o[i].grad, **tol), f"grad col {i}"
        assert float(cg.grad[i, g:].abs().sum()) == 0.0
    assert torch.allclose(Xg.grad.cpu().double(), Xo.grad, **(_tol(dt) if dt == torch.float64 else dict(rtol=1e-3, atol=1e-3)))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m,r,r2", [(1000, 25, 25), (4096, 128, 128), (777, 100, 1), (5000, 432, 432), (300, 64, 5),
                                    (2048, 512, 512)])
def test_panel_rmul_and_gram(dt, m, r, r2):
    ops = _ops()
    gen = torch.Generator().manual_seed(0)
    P = torch.randn(m, r, generator=gen, dtype=dt)
    M = torch.randn(r, r2, generator=gen, dtype=dt)
    Bp = torch.randn(m, r2, generator=gen, dtype=dt)
    out = ops.panel_rmul(P.to(DEV), M.to(DEV))
    ref = P.double() @ M.double()
    # fp32 bound is the dot-product backward-error form |err_ij| <= c * ||row_i|| * ||col_j||.  The operands are exact to
    # ~2^-22 (3xTF32); what remains is the tensor core's fp32 accumulation, which TRUNCATES each add instead of rounding,
    # so a length-K chain carries a one-sided bias of up to K/8 half-ulps of the running sum (csrc/test_gemm_tc.cu
    # measures 3e-7 of ||a|| ||b|| at K = 4096), and the remainders of the 3xTF32 split are themselves cut to tf32 (2^-21 of
    # each operand, one-sided).  c = 3e-6 is 50 fp32 ulps of the norm product.  fp64 to rounding.
    def within(got, want, left, right, c=3e-6):
        bound = c * left.double().norm(dim=1)[:, None] * right.double().norm(dim=0)[None, :]
        return bool(((got.cpu().double() - want).abs() <= bound + 1e-30).all())
    if dt == torch.float64:
        assert torch.allclose(out.cpu().double(), ref, rtol=1e-10, atol=1e-9)
    else:
        assert within(out, ref, P, M)
    G = ops.gram(P.to(DEV), Bp.to(DEV))
    refg = P.double().t() @ Bp.double()
    if dt == torch.float64:
        assert torch.allclose(G.cpu().double(), refg, rtol=1e-10, atol=1e-8)
    else:
        assert within(G, refg, P.t(), Bp)


@pytest.mark.parametrize("dt", DT)
def test_gram_rmul_autograd(dt):
    ops = _ops()
    gen = torch.Generator().manual_seed(1)
    A = torch.randn(300, 20, generator=gen, dtype=torch.float64)
    B = torch.randn(300, 12, generator=gen, dtype=torch.float64)
    M = torch.randn(12, 7, generator=gen, dtype=torch.float64)
    Ao, Bo, Mo = (t.clone().requires_grad_(True) for t in (A, B, M))
    ((Ao.t() @ (Bo @ Mo)) ** 2).sum().backward()
    Ag, Bg, Mg = (t.to(dt).to(DEV).requires_grad_(True) for t in (A, B, M))
    (ops.gram(Ag, ops.panel_rmul(Bg, Mg)) ** 2).sum().backward()
    for g, o in ((Ag, Ao), (Bg, Bo), (Mg, Mo)):
        tol = dict(rtol=1e-9, atol=1e-7) if dt == torch.float64 else dict(rtol=1e-4, atol=2e-5 * float(o.grad.abs().max()))
        assert torch.allclose(g.grad.cpu().double(), o.grad, **tol)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m,r,q", [(1000, 30, 1), (513, 432, 3), (2000, 512, 8), (100, 100, 32)])
def test_lowrank_update(dt, m, r, q):
    ops = _ops()
    gen = torch.Generator().manual_seed(2)
    P = torch.randn(m, r, generator=gen, dtype=dt)
    U = torch.randn(r, q, generator=gen, dtype=dt) / r ** 0.5
    Vt = torch.randn(q, r, generator=gen, dtype=dt)
    ref = P.double() + (P.double() @ U.double()) @ Vt.double()
    Pg = P.to(DEV).clone()
    ops.panel_lowrank_update_(Pg, U.to(DEV), Vt.to(DEV))
    assert torch.allclose(Pg.cpu().double(), ref, **(_tol(dt) if dt == torch.float64 else dict(rtol=1e-4, atol=1e-4)))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m,r,c", [(3000, 40, 1), (5000, 432, 1), (4096, 512, 2), (1500, 100, 3), (2000, 700, 1)])
def test_q_matvec_and_cg(dt, m, r, c):
    ops = _ops()
    gen = torch.Generator().manual_seed(4)
    L = torch.randn(m, r, generator=gen, dtype=dt) / m ** 0.5
    KL = (L + 0.1 * torch.randn(m, r, generator=gen, dtype=dt) / m ** 0.5)
    v = torch.randn(r, c, generator=gen, dtype=dt)
    Q = torch.eye(r, dtype=torch.float64) + L.double().t() @ KL.double()
    w = ops.q_matvec(L.to(DEV), KL.to(DEV), v.to(DEV))
    tol = dict(rtol=1e-10, atol=1e-10) if dt == torch.float64 else dict(rtol=1e-4, atol=1e-4)
    assert torch.allclose(w.cpu().double(), Q @ v.double(), **tol)
    # CG on the symmetric part: use KL = K L with K = 2 I so that Q is SPD
    KL2 = 2.0 * L
    Q2 = torch.eye(r, dtype=torch.float64) + 2.0 * L.double().t() @ L.double()
    x, iters, resid = ops.cg_solve(L.to(DEV), KL2.to(DEV), v.to(DEV), tol=1e-8 if dt == torch.float64 else 1e-5,
                                   max_iter=200, check_every=4)
    ref = torch.linalg.solve(Q2, v.double())
    assert iters <= 200
    assert torch.allclose(x.cpu().double(), ref, **(dict(rtol=1e-6, atol=1e-7) if dt == torch.float64 else dict(rtol=1e-3, atol=1e-4)))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m,r,q", [(1000, 32, 1), (513, 432, 3), (2003, 512, 8), (300, 112, 16), (100, 96, 32), (7, 48, 6)])
def test_lowrank_update_both_panels_one_launch(dt, m, r, q):
    """wiski_panel_lowrank_update2: L and B of the rank-q root update in one launch (coefficients in shared memory,
    several rows per warp), every q bucket and ragged row counts."""
    ops = _ops()
    gen = torch.Generator().manual_seed(5)
    P0 = torch.randn(m, r, generator=gen, dtype=dt)
    P1 = torch.randn(m, r, generator=gen, dtype=dt)
    U = torch.randn(r, q, generator=gen, dtype=dt) / r ** 0.5
    V0 = torch.randn(q, r, generator=gen, dtype=dt)
    V1 = torch.randn(q, r, generator=gen, dtype=dt)
    r0 = P0.double() + (P0.double() @ U.double()) @ V0.double()
    r1 = P1.double() + (P1.double() @ U.double()) @ V1.double()
    g0, g1 = P0.to(DEV).clone(), P1.to(DEV).clone()
    ops.panel_lowrank_update2_(g0, g1, U.to(DEV), V0.to(DEV), V1.to(DEV))
    tol = _tol(dt) if dt == torch.float64 else dict(rtol=1e-4, atol=1e-4)
    assert torch.allclose(g0.cpu().double(), r0, **tol) and torch.allclose(g1.cpu().double(), r1, **tol)
    # the same launch with its by-product: P0 @ U of the rows before the update
    g0, g1 = P0.to(DEV).clone(), P1.to(DEV).clone()
    _, _, T = ops.panel_lowrank_update2_(g0, g1, U.to(DEV), V0.to(DEV), V1.to(DEV), return_t=True)
    assert torch.allclose(g0.cpu().double(), r0, **tol) and torch.allclose(g1.cpu().double(), r1, **tol)
    assert torch.allclose(T.cpu().double(), P0.double() @ U.double(), **tol)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("m,c,q", [(4096, 224, 1), (1000, 64, 3), (777, 30, 8), (2048, 432, 32)])
def test_panel_outer_add(dt, m, c, q):
    ops = _ops()
    gen = torch.Generator().manual_seed(9)
    P = torch.randn(m, c, generator=gen, dtype=dt)
    T = torch.randn(m, q, generator=gen, dtype=dt)
    W = torch.randn(q, c, generator=gen, dtype=dt)
    got = ops.panel_outer_add_(P.clone().to(DEV), T.to(DEV), W.to(DEV)).cpu()
    ref = P.double() + T.double() @ W.double()
    assert torch.allclose(got.double(), ref, **(dict(rtol=1e-12, atol=1e-12) if dt == torch.float64 else dict(rtol=1e-5, atol=1e-5)))


@pytest.mark.parametrize("m,r", [(4096, 128), (5000, 432), (2048, 512), (3000, 256)])
def test_gram_symmetric_and_single_pass_rmul(m, r):
    """fp32 tensor-core options: the Gram of a symmetric product skips the tiles below the diagonal (mirrored), and the
    gradient-only panel GEMM runs ONE tf32 pass (relative error ~1e-3: a uniform scale bias plus noise)."""
    ops = _ops()
    gen = torch.Generator().manual_seed(6)
    L = torch.randn(m, r, generator=gen) / m ** 0.5
    # an exactly symmetric product: KL = W L with W = diag(w)  =>  L^T W L is symmetric
    w = torch.rand(m, 1, generator=gen) + 0.5
    KL = L * w
    ref = L.double().t() @ KL.double()
    G = ops._gram(L.to(DEV), KL.to(DEV), symmetric=True)
    G0 = ops._gram(L.to(DEV), KL.to(DEV), symmetric=False)
    scale = float(ref.abs().max())
    assert torch.allclose(G.cpu().double(), ref, rtol=0.0, atol=1e-4 * scale), float((G.cpu().double() - ref).abs().max()) / scale
    assert torch.allclose(G.cpu(), G.cpu().t(), rtol=0.0, atol=1e-4 * scale)
    assert torch.allclose(G.cpu(), G0.cpu(), rtol=0.0, atol=1e-4 * scale)
    M = torch.randn(r, r, generator=gen)
    refz = L.double() @ M.double()
    Z3 = ops._rmul(L.to(DEV), M.to(DEV), terms=3).cpu().double()
    Z1 = ops._rmul(L.to(DEV), M.to(DEV), terms=1).cpu().double()
    zs = float(refz.abs().max())
    assert float((Z3 - refz).abs().max()) <= 2e-5 * zs, float((Z3 - refz).abs().max()) / zs
    e1 = float((Z1 - refz).abs().max())
    assert 3e-5 * zs <= e1 <= 4e-3 * zs, e1 / zs                          # one tf32 pass: coarser, and really taken
    # the error of the single pass is dominated by a uniform scale (operands truncated toward zero)
    alpha = float((Z1 * refz).sum() / (refz * refz).sum())
    assert 0.998 < alpha < 1.0, alpha
    # two passes (the default for gradient-only products): the small operand M is exact (its tf32 head and remainder are
    # stacked along K), the panel is one tf32 value -> ~4x closer than one pass, and still one sweep over the panel
    Z2 = ops._rmul(L.to(DEV), M.to(DEV), terms=2).cpu().double()
    e2 = float((Z2 - refz).abs().max())
    assert 2e-5 * zs <= e2 <= 1.5e-3 * zs and e2 < 0.8 * e1, (e2 / zs, e1 / zs)


def test_tensor_core_pair_kernels_match_simt():
    """The tcgen05 pair kernels (csrc/kron_tc.cu) against the SIMT pair kernels (csrc/kron_fused.cu) on the same inputs:
    forward K X on 32^4 and the directional backward (both write-out variants), plain layouts."""
    ops = _ops()
    from online_gp_b200 import _lib
    lib = _lib.load()
    sizes, c = [32, 32, 32, 32], 32
    m = 32 ** 4
    gen = torch.Generator().manual_seed(11)
    ell = torch.tensor([0.5, 0.8, 0.6, 0.7])
    grid = torch.linspace(-1.17, 1.17, 32)
    rr = (grid - grid[0]).abs().unsqueeze(0) / ell.unsqueeze(-1)
    cols = (torch.exp(-0.5 * rr * rr) * 0.8).to(DEV)
    dirs = (torch.exp(-0.5 * rr * rr) * rr * rr / ell.unsqueeze(-1)).to(DEV)
    X = (torch.randn(m, c, generator=gen) / 10).to(DEV)
    Z = (torch.randn(m, c, generator=gen) / 10).to(DEV)
    res = {}
    for on in (1, 0):
        prev = lib.wiski_kron_tc_enable(on)
        try:
            cg = cols.clone().requires_grad_(True)
            Y = ops.kron_toeplitz_matmul(cg, sizes, X, dirs=dirs)
            (Y * Z).sum().backward()
            res[on] = (Y.detach().clone(), cg.grad.clone())
        finally:
            lib.wiski_kron_tc_enable(prev)
    (Y1, g1), (Y0, g0) = res[1], res[0]
    assert torch.allclose(Y1, Y0, rtol=0.0, atol=3e-6 * float(Y0.abs().max()))
    assert torch.allclose(g1, g0, rtol=2e-4, atol=2e-4 * float(g0.abs().max()))


@pytest.mark.parametrize("W", [2, 4, 8])
def test_pushing_kernels_write_the_exchange_layout(W):
    """The compute + exchange kernels of the row-sharded path (``ops.*_push``), with every "peer" region in local memory:
    what lands at dst[j] must be exactly block j of the plain kernel's result — column block j (mode 1: slab-local pair
    pass and the panel GEMM) or the rows of axis-0 range j (mode 2: the pair passes that contain grid axis 0)."""
    import ctypes
    ops = _ops()
    gen = torch.Generator().manual_seed(13 + W)
    ell = torch.tensor([0.5, 0.8, 0.6, 0.7])
    grid = torch.linspace(-1.17, 1.17, 32)
    rr = (grid - grid[0]).abs().unsqueeze(0) / ell.unsqueeze(-1)
    cols = (torch.exp(-0.5 * rr * rr) * 0.8).to(DEV)
    dirs = (torch.exp(-0.5 * rr * rr) * rr * rr / ell.unsqueeze(-1)).to(DEV)
    g0 = 32 // W
    slab, full = [g0, 32, 32, 32], [32, 32, 32, 32]
    m_loc, m = g0 * 32 ** 3, 32 ** 4
    c = 32 * W                       # panel columns; cw = 32 per rank
    cw = c // W

    def table(buf):                  # buf [W, ...] contiguous: one destination per leading index
        return (ctypes.c_void_p * W)(*[buf[j].data_ptr() for j in range(W)])

    # mode 1: slab-local pair (1, 2) on a row slab [m_loc, c]
    X = (torch.randn(m_loc, c, generator=gen) / 10).to(DEV)
    want = ops._fused_pair_apply(cols, slab, (1, 2), X)
    got = torch.full((W, m_loc, cw), float("nan"), device=DEV)
    ops._fused_pair_apply_push(cols, slab, (1, 2), X, table(got), W, 1)
    for j in range(W):
        assert torch.equal(got[j], want[:, j * cw:(j + 1) * cw]), j
    # mode 1, panel GEMM: Z = P M, column block j at dst[j]
    M = torch.randn(c, c, generator=gen).to(DEV)
    wantz = ops._rmul(X, M, terms=3)
    gotz = torch.full((W, m_loc, cw), float("nan"), device=DEV)
    ops.rmul_push(X, M, table(gotz), W, terms=3)
    for j in range(W):
        assert torch.equal(gotz[j], wantz[:, j * cw:(j + 1) * cw]), j
    del X, want, got, wantz, gotz
    # mode 2: pair (0, 3) on a column block [m, cw]: rows of axis-0 range j at dst[j]
    Xc = (torch.randn(m, cw, generator=gen) / 10).to(DEV)
    Zc = (torch.randn(m, cw, generator=gen) / 10).to(DEV)
    want = ops._fused_pair_apply(cols, full, (0, 3), Xc)
    got = torch.full((W, m_loc, cw), float("nan"), device=DEV)
    ops._fused_pair_apply_push(cols, full, (0, 3), Xc, table(got), W, 2)
    assert torch.equal(got.view(m, cw), want)
    # mode 3: the same result through 32-column store boxes (two 16-column tiles per CTA step, the first one's result
    # parked in TMEM); also on a wider block (4 column groups per row of tiles)
    got.fill_(float("nan"))
    ops._fused_pair_apply_push(cols, full, (0, 3), Xc, table(got), W, 3)
    assert torch.equal(got.view(m, cw), want)
    if W == 2:
        Xw = (torch.randn(m, 128, generator=gen) / 10).to(DEV)
        wantw = ops._fused_pair_apply(cols, full, (0, 3), Xw)
        gotw = torch.full((W, m_loc, 128), float("nan"), device=DEV)
        ops._fused_pair_apply_push(cols, full, (0, 3), Xw, table(gotw), W, 3)
        assert torch.equal(gotw.view(m, 128), wantw)
        del Xw, wantw, gotw
    o_want = torch.zeros(3, dtype=torch.float64, device=DEV)
    o_got = torch.zeros(3, dtype=torch.float64, device=DEV)
    wantz = ops._fused_pair_grad_dir(cols, dirs, full, (0, 3), Zc, Xc, o_want, store=True)
    gotz = torch.full((W, m_loc, cw), float("nan"), device=DEV)
    ops._fused_pair_grad_dir_push(cols, dirs, full, (0, 3), Zc, Xc, o_got, table(gotz), W)
    assert torch.equal(gotz.view(m, cw), wantz)
    assert torch.allclose(o_got, o_want, rtol=1e-6, atol=0.0)


def test_kron_directional_backward_matches_full_gradient():
    """32^4 fused path: the surrogate column gradient of the directional backward has the same inner products with
    dirs[i] and cols[i] as the full column gradient (which is all lengthscale / scale parameters see)."""
    ops = _ops()
    sizes, c = [32, 32, 32, 32], 16
    m = 32 ** 4
    gen = torch.Generator().manual_seed(7)
    ell = torch.tensor([0.4, 0.7, 0.5, 0.9])
    grid = torch.linspace(-1.17, 1.17, 32)
    r = (grid - grid[0]).abs().unsqueeze(0) / ell.unsqueeze(-1)
    cols = torch.exp(-0.5 * r * r) * 0.7
    dirs = torch.exp(-0.5 * r * r) * r * r / ell.unsqueeze(-1)
    X = torch.randn(m, c, generator=gen) / 30
    Z = torch.randn(m, c, generator=gen) / 30
    outs = []
    for use_dirs in (False, True):
        cg = cols.clone().to(DEV).requires_grad_(True)
        Y = ops.kron_toeplitz_matmul(cg, sizes, X.to(DEV), dirs=dirs.to(DEV) if use_dirs else None)
        (Y * Z.to(DEV)).sum().backward()
        outs.append(cg.grad.cpu().double())
    full, sur = outs
    for i in range(4):
        for v in (dirs[i].double(), cols[i].double()):
            a, b = float(full[i] @ v), float(sur[i] @ v)
            assert abs(a - b) <= 2e-3 * max(abs(a), float(full[i].abs().max() * v.abs().max())), (i, a, b)


# ---- tensor-core (tcgen05, 3xTF32) form of one Kronecker axis: g >= 64, fp32
TC_AXIS_CASES = [(64, 3, 80), (128, 1, 4096), (128, 5, 64), (256, 2, 528), (1024, 1, 2048), (1024, 3, 432),
                 (100, 2, 68), (72, 7, 4 * 37)]


@pytest.mark.parametrize("g,outer,inner", TC_AXIS_CASES)
def test_kron_axis_tensor_core_apply_and_contract(g, outer, inner):
    ops = _ops()
    from online_gp_b200 import _lib, settings as S
    assert _lib.load().wiski_kron_axis_tc_work_elems(g, outer, inner, 0) == g * g        # the fast path is taken
    assert _lib.load().wiski_kron_axis_tc_work_elems(g, outer, inner, 1) >= 2 * g * g
    gen = torch.Generator().manual_seed(g + outer + inner)
    col = (torch.exp(-0.002 * torch.arange(g, dtype=torch.float64) ** 2) + 0.05 * torch.rand(g, generator=gen, dtype=torch.float64))
    X = torch.randn(outer * g * inner, 1, generator=gen, dtype=torch.float32)
    Z = torch.randn(outer * g * inner, 1, generator=gen, dtype=torch.float32)
    ar = torch.arange(g)
    T = col[(ar.unsqueeze(0) - ar.unsqueeze(1)).abs()]
    X3, Z3 = X.double().reshape(outer, g, inner), Z.double().reshape(outer, g, inner)
    ref_apply = torch.einsum("ab,obw->oaw", T, X3).reshape(-1, 1)
    Sfull = torch.einsum("oaw,obw->ab", Z3, X3)
    ref_acc = torch.zeros(g, dtype=torch.float64).index_add_(0, (ar.unsqueeze(0) - ar.unsqueeze(1)).abs().reshape(-1),
                                                             Sfull.reshape(-1))
    Xd, Zd, cd = X.to(DEV), Z.to(DEV), col.float().to(DEV)
    out = ops.kron_axis_apply(Xd, cd, g, outer, inner)
    scale = float(ref_apply.abs().max())
    assert torch.allclose(out.cpu().double(), ref_apply, rtol=2e-4, atol=2e-5 * scale)
    acc = torch.zeros(g, dtype=torch.float64, device=DEV)
    acc += 1.0                                                # accumulates onto the caller's values
    ops.kron_axis_contract(Zd, Xd, g, outer, inner, acc)
    assert torch.allclose(acc.cpu() - 1.0, ref_acc, rtol=1e-3, atol=1e-4 * float(Sfull.abs().max()) * g ** 0.5)
    # and it agrees with the SIMT kernels of the same entry points
    with S.kron_tensor_core_axes(False):
        out_simt = ops.kron_axis_apply(Xd, cd, g, outer, inner)
        acc_simt = torch.zeros(g, dtype=torch.float64, device=DEV)
        ops.kron_axis_contract(Zd, Xd, g, outer, inner, acc_simt)
    assert torch.allclose(out, out_simt, rtol=2e-4, atol=2e-5 * scale)
    assert torch.allclose(acc - 1.0, acc_simt, rtol=1e-3, atol=1e-4 * float(Sfull.abs().max()) * g ** 0.5)


@pytest.mark.parametrize("sizes,c", [([128, 128], 64), ([64, 64, 64], 16), ([256, 256], 64), ([1024, 1024], 16)])
def test_kron_large_axes_forward_backward_tensor_core(sizes, c):
    """Full K X and its column gradient on grids with >= 64 points per axis (fp32): tensor-core path vs the fp64 oracle
    (forward) and vs the SIMT path (gradient)."""
    ops = _ops()
    from online_gp_b200 import settings as S
    d = len(sizes)
    m = int(np.prod(sizes))
    grid = create_grid(sizes, [(-1.1, 1.1)] * d)
    hyp = Hypers(d, kind="rbf")
    cols = [cc.detach() for cc in kuu_columns(grid, hyp)]
    gen = torch.Generator().manual_seed(m + c)
    X = torch.randn(m, c, generator=gen, dtype=torch.float32)
    Z = torch.randn(m, c, generator=gen, dtype=torch.float32)
    ref = o_kron([cc.double() for cc in cols], X.double())
    grads = []
    for tc in (True, False):
        with S.kron_tensor_core_axes(tc):
            cg = _pad_cols(cols, torch.float32).to(DEV).requires_grad_(True)
            out = ops.kron_toeplitz_matmul(cg, sizes, X.to(DEV))
            (out * Z.to(DEV)).sum().backward()
            grads.append(cg.grad.clone())
            assert torch.allclose(out.detach().cpu().double(), ref, rtol=2e-4, atol=2e-5 * float(ref.abs().max()))
            with torch.no_grad():
                out2 = ops.kron_toeplitz_matmul(cg.detach(), sizes, X.to(DEV))
            assert torch.allclose(out2, out.detach(), rtol=1e-5, atol=1e-5 * float(ref.abs().max()))
    assert torch.allclose(grads[0], grads[1], rtol=1e-3, atol=1e-4 * float(grads[1].abs().max()))


@pytest.mark.parametrize("slab,pair,c,W", [([8, 32, 32, 32], 1, 64, 4), ([32, 32, 32, 32], 1, 32, 2), ([32, 32], 0, 48, 3),
                                            ([4, 32, 32, 32], 1, 128, 8)])
def test_fused_pair_kernels_chunked_layouts(slab, pair, c, W):
    """Column-chunked operand layouts of the fused pair kernels (send / receive buffers of the sharded path's
    all-to-all): same numbers as the plain layout, bit for bit."""
    ops = _ops()
    d = len(slab)
    m = int(np.prod(slab))
    gen = torch.Generator().manual_seed(m + c + W)
    cols = (torch.rand(d, 32, generator=gen) * 0.1 + torch.exp(-0.05 * torch.arange(32.0) ** 2)).to(DEV)
    X = torch.randn(m, c, generator=gen).to(DEV)
    Z = torch.randn(m, c, generator=gen).to(DEV)
    Y = ops._fused_pair_apply(cols, slab, pair, X)
    Yc = ops._fused_pair_apply(cols, slab, pair, X, chunk_out=W)
    assert Yc.shape == (W, m, c // W)
    assert torch.equal(Yc, Y.view(m, W, c // W).permute(1, 0, 2).contiguous())
    acc0 = torch.zeros(d, 32, dtype=torch.float64, device=DEV)
    acc1 = torch.zeros(d, 32, dtype=torch.float64, device=DEV)
    Zo0 = ops._fused_pair_grad(cols, slab, pair, Z, X, acc0, store=True)
    Zch = Z.view(m, W, c // W).permute(1, 0, 2).contiguous()
    Zo1 = ops._fused_pair_grad(cols, slab, pair, Zch, X, acc1, store=True, chunk_z=W)
    assert torch.equal(Zo0, Zo1)
    assert torch.allclose(acc0, acc1, rtol=1e-12, atol=1e-12 * float(acc0.abs().max()))
    acc2 = torch.zeros(d, 32, dtype=torch.float64, device=DEV)
    assert ops._fused_pair_grad(cols, slab, pair, Zch, X, acc2, store=False, chunk_z=W) is None
    assert torch.allclose(acc0, acc2, rtol=1e-12, atol=1e-12 * float(acc0.abs().max()))


@pytest.mark.parametrize("m,r,nb", [(4096, 128, 2), (8192, 448, 2), (6000, 512, 8), (4096, 256, 4), (2048, 96, 2)])
def test_gram_and_rmul_on_column_blocks(m, r, nb):
    """Column-chunked Gram / panel right-multiply (one tensor-core launch over the all-to-all's block layout, or the
    per-block fallback when (r / nb) % 32 != 0) against fp64."""
    ops = _ops()
    gen = torch.Generator().manual_seed(m + r + nb)
    A = torch.randn(m, r, generator=gen)
    B = torch.randn(m, r, generator=gen)
    M = torch.randn(r, r, generator=gen)
    cwb = r // nb
    Bb = B.view(m, nb, cwb).permute(1, 0, 2).contiguous()
    G = ops.gram_blocks(A.to(DEV), Bb.to(DEV))
    ref = A.double().t() @ B.double()
    assert torch.allclose(G.cpu().double(), ref, rtol=2e-4, atol=2e-5 * float(ref.abs().max()))
    Ob = ops.rmul_blocks(A.to(DEV), M.to(DEV), nb)
    refO = (A.double() @ M.double()).view(m, nb, cwb).permute(1, 0, 2)
    assert Ob.shape == (nb, m, cwb)
    assert torch.allclose(Ob.cpu().double(), refO, rtol=2e-4, atol=2e-5 * float(refO.abs().max()))
