"""CPU, world_size 2, gloo: host-side logic of the row-sharded path (partition, exchanges, autograd completeness)
against the unsharded oracle loop.  CUDA entry points are mocked by the oracle (tests/cpu_ops_mock.py); the sharded
kernels themselves run in the -m gpu suite / bench.py --gpus N."""
import os
import sys
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, d, g, n0, steps, ret, dual=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_ops_mock
    from online_gp_b200 import settings as S
    from online_gp_b200.parallel import Comm, ShardedOnlineSKIRegression
    warnings.simplefilter("ignore")
    gen = torch.Generator().manual_seed(7)
    X = torch.rand(n0 + steps * 2, d, generator=gen) * 2 - 1
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps * 2, generator=gen)).unsqueeze(-1)
    out = []
    # max_cholesky_size(0): low-rank initial root AND every Q solve of predict() through the sharded CG driver (one
    # all-reduce per iteration); a tight tolerance makes it comparable with the oracle's dense solves
    with cpu_ops_mock.install(), S.max_cholesky_size(0), S.max_root_decomposition_size(64), S.eval_cg_tolerance(1e-13), \
            S.sharded_dual_layout(dual):
        model = ShardedOnlineSKIRegression(X[:n0], y[:n0], lr=1e-2, grid_size=g, grid_bound=1.0, comm=Comm())
        for t in range(steps):
            xt, yt = X[n0 + 2 * t:n0 + 2 * t + 2], y[n0 + 2 * t:n0 + 2 * t + 2]
            rmse, nll = model.evaluate(xt, yt)
            _, loss = model.update(xt, yt)
            out.append((rmse, nll, loss, float(model._noise())))
        ls = model.covar_module.base_kernel.base_kernel.lengthscale.detach().reshape(-1).tolist()
    assert model.last_cg[0] >= 10 and model.last_cg[1] < 1e-12          # the CG path really ran and converged
    assert (model.Lc is not None) == dual
    if dual:        # the column-sharded copy followed every rank-q update: it equals the exchanged row slab
        assert torch.allclose(model.Lc, model._rows_to_cols(model.L_loc), rtol=1e-10, atol=1e-12)
    ret[rank] = (out, ls, model.L_loc.shape)
    dist.barrier()
    dist.destroy_process_group()


def _oracle_loop(d, g, n0, steps):
    from oracle.gridkernel import Hypers
    from oracle.interp import create_grid
    from oracle.wiski_matfree import WiskiMatFree
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(7)
    X = torch.rand(n0 + steps * 2, d, generator=gen) * 2 - 1
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps * 2, generator=gen)).unsqueeze(-1)
    grid = create_grid([g] * d, [(-1.1, 1.1)] * d)
    hyp = Hypers(d, learn_noise=True)
    orc = WiskiMatFree(grid, hyp, X[:n0], y[:n0, 0], torch.ones(n0), max_cholesky_size=0, max_root=64, update_mode="svd")
    opt = torch.optim.Adam(hyp.params(), lr=1e-2)
    out = []
    for t in range(steps):
        xt, yt = X[n0 + 2 * t:n0 + 2 * t + 2], y[n0 + 2 * t:n0 + 2 * t + 2, 0]
        with torch.no_grad():
            mo, co = orc.predict(xt)
            var = co.diagonal() + hyp.noise
            rmse = float((mo - yt).pow(2).mean().sqrt())
            nll = float(-torch.distributions.Normal(mo, var.sqrt()).log_prob(yt).mean())
        opt.zero_grad()
        mll = orc.mll()
        (-mll).backward()
        opt.step()
        orc.condition_on_observations(xt, yt, torch.ones(2))
        out.append((rmse, nll, float(hyp.noise)))
    torch.set_default_dtype(torch.float32)
    return out, hyp.lengthscale.detach().tolist()


@pytest.mark.parametrize("d,g,n0", [(2, 8, 20), (3, 6, 30), (2, 10, 90)])   # last: n0 > root rank (batched fold-in)
@pytest.mark.parametrize("dual", [False, True])
def test_sharded_stream_matches_unsharded_oracle(d, g, n0, dual):
    """dual = True: ``settings.sharded_dual_layout`` (K L and its gradient computed on a column-sharded copy of the
    root panel: two exchanges per step instead of four)."""
    world, steps = 2, 3
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + (7 if dual else 0)
    mp.spawn(_worker, args=(world, port, d, g, n0, steps, ret, dual), nprocs=world, join=True)
    ref, ls_ref = _oracle_loop(d, g, n0, steps)
    for rank in range(world):
        out, ls, shape = ret[rank]
        assert shape[0] == g ** d // world                      # each rank holds half of the grid rows
        for (rmse, nll, loss, noise), (rmse_o, nll_o, noise_o) in zip(out, ref):
            assert abs(rmse - rmse_o) <= 1e-6 * max(1.0, abs(rmse_o))
            assert abs(nll - nll_o) <= 1e-6 * max(1.0, abs(nll_o))
            assert abs(noise - noise_o) <= 1e-8                  # identical Adam trajectory => complete gradients
        assert all(abs(a - b) <= 1e-8 for a, b in zip(ls, ls_ref))
    assert ret[0][0] == ret[1][0]                               # replicas agree exactly


def test_shard_plan_and_layout_roundtrip():
    sys.path.insert(0, ROOT)
    from online_gp_b200.parallel import Comm, ShardPlan, _to_cols, _to_rows
    plan = ShardPlan([8, 4, 6], world=4, rank=2)
    assert (plan.m, plan.m_loc, plan.row0, plan.g0_loc, plan.rest) == (192, 48, 96, 2, 24)
    idx = torch.tensor([[0, 95, 96, 143, 144]])
    il, vl = plan.localize(idx, torch.ones(1, 5))
    assert il.tolist() == [[0, 0, 0, 47, 0]] and vl.tolist() == [[0, 0, 1, 1, 0]]
    assert plan.local_axes(5) == [(1, 4, 2, 30), (2, 6, 8, 5)]
    with pytest.raises(ValueError):
        ShardPlan([6, 4], world=4, rank=0)
    one = ShardPlan([4, 4], world=1, rank=0)
    X = torch.randn(16, 3)
    assert torch.equal(_to_rows(_to_cols(X, one, Comm()), one, Comm()), X)


# ---- the fused (default multi-GPU) orchestration of _ShardedKronFn on CPU: the pair kernels are emulated for small
# axes (4 points instead of 32), including the column-chunked layouts of the all-to-all send / receive buffers
def _toeplitz(col, g):
    ar = torch.arange(g)
    return col[(ar.unsqueeze(0) - ar.unsqueeze(1)).abs()]


def _install_fused_emulation(ops, parallel, axis):
    def pair_views(sizes, pair, X):
        """X [m, c] viewed as [before, u, mid, v, after, c] for the axes of `pair` (index p or (axis_u, axis_v))."""
        au, av = ops._pair_axes(pair)
        c = X.shape[1]
        u, v = sizes[au], sizes[av]
        before = mid = 1
        for s_ in sizes[:au]:
            before *= s_
        for s_ in sizes[au + 1:av]:
            mid *= s_
        after = X.shape[0] // (before * u * mid * v)
        return X.reshape(before, u, mid, v, after, c), u, v, au, av

    def unchunk(Zb):                        # [W, m, cw] -> [m, W cw]
        W, m, cw = Zb.shape
        return Zb.permute(1, 0, 2).reshape(m, W * cw)

    def pair_apply(cols, sizes, pair, X, chunk_out=1, out=None):
        V, u, v, au, av = pair_views(sizes, pair, X)
        Tu, Tv = _toeplitz(cols[au], u), _toeplitz(cols[av], v)
        Y = torch.einsum("ab,cd,obmdwk->oamcwk", Tu, Tv, V).reshape(X.shape)
        if chunk_out > 1:
            m, c = X.shape
            Y = Y.view(m, chunk_out, c // chunk_out).permute(1, 0, 2).contiguous()
        return Y if out is None else out.copy_(Y)

    def pair_grad(cols, sizes, pair, Z, P, acc, store, chunk_z=1, zout=None):
        if chunk_z > 1:
            Z = unchunk(Z)
        Zv, u, v, au, av = pair_views(sizes, pair, Z)
        Pv = pair_views(sizes, pair, P)[0]
        Tu, Tv = _toeplitz(cols[au], u), _toeplitz(cols[av], v)
        TvP = torch.einsum("cd,obmdwk->obmcwk", Tv, Pv)
        TuZ = torch.einsum("ab,obmdwk->oamdwk", Tu, Zv)
        Su = torch.einsum("oamdwk,obmdwk->ab", Zv.double(), TvP.double())
        Sv = torch.einsum("oamdwk,oamewk->de", TuZ.double(), Pv.double())
        for S, g, slot in ((Su, u, au), (Sv, v, av)):
            ar = torch.arange(g)
            off = (ar.unsqueeze(0) - ar.unsqueeze(1)).abs().reshape(-1)
            acc[slot][:g] += torch.zeros(g, dtype=torch.float64).index_add_(0, off, S.reshape(-1))
        if not store:
            return None
        Zo = torch.einsum("cd,oamdwk->oamcwk", Tv, TuZ).reshape(P.shape)
        return Zo if zout is None else zout.copy_(Zo)

    def pair_grad_dir(cols, dirs, sizes, pair, Z, P, out3, store, chunk_z=1, zout=None):
        if chunk_z > 1:
            Z = unchunk(Z)
        Zv, u, v, au, av = pair_views(sizes, pair, Z)
        Pv = pair_views(sizes, pair, P)[0]
        Tu, Tv = _toeplitz(cols[au], u), _toeplitz(cols[av], v)
        Du, Dv = _toeplitz(dirs[au], u), _toeplitz(dirs[av], v)
        S = torch.einsum("cd,obmdwk->obmcwk", Tv, Pv).double()
        zu = torch.einsum("ab,obmdwk->oamdwk", Tu, Zv)
        zd = torch.einsum("ab,obmdwk->oamdwk", Du, Zv).double()
        zd2 = torch.einsum("cd,oamdwk->oamcwk", Dv, zu).double()
        out3[0] += (zd * S).sum()
        out3[1] += (zd2 * Pv.double()).sum()
        out3[2] += (zu.double() * S).sum()
        if not store:
            return None
        Zo = torch.einsum("cd,oamdwk->oamcwk", Tv, zu).reshape(P.shape)
        return Zo if zout is None else zout.copy_(Zo)

    # ---- the pushing kernels (results stored straight into the peers' regions): emulated by queueing the pieces per
    # destination on the fake push buffers; `barrier()` then delivers them with one gloo all-gather per queued push
    def split_for_push(Y, n_dst, mode):
        if mode == 1:                                        # column block j -> rank j
            cw = Y.shape[1] // n_dst
            return [Y[:, j * cw:(j + 1) * cw].contiguous() for j in range(n_dst)]
        rows = Y.shape[0] // n_dst                           # modes 2 / 3: rows of axis-0 range j -> rank j
        return [Y[j * rows:(j + 1) * rows].contiguous() for j in range(n_dst)]

    def pair_apply_push(cols, sizes, pair, X, dst, n_dst, mode):
        dst.queue(split_for_push(pair_apply(cols, sizes, pair, X), n_dst, mode))

    def pair_grad_dir_push(cols, dirs, sizes, pair, Z, P, out3, dst, n_dst):
        dst.queue(split_for_push(pair_grad_dir(cols, dirs, sizes, pair, Z, P, out3, store=True), n_dst, 2))

    def rmul_push(P, M, dst, n_dst, terms=3):
        dst.queue(split_for_push(P @ M, n_dst, 1))

    ops._fused_pair_apply_push, ops._fused_pair_grad_dir_push, ops.rmul_push = pair_apply_push, pair_grad_dir_push, rmul_push
    saved = (ops._fused_pair_apply, ops._fused_pair_grad, parallel._fused_ok, ops._fused_pair_grad_dir)
    ops._fused_pair_apply, ops._fused_pair_grad, ops._fused_pair_grad_dir = pair_apply, pair_grad, pair_grad_dir
    parallel._fused_ok = lambda plan, X: plan.d == 4 and all(s == axis for s in plan.sizes) and \
        X.shape[1] % (16 * plan.world) == 0
    return saved


class _FakePushBuffers:
    """CPU stand-in of parallel._PushBuffers: the same four regions, `dst()` hands the emulated pushing kernels a handle
    that queues their per-destination pieces, and `barrier()` delivers every queued push (piece j of rank p lands in part
    p of the region on rank j) — what the NVLink stores + the symmetric-memory barrier do on the GPUs."""
    REGIONS = ("A", "B", "C", "D")

    class _Dst:
        def __init__(self, owner, name, part):
            self.owner, self.name, self.part = owner, name, part

        def queue(self, pieces):
            self.owner.pending.append((self.name, self.part, pieces))

    def __init__(self, numel, dtype, rank, world):
        # one tensor per region: a tensor saved for the backward (region A) must not share its version counter with the
        # regions written later (the CUDA kernels write through raw pointers, torch ops here do not)
        self.regions = {n: torch.zeros(numel, dtype=dtype) for n in self.REGIONS}
        self.buf = self.regions["A"]            # (dtype probe of Comm.push_buffers)
        self.numel, self.rank, self.world = numel, rank, world
        self.pending, self.c_pushed, self.barriers = [], False, 0

    def local(self, name, numel=None):
        return self.regions[name][:self.numel if numel is None else numel]

    def dst(self, name, part):
        return self._Dst(self, name, part)

    def barrier(self):
        self.barriers += 1
        for name, part, pieces in self.pending:
            assert all(p.numel() == part for p in pieces), (name, part, [p.shape for p in pieces])
            mine = torch.stack([p.reshape(-1) for p in pieces])                 # [W, part]: row j goes to rank j
            allp = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allp, mine)
            reg = self.local(name)
            for src in range(self.world):
                reg[src * part:(src + 1) * part] = allp[src][self.rank]
        self.pending = []


def _fused_worker(rank, world, port, ret, directional=True, push=False, dual=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_ops_mock
    from online_gp_b200 import ops, parallel, settings as S
    warnings.simplefilter("ignore")
    gen = torch.Generator().manual_seed(3)
    d, g, n0, steps = 4, 4, 40, 3
    X = torch.rand(n0 + steps, d, generator=gen) * 2 - 1
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps, generator=gen)).unsqueeze(-1)
    out = []
    with cpu_ops_mock.install(), S.max_cholesky_size(0), S.max_root_decomposition_size(64), S.eval_cg_tolerance(1e-13), \
            S.kron_directional_grad(directional), S.sharded_dual_layout(dual):
        saved = _install_fused_emulation(ops, parallel, g)
        try:
            if directional:
                # through the public class: OnlineSKIRegression(..., comm=) drives the row-sharded engine
                from online_gp_b200.models import OnlineSKIRegression
                from online_gp_b200.models.stems import Identity
                front = OnlineSKIRegression(Identity(d), X[:n0], y[:n0], lr=1e-2, grid_size=g, grid_bound=1.0,
                                            comm=parallel.Comm())
                model = front.engine
                assert isinstance(model, parallel.ShardedOnlineSKIRegression) and front.gp is model
            else:
                front = model = parallel.ShardedOnlineSKIRegression(X[:n0], y[:n0], lr=1e-2, grid_size=g, grid_bound=1.0,
                                                                    comm=parallel.Comm())
            assert parallel._fused_ok(model.plan, model.L_loc)
            assert (model.Lc is not None) == dual
            if push:        # what enable_push sets up on GPUs with peer memory
                model.comm.push = _FakePushBuffers(model.L_loc.numel(), model.L_loc.dtype, rank, world)
            for t in range(steps):
                xt, yt = X[n0 + t:n0 + t + 1], y[n0 + t:n0 + t + 1]
                rmse, nll = front.evaluate(xt, yt)
                P = model.pieces()
                assert P["KL"].shape == (world, model.plan.m_loc, model.L_loc.shape[1] // world)     # column blocks
                _, loss = front.update(xt, yt)
                out.append((rmse, nll, loss, float(model._noise())))
            if push:        # pushes per step: dual layout 2 (K L, Z), single layout 4; every one followed by its barrier
                assert model.comm.push.barriers == steps * (2 if dual else 4) and not model.comm.push.pending
                assert model.comm.push.c_pushed is False
        finally:
            ops._fused_pair_apply, ops._fused_pair_grad, parallel._fused_ok, ops._fused_pair_grad_dir = saved
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("directional,push,dual", [(True, False, False), (False, False, False), (True, True, False),
                                                   (True, True, True)])
def test_fused_sharded_orchestration_matches_unsharded_oracle(directional, push, dual):
    """The multi-GPU paths on the fused pair kernels (emulated on a 4^4 grid) against the unsharded oracle:
    pull exchange (slab-local pair, all-to-all, column-sharded pair, column blocks through Gram / rmul / gathers, chunked
    operands in the backward) with the directional and the full column-gradient passes; and the pushing kernels — results
    stored into the peers' regions, barrier protocol, gradient panel pushed by the Gram backward — in the single-layout
    (4 pushes per step) and the dual-layout form (2 pushes per step: the GPU default)."""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + 13 + int(directional) + 2 * int(push) + 4 * int(dual)
    mp.spawn(_fused_worker, args=(world, port, ret, directional, push, dual), nprocs=world, join=True)
    from oracle.gridkernel import Hypers
    from oracle.interp import create_grid
    from oracle.wiski_matfree import WiskiMatFree
    torch.set_default_dtype(torch.float64)
    gen = torch.Generator().manual_seed(3)
    d, g, n0, steps = 4, 4, 40, 3
    X = torch.rand(n0 + steps, d, generator=gen) * 2 - 1
    y = (torch.sin(3 * X.sum(-1)) + 0.1 * torch.randn(n0 + steps, generator=gen)).unsqueeze(-1)
    hyp = Hypers(d, learn_noise=True)
    orc = WiskiMatFree(create_grid([g] * d, [(-1.1, 1.1)] * d), hyp, X[:n0], y[:n0, 0], torch.ones(n0),
                       max_cholesky_size=0, max_root=64, update_mode="svd")
    opt = torch.optim.Adam(hyp.params(), lr=1e-2)
    for t in range(steps):
        xt, yt = X[n0 + t:n0 + t + 1], y[n0 + t:n0 + t + 1, 0]
        with torch.no_grad():
            mo, co = orc.predict(xt)
            var = co.diagonal() + hyp.noise
            rmse_o = float((mo - yt).pow(2).mean().sqrt())
            nll_o = float(-torch.distributions.Normal(mo, var.sqrt()).log_prob(yt).mean())
        opt.zero_grad()
        (-orc.mll()).backward()
        opt.step()
        orc.condition_on_observations(xt, yt, torch.ones(1))
        for rank in range(world):
            rmse, nll, loss, noise = ret[rank][t]
            assert abs(rmse - rmse_o) <= 1e-6 * max(1.0, abs(rmse_o)) and abs(nll - nll_o) <= 1e-6 * max(1.0, abs(nll_o))
            assert abs(noise - float(hyp.noise)) <= 1e-8            # complete hyper-gradients on every rank
    torch.set_default_dtype(torch.float32)
    assert ret[0] == ret[1]
