"""Cases for the "next" rows of SURVEY §8f-1/2/3 — the BoTorch-facing wrapper, batched posteriors over candidate
sets, fantasies on shared panels / in predictive space, the exact predictive root under ``fast_pred_samples``, the BO
and active-learning loops in miniature, ``fit()`` and the trainable-stem streaming step.  Shared by
``test_next_rows_host_cpu.py`` (kernels mocked by the oracle) and ``test_next_rows_gpu.py`` (real kernels, ``-m gpu``);
``DEV`` is set by the importing module.  Random draws come from CPU generators (same numbers on both devices)."""
import warnings

import pytest
import torch

DEV = "cuda:0"
#: tolerance multiplier: the mocked CPU kernels are deterministic dense algebra; on the GPU the Cholesky root of the
#: rank-deficient W^T D^-1 W (+ 1e-8 jitter) amplifies the kernels' different summation orders to ~1e-8 relative
TOLX = 1.0


def _close(a, b, rtol, atol=0.0):
    ok = torch.allclose(a, b, rtol=min(rtol * TOLX, 1e-4), atol=min(atol * TOLX, 1e-6) if atol else 0.0)
    if not ok:          # shown by pytest on failure
        print("mismatch: max abs diff %.3e, max |b| %.3e, dtypes %s %s" % (float((a - b).abs().max()), float(b.abs().max()),
                                                                         a.dtype, b.dtype))
    return ok


def _rand(*size, generator=None):
    return torch.rand(*size, generator=generator).to(DEV)


def _randn(*size, generator=None):
    return torch.randn(*size, generator=generator).to(DEV)


def _ones(*size):
    return torch.ones(*size).to(DEV)


def _tensor(data):
    return torch.tensor(data).to(DEV)


@pytest.fixture(autouse=True)
def _fp64_default():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(prev)


def _model(t=1, learn=True, n0=25, d=2, g=8, seed=0):
    from online_gp_b200.models import OnlineSKIBotorchModel
    gen = torch.Generator().manual_seed(seed)
    X = _rand(n0, d, generator=gen)
    Y = torch.stack([torch.sin(3 * X.sum(-1) + o) for o in range(t)], dim=-1) + 0.05 * _randn(n0, t, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = OnlineSKIBotorchModel(X, Y, 0.1 * _ones(n0, t), grid_bounds=_tensor([[0.0, 1.0]] * d),
                                      grid_size=g, learn_additional_noise=learn)
    return model, X, Y, gen


@pytest.mark.parametrize("t,learn", [(1, True), (1, False), (2, True)])
def test_fantasy_batch_equals_one_model_per_draw(t, learn):
    """condition_on_observations(X, Y[nf, q, t]) == nf separately conditioned models: means per draw, one shared
    covariance (reference: get_fantasy_model expands every cache per fantasy, batched_fixed_noise_online_gp.py:287-332)."""
    model, X, Y, gen = _model(t=t, learn=learn)
    nf, q, d = 5, 3, X.shape[-1]
    Xn = _rand(q, d, generator=gen)
    Yf = _randn(nf, q, t, generator=gen)
    noise = 0.2 * _ones(q, t)
    Xs = _rand(6, d, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.eval()
        fm = model.condition_on_observations(Xn, Yf, noise)
        assert fm.num_fantasies == nf and fm.num_data == model.num_data + q
        dist = fm(Xs)
        for f in range(nf):
            one = model.condition_on_observations(Xn, Yf[f], noise, inplace=False)
            one.eval()
            ref = one(Xs)
            mean_f = dist.mean[f]
            assert _close(mean_f, ref.mean, 1e-9, 1e-11)
            assert _close(dist.variance[f], ref.variance, 1e-9, 1e-12)
        assert dist.covariance_matrix.shape[0] == nf
        # the base model is untouched
        assert model.num_data == X.shape[0]


def test_botorch_wrapper_posterior_and_fantasize():
    from online_gp_b200.models import FantasizedOnlineSKIGP, GPyTorchPosterior
    model, X, Y, gen = _model(t=1, learn=True)
    assert model._is_custom_likelihood is True
    Xs = _rand(4, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        post = model.posterior(Xs.float())                    # posterior() casts to the model dtype (:66)
        assert isinstance(post, GPyTorchPosterior) and post.mean.shape == (4, 1) and post.variance.shape == (4, 1)
        assert post.rsample(torch.Size([7])).shape == (7, 4, 1)
        same = model.posterior(Xs.unsqueeze(0))               # forward squeezes a leading singleton batch (:36-40)
        assert torch.allclose(same.mean, post.mean)
        batched = model.posterior(_rand(3, 4, 2, generator=gen))
        assert batched.mean.shape == (3, 4, 1) and batched.variance.shape == (3, 4, 1)

        def sampler(posterior):
            return posterior.rsample(torch.Size([6]))
        Xn = _rand(2, 2, generator=gen)
        fm = model.fantasize(Xn, sampler)
        assert isinstance(fm, FantasizedOnlineSKIGP) and fm.num_fantasies == 6
        fpost = fm.posterior(Xs)
        # a model re-hydrated from a kernel cache keeps the leading output dimension (reference: `_batch_shape` is then
        # an int, batched_fixed_noise_online_gp.py:88,247-250), so fantasies are [nf, t, q*] (+ BoTorch's trailing 1)
        assert fpost.mean.shape == (6, 1, 4, 1) and fpost.variance.shape == (6, 1, 4, 1)
        # conditioning shrinks the predictive variance at the fantasised inputs
        v_before = model.posterior(Xn).variance
        v_after = fm.posterior(Xn).variance[0, 0]
        assert bool((v_after <= v_before + 1e-12 * TOLX).all())
        # default fantasy noise = mean of the likelihood noise (:43-47)
        gf = model.get_fantasy_model(Xn, _randn(2, generator=gen))
        assert gf.num_data == model.num_data + 2
        with pytest.raises(RuntimeError):
            model.condition_on_observations(Xn, _randn(6, 2, 1), inplace=True)


def test_fast_pred_samples_root_is_exact():
    from online_gp_b200 import settings as S
    model, X, Y, gen = _model(t=1, learn=True)
    Xs = _rand(5, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.eval()
        exact = model(Xs).covariance_matrix
        with S.fast_pred_samples(True):
            dist = model(Xs)
        root = dist.lazy_covariance_matrix.root.evaluate()
        assert root.shape == (5, 5)
        assert _close(root @ root.t(), exact, 1e-8, 1e-10)
        assert torch.allclose(dist.mean, model(Xs).mean)


@pytest.mark.parametrize("t,learn", [(1, True), (1, False)])
def test_predictive_space_fantasies_for_batched_candidates(t, learn):
    """X [b, q, d] with different candidates per batch element: the predictive-space fantasy (exact Gaussian
    conditional) equals explicitly conditioning the WISKI caches per element when the root is exact (Cholesky
    regime) — means per draw and the look-ahead variances qNIPV integrates."""
    model, X, Y, gen = _model(t=t, learn=learn, g=6)            # m = 36 <= max_cholesky_size: r = m
    b, q, nf, d = 3, 2, 4, X.shape[-1]
    Xc = _rand(b, q, d, generator=gen)
    Yf = _randn(nf, b, q, t, generator=gen)
    noise = 0.3 * _ones(b, q, t)
    Xs = _rand(7, d, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.eval()
        fm = model.condition_on_observations(Xc, Yf, noise)
        out = fm(Xs)
        assert out.mean.shape == (nf, b, 7) and out.variance.shape == (nf, b, 7)
        for i in range(b):
            for f in range(nf):
                one = model.condition_on_observations(Xc[i], Yf[f, i], noise[i], inplace=False)
                one.eval()
                ref = one(Xs)
                assert _close(out.mean[f, i], ref.mean.reshape(-1), 1e-7, 1e-9), (i, f)
                assert _close(out.variance[f, i], ref.variance.reshape(-1), 1e-6, 1e-9), (i, f)
        # variance-only form (no targets): what qNegIntegratedPosteriorVariance needs
        vo = model.condition_on_observations(Xc, None, noise)(Xs)
        assert _close(vo.variance, out.variance[0], 1e-12)
        # differentiable w.r.t. the candidates (acquisition optimisation)
        Xg = Xc.clone().requires_grad_(True)
        val = model.condition_on_observations(Xg, None, noise)(Xs).variance.sum()
        val.backward()
        assert Xg.grad is not None and bool(torch.isfinite(Xg.grad).all()) and float(Xg.grad.abs().sum()) > 0


def test_fantasize_with_batched_candidates_end_to_end():
    """``fantasize(X [b, q, d], sampler)`` -> look-ahead posterior at MC points, shapes as BoTorch's qNIPV reads them."""
    from online_gp_b200.models.fantasy import PredictiveSpaceFantasy
    model, X, Y, gen = _model(t=1, learn=True, g=6)
    Xc = _rand(4, 3, 2, generator=gen)
    mc = _rand(11, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fm = model.fantasize(Xc, lambda post: post.rsample(torch.Size([5])))
        assert isinstance(fm, PredictiveSpaceFantasy) and fm.num_fantasies == 5
        post = fm.posterior(mc)
        assert post.mean.shape == (5, 4, 11, 1) and post.variance.shape == (5, 4, 11, 1)
        v0 = model.posterior(mc).variance.reshape(-1)
        assert bool((post.variance[0, :, :, 0] <= v0 + 1e-10 * TOLX).all())          # conditioning never increases the variance


def test_bayesopt_loop_plumbing_ackley3d():
    """BASELINE config 4 in miniature (experiments/bayesopt/bayesopt.py:65-101,176-230): Ackley-3D on the unit cube,
    WISKI with a 10^3 grid, Matern-5/2 product kernel with Gamma priors and Interval constraints, UCB over random
    candidates, q = 3, the model re-hydrated from the previous kernel cache every step and conditioned out of place."""
    import math
    from online_gp_b200 import settings as S
    from online_gp_b200.kernels import GammaPrior, Interval, MaternKernel, ScaleKernel
    from online_gp_b200.mlls import BatchedWoodburyMarginalLogLikelihood
    from online_gp_b200.models import OnlineSKIBotorchModel

    def ackley(u):                       # negated Ackley on [-32.768, 32.768]^3, inputs in the unit cube
        x = (u * 2 - 1) * 32.768
        a, b, c = 20.0, 0.2, 2 * math.pi
        val = -a * torch.exp(-b * x.pow(2).mean(-1).sqrt()) - torch.exp(torch.cos(c * x).mean(-1)) + a + math.e
        return -val

    gen = torch.Generator().manual_seed(0)
    torch.manual_seed(0)
    d, q = 3, 3
    X = _rand(10, d, generator=gen)
    raw = ackley(X)
    mu, sd = raw.mean(), raw.std()
    Y = ((raw - mu) / sd).unsqueeze(-1)
    noise = (4.0 / float(sd)) ** 2 * 1e-4 * torch.ones_like(Y)
    bounds = _tensor([[0.0, 1.0]] * d)
    model, best0 = None, float(Y.max())
    with warnings.catch_warnings(), S.cholesky_jitter(1e-3), S.max_cholesky_size(2048):
        warnings.simplefilter("ignore")
        for step in range(4):
            if model is None:
                covar = ScaleKernel(MaternKernel(nu=2.5, lengthscale_prior=GammaPrior(3.0, 6.0),
                                                 lengthscale_constraint=Interval(1e-4, 12.0)),
                                    outputscale_prior=GammaPrior(2.0, 0.15), outputscale_constraint=Interval(1e-4, 12.0))
                cache = None
            else:
                covar, cache = model.covar_module, model._kernel_cache
            model = OnlineSKIBotorchModel(X, Y, train_noise_term=noise, grid_bounds=bounds, grid_size=10,
                                          learn_additional_noise=True, kernel_cache=cache, covar_module=covar).to(X.device)
            mll = BatchedWoodburyMarginalLogLikelihood(model.likelihood, model, clear_caches_every_iteration=True)
            opt = torch.optim.Adam(model.parameters(), lr=0.05)
            model.train()
            for _ in range(5):                                  # stands in for fit_gpytorch_model (L-BFGS in BoTorch)
                opt.zero_grad()
                loss = -mll(model(X), Y).sum()
                loss.backward()
                opt.step()
                assert bool(torch.isfinite(loss))
            model.zero_grad()
            cand = _rand(64, q, d, generator=gen)           # 64 candidate sets of q points (batched posterior)
            post = model.posterior(cand)
            ucb = (post.mean + math.sqrt(2.0) * post.variance.clamp_min(0).sqrt()).squeeze(-1).max(-1)[0]
            new_x = cand[ucb.argmax()]
            new_y = ((ackley(new_x) - mu) / sd).unsqueeze(-1)
            new_noise = noise[:q]
            X, Y, noise = torch.cat([X, new_x]), torch.cat([Y, new_y]), torch.cat([noise, new_noise])
            model = model.condition_on_observations(X=new_x, Y=new_y, noise=new_noise)
            assert model.num_data == X.shape[0]
        assert model._kernel_cache["interpolation_cache"].shape[-2] == 1000
        assert float(Y.max()) >= best0
        # the streamed model still agrees with one built from scratch on all the data (same hyper-parameters)
        scratch = OnlineSKIBotorchModel(X, Y, train_noise_term=noise, grid_bounds=bounds, grid_size=10,
                                        learn_additional_noise=True, covar_module=model.covar_module).to(X.device)
        scratch.likelihood = model.likelihood
        Xs = _rand(9, d, generator=gen)
        a, b = model.posterior(Xs), scratch.posterior(Xs)
        assert _close(a.mean.reshape(-1), b.mean.reshape(-1), 1e-5, 1e-6)
        assert _close(a.variance.reshape(-1), b.variance.reshape(-1), 1e-4, 1e-7)


def test_qnipv_active_learning_plumbing():
    """BASELINE config 5 in miniature (experiments/active_learning/qnIPV_experiment.py:85-105,137-212): Matern-1/2,
    heteroskedastic fixed noise, no learnable noise; candidate sets of q = 6 scored by the negative integrated posterior
    variance over MC points through variance-only predictive-space fantasies; the chosen set is conditioned on and the
    integrated variance really drops to the look-ahead value."""
    from online_gp_b200 import settings as S
    from online_gp_b200.kernels import MaternKernel, ScaleKernel
    from online_gp_b200.models import OnlineSKIBotorchModel
    gen = torch.Generator().manual_seed(3)
    d, q = 2, 6
    X = _rand(10, d, generator=gen)
    Y = torch.sin(6 * X[:, :1]) * torch.cos(4 * X[:, 1:])
    D = _rand(10, 1, generator=gen) * 0.09 + 0.01 + 1e-6            # data.py:71
    with warnings.catch_warnings(), S.max_cholesky_size(2048), S.skip_posterior_variances(False):
        warnings.simplefilter("ignore")
        model = OnlineSKIBotorchModel(X, Y, train_noise_term=D, grid_bounds=_tensor([[0.0, 1.0]] * d), grid_size=16,
                                      learn_additional_noise=False, covar_module=ScaleKernel(MaternKernel(nu=0.5)))
        mc = _rand(200, d, generator=gen)
        ipv0 = float(model.posterior(mc).variance.mean())
        cand = _rand(12, q, d, generator=gen)
        cand_noise = 0.05 * _ones(12, q, 1)
        look = model.condition_on_observations(cand, None, cand_noise).posterior(mc).variance.mean(dim=(-2, -1))   # [12]
        assert look.shape == (12,) and bool((look <= ipv0 + 1e-12 * TOLX).all())
        best = int(look.argmin())
        new_y = torch.sin(6 * cand[best][:, :1]) * torch.cos(4 * cand[best][:, 1:])
        model2 = model.condition_on_observations(X=cand[best], Y=new_y, noise=cand_noise[best])
        ipv1 = float(model2.posterior(mc).variance.mean())
        assert abs(ipv1 - float(look[best])) <= 1e-8 * TOLX * max(1.0, ipv0)        # look-ahead value == realised value
        assert ipv1 < ipv0


def test_get_fantasy_model_with_batched_inputs():
    """``get_fantasy_model(inputs [b, q, d], targets [b, q], noise)`` (batched_fixed_noise_online_gp.py:287-332) through the
    predictive-space fantasy: per-element posterior == explicit conditioning (exact root)."""
    model, X, Y, gen = _model(t=1, learn=True, g=6)
    b, q = 3, 2
    Xc = _rand(b, q, 2, generator=gen)
    Yc = _randn(b, q, generator=gen)
    noise = 0.2 * _ones(b, q)
    Xs = _rand(5, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.eval()
        fm = super(type(model), model).get_fantasy_model(Xc, Yc, noise)
        out = fm(Xs)
        assert out.mean.shape == (1, b, 5)
        for i in range(b):
            one = model.condition_on_observations(Xc[i], Yc[i].unsqueeze(-1), noise[i].unsqueeze(-1), inplace=False)
            one.eval()
            ref = one(Xs)
            assert _close(out.mean[0, i], ref.mean.reshape(-1), 1e-7, 1e-9)
            assert _close(out.variance[0, i], ref.variance.reshape(-1), 1e-6, 1e-9)
        with pytest.raises(RuntimeError, match="Unsupported batch shapes"):
            model.get_fantasy_model(_rand(2, 3, 2, 2), _rand(2, 3, 2), _ones(2, 3, 2))
        with pytest.raises(RuntimeError, match="Unsupported batch shapes"):
            super(type(model), model).get_fantasy_model(Xc, _rand(2, 2, b, q), noise)


# ------------------------------------------------------------------ §8f-2 / f-3: fit() and the trainable stem
def _stem_problem(n=30, d_in=3, seed=2):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    X = _rand(n + 6, d_in, generator=gen) * 2 - 1
    Y = torch.sin(2 * X.sum(-1, keepdim=True)) + 0.05 * _randn(n + 6, 1, generator=gen)
    return X, Y, gen


@pytest.mark.parametrize("mcs", [2048, 0])
def test_fit_differentiates_the_caches_wrt_the_stem(mcs):
    """``fit()`` trains stem and GP jointly (online_ski_regression.py:80-111): the caches built by ``set_train_data``
    carry the graph back to the stem features (the reference differentiates its dense W^T), in the Cholesky regime and
    beyond it.  In the Cholesky regime (exact root) the gradient of the Woodbury MLL w.r.t. the features equals that of
    the dense exact GP on the SKI kernel (oracle/exact_gp.py), which shares no code with the model."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import LinearStem
    from oracle.exact_gp import DenseExactSKIGP
    from oracle.gridkernel import Hypers
    from oracle.interp import create_grid
    X, Y, gen = _stem_problem()
    n = 30
    with warnings.catch_warnings(), S.max_cholesky_size(mcs), S.max_root_decomposition_size(32):
        warnings.simplefilter("ignore")
        stem = LinearStem(3, 2).double().to(DEV)
        reg = OnlineSKIRegression(stem, X[:n], Y[:n], lr=1e-2, grid_size=8, grid_bound=1.0)
    with warnings.catch_warnings(), S.max_cholesky_size(max(mcs, 800)), S.max_root_decomposition_size(32), \
            S.skip_logdet_forward(False):
        warnings.simplefilter("ignore")
        reg.train()
        feats = reg.stem(X[:n])
        assert feats.requires_grad
        reg.set_train_data(feats, Y[:n])
        reg.gp.zero_grad()
        loss = -reg.mll(reg.gp(feats), Y[:n]).sum()
        (gfeat,) = torch.autograd.grad(loss, feats, retain_graph=True)
        loss.backward()
        for prm in reg.stem.parameters():
            assert prm.grad is not None and bool(torch.isfinite(prm.grad).all()) and float(prm.grad.abs().sum()) > 0
        if mcs > 0:
            fc = feats.detach().cpu().clone().requires_grad_(True)
            grid = create_grid([8, 8], [(-1.1, 1.1)] * 2)
            hyp = Hypers(2, learn_noise=True)
            ref = -DenseExactSKIGP(grid, hyp, fc, Y[:n, 0].cpu(), torch.ones(n)).mll()
            (gref,) = torch.autograd.grad(ref, fc)
            assert abs(float(loss) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))
            assert torch.allclose(gfeat.cpu(), gref, rtol=1e-4, atol=1e-6 * float(gref.abs().max()))
        # and fit() itself moves the stem
        w0 = reg.stem[0].weight.detach().clone()
        rec = reg.fit(X[:n], Y[:n], 3)
        assert len(rec) == 3 and all(r["train_loss"] == r["train_loss"] for r in rec)
        assert not torch.equal(w0, reg.stem[0].weight.detach())
        assert not reg.gp._kernel_cache["interpolation_cache"].requires_grad      # final refresh is detached (:108-110)


def test_trainable_stem_streaming_update():
    """``update()`` with a trainable stem (config/regression.yaml:5 ``stem: linear``): the stem step on the
    Sherman-Morrison partial MLL (online_ski_regression.py:148-162), then the GP step and the conditioning."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIRegression
    from online_gp_b200.models.stems import LinearStem
    X, Y, gen = _stem_problem()
    n = 30
    with warnings.catch_warnings(), S.max_cholesky_size(2048), S.max_root_decomposition_size(64):
        warnings.simplefilter("ignore")
        reg = OnlineSKIRegression(LinearStem(3, 2).double().to(DEV), X[:n], Y[:n], lr=1e-2, grid_size=8, grid_bound=1.0)
        reg.fit(X[:n], Y[:n], 2)
        reg.set_lr(1e-3, 1e-3, bn_mom=0.05)
        w0 = reg.stem[0].weight.detach().clone()
        for t in range(n, n + 6):
            with S.detach_interp_coeff(True):
                rmse, nll = reg.evaluate(X[t:t + 1], Y[t:t + 1])
            stem_loss, gp_loss = reg.update(X[t:t + 1], Y[t:t + 1], update_stem=True)
            assert all(v == v and abs(v) < 1e6 for v in (rmse, nll, stem_loss, gp_loss))
        assert reg.gp.num_data == n + 6
        assert not torch.equal(w0, reg.stem[0].weight.detach())
