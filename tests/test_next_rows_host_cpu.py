"""CPU (kernels mocked by the oracle, tests/cpu_ops_mock.py): host logic of the "next" rows of SURVEY §8f-1 — the
BoTorch-facing wrapper, batched fantasies on shared panels, and the exact predictive root under ``fast_pred_samples``.
GPU validation of these rows is pending (the kernels they call are the same ones the -m gpu suite checks)."""
import warnings

import pytest
import torch

import cpu_ops_mock


@pytest.fixture(autouse=True)
def _mock_ops():
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    with cpu_ops_mock.install():
        yield
    torch.set_default_dtype(prev)


def _model(t=1, learn=True, n0=25, d=2, g=8, seed=0):
    from online_gp_b200.models import OnlineSKIBotorchModel
    gen = torch.Generator().manual_seed(seed)
    X = torch.rand(n0, d, generator=gen)
    Y = torch.stack([torch.sin(3 * X.sum(-1) + o) for o in range(t)], dim=-1) + 0.05 * torch.randn(n0, t, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = OnlineSKIBotorchModel(X, Y, 0.1 * torch.ones(n0, t), grid_bounds=torch.tensor([[0.0, 1.0]] * d),
                                      grid_size=g, learn_additional_noise=learn)
    return model, X, Y, gen


@pytest.mark.parametrize("t,learn", [(1, True), (1, False), (2, True)])
def test_fantasy_batch_equals_one_model_per_draw(t, learn):
    """condition_on_observations(X, Y[nf, q, t]) == nf separately conditioned models: means per draw, one shared
    covariance (reference: get_fantasy_model expands every cache per fantasy, batched_fixed_noise_online_gp.py:287-332)."""
    model, X, Y, gen = _model(t=t, learn=learn)
    nf, q, d = 5, 3, X.shape[-1]
    Xn = torch.rand(q, d, generator=gen)
    Yf = torch.randn(nf, q, t, generator=gen)
    noise = 0.2 * torch.ones(q, t)
    Xs = torch.rand(6, d, generator=gen)
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
            assert torch.allclose(mean_f, ref.mean, rtol=1e-9, atol=1e-11)
            assert torch.allclose(dist.variance[f], ref.variance, rtol=1e-9, atol=1e-12)
        assert dist.covariance_matrix.shape[0] == nf
        # the base model is untouched
        assert model.num_data == X.shape[0]


def test_botorch_wrapper_posterior_and_fantasize():
    from online_gp_b200.models import FantasizedOnlineSKIGP, GPyTorchPosterior
    model, X, Y, gen = _model(t=1, learn=True)
    assert model._is_custom_likelihood is True
    Xs = torch.rand(4, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        post = model.posterior(Xs.float())                    # posterior() casts to the model dtype (:66)
        assert isinstance(post, GPyTorchPosterior) and post.mean.shape == (4, 1) and post.variance.shape == (4, 1)
        assert post.rsample(torch.Size([7])).shape == (7, 4, 1)
        same = model.posterior(Xs.unsqueeze(0))               # forward squeezes a leading singleton batch (:36-40)
        assert torch.allclose(same.mean, post.mean)
        batched = model.posterior(torch.rand(3, 4, 2, generator=gen))
        assert batched.mean.shape == (3, 4, 1) and batched.variance.shape == (3, 4, 1)

        def sampler(posterior):
            return posterior.rsample(torch.Size([6]))
        Xn = torch.rand(2, 2, generator=gen)
        fm = model.fantasize(Xn, sampler)
        assert isinstance(fm, FantasizedOnlineSKIGP) and fm.num_fantasies == 6
        fpost = fm.posterior(Xs)
        # a model re-hydrated from a kernel cache keeps the leading output dimension (reference: `_batch_shape` is then
        # an int, batched_fixed_noise_online_gp.py:88,247-250), so fantasies are [nf, t, q*] (+ BoTorch's trailing 1)
        assert fpost.mean.shape == (6, 1, 4, 1) and fpost.variance.shape == (6, 1, 4, 1)
        # conditioning shrinks the predictive variance at the fantasised inputs
        v_before = model.posterior(Xn).variance
        v_after = fm.posterior(Xn).variance[0, 0]
        assert bool((v_after <= v_before + 1e-12).all())
        # default fantasy noise = mean of the likelihood noise (:43-47)
        gf = model.get_fantasy_model(Xn, torch.randn(2, generator=gen))
        assert gf.num_data == model.num_data + 2
        with pytest.raises(RuntimeError):
            model.condition_on_observations(Xn, torch.randn(6, 2, 1), inplace=True)


def test_fast_pred_samples_root_is_exact():
    from online_gp_b200 import settings as S
    model, X, Y, gen = _model(t=1, learn=True)
    Xs = torch.rand(5, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model.eval()
        exact = model(Xs).covariance_matrix
        with S.fast_pred_samples(True):
            dist = model(Xs)
        root = dist.lazy_covariance_matrix.root.evaluate()
        assert root.shape == (5, 5)
        assert torch.allclose(root @ root.t(), exact, rtol=1e-8, atol=1e-10)
        assert torch.allclose(dist.mean, model(Xs).mean)


def _two_blobs(n, gen):
    y = (torch.rand(n, generator=gen) > 0.5).long()
    centers = torch.tensor([[-0.8, -0.5], [0.7, 0.6]])
    x = centers[y] + 0.35 * torch.randn(n, 2, generator=gen)
    return x.clamp(-2.9, 2.9), y


def test_dirichlet_ski_classifier_batch_and_online():
    """tests/classification/test_ski_classifier.py (batch fit with an Identity stem, then streaming updates) on a
    synthetic two-class problem (the Banana data set of the reference is not available offline)."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIClassifier
    from online_gp_b200.models.online_ski_classifier import dirichlet_transform
    from online_gp_b200.models.stems import Identity
    gen = torch.Generator().manual_seed(0)
    y_t, alpha, s2 = dirichlet_transform(torch.tensor([0, 1, 1]), 1e-2)
    assert y_t.shape == (3, 2) and torch.allclose(alpha, torch.tensor([[1.01, 0.01], [0.01, 1.01], [0.01, 1.01]]))
    assert torch.allclose(s2, torch.log(1.0 / alpha + 1.0)) and bool((y_t[:, 0] > y_t[:, 1]).tolist() == [True, False, False])
    train_x, train_y = _two_blobs(120, gen)
    test_x, test_y = _two_blobs(200, gen)
    with warnings.catch_warnings(), S.max_root_decomposition_size(512), S.max_cholesky_size(2048):
        warnings.simplefilter("ignore")
        clf = OnlineSKIClassifier(Identity(2), train_x[:60], train_y[:60], 1e-2, 1e-1, 12, 3.1)
        rec = clf.fit(train_x[:60], train_y[:60], 8)
        assert len(rec) == 8 and rec[-1]["epoch"] == 8
        acc0 = clf.predict(test_x).eq(test_y).float().mean().item()
        assert acc0 >= 0.9
        n_before = clf.gp.num_data
        for t in range(60, 120, 10):
            stem_loss, gp_loss = clf.update(train_x[t:t + 10], train_y[t:t + 10])
            assert stem_loss == 0 and gp_loss == gp_loss          # Identity stem: nothing to train
        assert clf.gp.num_data == n_before + 60
        acc1 = clf.predict(test_x).eq(test_y).float().mean().item()
        assert acc1 >= 0.9


def test_dirichlet_ski_classifier_online_learned_features():
    """Streaming updates with a trainable stem (LinearStem: Linear + BatchNorm + tanh): the stem step goes through the
    Sherman-Morrison partial MLL with two outputs (online_ski_classifier.py:103-117)."""
    from online_gp_b200 import settings as S
    from online_gp_b200.models import OnlineSKIClassifier
    from online_gp_b200.models.stems import LinearStem
    gen = torch.Generator().manual_seed(1)
    torch.manual_seed(1)
    train_x, train_y = _two_blobs(40, gen)
    with warnings.catch_warnings(), S.max_root_decomposition_size(512), S.max_cholesky_size(2048):
        warnings.simplefilter("ignore")
        clf = OnlineSKIClassifier(LinearStem(2, 2).double(), train_x[:8], train_y[:8], 1e-2, 1e-3, 10, 1.0)
        clf.set_lr(1e-3, 1e-3, bn_mom=0.05)
        clf.eval()                          # as after fit(): BatchNorm uses its running statistics for single points
        w0 = clf.stem[0].weight.detach().clone()
        for t in range(8, 40):              # one point per update: the Sherman-Morrison increment is a rank-one formula
            stem_loss, gp_loss = clf.update(train_x[t:t + 1], train_y[t:t + 1])
            assert stem_loss == stem_loss and gp_loss == gp_loss and abs(stem_loss) < 1e6
        assert clf.gp.num_data == 40
        assert not torch.equal(w0, clf.stem[0].weight.detach())           # the stem really trained
        assert clf.predict(train_x).shape == (40,)


@pytest.mark.parametrize("t,learn", [(1, True), (1, False)])
def test_predictive_space_fantasies_for_batched_candidates(t, learn):
    """X [b, q, d] with different candidates per batch element: the predictive-space fantasy (exact Gaussian
    conditional) equals explicitly conditioning the WISKI caches per element when the root is exact (Cholesky
    regime) — means per draw and the look-ahead variances qNIPV integrates."""
    model, X, Y, gen = _model(t=t, learn=learn, g=6)            # m = 36 <= max_cholesky_size: r = m
    b, q, nf, d = 3, 2, 4, X.shape[-1]
    Xc = torch.rand(b, q, d, generator=gen)
    Yf = torch.randn(nf, b, q, t, generator=gen)
    noise = 0.3 * torch.ones(b, q, t)
    Xs = torch.rand(7, d, generator=gen)
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
                assert torch.allclose(out.mean[f, i], ref.mean.reshape(-1), rtol=1e-7, atol=1e-9), (i, f)
                assert torch.allclose(out.variance[f, i], ref.variance.reshape(-1), rtol=1e-6, atol=1e-9), (i, f)
        # variance-only form (no targets): what qNegIntegratedPosteriorVariance needs
        vo = model.condition_on_observations(Xc, None, noise)(Xs)
        assert torch.allclose(vo.variance, out.variance[0], rtol=1e-12)
        # differentiable w.r.t. the candidates (acquisition optimisation)
        Xg = Xc.clone().requires_grad_(True)
        val = model.condition_on_observations(Xg, None, noise)(Xs).variance.sum()
        val.backward()
        assert Xg.grad is not None and bool(torch.isfinite(Xg.grad).all()) and float(Xg.grad.abs().sum()) > 0


def test_fantasize_with_batched_candidates_end_to_end():
    """``fantasize(X [b, q, d], sampler)`` -> look-ahead posterior at MC points, shapes as BoTorch's qNIPV reads them."""
    from online_gp_b200.models.fantasy import PredictiveSpaceFantasy
    model, X, Y, gen = _model(t=1, learn=True, g=6)
    Xc = torch.rand(4, 3, 2, generator=gen)
    mc = torch.rand(11, 2, generator=gen)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fm = model.fantasize(Xc, lambda post: post.rsample(torch.Size([5])))
        assert isinstance(fm, PredictiveSpaceFantasy) and fm.num_fantasies == 5
        post = fm.posterior(mc)
        assert post.mean.shape == (5, 4, 11, 1) and post.variance.shape == (5, 4, 11, 1)
        v0 = model.posterior(mc).variance.reshape(-1)
        assert bool((post.variance[0, :, :, 0] <= v0 + 1e-10).all())          # conditioning never increases the variance


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
    X = torch.rand(10, d, generator=gen)
    raw = ackley(X)
    mu, sd = raw.mean(), raw.std()
    Y = ((raw - mu) / sd).unsqueeze(-1)
    noise = (4.0 / float(sd)) ** 2 * 1e-4 * torch.ones_like(Y)
    bounds = torch.tensor([[0.0, 1.0]] * d)
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
                                          learn_additional_noise=True, kernel_cache=cache, covar_module=covar).to(X)
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
            cand = torch.rand(64, q, d, generator=gen)           # 64 candidate sets of q points (batched posterior)
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
                                        learn_additional_noise=True, covar_module=model.covar_module).to(X)
        scratch.likelihood = model.likelihood
        Xs = torch.rand(9, d, generator=gen)
        a, b = model.posterior(Xs), scratch.posterior(Xs)
        assert torch.allclose(a.mean.reshape(-1), b.mean.reshape(-1), rtol=1e-5, atol=1e-6)
        assert torch.allclose(a.variance.reshape(-1), b.variance.reshape(-1), rtol=1e-4, atol=1e-7)


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
    X = torch.rand(10, d, generator=gen)
    Y = torch.sin(6 * X[:, :1]) * torch.cos(4 * X[:, 1:])
    D = torch.rand(10, 1, generator=gen) * 0.09 + 0.01 + 1e-6            # data.py:71
    with warnings.catch_warnings(), S.max_cholesky_size(2048), S.skip_posterior_variances(False):
        warnings.simplefilter("ignore")
        model = OnlineSKIBotorchModel(X, Y, train_noise_term=D, grid_bounds=torch.tensor([[0.0, 1.0]] * d), grid_size=16,
                                      learn_additional_noise=False, covar_module=ScaleKernel(MaternKernel(nu=0.5)))
        mc = torch.rand(200, d, generator=gen)
        ipv0 = float(model.posterior(mc).variance.mean())
        cand = torch.rand(12, q, d, generator=gen)
        cand_noise = 0.05 * torch.ones(12, q, 1)
        look = model.condition_on_observations(cand, None, cand_noise).posterior(mc).variance.mean(dim=(-2, -1))   # [12]
        assert look.shape == (12,) and bool((look <= ipv0 + 1e-12).all())
        best = int(look.argmin())
        new_y = torch.sin(6 * cand[best][:, :1]) * torch.cos(4 * cand[best][:, 1:])
        model2 = model.condition_on_observations(X=cand[best], Y=new_y, noise=cand_noise[best])
        ipv1 = float(model2.posterior(mc).variance.mean())
        assert abs(ipv1 - float(look[best])) <= 1e-8 * max(1.0, ipv0)        # look-ahead value == realised value
        assert ipv1 < ipv0


def test_get_fantasy_model_with_batched_inputs():
    """``get_fantasy_model(inputs [b, q, d], targets [b, q], noise)`` (batched_fixed_noise_online_gp.py:287-332) through the
    predictive-space fantasy: per-element posterior == explicit conditioning (exact root)."""
    model, X, Y, gen = _model(t=1, learn=True, g=6)
    b, q = 3, 2
    Xc = torch.rand(b, q, 2, generator=gen)
    Yc = torch.randn(b, q, generator=gen)
    noise = 0.2 * torch.ones(b, q)
    Xs = torch.rand(5, 2, generator=gen)
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
            assert torch.allclose(out.mean[0, i], ref.mean.reshape(-1), rtol=1e-7, atol=1e-9)
            assert torch.allclose(out.variance[0, i], ref.variance.reshape(-1), rtol=1e-6, atol=1e-9)
        with pytest.raises(RuntimeError, match="Unsupported batch shapes"):
            model.get_fantasy_model(torch.rand(2, 3, 2, 2), torch.rand(2, 3, 2), torch.ones(2, 3, 2))
        with pytest.raises(RuntimeError, match="Unsupported batch shapes"):
            super(type(model), model).get_fantasy_model(Xc, torch.rand(2, 2, b, q), noise)
