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
