"""``OnlineSKIBotorchModel`` — the BoTorch-facing face of the WISKI GP (``online_gp/models/online_ski_botorch_model.py:11-68``).

BoTorch itself is not a dependency (it cannot be installed in the build environment): the class keeps the
reference's methods and argument meaning (``forward`` squeezing a leading singleton batch, ``posterior``,
``get_fantasy_model`` with the mean-noise default, ``fantasize(X, sampler)``, ``_is_custom_likelihood``) and returns a
small ``GPyTorchPosterior`` stand-in with the members acquisition functions read (``mvn``, ``mean``, ``variance``,
``rsample``).  With BoTorch installed the class can additionally be mixed with ``botorch.models.gpytorch.GPyTorchModel``.
"""
import torch

from .batched_fixed_noise_online_gp import FixedNoiseOnlineSKIGP
from .fantasy import PredictiveSpaceFantasy, condition_on_target_batch


class GPyTorchPosterior:
    """``botorch.posteriors.GPyTorchPosterior`` for a (batched) single-task MVN: trailing output dimension of 1."""

    def __init__(self, mvn):
        self.mvn = mvn

    @property
    def mean(self):
        return self.mvn.mean.unsqueeze(-1)

    @property
    def variance(self):
        return self.mvn.variance.unsqueeze(-1)

    @property
    def device(self):
        return self.mvn.mean.device

    @property
    def dtype(self):
        return self.mvn.mean.dtype

    @property
    def event_shape(self):
        return torch.Size((*self.mvn.mean.shape, 1))

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        if base_samples is not None:
            base_samples = base_samples.squeeze(-1)
        return self.mvn.rsample(sample_shape, base_samples=base_samples).unsqueeze(-1)

    sample = rsample


class OnlineSKIBotorchModel(FixedNoiseOnlineSKIGP):
    def __init__(self, train_inputs=None, train_targets=None, train_noise_term=None, covar_module=None,
                 kernel_cache=None, grid_bounds=None, grid_size=30, learn_additional_noise=False, **kwargs):
        super().__init__(train_inputs=train_inputs, train_targets=train_targets, train_noise_term=train_noise_term,
                         covar_module=covar_module, kernel_cache=kernel_cache, grid_bounds=grid_bounds,
                         grid_size=grid_size, learn_additional_noise=learn_additional_noise, **kwargs)
        self._is_custom_likelihood = True

    def forward(self, X, **kwargs):
        if X is not None:
            if X.shape[0] == 1 and X.dim() > 2:
                X = X[0]
        return super().forward(X, **kwargs)

    def condition_on_observations(self, X, Y, noise=None, inplace=False):
        # candidate batches that differ per element (X [b, q, d]): predictive-space fantasies, no m-sized copies
        if X.dim() == 3:
            if inplace:
                raise RuntimeError("a batch of candidate sets cannot be conditioned on in place")
            if noise is not None and noise.dim() == 4:
                noise = noise[0]
            return PredictiveSpaceFantasy(self, X, Y, noise)
        # a batch of target draws at the same inputs (what ``fantasize`` produces): shared panels, batched caches
        if Y.dim() == X.dim() + 1 and X.dim() == 2:
            if inplace:
                raise RuntimeError("a batch of target draws cannot be conditioned on in place")
            return condition_on_target_batch(self, X, Y, noise)
        return super().condition_on_observations(X, Y, noise=noise, inplace=inplace)

    def get_fantasy_model(self, inputs, targets, noise=None, **kwargs):
        if noise is None:
            noise = torch.ones_like(targets)
            noise = noise * self.likelihood.noise.mean().detach()
        return super().get_fantasy_model(inputs, targets, noise)

    def fantasize(self, X, sampler, observation_noise=True, **kwargs):
        kwargs.pop("propagate_grads", False)
        post_X = self.posterior(X, observation_noise=observation_noise, **kwargs)
        Y_fantasized = sampler(post_X)  # num_fantasies x batch_shape x n' x m
        # Use the mean of the previous noise values (as the reference does, :58-60)
        noise_shape = Y_fantasized.shape[1:]
        noise = self.likelihood.noise.mean().detach().expand(noise_shape)
        return self.condition_on_observations(X=X, Y=Y_fantasized, noise=noise)

    def posterior(self, X, observation_noise=False, **kwargs):
        self.eval()
        X = X.to(self._dtype)
        mvn = self(X)
        return GPyTorchPosterior(mvn)
