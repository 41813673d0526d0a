"""Fixed per-point noise times a learnable scalar — ``online_gp/likelihoods/fnmg_likelihood.py:12-38`` on top of a
minimal ``FixedNoiseGaussianLikelihood`` (``second_noise_covar = HomoskedasticNoise`` with
``GreaterThan(1e-4)``, raw init 0 => 0.6932)."""
import torch
from torch import nn

from ..kernels import GreaterThan, _PriorMixin


class HomoskedasticNoise(nn.Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size()):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(tuple(batch_shape) + (1,)))
        self.raw_noise_constraint = noise_constraint if noise_constraint is not None else GreaterThan(1e-4)
        self._priors = {}
        if noise_prior is not None:
            self._priors["noise_prior"] = (noise_prior, lambda m: m.noise)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value, dtype=self.raw_noise.dtype, device=self.raw_noise.device)
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(value).expand_as(self.raw_noise))


class FixedNoise(nn.Module):
    def __init__(self, noise):
        super().__init__()
        self.noise = noise


class FixedNoiseGaussianLikelihood(nn.Module, _PriorMixin):
    def __init__(self, noise, learn_additional_noise=False, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = FixedNoise(noise)
        self.second_noise_covar = None
        if learn_additional_noise:
            self.second_noise_covar = HomoskedasticNoise(noise_prior=kwargs.get("noise_prior"),
                                                         noise_constraint=kwargs.get("noise_constraint"),
                                                         batch_shape=batch_shape)

    @property
    def second_noise(self):
        return 0.0 if self.second_noise_covar is None else self.second_noise_covar.noise

    @second_noise.setter
    def second_noise(self, value):
        if self.second_noise_covar is None:
            raise RuntimeError("Attempting to set secondary learned noise for FixedNoiseGaussianLikelihood, "
                               "but learn_additional_noise must have been False!")
        self.second_noise_covar.noise = value

    @property
    def noise(self):
        return self.noise_covar.noise + self.second_noise


class FNMGLikelihood(FixedNoiseGaussianLikelihood):
    """Fixed-Noise w/ multiplicative learnable second noise term Gaussian likelihood (fnmg_likelihood.py:12)."""

    @property
    def noise(self):
        return self.noise_covar.noise * self.second_noise     # :16-18
