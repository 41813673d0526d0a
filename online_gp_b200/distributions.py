"""``MultivariateNormal`` with a lazily evaluated covariance — the slice of ``gpytorch.distributions`` the WISKI
wrappers read (``.mean``, ``.variance``, ``.covariance_matrix``, ``.lazy_covariance_matrix``, ``.stddev``,
``.rsample``; ``online_gp/models/online_ski_regression.py:56-62``)."""
import torch

from .lazy.lazy_tensor import LazyTensor, lazify


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self._mean = mean.evaluate() if isinstance(mean, LazyTensor) else mean
        self._covar = lazify(covariance_matrix)

    @property
    def mean(self):
        return self._mean

    loc = mean

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.evaluate()

    @property
    def variance(self):
        return self._covar.diag()

    @property
    def stddev(self):
        return self.variance.clamp_min(1e-12).sqrt()

    @property
    def event_shape(self):
        return self._mean.shape[-1:]

    @property
    def batch_shape(self):
        return self._mean.shape[:-1]

    def confidence_region(self):
        s = self.stddev * 2
        return self.mean - s, self.mean + s

    def rsample(self, sample_shape=torch.Size(), base_samples=None):
        cov = self.covariance_matrix
        n = cov.shape[-1]
        jitter = 1e-6 if cov.dtype == torch.float32 else 1e-8
        Lc = torch.linalg.cholesky(cov + jitter * torch.eye(n, dtype=cov.dtype, device=cov.device))
        if base_samples is None:
            base_samples = torch.randn(*sample_shape, *self._mean.shape, dtype=cov.dtype, device=cov.device)
        return self._mean + (Lc @ base_samples.unsqueeze(-1)).squeeze(-1)

    sample = rsample
