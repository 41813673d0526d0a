"""ORACLE (test infrastructure, CPU torch fp64).  Dense exact GP on the SKI kernel — the independent check.

This is the quantity the reference's only live numerical test compares WISKI against
(``tests/mlls/test_batched_woodbury_marginal_log_likelihood.py:45-73``: ``SingleTaskGP`` +
``FixedNoiseGaussianLikelihood`` + ``deepcopy(model.covar_module)``, ``ZeroMean``): by the Woodbury / Sylvester
identities the WISKI marginal log-likelihood, posterior mean and covariance equal those of
``y ~ N(0, W K W^T + sigma^2 D)`` exactly whenever the root of ``W^T D^-1 W`` is exact (r = m), SURVEY.md §8c.
It shares only the interpolation and Toeplitz-column restatements with the WISKI oracles; all GP algebra is a
plain dense Cholesky.
"""
import math
import torch

from .interp import interpolate, dense_wt
from .gridkernel import kuu_columns, kron_dense


class DenseExactSKIGP:
    def __init__(self, grid, hyp, X, y, noise_diag):
        """X (n x d), y (n,), noise_diag (n,) fixed per-point noise D; hyp.learn_noise multiplies D by sigma^2."""
        self.grid, self.hyp = grid, hyp
        self.m = 1
        for g in grid:
            self.m *= len(g)
        self.X, self.y, self.D = X, y, noise_diag

    def _W(self, X):
        idx, val = interpolate(self.grid, X)
        return dense_wt(idx, val.to(self.y.dtype), self.m).t()

    def _K(self):
        return kron_dense(kuu_columns(self.grid, self.hyp))

    def _noise(self):
        return self.D * self.hyp.noise if self.hyp.learn_noise else self.D

    def mll(self):
        """Exact MLL divided by n (GPyTorch ``ExactMarginalLogLikelihood`` convention)."""
        W = self._W(self.X)
        S = W @ self._K() @ W.t() + torch.diag(self._noise())
        Lc = torch.linalg.cholesky(S)
        alpha = torch.cholesky_solve(self.y.unsqueeze(-1), Lc).squeeze(-1)
        n = self.y.shape[0]
        val = -0.5 * (self.y @ alpha + 2 * Lc.diagonal().log().sum() + n * math.log(2 * math.pi))
        return val / n

    def posterior(self, Xs):
        """Latent posterior mean (q,) and covariance (q x q) at Xs."""
        W, Ws, K = self._W(self.X), self._W(Xs), self._K()
        S = W @ K @ W.t() + torch.diag(self._noise())
        Lc = torch.linalg.cholesky(S)
        Ksx = Ws @ K @ W.t()
        mean = Ksx @ torch.cholesky_solve(self.y.unsqueeze(-1), Lc).squeeze(-1)
        cov = Ws @ K @ Ws.t() - Ksx @ torch.cholesky_solve(Ksx.t(), Lc)
        return mean, cov
