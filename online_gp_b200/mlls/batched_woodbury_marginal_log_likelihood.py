"""Woodbury marginal log-likelihood from the WISKI caches.

Same class and call signature as ``online_gp/mlls/batched_woodbury_marginal_log_likelihood.py:6-52``:
``-1/2 [ (y^T D^-1 y - b^T K b + c^T Q^-1 c) / sigma^2 + log|Q| + log|D| + n log sigma^2 + n log 2 pi ] / n``
with K = K_uu / sigma^2, Q = I + L^T K L, b = W^T D^-1 y, c = L^T K b.  The r x r solve / logdet dispatch
(Cholesky up to ``max_cholesky_size``) lives in ``PanelGramLazyTensor``; gradients w.r.t. the kernel
hyper-parameters flow through the CUDA kernels' autograd Functions (``ops._KronFn``, ``ops._GramFn``).
"""
import math

import torch
from torch import nn


class BatchedWoodburyMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model, clear_caches_every_iteration=False):
        super().__init__()
        self.likelihood = likelihood
        self.model = model
        self.has_learnable_noise = self.likelihood.second_noise_covar is not None
        self.clear_caches_every_iteration = clear_caches_every_iteration

    def named_priors(self):
        yield from self.model.named_priors()

    def forward(self, distro, targets, *args):
        if self.clear_caches_every_iteration:
            self.model.zero_grad()  # for the time being to clear caches

        current_cache = self.model._kernel_cache

        # I + L'KL
        inner_qmat = self.model.current_qmatrix
        inner_qform, inner_logdet = inner_qmat.inv_quad_logdet(
            inv_quad_rhs=self.model.root_space_projection, logdet=True
        )
        inducing_qform = current_cache["interpolation_cache"].transpose(-1, -2).matmul(self.model.Kuu_response)
        inv_quad_term = (current_cache["response_cache"] - inducing_qform).sum((-2, -1)) + inner_qform
        logdet_term = inner_logdet + current_cache["D_logdet"]

        num_data = self.model.num_data
        if getattr(self.model, "_num_data_t", None) is not None:
            num_data = self.model._num_data_t       # device-side counter (CUDA-graph mode): same value, no host constant

        # add in add'l noise
        final_term = num_data * math.log(2 * math.pi)
        if self.has_learnable_noise:
            noise = self.likelihood.second_noise_covar.noise.to(inv_quad_term.dtype).reshape(-1)   # one per output
            inv_quad_term = inv_quad_term / noise
            # should only be an `n` term here when the logdet is calculated b/c
            # \log |\sigma^{-2} Kuu| = \log Kuu - m \log \sigma^{-2}
            # which is computed in the `logdet_term` in the forwards
            final_term = num_data * noise.log() + final_term

        res = -0.5 * (inv_quad_term + logdet_term + final_term)

        for _, prior, closure, _ in self.named_priors():
            res = res + prior.log_prob(closure()).sum()

        return res / num_data
