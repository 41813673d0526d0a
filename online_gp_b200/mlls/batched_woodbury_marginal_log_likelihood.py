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
    """``mll(distro, targets)`` -> tensor of shape ``batch_shape`` (one value per GP output), already divided by the
    number of observations.  Both arguments are ignored: everything comes from the model's caches."""

    def __init__(self, likelihood, model, clear_caches_every_iteration=False):
        super().__init__()
        self.likelihood, self.model = likelihood, model
        self.clear_caches_every_iteration = clear_caches_every_iteration
        self.has_learnable_noise = likelihood.second_noise_covar is not None

    def named_priors(self):
        yield from self.model.named_priors()

    def forward(self, distro, targets, *args):
        gp = self.model
        if self.clear_caches_every_iteration:
            gp.zero_grad()                       # the BO loop refits from scratch every step (bayesopt.py:98-100)
        caches = gp._kernel_cache

        # quadratic form  y^T D^-1 y - b^T K b + c^T Q^-1 c  and  log|Q| + log|D|,  Q = I + L^T K L
        quad_q, logdet_q = gp.current_qmatrix.inv_quad_logdet(inv_quad_rhs=gp.root_space_projection, logdet=True)
        b = caches["interpolation_cache"]
        b_K_b = b.transpose(-1, -2).matmul(gp.Kuu_response)
        quad = (caches["response_cache"] - b_K_b).sum((-2, -1)) + quad_q
        logdet = logdet_q + caches["D_logdet"]

        # n: a Python int, or the device-side counter in CUDA-graph mode (same value, no host constant in the graph)
        n = gp.num_data if getattr(gp, "_num_data_t", None) is None else gp._num_data_t
        const = n * math.log(2 * math.pi)
        if self.has_learnable_noise:
            # K is stored as K_uu / sigma^2, so only the n log sigma^2 term and the 1 / sigma^2 on the quadratic form
            # remain here (one noise value per output)
            sigma2 = self.likelihood.second_noise_covar.noise.to(quad.dtype).reshape(-1)
            quad = quad / sigma2
            const = n * sigma2.log() + const

        value = -0.5 * (quad + logdet + const)
        for _, prior, closure, _ in self.named_priors():
            value = value + prior.log_prob(closure()).sum()
        return value / n
