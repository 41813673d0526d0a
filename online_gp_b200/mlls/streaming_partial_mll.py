"""Sherman-Morrison one-point MLL increment used to train the stem online —
``online_gp/mlls/streaming_partial_mll.py:6-62`` with w(x') kept as stencils (sparse) instead of a dense m-vector
where possible; differentiable w.r.t. the new features through the interpolation values."""
import torch

from .. import ops
from ..lazy.lazy_tensor import _scatter_dense
from ..settings import skip_posterior_variances


def sm_partial_mll(ski_gp, new_x, new_y, num_seen):
    # inner covariance M = (K_uu^-1 + W^T W)^-1, applied through its Woodbury form K_uu - K_uu L Q^-1 L^T K_uu
    with skip_posterior_variances(False):
        M = ski_gp.prediction_cache["pred_cov"].detach()
    W_y = ski_gp._kernel_cache["interpolation_cache"].detach()         # [t,m,1]

    # w := w(x')
    lazy_kernel = ski_gp.covar_module(new_x).evaluate_kernel()
    idx, val = lazy_kernel.left_interp_indices, lazy_kernel.left_interp_values
    m = W_y.shape[-2]
    out = []
    for o in range(W_y.shape[0]):
        one = torch.ones(idx.shape[0], 1, dtype=val.dtype, device=val.device)
        w = _scatter_dense(idx, val, one, m) if val.requires_grad else ops.left_t_interp(idx, val, one, m)   # [m,1]
        y_o = new_y[o] if new_y.dim() == 3 else new_y
        new_W_y = W_y[o] + w * y_o.reshape(1, 1)

        rhs = torch.cat([w, new_W_y], dim=-1)
        solves = M[o].matmul(rhs)                                      # :28-29

        # v := Mw
        v = solves[..., :1]
        sm_divisor = 1 + v.t() @ w                                     # :36

        M_W_y = solves[..., 1:]
        quad_term_1 = new_W_y.t() @ M_W_y                              # :47
        quad_term_3 = (v.t() @ new_W_y) ** 2 / sm_divisor              # :49
        quad_term = quad_term_1 - quad_term_3
        if ski_gp.has_learnable_noise:
            quad_term = quad_term / ski_gp._second_noise(o).detach()   # :54-55

        # matrix determinant lemma: the log-determinant moves by -log(1 + v^T w)
        logdet_term = torch.log(sm_divisor)                            # :59
        out.append((quad_term - logdet_term) / 2)
    partial_mll = torch.stack(out)
    return partial_mll / (num_seen + 1)
