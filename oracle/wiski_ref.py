"""ORACLE (test infrastructure, CPU torch).  Reference-literal restatement of the WISKI model core (dense).

Line-by-line restatement (no GPyTorch) of
  * ``online_gp/lazy/updated_root_lazy_tensor.py:10-159``      -> ``UpdatedRootRef``
  * ``online_gp/models/batched_fixed_noise_online_gp.py:22-60`` (``_get_wmat_from_kernel``, ``_initialize_caches``),
    ``:155-171`` (``_update_cache_dicts``), ``:204-256`` (eval ``forward``), ``:258-285``
    (``condition_on_observations``), ``:334-404`` (cached properties, ``_make_predictive_covar``) -> ``WiskiRef``
  * ``online_gp/mlls/batched_woodbury_marginal_log_likelihood.py:19-52``  -> ``WiskiRef.mll``
  * ``online_gp/mlls/streaming_partial_mll.py:6-62``                       -> ``WiskiRef.sm_partial_mll``
  * ``online_gp/likelihoods/fnmg_likelihood.py:12-38`` (noise = D * sigma^2)
with the GPyTorch dispatch rules of SURVEY.md Appendix A.5 on the Cholesky path (``psd_safe_cholesky`` with jitter
escalation; ``root_decomposition`` = Cholesky, ``root_inv_decomposition`` = L^-T) — i.e. what the reference runs in
every shipped config.  One output dimension per instance (the reference's t > 1 batch is t independent
copies with separate hyper-parameters; build t instances).  Dense m x m state: only for m up to a few thousand.
"""
import math
import warnings
import torch

from .interp import interpolate, dense_wt, left_interp
from .gridkernel import kuu_columns, kron_dense


def psd_safe_cholesky(A, jitter=None, max_tries=3):
    """A.5: try plain Cholesky, then add jitter * 10**i to the diagonal (i < max_tries)."""
    L, info = torch.linalg.cholesky_ex(A)
    if info.item() == 0:
        return L
    if jitter is None:
        jitter = 1e-6 if A.dtype == torch.float32 else 1e-8
    Ap, prev = A.clone(), 0.0
    for i in range(max_tries):
        jn = jitter * (10 ** i)
        Ap.diagonal().add_(jn - prev)
        prev = jn
        L, info = torch.linalg.cholesky_ex(Ap)
        if info.item() == 0:
            warnings.warn(f"A not p.d., added jitter of {jn:.1e} to the diagonal", RuntimeWarning)
            return L
    raise RuntimeError(f"Matrix not positive definite after repeatedly adding jitter up to {jn:.1e}.")


class UpdatedRootRef:
    """``UpdatedRootLazyTensor`` (updated_root_lazy_tensor.py:10): dense A with an explicitly updated root pair."""

    def __init__(self, tensor, root=None, inv_root=None):
        self.tensor, self.root, self.inv_root = tensor, root, inv_root

    def root_decomposition(self):          # :121-126 -> A.5 Cholesky path
        if self.root is None:
            self.root = psd_safe_cholesky(self.tensor)
        return self.root

    def root_inv_decomposition(self):      # :128-133 -> A.5: L^-T via triangular solve with I
        if self.inv_root is None:
            L = psd_safe_cholesky(self.tensor) if self.root is None else self.root
            if self.root is None:
                self.root = L
            eye = torch.eye(L.shape[-1], dtype=L.dtype)
            self.inv_root = torch.linalg.solve_triangular(L, eye, upper=False).t()
        return self.inv_root

    def collect_vector(self, vector):      # :69-119
        L = self.root_decomposition()
        Bt = self.root_inv_decomposition().t()
        p = Bt @ vector                                    # :79
        U, S, _ = torch.linalg.svd(p, full_matrices=True)  # :82 (some=False)
        pad = torch.ones(U.shape[-2] - S.shape[-1], dtype=S.dtype)
        rs = (S ** 2 + 1.0) ** 0.5
        new_root = L @ (U @ torch.diag(torch.cat([rs, pad])))            # :97-100
        new_inv = Bt.t() @ (U @ torch.diag(torch.cat([1.0 / rs, pad])))  # :111-117
        return new_root, new_inv

    def update(self, vector):              # :53-67
        if vector.dim() == 1:
            vector = vector.view(-1, 1)
        tensor = self.tensor + vector @ vector.t()
        root, inv_root = self.collect_vector(vector)
        return UpdatedRootRef(tensor, root, inv_root)


class WiskiRef:
    """``FixedNoiseOnlineSKIGP`` for one output (batched_fixed_noise_online_gp.py:63), dense restatement."""

    def __init__(self, grid, hyp, X, y, noise_diag):
        self.grid, self.hyp = grid, hyp
        self.m = 1
        for g in grid:
            self.m *= len(g)
        self.dtype = y.dtype
        Wt = self._wt(X)                                   # :141-143
        dinv_y = y / noise_diag                            # :42
        self.response_cache = y @ dinv_y                   # :45
        self.interpolation_cache = Wt @ dinv_y             # :46
        self.WtW = UpdatedRootRef(Wt @ (Wt / noise_diag).t())   # :49-53
        self.D_logdet = noise_diag.log().sum()             # :55
        self.num_data = X.shape[0]

    def _wt(self, X):                                      # :22-28
        idx, val = interpolate(self.grid, X)
        return dense_wt(idx, val.to(self.dtype), self.m)

    def condition_on_observations(self, X, y, noise_diag):  # :258-273 (inplace=True) + :155-171
        Wt = self._wt(X)
        dinv_y = y / noise_diag
        self.response_cache = self.response_cache + y @ dinv_y
        self.interpolation_cache = self.interpolation_cache + Wt @ dinv_y
        self.D_logdet = self.D_logdet + noise_diag.log().sum()
        v = Wt / noise_diag.clamp_min(1e-7) ** 0.5          # :163-168
        self.WtW = self.WtW.update(v)                        # :169
        self.num_data += X.shape[0]

    # ---- cached properties :334-383 (recomputed on demand here)
    def Kuu(self):                                         # :334-341
        K = kron_dense(kuu_columns(self.grid, self.hyp))
        if self.hyp.learn_noise:
            K = K / self.hyp.noise
        return K

    def pieces(self):
        K = self.Kuu()
        L = self.WtW.root_decomposition().detach()
        KL = K @ L                                          # :348
        Q = L.t() @ KL + torch.eye(L.shape[-1], dtype=L.dtype)   # :352-355
        Kb = K @ self.interpolation_cache                   # :366
        c = L.t() @ Kb                                      # :360-361
        return K, L, KL, Q, Kb, c

    def prediction_cache(self):                            # :368-383, :385-404 (fast_pred_var off)
        K, L, KL, Q, Kb, c = self.pieces()
        Lq = torch.linalg.cholesky(Q)
        a = torch.cholesky_solve(c.unsqueeze(-1), Lq).squeeze(-1)
        pred_mean = Kb - KL @ a                             # :376
        pred_cov = K - KL @ torch.cholesky_solve(KL.t(), Lq)    # :399-403
        return pred_mean, pred_cov

    def predict(self, Xs):                                 # eval forward :204-256
        idx, val = interpolate(self.grid, Xs)
        val = val.to(self.dtype)
        pm, pc = self.prediction_cache()
        mean = left_interp(idx, val, pm.unsqueeze(-1)).squeeze(-1)    # :206-210
        Wst = dense_wt(idx, val, self.m)
        cov = Wst.t() @ pc @ Wst                                       # :222-225
        if self.hyp.learn_noise:
            cov = cov * self.hyp.noise                                 # :227-228
        return mean, cov

    def mll(self):                                         # batched_woodbury_marginal_log_likelihood.py:19-52
        K, L, KL, Q, Kb, c = self.pieces()
        Lq = torch.linalg.cholesky(Q)
        inner_qform = c @ torch.cholesky_solve(c.unsqueeze(-1), Lq).squeeze(-1)     # :27-30
        inner_logdet = 2 * Lq.diagonal().log().sum()
        inducing_qform = self.interpolation_cache @ Kb                              # :31
        inv_quad = (self.response_cache - inducing_qform) + inner_qform             # :32
        logdet = inner_logdet + self.D_logdet                                       # :33
        n = self.num_data
        final = n * math.log(2 * math.pi)                                           # :38
        if self.hyp.learn_noise:
            inv_quad = inv_quad / self.hyp.noise                                    # :40
            final = n * self.hyp.noise.log() + final                                # :45
        return -0.5 * (inv_quad + logdet + final) / n                               # :47,52

    def sm_partial_mll(self, new_x, new_y, num_seen):      # streaming_partial_mll.py:6-62
        with torch.no_grad():
            _, M = self.prediction_cache()
        W_y = self.interpolation_cache.detach().unsqueeze(-1)
        idx, val = interpolate(self.grid, new_x)
        w = dense_wt(idx, val.to(self.dtype), self.m)            # m x 1
        new_W_y = W_y + w * new_y
        solves = M @ torch.cat([w, new_W_y], dim=-1)             # :28-29
        v = solves[:, :1]
        sm_div = 1 + v.t() @ w                                    # :36
        quad1 = new_W_y.t() @ solves[:, 1:]                       # :47
        quad3 = (v.t() @ new_W_y) ** 2 / sm_div                   # :49
        quad = quad1 - quad3
        if self.hyp.learn_noise:
            quad = quad / self.hyp.noise.detach()                 # :54-55
        return ((quad - torch.log(sm_div)) / 2 / (num_seen + 1)).squeeze()   # :59-62
