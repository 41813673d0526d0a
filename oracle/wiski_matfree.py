"""ORACLE (test infrastructure, CPU torch).  Matrix-free restatement of the WISKI model core (m x r panels).

Same algebra as ``oracle.wiski_ref`` / the reference files it cites, but without the dense m x m
``W^T D^-1 W`` that the reference stores (``online_gp/lazy/updated_root_lazy_tensor.py:42,58``), so it
runs at BASELINE.json grid sizes (m ~ 1e6).  State = root panel ``L`` and inverse-root panel ``B`` (m x r,
``B^T L = I_r``), ``interpolation_cache`` (m), scalars.  Used (i) as the parity checker for the CUDA path beyond
the Cholesky regime and (ii) as the CPU baseline / ``--impl reference`` arm in ``bench.py``.

Rules that the reference leaves to GPyTorch's non-deterministic Lanczos (random probe,
``root_decomposition`` A.5) and that are fixed deterministically here — the product implements the *same* rules
independently (DESIGN.md "Initial root"):
  * m <= max_cholesky_size: ``L = chol(A + jitter)``, ``B = L^-T`` (identical to ``wiski_ref``).
  * otherwise: with V1 = W1^T D1^-1/2 for the first n1 = min(n0, max_root) initial points, G = V1^T V1 = U diag(lam) U^T,
    keep lam_j > tol * lam_max:  L = V1 U,  B = V1 U diag(1/lam)  (exact root of A1 on its own span); the remaining
    initial points are folded in with the reference's projected rank-q update
    (``updated_root_lazy_tensor.py:69-119``; lossy when r < m) in chunks of ``chunk`` points.
  * update: ``collect_vector`` literally — p = B^T v, full SVD, L <- L U S~, B <- B U S~^-1 (two m x r x r GEMMs).
"""
import math
import torch

from .interp import interpolate, left_interp
from .gridkernel import kuu_columns, toeplitz_dense, toeplitz_matmul_fft
from .wiski_ref import psd_safe_cholesky

#: how K X is evaluated per grid axis: "dense" (Toeplitz matrix GEMM), "fft" (GPyTorch's ToeplitzLazyTensor product, A.4)
#: or "auto" (bench.py: whichever is faster on this host for the axis size, decided once by ``pick_kron_mode``)
KRON_MODE = {"mode": "dense", "per_size": {}}


def pick_kron_mode(sizes, ncol, dtype, reps=2):
    """Time the dense and the FFT form of one axis product on a slice of the workload's shape and remember the faster
    one per axis size (so that the CPU baseline is the best the reference's algorithm does on this host)."""
    import time
    for g in sorted(set(int(s) for s in sizes)):
        if g in KRON_MODE["per_size"]:
            continue
        inner = max(ncol, (1 << 25) // g)                # ~128 MB in fp32: far beyond the caches, like the real panels
        X = torch.randn(g, inner, dtype=dtype)
        col = torch.exp(-0.01 * torch.arange(g, dtype=dtype) ** 2)
        best = {}
        for mode in ("dense", "fft"):
            (torch.tensordot(toeplitz_dense(col), X, dims=([1], [0])) if mode == "dense" else toeplitz_matmul_fft(col, X, 0))
            t0 = time.perf_counter()
            for _ in range(reps):
                if mode == "dense":
                    torch.tensordot(toeplitz_dense(col), X, dims=([1], [0]))
                else:
                    toeplitz_matmul_fft(col, X, 0)
            best[mode] = time.perf_counter() - t0
        KRON_MODE["per_size"][g] = min(best, key=best.get)
    KRON_MODE["mode"] = "auto"
    return dict(KRON_MODE["per_size"])


def kron_mm(cols, X):
    """K X, one product per grid axis (A.4), X (m x c): dense Toeplitz GEMM, or the reference's FFT form."""
    sizes = [c.shape[0] for c in cols]
    ncol = X.shape[-1]
    Y = X.reshape(*sizes, ncol)
    for i, c in enumerate(cols):
        mode = KRON_MODE["mode"]
        if mode == "auto":
            mode = KRON_MODE["per_size"].get(int(c.shape[0]), "dense")
        if mode == "fft":
            Y = toeplitz_matmul_fft(c, Y, i)
        else:
            Y = torch.tensordot(toeplitz_dense(c), Y, dims=([1], [i])).movedim(0, i)
    return Y.reshape(-1, ncol)


def scatter_wt(idx, val, src, m):
    """W^T src : (m x c) from src (N x c)."""
    out = torch.zeros(m, src.shape[-1], dtype=src.dtype)
    out.index_add_(0, idx.reshape(-1), (val.unsqueeze(-1) * src.unsqueeze(1)).reshape(-1, src.shape[-1]))
    return out


class WiskiMatFree:
    def __init__(self, grid, hyp, X, y, noise_diag, max_cholesky_size=2048, max_root=512, chunk=64,
                 dtype=torch.float64, update_mode="svd", fold="sequential", root_tol=None):
        self.grid, self.hyp, self.dtype = grid, hyp, dtype
        self.sizes = [len(g) for g in grid]
        self.m = 1
        for g in self.sizes:
            self.m *= g
        self.update_mode = update_mode
        y, noise_diag = y.to(dtype), noise_diag.to(dtype)
        idx, val = self._interp(X)
        self.response_cache = (y * y / noise_diag).sum()
        self.interpolation_cache = scatter_wt(idx, val, (y / noise_diag).unsqueeze(-1), self.m).squeeze(-1)
        self.D_logdet = noise_diag.log().sum()
        self.num_data = X.shape[0]
        n0 = X.shape[0]
        if self.m <= max_cholesky_size:
            Vt = scatter_wt(idx, val / noise_diag.sqrt().unsqueeze(-1), torch.eye(n0, dtype=dtype), self.m)
            Lc = psd_safe_cholesky(Vt @ Vt.t())
            self.L = Lc
            self.B = torch.linalg.solve_triangular(Lc, torch.eye(self.m, dtype=dtype), upper=False).t().contiguous()
        else:
            n1 = min(n0, max_root)
            V1 = scatter_wt(idx[:n1], val[:n1] / noise_diag[:n1].sqrt().unsqueeze(-1), torch.eye(n1, dtype=dtype), self.m)
            lam, U = torch.linalg.eigh(V1.t() @ V1)
            tol = root_tol if root_tol is not None else (1e-10 if dtype == torch.float64 else 1e-5)
            keep = lam > tol * lam.max()
            lam, U = lam[keep].flip(0), U[:, keep].flip(1)
            self.L = V1 @ U
            self.B = self.L / lam
            if fold == "batched" and n0 > n1:
                # the projected updates telescope (L_k = L_0 G_k, p_k = G_k^-1 B_0^T v_k):  L_n L_n^T = L_0 (I + P P^T) L_0^T
                # with P = B_0^T [v_{n1} .. v_{n0}] — one r x r factorisation and one panel GEMM per panel instead of
                # (n0 - n1) / chunk passes.  Same L L^T / B B^T as the sequential rule below (tests compare the two);
                # used by bench.py's CPU legs so that an initial set of 2e4 points does not take hours on host cores.
                r = self.L.shape[1]
                M = torch.eye(r, dtype=torch.float64)
                for s in range(n1, n0, 4096):
                    e = min(s + 4096, n0)
                    Pc = left_interp(idx[s:e], val[s:e] / noise_diag[s:e].clamp_min(1e-7).sqrt().unsqueeze(-1), self.B).double()
                    M = M + Pc.t() @ Pc
                F = torch.linalg.cholesky(M)
                Finv_t = torch.linalg.solve_triangular(F, torch.eye(r, dtype=torch.float64), upper=False).t()
                self.L = self.L @ F.to(dtype)
                self.B = self.B @ Finv_t.to(dtype)
            else:
                for s in range(n1, n0, chunk):
                    e = min(s + chunk, n0)
                    self._root_update(idx[s:e], val[s:e] / noise_diag[s:e].clamp_min(1e-7).sqrt().unsqueeze(-1))

    def _interp(self, X):
        idx, val = interpolate(self.grid, X)
        return idx, val.to(self.dtype)

    def _root_update(self, idx, vval):
        """collect_vector (updated_root_lazy_tensor.py:69-119) with sparse v = W^T D^-1/2 given as (idx, vval)."""
        p = left_interp(idx, vval, self.B).t()             # B^T v  (r x q)
        if self.update_mode == "svd":
            U, S, _ = torch.linalg.svd(p, full_matrices=True)
            pad = torch.ones(U.shape[-2] - S.shape[-1], dtype=S.dtype)
            rs = (S ** 2 + 1.0) ** 0.5
            self.L = self.L @ (U * torch.cat([rs, pad]))
            self.B = self.B @ (U * torch.cat([1.0 / rs, pad]))
        else:   # symmetric square-root form of the same update: (I + p p^T)^(+-1/2) = I + P f(S) P^T
            Pq, S, _ = torch.linalg.svd(p, full_matrices=False)
            rs = (S ** 2 + 1.0) ** 0.5
            self.L = self.L + (self.L @ Pq) * (rs - 1.0) @ Pq.t()
            self.B = self.B + (self.B @ Pq) * (1.0 / rs - 1.0) @ Pq.t()

    def condition_on_observations(self, X, y, noise_diag):
        y, noise_diag = y.to(self.dtype), noise_diag.to(self.dtype)
        idx, val = self._interp(X)
        self.response_cache = self.response_cache + (y * y / noise_diag).sum()
        self.interpolation_cache = self.interpolation_cache + scatter_wt(
            idx, val, (y / noise_diag).unsqueeze(-1), self.m).squeeze(-1)
        self.D_logdet = self.D_logdet + noise_diag.log().sum()
        self._root_update(idx, val / noise_diag.clamp_min(1e-7).sqrt().unsqueeze(-1))
        self.num_data += X.shape[0]

    def cols(self):
        cols = [c.to(self.dtype) for c in kuu_columns(self.grid, self.hyp)]
        if self.hyp.learn_noise:
            cols[0] = cols[0] / self.hyp.noise.to(self.dtype)
        return cols

    def pieces(self):
        cols = self.cols()
        KL = kron_mm(cols, self.L)
        Q = self.L.t() @ KL + torch.eye(self.L.shape[-1], dtype=self.dtype)
        Kb = kron_mm(cols, self.interpolation_cache.unsqueeze(-1)).squeeze(-1)
        c = self.L.t() @ Kb
        return cols, KL, Q, Kb, c

    def predict(self, Xs, pieces=None):
        """Posterior mean (q,) and marginal variance (q,) (latent; * sigma^2 if learnable) — eval forward."""
        cols, KL, Q, Kb, c = self.pieces() if pieces is None else pieces
        idx, val = self._interp(Xs)
        Lq = torch.linalg.cholesky(Q)
        a = torch.cholesky_solve(c.unsqueeze(-1), Lq).squeeze(-1)
        mu_u = Kb - KL @ a
        mean = left_interp(idx, val, mu_u.unsqueeze(-1)).squeeze(-1)
        q = Xs.shape[0]
        Wst = scatter_wt(idx, val, torch.eye(q, dtype=self.dtype), self.m)
        KWs = kron_mm(cols, Wst)
        t = left_interp(idx, val, KL).t()                   # (KL)^T W*^T : r x q  (row-gather of KL)
        cov = Wst.t() @ KWs - t.t() @ torch.cholesky_solve(t, Lq)
        if self.hyp.learn_noise:
            cov = cov * self.hyp.noise.to(self.dtype)
        return mean, cov

    def mll(self, pieces=None, skip_logdet_forward=False):
        """``skip_logdet_forward``: GPyTorch's setting of the same name, which the reference switches on around the
        streaming hyper-parameter step (online_ski_regression.py:137): the *value* of log|Q| is dropped from the loss,
        its gradient is kept."""
        cols, KL, Q, Kb, c = self.pieces() if pieces is None else pieces
        Lq = torch.linalg.cholesky(Q)
        inner_qform = c @ torch.cholesky_solve(c.unsqueeze(-1), Lq).squeeze(-1)
        inv_quad = self.response_cache - self.interpolation_cache @ Kb + inner_qform
        logdet_q = 2 * Lq.diagonal().log().sum()
        if skip_logdet_forward:
            logdet_q = logdet_q - logdet_q.detach()
        logdet = logdet_q + self.D_logdet
        n = self.num_data
        final = n * math.log(2 * math.pi)
        if self.hyp.learn_noise:
            noise = self.hyp.noise.to(self.dtype)
            inv_quad = inv_quad / noise
            final = n * noise.log() + final
        return -0.5 * (inv_quad + logdet + final) / n
