"""``UpdatedRootLazyTensor`` — A = W^T D^-1 W carried as an explicitly updated root / inverse-root panel pair.

Same class name, constructor signature and protocol as ``online_gp/lazy/updated_root_lazy_tensor.py:10-159``
(``update``, ``collect_vector``, ``root_decomposition``, ``root_inv_decomposition``, ``_matmul``, ``evaluate``,
``_expand_batch``), re-designed matrix-free: the dense m x m ``tensor`` the reference stores (``:42,58``) is
optional and only kept in the Cholesky regime (m <= max_cholesky_size); beyond it the operator is the panel pair
``root`` (L, m x r) / ``inv_root`` (B, m x r, B^T L = I) alone and ``_matmul`` is L (L^T x).

Rank-q update (``collect_vector``, ``:69-119``): p = B^T v, then L <- L (I + p p^T)^(1/2)', B <- B (I + p p^T)^(-1/2)'.
  * ``root_update_mode("svd")``: the reference's literal factors U S~ / U S~^-1 from the full SVD of p
    (two m x r x r panel GEMMs);
  * ``root_update_mode("sym")`` (default): the symmetric square roots I + p C p^T, C = (I + (I+G)^(1/2))^-1,
    C' = -(I+G)^(-1/2) C with G = p^T p (q x q) — the same L L^T and B B^T (they differ from the literal factors
    by an orthogonal right factor U^T), applied row-locally in two passes over each panel.
When r < m both drop the component of v outside span(L), exactly like the reference.

A leading batch dimension (one element per GP output) is supported by looping; panels are [t, m, r].
"""
import torch

from .. import ops, settings
from .lazy_tensor import LazyTensor, NonLazyTensor, PanelLazyTensor, RootLazyTensor, psd_safe_cholesky

_DENSE_LIMIT = 16384


def _sym_factors(p):
    """p [r,q] -> (C, C') [q,q] with (I + p p^T)^(1/2) = I + p C p^T and (I + p p^T)^(-1/2) = I + p C' p^T."""
    q = p.shape[1]
    if q == 1:
        lam = (p * p).sum().reshape(1, 1)
        s = torch.sqrt(1.0 + lam)
        return 1.0 / (s + 1.0), -1.0 / (s * (s + 1.0))
    if p.is_cuda and torch.cuda.is_current_stream_capturing():
        return _sym_factors_iterative(p)        # eigh reads its status flag back on the host: not capturable
    G = p.t() @ p
    lam, V = torch.linalg.eigh(G)
    lam = lam.clamp_min(0.0)
    s = torch.sqrt(1.0 + lam)
    C = (V / (s + 1.0)) @ V.t()
    Cp = (V * (-1.0 / (s * (s + 1.0)))) @ V.t()
    return C, Cp


def _sym_factors_iterative(p, iters=18):
    """Same factors without an eigendecomposition (no host read, so it can be captured in a CUDA graph): the
    Denman-Beavers iteration  Y <- (Y + Z^-1) / 2,  Z <- (Z + Y^-1) / 2  on A / ||A||_F, A = I + p^T p, gives
    Y -> A^(1/2), Z -> A^(-1/2) (Higham, Functions of Matrices, §6.3).  Convergence is quadratic only once every
    eigenvalue of the iterate is near its limit; before that the small eigenvalues (1 / ||A||_F after the scaling) halve
    their log-distance per step, so the count must cover ~0.5 log2 ||A||_F + 5 steps: 18 handles ||p||^2 up to ~6e7 (a
    leverage no stream of unit-noise observations reaches), and the count cannot depend on device data under capture."""
    q = p.shape[1]
    eye = torch.eye(q, dtype=p.dtype, device=p.device)
    A = eye + p.t() @ p
    c = torch.linalg.matrix_norm(A)
    Y, Z = A / c, eye
    for _ in range(iters):
        Yi, _ = torch.linalg.inv_ex(Y)
        Zi, _ = torch.linalg.inv_ex(Z)
        Y, Z = 0.5 * (Y + Zi), 0.5 * (Z + Yi)
    rc = c.sqrt()
    C, _ = torch.linalg.inv_ex(eye + Y * rc)
    return C, -((Z / rc) @ C)


class UpdatedRootLazyTensor(LazyTensor):
    def __init__(self, initial_tensor=None, n_shape=None, initial_is_root=True, root=None, inv_root=None):
        r"""
        initial_tensor: initial matrix (or its root if ``initial_is_root``); optional when root/inv_root are given
        n_shape: size of the matrix if building from zero rows
        root / inv_root: panels with  root root^T = A,  inv_root inv_root^T = A^+
        """
        if initial_tensor is None and root is None:
            initial_tensor = torch.zeros(1, n_shape)
        if initial_tensor is not None and initial_is_root:
            initial_tensor = initial_tensor.transpose(-1, -2) @ initial_tensor
        self.tensor = initial_tensor
        self.root = root
        self.inv_root = inv_root

    # ---- LazyTensor protocol
    def _size(self):
        if self.tensor is not None:
            return self.tensor.shape
        m = self.root.shape[-2]
        return torch.Size((*self.root.shape[:-2], m, m))

    @property
    def dtype(self):
        return self.tensor.dtype if self.tensor is not None else self.root.dtype

    @property
    def device(self):
        return self.tensor.device if self.tensor is not None else self.root.device

    def _matmul(self, rhs):
        if self.tensor is not None:
            return self.tensor.matmul(rhs)
        outs = []
        for L, x in zip(self._panels(self.root), self._batched(rhs)):
            outs.append(ops.panel_rmul(L, ops.gram(L, x.contiguous())))
        return self._restack(outs, self.root)

    def _transpose_nonbatch(self):
        return self

    def evaluate(self):
        if self.tensor is not None:
            return self.tensor
        if self.shape[-1] > _DENSE_LIMIT:
            raise RuntimeError(f"UpdatedRootLazyTensor: refusing to densify a {self.shape[-1]}^2 matrix (matrix-free "
                               f"mode); use matmul / root_decomposition")
        return self.root @ self.root.transpose(-1, -2)

    # ---- batching helpers (panels are [m,r] or [t,m,r])
    @staticmethod
    def _panels(P):
        return [P] if P.dim() == 2 else list(P.reshape(-1, *P.shape[-2:]))

    def _batched(self, x):
        nb = 1 if self.dim() == 2 else self.shape[:-2].numel()
        if x.dim() == 2:
            return [x] * nb
        return list(x.reshape(-1, *x.shape[-2:]))

    @staticmethod
    def _restack(items, like):
        if like.dim() == 2:
            return items[0]
        return torch.stack(items).reshape(*like.shape[:-2], *items[0].shape)

    # ---- root / inverse root  (:121-133 + GPyTorch dispatch, App. A.5)
    def _compute_roots(self):
        A = self.evaluate()
        m = A.shape[-1]
        if m <= settings.max_cholesky_size.value():
            Lc = psd_safe_cholesky(A)
            eye = torch.eye(m, dtype=A.dtype, device=A.device)
            self.root = Lc.contiguous()
            self.inv_root = torch.linalg.solve_triangular(Lc, eye, upper=False).transpose(-1, -2).contiguous()
        else:
            # deterministic stand-in for GPyTorch's randomly-started Lanczos: top-r eigenpairs of the dense matrix
            r = min(settings.max_root_decomposition_size.value(), m)
            lam, V = torch.linalg.eigh(A)
            lam, V = lam[..., -r:].flip(-1), V[..., -r:].flip(-1)
            tol = (1e-10 if A.dtype == torch.float64 else 1e-5) * lam[..., :1]
            keep = lam > tol
            sq = torch.where(keep, lam, torch.ones_like(lam)).sqrt()
            self.root = (V * torch.where(keep, sq, torch.zeros_like(sq)).unsqueeze(-2)).contiguous()
            self.inv_root = (V * torch.where(keep, 1.0 / sq, torch.zeros_like(sq)).unsqueeze(-2)).contiguous()

    def root_decomposition(self, **kwargs):
        if self.root is None:
            self._compute_roots()
        return RootLazyTensor(self._as_lazy(self.root))

    def root_inv_decomposition(self, **kwargs):
        if self.inv_root is None:
            self._compute_roots()
        return RootLazyTensor(self._as_lazy(self.inv_root))

    @staticmethod
    def _as_lazy(P):
        return PanelLazyTensor(P) if P.dim() == 2 else NonLazyTensor(P)

    # ---- updates  (:53-119)
    def update(self, vector):
        """Dense-vector form of the reference API: ``vector`` [..., m, q] (or [m])."""
        if vector.dim() == 1:
            vector = vector.view(-1, 1)
        self.root_decomposition()
        self.root_inv_decomposition()
        tensor = None
        if self.tensor is not None:
            tensor = self.tensor + vector @ vector.transpose(-1, -2)
        root, inv_root = self.collect_vector(vector)
        return UpdatedRootLazyTensor(tensor, initial_is_root=False, root=root, inv_root=inv_root)

    # ---- settings.overlap_root_update: inverse-root half of an in-place rank-q update on a side stream
    def prestart_update_sparse(self, idx, vval):
        """First half of ``update_sparse(idx, vval, inplace=True)``: projections p = B^T v and the square-root factors on
        the current stream, then the in-place update of the inverse-root panel on a side stream.  The caller MUST follow
        up with ``update_sparse(idx, vval, inplace=True)`` (same stencils), which joins the side stream and updates the
        root panel.  Returns False (nothing started) when the update does not have the one-block row-local form."""
        self.root_decomposition()
        self.root_inv_decomposition()
        if (getattr(self, "_pending", None) is not None or self.tensor is not None or not ops.overlap_capable(self.inv_root)
                or settings.root_update_mode.value() != "sym" or not 1 <= idx.shape[0] <= 32):
            return False
        vval = vval.detach()
        Bs = self._panels(self.inv_root)
        vvs = [vval] * len(Bs) if vval.dim() == 2 else list(vval.reshape(-1, *vval.shape[-2:]))
        ps = [ops.left_interp(idx, vv, B).t().contiguous() for B, vv in zip(Bs, vvs)]
        facs = [_sym_factors(p) for p in ps]
        coef = [(p, (C @ p.t()).contiguous(), (Cp @ p.t()).contiguous()) for p, (C, Cp) in zip(ps, facs)]
        with ops.side_section(self.inv_root.device):
            for B, (p, _, CppT) in zip(Bs, coef):
                ops.panel_lowrank_update1_(B, p, CppT)
        self._pending = (idx, coef)
        return True

    def _finish_pending(self, idx):
        pend_idx, coef = self._pending
        self._pending = None
        ops.join_side(self.inv_root.device)
        if pend_idx.shape != idx.shape or pend_idx.data_ptr() != idx.data_ptr():
            raise RuntimeError("update_sparse: a pre-started update is pending for other stencils")
        for L, (p, CpT, _) in zip(self._panels(self.root), coef):
            ops.panel_lowrank_update1_(L, p, CpT)
        return self

    def update_sparse(self, idx, vval, inplace=False):
        """v = W^T D^-1/2 given by its stencils: idx [q,s] (shared by all outputs), vval [q,s] or [t,q,s]."""
        if getattr(self, "_pending", None) is not None:
            if not inplace:
                raise RuntimeError("update_sparse: a pre-started in-place update is pending")
            return self._finish_pending(idx)
        self.root_decomposition()
        self.root_inv_decomposition()
        vval = vval.detach()             # the panels are state, not part of any autograd graph
        Bs = self._panels(self.inv_root)
        vvs = [vval] * len(Bs) if vval.dim() == 2 else list(vval.reshape(-1, *vval.shape[-2:]))
        tensor = self.tensor
        if tensor is not None:
            m = tensor.shape[-1]
            dense = [ops.left_t_interp(idx, vv, torch.eye(idx.shape[0], dtype=vv.dtype, device=vv.device), m) for vv in vvs]
            vvt = self._restack([d @ d.t() for d in dense], self.inv_root)
            tensor = tensor.add_(vvt) if inplace else tensor + vvt
        q = idx.shape[0]
        step = 32 if settings.root_update_mode.value() == "sym" else q
        first = True
        root, inv_root = self.root, self.inv_root
        for s0 in range(0, q, step):
            sl = slice(s0, s0 + step)
            ps = [ops.left_interp(idx[sl], vv[sl], B).t().contiguous()
                  for B, vv in zip(self._panels(inv_root), vvs)]
            root, inv_root = self._apply(ps, inplace=inplace or not first, root=root, inv_root=inv_root)
            first = False
        if inplace:
            self.root, self.inv_root, self.tensor = root, inv_root, tensor
            return self
        return UpdatedRootLazyTensor(tensor, initial_is_root=False, root=root, inv_root=inv_root)

    def fold_in_sparse(self, idx, vval, chunk=4096):
        """Projected update with MANY stencil vectors at once (initial data beyond the root rank), in place.

        The sequence of projected rank-q updates telescopes: with L_k = L_0 G_k, B_k = B_0 G_k^-T one has
        p_k = B_k^T v_k = G_k^-1 (B_0^T v_k), hence  L_n L_n^T = L_0 (I + P P^T) L_0^T  with  P = B_0^T [v_1 .. v_n]
        (r x n).  So instead of n / 32 passes over both panels: gather P in chunks (row gathers of B_0), form
        M = I + P P^T (r x r, accumulated in double), factor M = F F^T and apply  L <- L_0 F,  B <- B_0 F^-T  with ONE
        tensor-core panel GEMM each.  Same L L^T / B B^T as ``update_sparse`` applied point by point."""
        self.root_decomposition()
        self.root_inv_decomposition()
        if self.tensor is not None or idx.shape[0] == 0:
            return self.update_sparse(idx, vval, inplace=True)
        Ls, Bs = self._panels(self.root), self._panels(self.inv_root)
        vvs = [vval] * len(Bs) if vval.dim() == 2 else list(vval.reshape(-1, *vval.shape[-2:]))
        newL, newB = [], []
        for L, B, vv in zip(Ls, Bs, vvs):
            r = B.shape[-1]
            M = torch.eye(r, dtype=torch.float64, device=B.device)
            for s0 in range(0, idx.shape[0], chunk):
                Pc = ops.left_interp(idx[s0:s0 + chunk], vv[s0:s0 + chunk].contiguous(), B).double()     # [chunk, r]
                M = M + Pc.t() @ Pc
            F = torch.linalg.cholesky(M)                                        # M >= I: always positive definite
            Finv_t = torch.linalg.solve_triangular(F, torch.eye(r, dtype=torch.float64, device=B.device),
                                                   upper=False).t()
            newL.append(ops.panel_rmul(L, F.to(L.dtype).contiguous()))
            newB.append(ops.panel_rmul(B, Finv_t.to(B.dtype).contiguous()))
        self.root = self._restack(newL, self.root)
        self.inv_root = self._restack(newB, self.inv_root)
        return self

    def collect_vector(self, vector):
        """(updated_root, updated_inv_root) for a dense ``vector`` — reference signature (:69)."""
        self.root_decomposition()
        self.root_inv_decomposition()
        vs = self._batched(vector)
        q = vs[0].shape[-1]
        step = 32 if settings.root_update_mode.value() == "sym" else q      # the row-local kernel takes q <= 32
        root, inv_root = self.root, self.inv_root
        for s0 in range(0, max(q, 1), step):
            # B^T v_k with the inverse root as updated by the earlier column blocks (the same sequential rule as
            # ``update_sparse``; any q, like the reference's single SVD)
            ps = [ops.gram(B, v[:, s0:s0 + step].contiguous()) for B, v in zip(self._panels(inv_root), vs)]
            root, inv_root = self._apply(ps, inplace=s0 > 0, root=root, inv_root=inv_root)
        return root, inv_root

    def _apply(self, ps, inplace, root=None, inv_root=None):
        root = self.root if root is None else root
        inv_root = self.inv_root if inv_root is None else inv_root
        mode = settings.root_update_mode.value()
        Ls, Bs = self._panels(root), self._panels(inv_root)
        if mode == "svd":
            newL, newB = [], []
            for L, B, p in zip(Ls, Bs, ps):
                U, S, _ = torch.linalg.svd(p, full_matrices=True)        # :82 torch.svd(some=False)
                pad = torch.ones(U.shape[-2] - S.shape[-1], dtype=S.dtype, device=S.device)
                rs = (S ** 2 + 1.0) ** 0.5
                newL.append(ops.panel_rmul(L, U * torch.cat([rs, pad])))          # :97-100
                newB.append(ops.panel_rmul(B, U * torch.cat([1.0 / rs, pad])))    # :111-117
            return self._restack(newL, root), self._restack(newB, inv_root)
        if mode != "sym":
            raise ValueError(f"unknown root_update_mode {mode!r}")
        if not inplace:
            root, inv_root = root.clone(), inv_root.clone()
            Ls, Bs = self._panels(root), self._panels(inv_root)
        for L, B, p in zip(Ls, Bs, ps):
            # the row-local kernel takes q <= 32 columns per call; callers chunk wider updates and re-project each
            # block on the inverse root as updated so far (``update_sparse``, ``collect_vector``)
            if p.shape[1] > 32:
                raise ValueError("rank-q panel update: q <= 32 per call (callers chunk the columns and re-project)")
            C, Cp = _sym_factors(p)
            ops.panel_lowrank_update2_(L, B, p, C @ p.t(), Cp @ p.t())      # both panels in one launch
        return root, inv_root

    # ---- misc
    def _expand_batch(self, batch_shape):
        """Expand along batch dimensions by repetition (:139-159)."""
        cur = torch.Size([1] * (len(batch_shape) - self.dim() + 2) + list(self.batch_shape))
        rep = [e // c for e, c in zip(batch_shape, cur)]
        tensor = None if self.tensor is None else self.tensor.repeat(*rep, 1, 1)
        if self.root is not None:
            return UpdatedRootLazyTensor(tensor, initial_is_root=False, root=self.root.repeat(*rep, 1, 1),
                                         inv_root=self.inv_root.repeat(*rep, 1, 1))
        return UpdatedRootLazyTensor(tensor, initial_is_root=False)

    def expand(self, *sizes):
        if len(sizes) == 1 and not isinstance(sizes[0], int):
            sizes = tuple(sizes[0])
        return self._expand_batch(torch.Size(sizes[:-2]))

    def to(self, device):
        mv = lambda t: None if t is None else t.to(device)
        return UpdatedRootLazyTensor(mv(self.tensor), initial_is_root=False, root=mv(self.root), inv_root=mv(self.inv_root))

    def detach(self):
        return self
