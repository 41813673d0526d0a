"""Feature flags and value contexts steering the hot path.

Mirrors ``online_gp/settings.py:3-7`` (``check_decomposition``, ``detach_interp_coeff``) and the subset of
``gpytorch.settings`` the reference path reads (SURVEY.md §5 "Config / flags", App. A.5 for the defaults):
``max_cholesky_size`` (800), ``max_root_decomposition_size`` (100), ``cg_tolerance`` (1.0), ``eval_cg_tolerance``
(0.01), ``max_cg_iterations`` (1000), ``cholesky_jitter``, ``skip_logdet_forward``, ``skip_posterior_variances``,
``fast_pred_var``, ``fast_pred_samples``, ``use_toeplitz``, ``detach_test_caches``.  Same usage:
``with max_cholesky_size(2048), cg_tolerance(1e-2): ...`` / ``flag.on()`` / ``value_ctx.value()``.
"""
import torch


class _feature_flag:
    _state = False

    @classmethod
    def on(cls):
        return cls._state

    @classmethod
    def off(cls):
        return not cls._state

    @classmethod
    def _set_state(cls, state):
        cls._state = state

    def __init__(self, state=True):
        self.prev = self.__class__.on()
        self.state = state

    def __enter__(self):
        self.__class__._set_state(self.state)

    def __exit__(self, *args):
        self.__class__._set_state(self.prev)
        return False


class _value_context:
    _global_value = None

    @classmethod
    def value(cls):
        return cls._global_value

    @classmethod
    def _set_value(cls, value):
        cls._global_value = value

    def __init__(self, value):
        self._orig_value = self.__class__.value()
        self._instance_value = value

    def __enter__(self):
        self.__class__._set_value(self._instance_value)

    def __exit__(self, *args):
        self.__class__._set_value(self._orig_value)
        return False


# --- online_gp/settings.py
class check_decomposition(_feature_flag):
    _state = False


class detach_interp_coeff(_feature_flag):
    _state = False


# --- gpytorch.settings subset
class max_cholesky_size(_value_context):
    _global_value = 800


class max_root_decomposition_size(_value_context):
    _global_value = 100


class cg_tolerance(_value_context):
    _global_value = 1.0


class eval_cg_tolerance(_value_context):
    _global_value = 0.01


class max_cg_iterations(_value_context):
    _global_value = 1000


class cholesky_jitter(_value_context):
    _global_value = None     # None -> dtype default (1e-6 fp32, 1e-8 fp64)

    @classmethod
    def value(cls, dtype=None):
        if cls._global_value is not None:
            return cls._global_value
        return 1e-6 if dtype == torch.float32 else 1e-8


class skip_logdet_forward(_feature_flag):
    _state = False


class skip_posterior_variances(_feature_flag):
    _state = False


class fast_pred_var(_feature_flag):
    _state = False


class fast_pred_samples(_feature_flag):
    _state = False


class use_toeplitz(_feature_flag):
    _state = True


class detach_test_caches(_feature_flag):
    _state = True


# --- additions of this implementation
class root_update_mode(_value_context):
    """How ``UpdatedRootLazyTensor.collect_vector`` applies the rank-q update (updated_root_lazy_tensor.py:69-119).

    "sym": row-local  P <- P + (P U) V^T  with  (I + p p^T)^(+-1/2) = I + U f(S) U^T  (two panel passes, no GEMM);
    "svd": the reference's literal form — full SVD of p, two m x r x r panel GEMMs.  Both give the same
    L L^T and B B^T; they differ by an orthogonal right factor, which no downstream quantity sees.
    """
    _global_value = "sym"


class check_interp_bounds(_feature_flag):
    """Raise GPyTorch's out-of-bounds RuntimeError eagerly (costs one device->host flag read per call)."""
    _state = True


class kron_tensor_core_axes(_feature_flag):
    """Run Kronecker axes with g >= 64 grid points (fp32) as batched tcgen05 GEMMs with 3xTF32 split operands instead
    of the direct SIMT kernels (the 128^3, 256^2 and 1024^2 grids of BASELINE.json); off = SIMT everywhere."""
    _state = True


class sharded_dual_layout(_feature_flag):
    """Row-sharded model: additionally keep the root panel in the column-sharded layout (all rows of r / world
    columns, maintained by the rank-q update with one all-gather of an m x q vector).  ``K L`` is then a purely local
    Kronecker MVM and the hyper-gradient a purely local pair of gradient passes, which halves the row <-> column
    exchanges per step (2 instead of 4) at the price of one more slab-sized panel per rank.  On by default: measured
    on 2 / 4 / 8 B200 with the pushing kernels it is the faster form everywhere (8 GPUs: 213.6 vs 191.6 updates/s on
    BASELINE config 2; profiles/r02_scaling.md)."""
    _state = True


class defer_interp_bounds_check(_feature_flag):
    """Queue the out-of-bounds flag of ``ops.interpolate`` (async copy to pinned memory + event) instead of reading
    it back at once; ``ops.flush_bounds_checks()`` raises later without stalling the stream.  Used by
    ``OnlineSKIRegression.evaluate``, which flushes before it returns (no model state is modified in between)."""
    _state = False


class kron_directional_grad(_feature_flag):
    """Directional (JVP) form of the fused Kronecker column-gradient pass: the loss depends on grid column i only through
    <grad_i, d col_i / d lengthscale_i> and <grad_i, col_i> (every scale-type parameter), so the pass applies the
    direction matrix T'_i and takes dot products instead of forming the 32-entry column gradient
    (``ops._surrogate_col_grad``).  Exact for stationary product kernels with one lengthscale per dimension and scalar
    scales (RBF / Matern + ScaleKernel + noise: every shipped config) and parity-tested against the full contraction.
    On (default) it runs on the tensor pipe (``csrc/kron_tc.cu``); kernels without ``grid_column_dirs`` and
    ``kron_directional_grad(False)`` use the full column-gradient pass."""
    _state = True


class backward_gemm_tf32_passes(_value_context):
    """tcgen05 passes of the fp32 panel GEMM that produces the *gradient* panel ``Z = L grad_Q`` in the backward of
    ``Q = I + L^T K L`` (``online_gp/models/online_ski_regression.py:141``).  3 (default) = the 3xTF32 split every value
    GEMM uses; a 500-step stream then follows the fp64 oracle within 1e-2 (tests/model_cases.py).  2 = grad_Q exact (big
    + remainder stacked along K), the panel L cut to tf32 by the tensor core; 1 = one raw tf32 pass.  Both cheaper forms
    are opt-in only: the tensor core TRUNCATES its fp32 inputs, so the gradient carries a one-sided ~5e-4 bias and the
    Adam trajectory walks away from the oracle (noise 8 % off after 500 steps with 2 passes, first step outside the
    1e-2 band at 177 / 145 with 2 / 1 passes; tools/diag_drift.py, B200).  On C2 the 2-pass form is not even faster
    (2.27 ms vs 2.09 ms: its 128-byte K slabs halve the bytes in flight per pipeline stage).  fp64 is unaffected."""
    _global_value = 3


class overlap_root_update(_feature_flag):
    """Streaming ``update()``: the in-place rank-q update of the INVERSE-root panel B (HBM-bound, 0.8 ms at m = 2^20,
    r = 432) does not depend on the hyper-parameter step, and nothing in that step reads B.  With this flag it is issued on
    a side stream before the MLL backward, whose first big kernel — the 3xTF32 panel GEMM — is tensor-bound and leaves 3/4
    of the HBM bandwidth idle; the root panel L (read by the backward) is updated afterwards as before.  Same kernels, same
    operands, same order per panel: results are bit-identical to the serial schedule.  The fork / join is by stream events,
    so it is captured into the CUDA graph of ``update`` like everything else."""
    _state = True


class kron_outer_inner_pairing(_feature_flag):
    """32^4 grids, tensor-core pair kernels: pair the grid axes as (1,2) + (0,3) instead of (0,1) + (2,3).  Every tile of the
    pair (0,1) is 1024 rows that lie 1024 rows apart (one 64-byte piece per 2 MB page); the tiles of (0,3) are 32 runs of 32
    consecutive rows and those of (1,2) rows 32 apart inside 32 pages, which the memory system streams about twice as
    fast (measured on B200, profiles/r02_pair_kernels.md).  The Kronecker factors commute, so the result is the same."""
    _state = True
