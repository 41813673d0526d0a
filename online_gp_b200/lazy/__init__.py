from .lazy_tensor import (AddedDiagLazyTensor, BatchLazyTensor, DiagLazyTensor, InterpolatedLazyTensor,
                          KroneckerToeplitzLazyTensor, LazyTensor, MatmulLazyTensor, NonLazyTensor, NotPSDError,
                          NumericalWarning, PanelGramLazyTensor, PanelLazyTensor, PanelTLazyTensor, RootLazyTensor,
                          ZeroLazyTensor, delazify, lazify, psd_safe_cholesky)
from .updated_root_lazy_tensor import UpdatedRootLazyTensor

__all__ = ["UpdatedRootLazyTensor", "LazyTensor", "RootLazyTensor", "lazify", "delazify"]
