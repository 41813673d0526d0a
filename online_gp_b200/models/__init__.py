from .batched_fixed_noise_online_gp import FixedNoiseOnlineSKIGP
from .online_ski_regression import OnlineSKIRegression
from . import stems

__all__ = ["FixedNoiseOnlineSKIGP", "OnlineSKIRegression", "stems"]
