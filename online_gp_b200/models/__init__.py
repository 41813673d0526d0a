from .batched_fixed_noise_online_gp import FixedNoiseOnlineSKIGP
from .online_ski_regression import OnlineSKIRegression
from .online_ski_botorch_model import GPyTorchPosterior, OnlineSKIBotorchModel
from .fantasy import FantasizedOnlineSKIGP
from .online_ski_classifier import DirichletGPClassifier, OnlineSKIClassifier
from . import stems

__all__ = ["FixedNoiseOnlineSKIGP", "OnlineSKIRegression", "OnlineSKIBotorchModel", "GPyTorchPosterior",
           "FantasizedOnlineSKIGP", "DirichletGPClassifier", "OnlineSKIClassifier", "stems"]
