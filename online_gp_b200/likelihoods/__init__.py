from .fnmg_likelihood import FNMGLikelihood, FixedNoiseGaussianLikelihood, HomoskedasticNoise

__all__ = ["FNMGLikelihood"]
