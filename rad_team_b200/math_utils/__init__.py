"""Mirrors the reference's ``src/math_utils`` (a namespace package there): same module, class and
method names, every call computed by the CUDA kernels behind include/rt_b200.h."""
from .estimator import IntensityResamplingEstimator
from .normalize import BensLogNormalizer, Normalizer
from .standardize import BatchedWelford, RunningMoments, WelfordsOnlineStandardization

__all__ = ["IntensityResamplingEstimator", "BensLogNormalizer", "Normalizer", "WelfordsOnlineStandardization",
           "BatchedWelford", "RunningMoments"]
