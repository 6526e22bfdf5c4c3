"""cauchyfriendly_b200 -- B200-native per-step term propagation of the Multivariate Cauchy Estimator.

The package is a thin host-side mirror of the reference's CauchyEstimator interface over libmce_b200.so
(hand-written sm_100a CUDA behind the C ABI in include/mce_b200.h).  There is no CPU implementation here."""
from .estimator import CauchyEstimator  # noqa: F401
from . import _capi  # noqa: F401

__all__ = ["CauchyEstimator"]
