"""gpfanova_b200 -- B200-native Gibbs-sampling hot path of ptonner/gpfanova behind the reference's Python API.

    from gpfanova_b200 import fanova
    m = fanova.FANOVA(x, y, effect, interactions=True, chains=4096, seed=1)
    m.sample(1000, thin=10)
    m.parameter_history            # chain 0, the reference's schema
    m.chain_history()              # every chain

Importing the package never needs a GPU; every compute call does (there is no CPU fallback).
"""
import logging

from . import _capi, engine  # noqa: F401
from . import linalg, kernel, sample, base, mean, fanova, prior, interval  # noqa: F401

logging.basicConfig()
logger = logging.getLogger(__name__)
logger.setLevel(logging.INFO)

__version__ = "0.1.0"
