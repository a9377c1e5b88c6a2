"""gpfanova_b200 -- B200-native Gibbs-sampling hot path of ptonner/gpfanova behind the reference's Python API.

Importing the package never needs a GPU; every compute call does (there is no CPU fallback).
"""
from . import _capi, engine  # noqa: F401

__version__ = "0.1.0"
