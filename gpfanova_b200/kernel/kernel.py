"""Kernel base class -- same parameter-resolution protocol as gpfanova/kernel/kernel.pyx:9-70, with the matrix
work (K, K_inv) done by the CUDA operators behind the C ABI."""
import logging
import math

import numpy as np

from .. import engine

OFFSET = 1e-9   # kernel.pyx:7


class Kernel(object):

    def __init__(self, model, params, logspace, *args, **kwargs):
        self.model = model
        self.parameters = params
        self.logspace = logspace

    def resolve_params(self, *args, **kwargs):
        """kernel.pyx:16-33: positional args, then kwargs by parameter name, then the model's parameter_cache.
        Returns the raw (un-exponentiated) values."""
        params = [None] * len(self.parameters)
        for i in range(len(args)):
            if args[i] is None:
                continue
            params[i] = args[i]
        for p in self.parameters:
            if p in kwargs and kwargs[p] is not None:
                params[self.parameters.index(p)] = kwargs[p]
        while None in params:
            i = params.index(None)
            params[i] = self.model.parameter_cache[self.parameters[i]]
        return [float(v) for v in params]

    def build_params(self, *args, **kwargs):
        """kernel.pyx:16-39 (values after the optional 10**)"""
        params = self.resolve_params(*args, **kwargs)
        if self.logspace:
            params = [pow(10, v) for v in params]
        return params

    def _log10_params(self, *args, **kwargs):
        params = self.resolve_params(*args, **kwargs)
        if not self.logspace:
            params = [math.log10(v) if v > 0 else float("nan") for v in params]
        return params

    def K(self, X, *args, **kwargs):
        """covariance matrix for inputs X (kernel.pyx:41-44); numpy array (n, n)"""
        return self._K_log10(np.asarray(X, dtype=np.float64), *self._log10_params(*args, **kwargs))

    def dK(self, X, cross=False, *args, **kwargs):
        params = self.build_params(*args, **kwargs)
        return self._dK(X, cross, *params)

    def K_inv(self, X, *args, **kwargs):
        """(K + 1e-9 I)^-1 through jitchol (kernel.pyx:51-64)"""
        K = self.K(X, *args, **kwargs)
        try:
            return engine.kinv(K, OFFSET).cpu().numpy()
        except np.linalg.LinAlgError as e:
            logging.getLogger(__name__).error('Kernel inversion error: %s' % str(self.parameters))
            raise e

    def _K_log10(self, X, *log10_params):
        raise NotImplementedError()

    def _dK(self, X, cross, *args):
        raise NotImplementedError()
