"""RBF kernel (gpfanova/kernel/rbf.pyx:4-38): sigma * exp(-|x - x'|^2 / (2 l^2)) built by the CUDA operator."""
import numpy as np

from .. import engine
from .kernel import Kernel


class RBF(Kernel):
    """Radial basis function kernel; parameters = [sigma name, lengthscale name]"""

    @staticmethod
    def dist(X, lengthscale):
        """rbf.pyx:8-14.  Scaled distances, from the device kernel: d = sqrt(-2 log K) with sigma = 1."""
        K = engine.rbf_K(np.asarray(X, dtype=np.float64), 0.0, float(np.log10(lengthscale)))[0]
        return np.sqrt(np.clip(-2.0 * np.log(K.cpu().numpy()), 0, np.inf))

    def _K_log10(self, X, log10_sigma, log10_lengthscale):
        return engine.rbf_K(X, log10_sigma, log10_lengthscale)[0].cpu().numpy()

    def _K(self, X, sigma, lengthscale):
        """rbf.pyx:16-19 with natural-scale parameters"""
        return self._K_log10(np.asarray(X, dtype=np.float64), float(np.log10(sigma)), float(np.log10(lengthscale)))

    def _dK(self, X, cross, sigma, lengthscale):
        """rbf.pyx:21-38 (only used by FunctionDerivative, `derivatives=True`; p = 1 inputs as in the reference)"""
        X = np.asarray(X, dtype=np.float64)
        k = self._K(X, sigma, lengthscale)
        diff = X[:, 0][:, None] - X[:, 0][None, :]
        if cross:
            return (-1. / lengthscale) * diff * k
        return (1. / lengthscale) * (1 - (1. / lengthscale) * (diff ** 2)) * k
