"""RBF kernel (gpfanova/kernel/rbf.pyx:4-38): sigma * exp(-|x - x'|^2 / (2 l^2)), its scaled distances and its derivative
kernels, all built by the CUDA operators (gpf_rbf_K, gpf_rbf_dist, gpf_rbf_dK)."""
import numpy as np

from .. import engine
from .kernel import Kernel


class RBF(Kernel):
    """Radial basis function kernel; parameters = [sigma name, lengthscale name]"""

    @staticmethod
    def dist(X, lengthscale):
        """rbf.pyx:8-14: sqrt(clip(-2 X X^T + |X_i|^2 + |X_j|^2, 0)) of the inputs scaled by the lengthscale"""
        return engine.rbf_dist(np.asarray(X, dtype=np.float64), float(np.log10(lengthscale)))[0].cpu().numpy()

    def _K_log10(self, X, log10_sigma, log10_lengthscale):
        return engine.rbf_K(X, log10_sigma, log10_lengthscale)[0].cpu().numpy()

    def _K(self, X, sigma, lengthscale):
        """rbf.pyx:16-19 with natural-scale parameters"""
        return self._K_log10(np.asarray(X, dtype=np.float64), float(np.log10(sigma)), float(np.log10(lengthscale)))

    def _dK(self, X, cross, sigma, lengthscale):
        """rbf.pyx:21-38, the derivative kernels FunctionDerivative uses (first input dimension, like the reference)"""
        return engine.rbf_dK(np.asarray(X, dtype=np.float64), float(np.log10(sigma)), float(np.log10(lengthscale)),
                             cross=cross)[0].cpu().numpy()
