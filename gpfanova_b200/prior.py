"""Prior over a group of latent functions (gpfanova/prior.py:6-111).

The reference's class is an unfinished refactor (it is not imported by gpfanova/__init__.py and calls methods that
do not exist, prior.py:46,50,94), so there is no behaviour to be bug-compatible with; this keeps its shape -- one RBF
kernel shared by a set of functions, `loglikelihood` and `sample` -- backed by the same device operators as the model.
"""
import numpy as np
import torch

from . import engine
from .kernel import RBF


class Prior(object):

    def __init__(self, x, kernel, functions, base=None, name='prior'):
        """x: (n, p) inputs; kernel: Kernel object (e.g. RBF(model, [...])); functions: list of function indices"""
        self.x = np.asarray(x, dtype=np.float64)
        self.n = self.x.shape[0]
        self.kernel = kernel
        self.functions = list(functions)
        self.base = base
        self.name = name

    def loglikelihood(self, F, *args, **kwargs):
        """sum_f log N(F[:, f]; 0, K + mean(K) 1e-6 I)  (the MVN part of base.py:216-243) through the device
        Cholesky: log-determinant from the pivots, quadratic forms from triangular solves."""
        F = np.asarray(F, dtype=np.float64)
        if F.ndim == 1:
            F = F[:, None]
        K = self.kernel.K(self.x, *args, **kwargs)
        cov = K + K.mean() * np.eye(self.n) * 1e-6
        L, logdet, _ = engine.jitchol(cov, return_info=True)
        V = torch.linalg.solve_triangular(L, torch.as_tensor(F).to(L.device), upper=False)
        quad = (V * V).sum(0)
        ll = -0.5 * (self.n * np.log(2 * np.pi) + logdet + quad)
        return float(ll.sum().item())

    def sample(self, size=1, *args, **kwargs):
        """draws from N(0, K) (base.py:245-254) using the device kernel matrix and factor"""
        K = self.kernel.K(self.x, *args, **kwargs)
        L = engine.jitchol(K)
        z = torch.as_tensor(np.random.standard_normal((self.n, size))).to(L.device)
        return (L @ z).cpu().numpy()
