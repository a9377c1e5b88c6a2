"""White noise kernel (gpfanova/kernel/white.py:4-9)."""
import numpy as np
import torch

from .. import engine
from .kernel import Kernel


class White(Kernel):

    def _K_log10(self, X, log10_sigma):
        dev = engine._require_cuda(None)
        n = X.shape[0]
        sigma = torch.pow(torch.tensor(10.0, dtype=torch.float64, device=dev), log10_sigma)
        return (sigma * torch.eye(n, dtype=torch.float64, device=dev)).cpu().numpy()

    def _K(self, X, sigma):
        return self._K_log10(np.asarray(X), float(np.log10(sigma)))
