"""Prior over a group of latent functions (gpfanova/prior.py:6-111).

The reference's class is an unfinished refactor (it is not imported by gpfanova/__init__.py and calls methods that do not
exist, prior.py:46,50,94), so there is no behaviour to be bug-compatible with; this keeps its shape -- one RBF kernel
shared by a set of functions, `loglikelihood` and `sample` -- on the batched device operators gpf_gp_loglik / gpf_gp_sample
(jitchol with the reference's jitter ladder, log-determinant from the pivots, quadratic forms from rows riding along the
factorisation, draws L z)."""
import numpy as np

from . import engine


class Prior(object):

    def __init__(self, x, kernel, functions, base=None, name='prior'):
        """x: (n, p) inputs; kernel: Kernel object (e.g. RBF(model, [...])); functions: list of function indices"""
        self.x = np.asarray(x, dtype=np.float64)
        self.n = self.x.shape[0]
        self.kernel = kernel
        self._functions = [int(f) for f in functions]
        self.base = base
        self.name = name

    def functions(self):
        return self._functions

    def _log10(self, *args, **kwargs):
        return self.kernel._log10_params(*args, **kwargs)

    def loglikelihood(self, F=None, *args, **kwargs):
        """sum_f log N(F[:, f]; 0, L L^T), L = jitchol(K)  (prior.py:86-99).  F: (n, #functions) values of this prior's
        functions; default: the current values in the model's parameter_cache.  Kernel parameters may be overridden by
        position or name like Kernel.K (log10 values for logspace kernels)."""
        if F is None:
            F = self.base.functionMatrix(only=self._functions)
        F = np.asarray(F, dtype=np.float64)
        if F.ndim == 1:
            F = F[:, None]
        ls, ll = self._log10(*args, **kwargs)
        return float(engine.gp_loglik(self.x, ls, ll, F.T[None])[0].item())

    def loglikelihoods(self, log10_sigma, log10_lengthscale, F):
        """batched: one value per (sigma, lengthscale) pair; F (batch, #functions, n) or (#functions, n) shared"""
        return engine.gp_loglik(self.x, log10_sigma, log10_lengthscale, F).cpu().numpy()

    def sample(self, size=None, seed=None, *args, **kwargs):
        """draws from N(0, L L^T), L = jitchol(K) (prior.py:67-84): (n, #functions), one draw per function of this prior;
        size=S: (S, n) draws"""
        ls, ll = self._log10(*args, **kwargs)
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        k = len(self._functions) if size is None else int(size)
        out = engine.gp_sample(self.x, ls, ll, nsamp=k, seed=seed)[0].cpu().numpy()
        return out.T if size is None else out
