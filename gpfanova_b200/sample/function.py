"""Function samplers (gpfanova/sample/function.py:6-66): conditional posterior draw of one latent function, and of its
derivative given the function.

The numerical work of Function -- residual projection and a draw from N(A^-1 b, A^-1), A = diag(n_t)/sigma_y + K^-1,
evaluated in observation form without inverting K -- is the CUDA kernel behind gpf_function_step; FunctionDerivative is
gpf_function_derivative.  These classes are the reference-shaped handles on them."""
import numpy as np

from .sampler import Sampler


class Function(Sampler):

    derivative = False

    def __init__(self, name, parameters, base, f, kernel):
        Sampler.__init__(self, name, 'Function', parameters, current_param_dependent=False)
        self.base = base
        self.f = f
        self.kernel = kernel

    def _sample(self):
        """one conditional draw of function f for every chain on the device; returns chain 0's values (n,)"""
        return self.base._device_function_step(self.f)

    def conditional(self, z=None):
        """(mu, Sigma) of the conditional N(mu, Sigma) for chain 0 -- the mean and covariance function.py:34-38 builds
        (A^-1 b, A^-1); z: optional injected standard normals (2n,) [z1 | z2] used for the draw that is written
        into the state."""
        return self.base._device_function_step(self.f, z=z, want_conditional=True)


class FunctionDerivative(Sampler):
    """beta'_f | beta_f  ~  N(K_ba K_a^-1 beta_f, K_b - K_ba K_a^-1 K_ba^T)  (function.py:43-66); `derivatives=True`"""

    derivative = True

    def __init__(self, name, parameters, base, f, kernel):
        Sampler.__init__(self, name, 'Function', parameters, current_param_dependent=False)
        self.base = base
        self.f = f
        self.kernel = kernel

    def _state(self):
        b = self.base
        vals = b.parameter_cache
        hyper = np.array([[float(vals[nm]) for nm in b.hyperNames()]])
        F = np.stack([vals[b.functionIndex(q)].values.astype(float) for q in range(b.f)])[None]
        return hyper, F

    def _params(self):
        """(mu (n, 1), cov (n, n)) at the current state of chain 0, like function.py:51-60"""
        hyper, F = self._state()
        _, mom = self.base._device_derivatives(hyper, F, stream_id=self.base._next_manual_sweep(), want_moments=True)
        g = self.base.functionPrior(self.f)
        mu, cov = mom[g]
        return mu[0, self.base._groups[g].index(self.f)][:, None], cov[0]

    def _sample(self):
        hyper, F = self._state()
        dF = self.base._device_derivatives(hyper, F, stream_id=self.base._next_manual_sweep())
        return dF[0, self.f]
