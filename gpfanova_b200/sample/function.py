"""Function sampler (gpfanova/sample/function.py:6-40): conditional posterior draw of one latent function.

The numerical work -- residual projection and a draw from N(A^-1 b, A^-1), A = diag(n_t)/sigma_y + K^-1, evaluated in
observation form without inverting K -- is the fused CUDA kernel behind gpf_function_step; this class is the
reference-shaped handle on it."""
from .sampler import Sampler


class Function(Sampler):

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
