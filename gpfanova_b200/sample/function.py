"""Function sampler (gpfanova/sample/function.py:6-40): conditional posterior draw of one latent function.

The numerical work -- residual projection, A = diag(n_t)/sigma_y + K^-1, jitchol(A), mu = A^-1 b and the Gaussian
draw -- is the fused CUDA kernel behind gpf_function_step; this class is the reference-shaped handle on it."""
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
        """(mu, L_A) of the conditional N(mu, (L_A L_A^T)^-1) for chain 0 -- the quantities function.py:34-38 builds;
        z: optional injected standard normals (n,) used for the draw (beta = mu + L_A^-T z)."""
        return self.base._device_function_step(self.f, z=z, want_conditional=True)
