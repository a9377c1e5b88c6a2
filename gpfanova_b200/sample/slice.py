"""Slice sampler (gpfanova/sample/slice.pyx:6-62), Neal (2003) step-out + shrinkage.

The model's own hyper-parameter slices (y_sigma, prior*_sigma, prior*_lengthscale) run entirely on the device
(gpf_slice_step: the step-out/shrink loop around the fused GP log-likelihood kernels).  A Slice built by user code
around an arbitrary Python log-density keeps the reference's host-side control loop, which only sequences calls to
that closure."""
import numpy as np

from .sampler import Sampler


class Slice(Sampler):

    def __init__(self, name, parameters, logdensity_fxn, w, m, base=None, hyper_index=None):
        """logdensity_fxn: log conditional density of the parameter; w: interval step; m: max interval = w*m"""
        Sampler.__init__(self, name, 'Slice', parameters, current_param_dependent=True)
        self.logdensity_fxn = logdensity_fxn
        self.w = w
        self.m = m
        self._base = base
        self._hyper_index = hyper_index

    def _sample(self, x):
        if self._base is not None:
            return self._base._device_slice_step(self._hyper_index, x)
        return self._sample_closure(x)

    def _sample_closure(self, x):
        """slice.pyx:24-62 around a caller-supplied closure (control flow only)"""
        f0 = self.logdensity_fxn(x)
        z = f0 - (1.0 + np.random.exponential(1.0))     # scipy.stats.expon.rvs(1): loc = 1 (slice.pyx:30)
        u = np.random.uniform(0, 1)
        l = x - self.w * u
        r = l + self.w
        v = np.random.uniform(0, 1)
        j = int(self.m * v)
        k = self.m - 1 - j
        while j > 0 and z < self.logdensity_fxn(l):
            j -= 1
            l -= self.w
        while k > 0 and z < self.logdensity_fxn(r):
            k -= 1
            r += self.w
        u = np.random.uniform(0, 1)
        x1 = l + u * (r - l)
        while z > self.logdensity_fxn(x1):
            if x1 < x:
                l = x1
            else:
                r = x1
            u = np.random.uniform(0, 1)
            x1 = l + u * (r - l)
        return x1
