"""Transform pseudo-sampler (gpfanova/sample/transform.py:3-11): a deterministic map of other parameters."""
from .sampler import Sampler


class Transform(Sampler):

    def __init__(self, name, parameters, transform_fxn):
        Sampler.__init__(self, name, 'Transform', parameters)
        self.transform_fxn = transform_fxn

    def _sample(self, *args, **kwargs):
        return self.transform_fxn()
