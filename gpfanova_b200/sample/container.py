"""SamplerContainer -- the state store and sweep driver API of gpfanova/sample/container.py:6-105: an ordered list of
samplers, `parameter_cache` (current value of every label), `parameter_history` (one row per stored sweep), thinning,
CSV save / load.  Base overrides sample() with the fused device sweep; the generic loop below serves containers
assembled from arbitrary user samplers."""
import logging
import random as pyrandom
import time

import numpy as np
import pandas as pd

from .sampler import Sampler

LOG = logging.getLogger(__name__)


class SamplerContainer(object):

    def __init__(self, samplers, parameterFile=None, *args, **kwargs):
        self.samplers = [s for s in samplers if isinstance(s, Sampler)]
        self.sampler_dict = {s.name: s for s in self.samplers}
        labels = self.build_index()
        self.parameter_cache = pd.Series(np.zeros(len(labels)), index=labels)
        self.parameter_history = pd.DataFrame(columns=labels)
        # keyword arguments named after a sampler give its initial value (container.py:16-18)
        for name in set(kwargs) & set(self.sampler_dict):
            self.parameter_cache[self.sampler_dict[name].parameters] = kwargs[name]
        self.load(parameterFile)

    def build_index(self):
        return [lab for s in self.samplers for lab in s.parameters]

    def _sample(self, random=False):
        """one sweep: every sampler once, in list order or shuffled (container.py:29-53)"""
        todo = list(self.samplers)
        if random:
            pyrandom.shuffle(todo)
        for s in todo:
            LOG.debug('sampling %s' % str(s))
            args = ()
            if s.current_param_dependent:
                cur = self.parameter_cache[s.parameters]
                args = (cur.iloc[0] if len(s.parameters) == 1 else cur,)
            self.parameter_cache[s.parameters] = s.sample(*args)

    def sample(self, n=1, thin=0, verbose=False, random=False):
        t0 = time.time()
        stored = 0
        for i in range(n):
            self._sample(random=random)
            if thin == 0 or (i + 1) % thin == 0:
                self.store()
                stored += 1
                if verbose:
                    LOG.info("%d/%d iterations (%.2lf%s) finished in %.2lf minutes" % (
                        stored, n, 100. * stored / n, '%', (time.time() - t0) / 60))
        if verbose:
            LOG.info("%d samples finished in %.2lf minutes" % (n, (time.time() - t0) / 60))

    def parameterSamples(self, name):
        s = self.sampler_dict.get(name)
        return None if s is None else self.parameter_history[s.parameters]

    def parameterCurrent(self, name):
        s = self.sampler_dict.get(name)
        return None if s is None else self.parameter_cache[s.parameters]

    def store(self):
        """append the current parameter_cache as a new row of parameter_history (container.py:90-92)"""
        row = pd.DataFrame([self.parameter_cache.to_numpy()], columns=self.parameter_cache.index)
        hist = self.parameter_history
        self.parameter_history = row if hist.shape[0] == 0 else pd.concat([hist, row], ignore_index=True)

    def save(self, f):
        self.parameter_history.to_csv(f, index=False)

    def load(self, f):
        if f is not None:
            self.parameter_history = pd.read_csv(f, index_col=None)

    def resume_from_last(self):
        """copy the last stored row back into parameter_cache (what analysis/hsalTF/hsalTF.py:123 does by hand)"""
        if self.parameter_history.shape[0]:
            last = self.parameter_history.iloc[-1]
            self.parameter_cache[last.index] = last.values.astype(float)

    def __repr__(self):
        return "\n".join(['SamplerContainer:'] + [str(s) for s in self.samplers])
