"""SamplerContainer (gpfanova/sample/container.py:6-105): ordered sampler list, parameter_cache / parameter_history,
thinning, CSV save / load.  Base overrides sample() with the fused device sweep; this generic loop stays for
containers assembled from user samplers."""
import logging
import random as pyrandom
import time

import pandas as pd

from .sampler import Sampler


class SamplerContainer(object):

    def __init__(self, samplers, parameterFile=None, *args, **kwargs):
        self.samplers = [a for a in samplers if issubclass(type(a), Sampler)]
        self.sampler_dict = {a.name: a for a in self.samplers}
        ind = self.build_index()
        self.parameter_cache = pd.Series([0.0] * len(ind), index=ind)
        self.parameter_history = pd.DataFrame(columns=ind)
        for k, v in kwargs.items():
            if k in self.sampler_dict:
                self.parameter_cache[self.sampler_dict[k].parameters] = v
        self.load(parameterFile)

    def build_index(self):
        ind = []
        for s in self.samplers:
            ind += s.parameters
        return ind

    def _sample(self, random=False):
        logger = logging.getLogger(__name__)
        ind = list(range(len(self.samplers)))
        if random:
            pyrandom.shuffle(ind)
        for i in ind:
            sampler = self.samplers[i]
            logger.debug('sampling %s' % str(sampler))
            args = []
            if sampler.current_param_dependent:
                param = self.parameter_cache[sampler.parameters]
                if len(sampler.parameters) == 1:
                    param = param.iloc[0]
                args += [param]
            sample = sampler.sample(*args)
            self.parameter_cache[sampler.parameters] = sample

    def sample(self, n=1, thin=0, verbose=False, random=False):
        logger = logging.getLogger(__name__)
        start = self.parameter_history.shape[0]
        start_time = time.time()
        for i in range(n):
            self._sample(random=random)
            if thin == 0 or (i + 1) % thin == 0:
                self.store()
                j = self.parameter_history.shape[0] - start
                if verbose:
                    logger.info("%d/%d iterations (%.2lf%s) finished in %.2lf minutes" % (
                        j, n, 100. * j / n, '%', (time.time() - start_time) / 60))
        if verbose:
            logger.info("%d samples finished in %.2lf minutes" % (n, (time.time() - start_time) / 60))

    def parameterSamples(self, name):
        if name not in self.sampler_dict:
            return None
        return self.parameter_history[self.sampler_dict[name].parameters]

    def parameterCurrent(self, name):
        if name not in self.sampler_dict:
            return None
        return self.parameter_cache[self.sampler_dict[name].parameters]

    def store(self):
        row = self.parameter_cache.to_frame().T
        if self.parameter_history.shape[0] == 0:
            self.parameter_history = row.reset_index(drop=True)
        else:
            self.parameter_history = pd.concat([self.parameter_history, row], ignore_index=True)
        self.parameter_history.index = range(self.parameter_history.shape[0])

    def save(self, f):
        self.parameter_history.to_csv(f, index=False)

    def load(self, f):
        if f is None:
            return
        self.parameter_history = pd.read_csv(f, index_col=None)

    def resume_from_last(self):
        """copy the last stored row back into parameter_cache (what analysis/hsalTF/hsalTF.py:123 does by hand)"""
        if self.parameter_history.shape[0] > 0:
            last = self.parameter_history.iloc[-1]
            self.parameter_cache[last.index] = last.values.astype(float)

    def __repr__(self):
        return "\n".join(['SamplerContainer:'] + [str(samp) for samp in self.samplers])
