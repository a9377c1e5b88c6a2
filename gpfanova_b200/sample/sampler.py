"""Sampler protocol (gpfanova/sample/sampler.py:5-30)."""


class Sampler(object):

    def __init__(self, name, _type, parameters, current_param_dependent=False, *args, **kwargs):
        """current_param_dependent: does _sample need the current parameter value (slice sampling)?"""
        self.name = name
        self.type = _type
        self.current_param_dependent = current_param_dependent
        if type(parameters) == str:
            self.parameters = [parameters]
        else:
            self.parameters = parameters

    def sample(self, *args, **kwargs):
        return self._sample(*args, **kwargs)

    def _sample(self, *args, **kwargs):
        raise NotImplementedError("implement this function for your sampler!")

    def __repr__(self):
        if len(self.parameters) < 3:
            return "%s (%s): %s" % (self.name, self.type, ', '.join(self.parameters))
        return "%s (%s): %s, ..." % (self.name, self.type, ', '.join(self.parameters[:3]))
