"""Credible intervals from stored samples -- the API of gpfanova/interval.py:3-184 (ScalarInterval: Chen & Shao HPD region;
FunctionInterval: mean +- sd widened to joint coverage alpha), computed by the device operators gpf_row_sort /
gpf_function_interval, plus batched entry points for many parameters / functions / chains at once."""
import numpy as np

from . import engine


class Interval(object):

    def __init__(self, samples, alpha, ndim=1):
        self.samples = np.asarray(samples, dtype=np.float64)
        self.n = self.samples.shape[0]
        self.alpha = alpha
        if self.alpha < 0 or self.alpha > 1:
            raise ValueError("must provide alpha between 0 and 1")
        if self.samples.ndim != ndim:
            raise ValueError("sample dimensions does not match %d" % ndim)

    def contains(self, x):
        raise NotImplementedError("implement this for your interval!")


def hpd_thresholds(logdens, alpha):
    """Chen & Shao (1999) thresholds for a batch: logdens (batch, S) = pi evaluated at every sample; returns the density
    value above which a point lies in the 1 - alpha ... region, element int(S alpha) of the sorted row (interval.py:32-35)"""
    v = engine.row_sort(np.atleast_2d(np.asarray(logdens, dtype=np.float64)))
    return v[:, int(v.shape[1] * alpha)].cpu().numpy()


class ScalarInterval(Interval):

    def __init__(self, samples, alpha, pi):
        """pi: density (or log density) of the parameter given the data, callable on a scalar"""
        Interval.__init__(self, samples, alpha, ndim=1)
        self.pi = pi
        dens = np.array([self.pi(t) for t in self.samples], dtype=np.float64)
        self.epsilon = engine.row_sort(dens[None])[0].cpu().numpy()
        self.epj = self.epsilon[int(self.n * self.alpha)]      # any observation with pi(x|D) > epj is in the region

    def contains(self, x):
        return self.pi(x) > self.epj


def function_intervals(samples, alpha=.95, start=1e-6, tol=1e-6, maxiter=100):
    """FunctionInterval for a batch of sample sets (batch, S, n) in one launch: returns dict(mean, std, epsilon, lb, ub,
    evaluations) of numpy arrays (lb/ub = mean -+ std -+ epsilon, interval.py:114-115)"""
    mean, std, eps, it = (t.cpu().numpy() for t in engine.function_interval(samples, alpha, start, tol, maxiter))
    return dict(mean=mean, std=std, epsilon=eps, lb=mean - std - eps[:, None], ub=mean + std + eps[:, None], evaluations=it)


class FunctionInterval(Interval):

    def __init__(self, samples, alpha, start=1e-6, tol=1e-6, maxiter=100):
        Interval.__init__(self, samples, alpha, ndim=2)
        r = function_intervals(self.samples[None], alpha, start, tol, maxiter)
        grid = np.arange(self.n) / float(self.n)               # alpha moves to the closest k/S (interval.py:66-67)
        self.alpha = grid[np.argmin(np.abs(grid - alpha))]
        self.mean, self.std, self.epsilon = r["mean"][0], r["std"][0], float(r["epsilon"][0])
        self.lb, self.ub = r["lb"][0], r["ub"][0]
        self.evaluations = int(r["evaluations"][0])

    def contains(self, x):
        return bool((x > self.lb).all() & (x < self.ub).all())


def functionInterval(samples, start=1e-6, alpha=.95, tol=1e-6, maxiter=100):
    """interval.py:134-184: the epsilon of FunctionInterval (the list of bisection bounds the reference also returns is a
    trace of its Python loop and is not reproduced: an empty list)"""
    r = function_intervals(np.asarray(samples, dtype=np.float64)[None], alpha, start, tol, maxiter)
    return float(r["epsilon"][0]), []
