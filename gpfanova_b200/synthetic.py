"""Synthetic FANOVA problems in the shapes of BASELINE.json `configs`.  Host-side data generation only
(numpy); nothing here is on the sampling path."""
import itertools

import numpy as np

CONFIGS = {
    # name: (levels per effect, replicates, n time points, interactions, helmertConvert, chains)
    "cfg1_oneway_n50": ((3,), 5, 50, False, False, 1),
    "cfg2_twoway_n100": ((3, 3), 5, 100, True, False, 64),
    "cfg3_missing_n50": ((3,), 5, 50, False, False, 256),
    "cfg4_cokol_n200": ((20, 20), 1, 200, True, True, 1024),
    "cfg5_twoway_n256": ((3, 3), 5, 256, True, False, 4096),
}


def effects(levels, reps):
    """test_convergence.py:59 / test_convergence_interactions.py:8: full factorial, replicated"""
    return np.array(list(itertools.product(*[range(l) for l in levels])) * reps)


def _rbf(x, sigma, ls):
    d = (x[:, None, :] - x[None, :, :]) / ls
    return sigma * np.exp(-0.5 * (d ** 2).sum(-1))


def prior_draw(x, design_matrix, groups, y_sigma=-2.0, seed=123):
    """y = samplePrior() at y_sigma, all prior sigma = lengthscale = 10**0 (base.py:245-259, test_convergence.py:180)"""
    rng = np.random.RandomState(seed)
    n = x.shape[0]
    f = design_matrix.shape[1]
    K = _rbf(x, 1.0, 1.0)
    w, V = np.linalg.eigh(K)
    A = V * np.sqrt(np.clip(w, 0, None))           # PSD square root, like the SVD map of multivariate_normal.rvs
    F = np.zeros((n, f))
    for i in range(f):
        F[:, i] = A @ rng.standard_normal(n)
    y = F @ design_matrix.T + rng.normal(0, np.sqrt(10.0 ** y_sigma), size=(n, design_matrix.shape[0]))
    return y, F


def chain_starts(chains, P, y_sigma=-2.0, first=0):
    """per-chain starting hyper-parameters: prior sigma / lengthscale ~ U(-1, 1) from default_rng(1000 + c)
    (test_convergence.py:166-175), y_sigma fixed; functions start at 0"""
    H = 1 + 2 * P
    hyper = np.empty((chains, H))
    for c in range(chains):
        rng = np.random.default_rng(1000 + first + c)
        hyper[c, 0] = y_sigma
        hyper[c, 1:] = rng.uniform(-1, 1, size=H - 1)
    return hyper


def add_missing(y, n, seed=7):
    """cfg3: whole time rows {2,7}*(n/10) missing (test_convergence_missingData.py:41) + 5% scattered NaNs"""
    y = y.copy()
    rows = [2 * n // 10, 7 * n // 10]
    y[rows, :] = np.nan
    rng = np.random.default_rng(seed)
    y[rng.random(y.shape) < 0.05] = np.nan
    return y
