#!/usr/bin/env python
"""Mint tests/golden/*.npz by running the UNMODIFIED reference (py3-importable copy
emitted by oracle/build_ref.py into oracle/_ref).  TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference); the .npz files are
committed so the GPU box, which has no /root/reference, can check against them.

Captured per case: inputs (x, y, effect, state) and the reference's own outputs for
every stage of the hot path: design matrix, K, K_inv, functionResidual-derived
(m_t, n_t are recomputed from the reference's residual exactly as function.py:16-25),
the (mu, cov) handed to multivariate_normal.rvs at function.py:40, observation and
prior log-likelihoods, Slice._sample trajectories under an injected uniform stream,
jitchol outputs, and the FANOVA transforms.
"""
import itertools
import os
import sys

import numpy as np
import scipy.stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_ref  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


class Capture(object):
    """intercept scipy.stats.multivariate_normal.rvs(mu, cov) (function.py:40)"""

    def __enter__(self):
        self.orig = scipy.stats.multivariate_normal.rvs
        self.calls = []

        def rvs(mean=None, cov=1, *a, **k):
            self.calls.append((np.array(mean, copy=True), np.array(cov, copy=True)))
            return np.zeros_like(mean)
        scipy.stats.multivariate_normal.rvs = rvs
        return self

    def __exit__(self, *a):
        scipy.stats.multivariate_normal.rvs = self.orig


class InjectUniforms(object):
    """feed Slice._sample a fixed stream: expon.rvs(1) -> 1 - log(u0); uniform.rvs(0,1) -> u1, u2, ..."""

    def __init__(self, us):
        self.us = list(us)
        self.i = 0

    def __enter__(self):
        self.oe, self.ou = scipy.stats.expon.rvs, scipy.stats.uniform.rvs

        def ex(loc=0, *a, **k):
            u = self.us[self.i]
            self.i += 1
            return loc - np.log(u)

        def un(loc=0, scale=1, *a, **k):
            u = self.us[self.i]
            self.i += 1
            return loc + scale * u
        scipy.stats.expon.rvs = ex
        scipy.stats.uniform.rvs = un
        return self

    def __exit__(self, *a):
        scipy.stats.expon.rvs, scipy.stats.uniform.rvs = self.oe, self.ou


def set_state(m, F, hyper):
    """write function values / hyper-parameters into the reference's parameter_cache"""
    for f in range(m.f):
        m.parameter_cache[m.functionIndex(f)] = F[:, f]
    m.parameter_cache['y_sigma'] = hyper[0]
    for g in range(len(m.priorGroups())):
        m.parameter_cache['prior%d_sigma' % g] = hyper[1 + 2 * g]
        m.parameter_cache['prior%d_lengthscale' % g] = hyper[2 + 2 * g]


def stage_case(gp, name, x, y, effect, F, hyper, cands, **kw):
    m = gp.fanova.FANOVA(x, y, effect, **kw)
    set_state(m, F, hyper)
    out = dict(x=x, y=y, effect=effect, F=F, hyper=np.asarray(hyper, float),
               design_matrix=m.design_matrix,
               helmertConvert=np.array(bool(kw.get('helmertConvert', False))),
               interactions=np.array(bool(kw.get('interactions', False))),
               groups_flat=np.concatenate([np.asarray(g) for g in m.priorGroups()]),
               groups_len=np.array([len(g) for g in m.priorGroups()]))
    P = len(m.priorGroups())
    for g in range(P):
        out['K_%d' % g] = m.kernels[g].K(m.x)
        out['Kinv_%d' % g] = m.kernels[g].K_inv(m.x)
    out['yinv00'] = m.y_k.K_inv(m.x)[0, 0]
    for f in range(m.f):
        r = m.functionResidual(f)
        w = np.power(m.design_matrix[:, f], 2)
        w = w[w != 0]
        miss = np.isnan(r)
        out['m_t_%d' % f] = np.nansum((r.T * w).T, 0)
        out['n_t_%d' % f] = np.sum(((~miss).T * w).T, 0)
        with Capture() as cap:
            m.samplers[[s.name for s in m.samplers].index(m.functionNames().get(f, 'f%d' % f))]._sample()
        mu, cov = cap.calls[0]
        out['mu_%d' % f] = mu
        out['cov_%d' % f] = cov
    # log densities at candidate points
    out['cands'] = np.asarray(cands, float)
    out['obs_ll'] = np.array([m.observationLikelihood(sigma=c, prior_lb=-2, prior_ub=2) for c in cands])
    for g in range(P):
        out['prior_ll_sigma_%d' % g] = np.array([m.prior_likelihood(p=g, sigma=c, prior_lb=-2, prior_ub=2) for c in cands])
        out['prior_ll_ls_%d' % g] = np.array([m.prior_likelihood(p=g, lengthscale=c, prior_lb=-2, prior_ub=2) for c in cands])
    # slice trajectories with an injected stream (one per hyper-parameter, 3 stream seeds)
    names = ['y_sigma'] + [s for g in range(P) for s in ('prior%d_sigma' % g, 'prior%d_lengthscale' % g)]
    rng = np.random.default_rng(99)
    us = rng.uniform(size=(3, len(names), 64))
    res = np.zeros((3, len(names)))
    used = np.zeros((3, len(names)), dtype=np.int64)
    for t in range(3):
        for h, nm in enumerate(names):
            s = m.sampler_dict[nm]
            with InjectUniforms(us[t, h]) as inj:
                res[t, h] = s._sample(float(m.parameter_cache[nm]))
                used[t, h] = inj.i
    out['slice_u'] = us
    out['slice_x1'] = res
    out['slice_used'] = used
    # transforms (fanova.py:58-124): run every Transform sampler on the current cache
    tnames, tvals = [], []
    for s in m.samplers:
        if s.type == 'Transform':
            v = np.asarray(s._sample())
            m.parameter_cache[s.parameters] = v
            tnames.append(s.name)
            tvals.append(v)
    if tvals:
        out['transform_names'] = np.array(tnames)
        out['transform_vals'] = np.array(tvals)
    if kw.get('derivatives', False):
        # FunctionDerivative (sample/function.py:43-66): the (mu, cov) of every derivative sampler, and the derivative kernels
        for g in range(P):
            out['dK_%d' % g] = m.kernels[g].dK(m.x)
            out['dKcross_%d' % g] = m.kernels[g].dK(m.x, cross=True)
        names_ = m.functionNames()
        for f in range(m.f):
            s_ = m.sampler_dict['d' + names_.get(f, 'f%d' % f)]
            mu_, cov_ = s_._params()
            out['dmu_%d' % f] = np.asarray(mu_)[:, 0]
            out['dcov_%d' % f] = np.asarray(cov_)
    out['sampler_names'] = np.array([s.name for s in m.samplers])
    out['sampler_types'] = np.array([s.type for s in m.samplers])
    out['index'] = np.array(list(m.parameter_cache.index))
    np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)
    print('wrote', name, 'f=%d P=%d n=%d m=%d' % (m.f, P, m.n, m.m))
    return m


def prior_data(gp, x, effect, seed, hyper, **kw):
    """y = samplePrior() (base.py:245-259) of the reference at the given hyper-parameters"""
    n = x.shape[0]
    m = gp.fanova.FANOVA(x, np.zeros((n, effect.shape[0])), effect, **kw)
    set_state(m, np.zeros((n, m.f)), hyper)
    np.random.seed(seed)
    y, fs = m.samplePrior()
    return y, fs


def main():
    build_ref.build(quiet=True)
    gp = build_ref.import_ref()
    os.makedirs(GOLD, exist_ok=True)
    cands = [-2.5, -2.0, -1.3, -0.5, 0.0, 0.25, 0.5, 0.7, 1.9, 2.0, 2.5]
    only = set(sys.argv[1:])          # optional: names of the cases to (re)mint; default all
    global stage_case
    stage_all = stage_case

    def stage_case(gp_, name, *a, **k):
        if only and name not in only:
            return None
        return stage_all(gp_, name, *a, **k)

    # KAT-0: n=5 one-way, closed-form data
    n = 5
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array([0, 1, 2] * 2)[:, None]
    y = np.zeros((n, 6))
    for o in range(6):
        y[:, o] = np.sin(3 * x[:, 0]) + 0.1 * o
    F = np.zeros((n, 3))
    F[:, 0] = np.sin(3 * x[:, 0])
    stage_case(gp, 'kat0_zeroF', x, y, effect, np.zeros((n, 3)), [-2, 0, 0, 0, 0], cands)
    stage_case(gp, 'kat0_sinF', x, y, effect, F, [-2, 0, 0, 0, 0], cands)

    # cfg-1 shape: one-way, 3 levels x 5 reps, n=50  (test_convergence.py:56-62 with k=3)
    n = 50
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(range(3)) * 5)[:, None]
    hyper_true = [-2, 0, 0, 0, 0]
    y, fs = prior_data(gp, x, effect, 123, hyper_true)
    rng = np.random.default_rng(5)
    F = fs + 0.05 * rng.standard_normal(fs.shape)
    hyper = [-1.7, 0.3, -0.2, -0.4, 0.1]
    stage_case(gp, 'oneway_n50', x, y, effect, F, hyper, cands)

    # cfg-3 shape: missing time rows {10,35} + 5% scattered NaNs (rng seed 7)
    ym = y.copy()
    ym[[10, 35], :] = np.nan
    r7 = np.random.default_rng(7)
    ym[r7.uniform(size=ym.shape) < 0.05] = np.nan
    stage_case(gp, 'oneway_n50_missing', x, ym, effect, F, hyper, cands)

    # cfg-2 shape (small): two-way 3x3 + interaction, r=2, n=24
    n = 24
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(itertools.product(range(3), range(3))) * 2)
    hyper_true = [-2] + [0] * 8
    y, fs = prior_data(gp, x, effect, 123, hyper_true, interactions=True)
    F = fs + 0.05 * np.random.default_rng(6).standard_normal(fs.shape)
    hyper = [-1.5, 0.2, -0.3, -0.1, 0.4, 0.3, -0.5, -0.6, 0.2]
    stage_case(gp, 'twoway_n24', x, y, effect, F, hyper, cands, interactions=True)
    stage_case(gp, 'twoway_n24_hc', x, y, effect, F, hyper, cands, interactions=True, helmertConvert=True)

    # unbalanced two-way without interaction, 2-d input x (p=2), n=12
    n = 12
    x = np.random.default_rng(8).uniform(-1, 1, size=(n, 2))
    effect = np.array([[0, 0], [0, 1], [1, 0], [1, 1], [2, 0], [2, 1], [0, 0], [1, 1], [2, 1], [0, 1], [0, 0]])
    hyper_true = [-2] + [0] * 6
    y, fs = prior_data(gp, x, effect, 11, hyper_true)
    F = fs + 0.1 * np.random.default_rng(9).standard_normal(fs.shape)
    hyper = [-1.0, 0.1, 0.2, -0.3, 0.3, 0.2, -0.1]
    stage_case(gp, 'unbalanced_p2_n12', x, y, effect, F, hyper, cands)

    # cfg-4 shape (small): 5x5 levels + interaction, one replicate, helmertConvert=True as analysis/cokol/cokol.py:74
    # (f = 25 functions, interaction group of 16 functions sharing one conditional precision), n = 20
    n = 20
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(itertools.product(range(5), range(5))))
    hyper_true = [-2] + [0] * 8
    y, fs = prior_data(gp, x, effect, 321, hyper_true, interactions=True, helmertConvert=True)
    F = fs + 0.05 * np.random.default_rng(12).standard_normal(fs.shape)
    hyper = [-1.8, 0.1, -0.2, 0.3, 0.2, -0.2, -0.4, -0.3, 0.1]
    stage_case(gp, 'cokol5x5_hc_n20', x, y, effect, F, hyper, cands, interactions=True, helmertConvert=True)

    # cfg-1 shape on a uniform grid with an even number >= 64 of time points: the centrosymmetric split is active
    n = 64
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(range(3)) * 5)[:, None]
    y, fs = prior_data(gp, x, effect, 77, [-2, 0, 0, 0, 0])
    F = fs + 0.05 * np.random.default_rng(13).standard_normal(fs.shape)
    stage_case(gp, 'oneway_n64', x, y, effect, F, [-1.6, 0.2, -0.1, -0.3, 0.15], cands)

    # derivatives=True (base.py:72-73): one-way, n = 20, lengthscales away from 1 (the reference's dK scaling shows)
    n = 20
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(range(3)) * 3)[:, None]
    y, fs = prior_data(gp, x, effect, 55, [-2, 0, 0, 0, 0])
    F = fs + 0.02 * np.random.default_rng(14).standard_normal(fs.shape)
    stage_case(gp, 'oneway_n20_deriv', x, y, effect, F, [-1.5, 0.1, 0.15, -0.2, -0.1], cands, derivatives=True)
    # lengthscale 1: the only case in which the reference's derivative kernels are a consistent joint covariance
    stage_case(gp, 'oneway_n20_deriv_l1', x, y, effect, F, [-1.5, 0.1, 0.0, -0.2, 0.0], cands, derivatives=True)

    if only and 'jitchol' not in only:
        return
    # jitchol (linalg.py:35-54): PD, PSD-needs-jitter, and failing inputs
    rng = np.random.default_rng(3)
    B = rng.standard_normal((16, 16))
    A_pd = B @ B.T + 16 * np.eye(16)
    v = rng.standard_normal((16, 3))
    A_psd = v @ v.T                       # rank 3: dpotrf fails, jitter rescues
    A_neg = A_pd.copy()
    A_neg[4, 4] = -1.0                    # non-positive diagonal -> LinAlgError at once
    out = dict(A_pd=A_pd, A_psd=A_psd, A_neg=A_neg)
    out['L_pd'] = gp.linalg.jitchol(A_pd)
    out['L_psd'] = gp.linalg.jitchol(A_psd)
    try:
        gp.linalg.jitchol(A_neg)
        out['neg_raises'] = np.array(False)
    except np.linalg.LinAlgError:
        out['neg_raises'] = np.array(True)
    np.savez_compressed(os.path.join(GOLD, 'jitchol.npz'), **out)
    print('wrote jitchol')


if __name__ == '__main__':
    main()
