"""Mint tests/golden/truth.npz: 60-digit mpmath evaluations of the conditioning-limited stages (K_inv, conditional
mean/covariance, prior log-likelihood) for selected golden cases.  TEST INFRASTRUCTURE ONLY.

Why: at cond(K + 1e-9 I) ~ 1e10..1e11 the reference's own FP64 path carries 1e-10..1e-9 relative error in mu / Sigma
and up to 2e-9 in its eigh-based log-pdf.  Two correct FP64 implementations therefore differ by more
than BASELINE's 1e-10; the arbiter lets the tests require  |device - truth| <= max(1e-10, c * |reference - truth|).

    python oracle/make_truth.py        (about a minute, CPU only)
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import gpfanova_oracle as orc  # noqa: E402
from helpers import load_case, oracle_model  # noqa: E402

mp.mp.dps = 60
SEL_FUNCS = [("oneway_n50", [0, 1, 2]), ("oneway_n50_missing", [0, 1]), ("twoway_n24", [0, 3, 5]),
             ("twoway_n24_hc", [0, 3, 5, 8]), ("unbalanced_p2_n12", [1]), ("oneway_n64", [0, 2]),
             ("cokol5x5_hc_n20", [0, 2, 9, 24])]
SEL_LL = [("oneway_n50", [0, 1]), ("twoway_n24", [0, 2, 3]), ("unbalanced_p2_n12", [1]), ("oneway_n64", [1]),
          ("cokol5x5_hc_n20", [3])]


def mpf(v):
    return mp.mpf(float(v))


def K_exact(x, lsig, lls, jitter_abs=0.0):
    n = x.shape[0]
    s = mp.mpf(10) ** mpf(lsig)
    l = mp.mpf(10) ** mpf(lls)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(i + 1):
            r2 = sum(((mpf(x[i, d]) - mpf(x[j, d])) / l) ** 2 for d in range(x.shape[1]))
            K[i, j] = K[j, i] = s * mp.exp(-r2 / 2)
    return K


def tonp(M):
    return np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])


def main():
    """python oracle/make_truth.py [case ...]: with case names, only those are (re)computed and merged into the file"""
    out = {}
    only = set(sys.argv[1:])
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "truth.npz")
    if only and os.path.exists(path):
        out = dict(np.load(path))
    for name, fs in SEL_FUNCS:
        if only and name not in only:
            continue
        c = load_case(name)
        m = oracle_model(c)
        kinv_cache = {}
        for f in fs:
            g = [i for i, grp in enumerate(m.groups) if f in grp][0]
            if g not in kinv_cache:
                K = K_exact(m.x, m.hyper[1 + 2 * g], m.hyper[2 + 2 * g])
                for t in range(m.n):
                    K[t, t] += mp.mpf('1e-9')                       # kernel.pyx:7,53
                kinv_cache[g] = K ** -1
                out["%s/Kinv_%d" % (name, g)] = tonp(kinv_cache[g])
            m_t, n_t = orc.function_moments(m.y, m.X, m.F, f)
            yinv = 1 / (mp.mpf(10) ** mpf(m.hyper[0]) + mp.mpf('1e-9'))
            A = kinv_cache[g].copy()
            for t in range(m.n):
                A[t, t] += mpf(n_t[t]) * yinv                       # function.py:30
            cov = A ** -1
            b = mp.matrix([mpf(v) * yinv for v in m_t])             # function.py:32
            out["%s/mu_%d" % (name, f)] = tonp(cov * b).ravel()
            out["%s/cov_%d" % (name, f)] = tonp(cov)
            print(name, "f", f, "done", flush=True)
    for name, gs in SEL_LL:
        if only and name not in only:
            continue
        c = load_case(name)
        m = oracle_model(c)
        cands = c["cands"]
        for g in gs:
            for which in ("sigma", "lengthscale"):
                vals = []
                for cand in cands:
                    if cand < -2 or cand > 2:
                        vals.append(-np.inf)
                        continue
                    ls_, ll_ = m.hyper[1 + 2 * g], m.hyper[2 + 2 * g]
                    if which == "sigma":
                        ls_ = cand
                    else:
                        ll_ = cand
                    K = K_exact(m.x, ls_, ll_)
                    mean = sum(K[i, j] for i in range(m.n) for j in range(m.n)) / (m.n * m.n)
                    for t in range(m.n):
                        K[t, t] += mean * mp.mpf('1e-6')             # base.py:222
                    L = mp.cholesky(K)
                    logdet = 2 * sum(mp.log(L[t, t]) for t in range(m.n))
                    ll = mp.mpf(1)
                    for f in m.groups[g]:
                        v = mp.lu_solve(K, mp.matrix([mpf(q) for q in m.F[:, f]]))
                        quad = sum(mpf(m.F[t, f]) * v[t] for t in range(m.n))
                        ll += -(m.n * mp.log(2 * mp.pi) + logdet + quad) / 2
                    ll += -mp.log(4)
                    vals.append(float(ll))
                out["%s/prior_ll_%s_%d" % (name, which, g)] = np.array(vals)
                print(name, "ll", g, which, "done", flush=True)
    np.savez_compressed(path, **out)


if __name__ == "__main__":
    main()
