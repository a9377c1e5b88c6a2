#!/usr/bin/env python
"""Mint tests/golden/posterior_<case>.npz: posterior summaries from LONG chains of the UNMODIFIED reference
(oracle/_ref), for the statistical-parity tests of the GPU sampler (BASELINE north_star: "posterior summaries over
long chains must agree statistically").  TEST INFRASTRUCTURE ONLY.

    python oracle/make_posterior_summary.py [--case oneway_n16] [--chains 8] [--sweeps 4000] [--burn 300]   (~10 min, 8 cores)
    python oracle/make_posterior_summary.py --case twoway_n100 --chains 8 --sweeps 500 --burn 100            (~15 min, 8 cores)
    python oracle/make_posterior_summary.py --case twoway_n256 --chains 8 --sweeps 500 --burn 100            (~10 min, 8 cores)

Cases: oneway_n16 (one-way, 3 levels x 3 replicates, n = 16); twoway_n100 (BASELINE config 2: 3 x 3 levels with
interaction, 5 replicates, n = 100 time points); twoway_n256 (BASELINE config 5, the bench workload: the same design at
n = 256 time points).
"""
import argparse
import multiprocessing as mp
import os
import sys

for _v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):      # one chain per core: BLAS threads set before
    os.environ.setdefault(_v, "1")                                            # numpy loads (8 threads x 8 chains is 5x slower)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def problem(case="oneway_n16"):
    from gpfanova_b200 import synthetic
    import gpfanova_oracle as orc
    if case == "oneway_n16":
        n, eff, inter = 16, synthetic.effects((3,), 3), False
    elif case == "twoway_n100":
        n, eff, inter = 100, synthetic.effects((3, 3), 5), True
    elif case == "twoway_n256":
        n, eff, inter = 256, synthetic.effects((3, 3), 5), True
    else:
        raise ValueError(case)
    x = np.linspace(-1, 1, n)[:, None]
    d = orc.Design(eff, interactions=inter)
    y, F = synthetic.prior_draw(x, d.X, d.prior_groups(), y_sigma=-2.0, seed=321)
    return x, eff, y, F, d, inter


def chain(args):
    c, sweeps, burn, case = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    import logging
    logging.disable(logging.CRITICAL)
    import build_ref
    gp = build_ref.import_ref()
    x, eff, y, F, d, inter = problem(case)
    np.random.seed(5000 + c)
    m = gp.fanova.FANOVA(x, y, eff, interactions=inter)
    m.parameter_cache['y_sigma'] = -2.0
    m.sample(burn, thin=burn)                       # burn-in (one stored row, dropped)
    m.parameter_history = m.parameter_history.iloc[0:0]
    m.sample(sweeps, thin=1)
    H = m.parameter_history
    names = ['y_sigma'] + [s % g for g in range(len(m.priorGroups())) for s in ('prior%d_sigma', 'prior%d_lengthscale')]
    hyper = H[names].values.astype(float)
    Fs = np.stack([H[m.functionIndex(f)].values.astype(float) for f in range(m.f)], axis=1)   # (S, f, n)
    return hyper.mean(0), hyper.var(0), Fs.mean(0), Fs.var(0), np.median(hyper, axis=0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--sweeps", type=int, default=4000)
    ap.add_argument("--burn", type=int, default=300)
    ap.add_argument("--case", default="oneway_n16")
    a = ap.parse_args()
    import build_ref
    if os.path.isdir("/root/reference/gpfanova"):
        build_ref.build(quiet=True)
    with mp.Pool(min(a.chains, os.cpu_count())) as pool:
        res = pool.map(chain, [(c, a.sweeps, a.burn, a.case) for c in range(a.chains)])
    hm = np.stack([r[0] for r in res]); hv = np.stack([r[1] for r in res])
    fm = np.stack([r[2] for r in res]); fv = np.stack([r[3] for r in res])
    hmed = np.stack([r[4] for r in res])
    x, eff, y, F, d, inter = problem(a.case)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "posterior_%s.npz" % a.case), interactions=inter,
                        x=x, y=y, effect=eff, F_true=F, chains=a.chains, sweeps=a.sweeps, burn=a.burn,
                        hyper_chain_means=hm, hyper_chain_vars=hv, hyper_chain_medians=hmed,
                        F_chain_means=fm, F_chain_vars=fv)
    print("hyper mean", hm.mean(0), "between-chain sd", hm.std(0, ddof=1))
