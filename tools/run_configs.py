#!/usr/bin/env python
"""Run every BASELINE.json config shape through the public API for a few sweeps (sanity + draws/s).
   python tools/run_configs.py [--sweeps 10] [--scale 1.0]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpfanova_b200 import fanova, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--sweeps", type=int, default=10)
ap.add_argument("--scale", type=float, default=1.0, help="scale the chain counts")
ap.add_argument("--only", default="")
a = ap.parse_args()
out = {}
for name, (levels, reps, n, inter, hc, chains) in synthetic.CONFIGS.items():
    if a.only and a.only not in name:
        continue
    chains = max(1, int(chains * a.scale))
    x = np.linspace(-1, 1, n)[:, None]
    eff = synthetic.effects(levels, reps)
    t0 = time.perf_counter()
    m = fanova.FANOVA(x, np.zeros((n, eff.shape[0])), eff, interactions=inter, helmertConvert=hc, chains=chains, seed=7,
                      store_chains=False)
    y, F = synthetic.prior_draw(x, m.design_matrix, m.priorGroups())
    if "missing" in name:
        y = synthetic.add_missing(y, n)
    m.y = y
    m.set_chain_state(hyper=synthetic.chain_starts(chains, len(m.priorGroups())))
    t1 = time.perf_counter()
    m.sample(2, thin=2)                       # warm-up (context creation, first kernels)
    t2 = time.perf_counter()
    m.sample(a.sweeps, thin=a.sweeps)
    t3 = time.perf_counter()
    h = m.parameter_history.values.astype(float)
    fit = np.stack([m.parameter_cache[m.functionIndex(f)].values for f in range(m.f)], 1) @ m.design_matrix.T
    obs = ~np.isnan(y)
    out[name] = dict(n=n, m=int(eff.shape[0]), f=int(m.f), P=len(m.priorGroups()), chains=chains, sweeps=a.sweeps,
                     construct_s=round(t1 - t0, 3), warm_s=round(t2 - t1, 3), ms_per_sweep=round(1e3 * (t3 - t2) / a.sweeps, 3),
                     draws_per_s=round(chains * m.f * a.sweeps / (t3 - t2), 1), finite=bool(np.isfinite(h).all()),
                     y_sigma=round(float(m.parameter_cache["y_sigma"]), 3),
                     rmse_fit=round(float(np.sqrt(np.mean((fit - y)[obs] ** 2))), 4), counters=m.counters())
    print(name, json.dumps(out[name]), flush=True)
