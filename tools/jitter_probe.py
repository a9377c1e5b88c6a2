"""Does the Schur function step ever need the jitter ladder of linalg.py:41-54, and does it stay with the oracle (dense LAPACK
factorisations) at the edge of the prior box?  One sweep at extreme (sigma, lengthscale); prints retries and deviations.
Measured on B200: no retries up to sigma = lengthscale = 10^1.95, deviation <= 2.4e-7 of |F| (conditioning-limited)."""
import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import gpfanova_oracle as orc
from gpfanova_b200 import engine
from test_gpu_configs import make_problem
for n in (64, 256):
    x, y, d, groups, F = make_problem((3,), 2, n, False, seed=3)
    for ll, ls_ in ((0.8, 1.0), (1.6, 1.5), (1.95, 1.95), (1.0, 1.95), (0.0, 1.95)):
        C = 2
        e = engine.ChainEngine(x, y, d.X, groups, chains=C, seed=5)
        hyper = np.tile(np.array([-2.0, ls_, ll, ls_, ll]), (C, 1))
        e.set_state(F=np.zeros((C, e.f, n)), hyper=hyper)
        hist = torch.zeros((1, C, e.H + e.f * n), dtype=torch.float64, device="cuda")
        e.sweep(1, history=hist); e.sync()
        try:
            e.status(); st = "ok"
        except Exception as ex:
            st = "ERR " + str(ex)[:40]
        cnt = e.counters()
        h = hist.cpu().numpy()
        m = orc.Model(x, y, d.X, groups); m.split = bool(e.get_option(3)); m.hyper = hyper[0].copy()
        try:
            m.sweep(orc.Stream(5), 0, 0)
            dev = np.max(np.abs(h[0, 0, e.H:].reshape(e.f, n) - m.F.T)) / max(1.0, np.max(np.abs(m.F)))
            dh = np.max(np.abs(h[0, 0, :e.H] - m.hyper))
        except Exception as ex:
            dev, dh = str(ex)[:40], None
        print("n", n, "log10 ls", ll, "log10 sigma", ls_, st, "jitter retries", cnt["jitter"], "F dev", dev, "hyper dev", dh, flush=True)
        e.close()
