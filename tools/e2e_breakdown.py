"""time breakdown of the public-API path (FANOVA.sample) on cfg5"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from gpfanova_b200 import fanova, synthetic, base as B
levels, reps, n, inter, hc, chains = synthetic.CONFIGS["cfg5_twoway_n256"]
x = np.linspace(-1, 1, n)[:, None]; eff = synthetic.effects(levels, reps)
t = time.perf_counter()
m = fanova.FANOVA(x, np.zeros((n, eff.shape[0])), eff, interactions=inter, chains=chains, seed=1)
print("construct %.1f ms" % (1e3 * (time.perf_counter() - t)))
y, _ = synthetic.prior_draw(x, m.design_matrix, m.priorGroups()); m.y = y
m.set_chain_state(hyper=synthetic.chain_starts(chains, 4))
import cProfile, pstats
for rep in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    m.sample(10, thin=1)
    pr.disable(); torch.cuda.synchronize()
    print("sample(10) %.1f ms" % (1e3 * (time.perf_counter() - t)))
    if rep == 1:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
