import sys, numpy as np
sys.path.insert(0, '.')
import bench
from gpfanova_b200 import engine, fanova, synthetic
x, eff, inter, hc, C = bench.problem(256)
model = fanova.FANOVA(x, np.zeros((256, eff.shape[0])), eff, interactions=inter, helmertConvert=hc, chains=C, seed=1, device=0)
y, _ = synthetic.prior_draw(x, model.design_matrix, model.priorGroups())
groups = [list(g) for g in model.priorGroups()]
eng = engine.ChainEngine(x, y, model.design_matrix, groups, chains=C, seed=1, device=0)
eng.set_state(hyper=synthetic.chain_starts(C, len(groups)))
eng.sweep(5)
for opt in (0, 1):
    eng.set_option(3, opt)
    eng.set_profile(1)
    eng.sweep(4); eng.sync()
    prof, _ = eng.profile()
    print("centro option", opt, {k: round(v / 4, 2) for k, v in prof.items()})
    eng.set_option(3, 1); eng.sweep(1)
