"""per-kernel time of one BASELINE config shape (gpf_set_profile(1)):  python tools/cfg_profile.py cfg4 [sweeps]"""
import sys, numpy as np
sys.path.insert(0, '.')
from gpfanova_b200 import engine, fanova, synthetic
key = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
name = [k for k in synthetic.CONFIGS if key in k][0]
levels, reps, n, inter, hc, chains = synthetic.CONFIGS[name]
x = np.linspace(-1, 1, n)[:, None]
eff = synthetic.effects(levels, reps)
m = fanova.FANOVA(x, np.zeros((n, eff.shape[0])), eff, interactions=inter, helmertConvert=hc, chains=chains, seed=7, store_chains=False)
y, F = synthetic.prior_draw(x, m.design_matrix, m.priorGroups())
if "missing" in name:
    y = synthetic.add_missing(y, n)
groups = [list(g) for g in m.priorGroups()]
eng = engine.ChainEngine(x, y, m.design_matrix, groups, chains=chains, seed=7)
eng.set_state(hyper=synthetic.chain_starts(chains, len(groups)))
eng.sweep(3)
eng.set_profile(1)
eng.sweep(sweeps); eng.sync()
prof, launches = eng.profile()
print(name, "groups", [len(g) for g in groups], {k: round(v / sweeps, 2) for k, v in prof.items()}, "launches/sweep", launches)
