import numpy as np, sys
sys.path.insert(0, '.')
from gpfanova_b200 import engine as E
from oracle import gpfanova_oracle as O
import mpmath as mp
for n in (64, 100, 128, 256, 192):
    x = np.linspace(-1, 1, n)[:, None]
    rng = np.random.default_rng(0)
    C = 6
    y = rng.normal(size=(n, 4)); X = np.ones((4, 1))
    eng = E.ChainEngine(x, y, X, [[0]], chains=C, seed=1)
    hyper = np.zeros((C, 3)); hyper[:, 1] = rng.uniform(-1, 1, C); hyper[:, 2] = rng.uniform(-1.0, 0.3, C)
    eng.set_state(F=np.zeros((C, 1, n)), hyper=hyper)
    outs = []
    for opt in (0, 1):
        eng.set_option(3, opt)
        outs.append(eng.group_kinv(0).cpu().numpy().copy())
    for ch in range(C):
        s, l = 10 ** hyper[ch, 1], 10 ** hyper[ch, 2]
        d = x - x.T
        K = s * np.exp(-0.5 * d * d / l / l) + 1e-9 * np.eye(n)
        Ko = np.linalg.inv(K)
        mp.mp.dps = 50
        Km = mp.matrix(K.tolist()); Kt = np.array((Km ** -1).tolist(), dtype=float) if n <= 100 else Ko
        f = lambda A: np.abs(A - Kt).max() / np.abs(Kt).max()
        print(n, ch, "ls %.3f" % l, "cond %.1e" % np.linalg.cond(K), "dense %.2e centro %.2e numpy %.2e" % (f(outs[0][ch]), f(outs[1][ch]), f(Ko)),
              "resid d %.2e c %.2e" % (np.abs(outs[0][ch] @ K - np.eye(n)).max(), np.abs(outs[1][ch] @ K - np.eye(n)).max()))
    eng.close()
