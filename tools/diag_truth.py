"""GPU / reference error against a 50-digit mpmath arbiter (run on the GPU box)"""
import os, sys
import numpy as np
import mpmath as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpfanova_oracle as orc
from helpers import CASES, load_case, oracle_model, relerr
from gpfanova_b200 import engine as eng
mp.mp.dps = 60

def truth(c, m, f, g):
    n = m.n
    x = c["x"]
    s = mp.mpf(10) ** mp.mpf(float(m.hyper[1 + 2 * g])); l = mp.mpf(10) ** mp.mpf(float(m.hyper[2 + 2 * g]))
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            r2 = sum(((mp.mpf(float(x[i, d])) - mp.mpf(float(x[j, d]))) / l) ** 2 for d in range(x.shape[1]))
            K[i, j] = s * mp.exp(-r2 / 2)
        K[i, i] += mp.mpf('1e-9')
    Kinv = K ** -1
    m_t, n_t = orc.function_moments(m.y, m.X, m.F, f)
    yinv = 1 / (mp.mpf(10) ** mp.mpf(float(m.hyper[0])) + mp.mpf('1e-9'))
    A = Kinv.copy()
    for t in range(n):
        A[t, t] += mp.mpf(float(n_t[t])) * yinv
    cov = A ** -1
    b = mp.matrix([mp.mpf(float(v)) * yinv for v in m_t])
    mu = cov * b
    tonp = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
    return tonp(Kinv), tonp(mu).ravel(), tonp(cov)

for name, fs in [("oneway_n50", [0, 1, 2]), ("twoway_n24_hc", [0, 3, 5]), ("oneway_n50_missing", [1])]:
    c = load_case(name); m = oracle_model(c)
    for f in fs:
        g = [i for i, grp in enumerate(m.groups) if f in grp][0]
        Kt, mut, covt = truth(c, m, f, g)
        e = eng.ChainEngine(c["x"], c["y"], c["design_matrix"], c["groups"], chains=1); e.set_state(F=c["F"].T, hyper=c["hyper"])
        Kd = e.group_kinv(g).cpu().numpy()[0]
        out = e.function_step(f, z=np.zeros((1, m.n)))
        LA = out["LA"].cpu().numpy()[0]; mu = out["mu"].cpu().numpy()[0]
        cov = np.linalg.inv(LA @ LA.T)
        K = orc.rbf_K(m.x, 10.0 ** m.hyper[1 + 2 * g], 10.0 ** m.hyper[2 + 2 * g])
        Ko = orc.K_inv(K)[0]
        mt, nt = orc.function_moments(m.y, m.X, m.F, f)
        so = orc.function_conditional(mt, nt, m.hyper[0], Ko)
        print(name, f, "Kinv err: gpu %.1e oracle %.1e ref %.1e | mu err: gpu %.1e oracle %.1e ref %.1e | cov err: gpu %.1e oracle %.1e ref %.1e" % (
            relerr(Kd, Kt), relerr(Ko, Kt), relerr(c["Kinv_%d" % g], Kt),
            relerr(mu, mut), relerr(so["mu"], mut), relerr(c["mu_%d" % f], mut),
            relerr(cov, covt), relerr(so["cov"], covt), relerr(c["cov_%d" % f], covt)))
        e.close()
