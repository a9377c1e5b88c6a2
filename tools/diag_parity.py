"""diagnostic dump of parity errors (run on the GPU box)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpfanova_oracle as orc
from helpers import CASES, load_case, oracle_model, relerr
from gpfanova_b200 import engine as eng
import torch

print("== kinv operator")
for n in [5, 50, 100, 256]:
    x = np.linspace(-1, 1, n)[:, None]
    hp = [(0, 0), (0.5, -0.7), (-1, 0.3), (1.0, -1.5)]
    Ks = np.stack([orc.rbf_K(x, 10.0 ** s, 10.0 ** l) for s, l in hp])
    inv = eng.kinv(Ks).cpu().numpy()
    for b in range(4):
        A = Ks[b] + 1e-9 * np.eye(n)
        ref = orc.K_inv(Ks[b])[0]
        print(n, hp[b], "cond %.2e" % np.linalg.cond(A), "r_gpu %.2e r_ref %.2e" % (np.max(np.abs(A @ inv[b] - np.eye(n))), np.max(np.abs(A @ ref - np.eye(n)))),
              "asym %.2e" % (np.abs(inv[b] - inv[b].T).max() / np.abs(inv[b]).max()), "rel %.2e" % relerr(inv[b], ref))

print("== function stages")
for name in CASES:
    c = load_case(name); m = oracle_model(c); rng = np.random.default_rng(5); C = 2
    for f in range(m.f):
        e = eng.ChainEngine(c["x"], c["y"], c["design_matrix"], c["groups"], chains=C); e.set_state(F=c["F"].T, hyper=c["hyper"])
        g = [i for i, grp in enumerate(m.groups) if f in grp][0]
        Kd = e.group_kinv(g).cpu().numpy()
        z = rng.standard_normal((C, m.n))
        out = e.function_step(f, z=z)
        st_ = e.status(raise_on_error=False)
        mt, nt = out["mt"].cpu().numpy(), out["nt"].cpu().numpy()
        st = orc.function_conditional(mt[0], nt[0], m.hyper[0], Kd[0])
        LA = out["LA"].cpu().numpy(); mu = out["mu"].cpu().numpy(); F1, _ = e.get_state()
        beta = orc.function_draw(st["mu"], st["L_A"], z[0])
        print(name, f, "status", st_, "Kinv %.1e mt %.1e nt %.1e LA %.1e mu %.1e cov %.1e beta %.1e | golden mu %.1e cov %.1e condA %.1e" % (
            relerr(Kd[0], c["Kinv_%d" % g]), relerr(mt[0], c["m_t_%d" % f]), relerr(nt[0], c["n_t_%d" % f]), relerr(LA[0], st["L_A"]), relerr(mu[0], st["mu"]),
            relerr(np.linalg.inv(LA[0] @ LA[0].T), st["cov"]), relerr(F1[0, f], beta), relerr(mu[0], c["mu_%d" % f]),
            relerr(np.linalg.inv(LA[0] @ LA[0].T), c["cov_%d" % f]), np.linalg.cond(st["A"])))
        e.close()

print("== logliks")
for name in CASES:
    c = load_case(name); m = oracle_model(c); cands = c["cands"]
    e = eng.ChainEngine(c["x"], c["y"], c["design_matrix"], c["groups"], chains=len(cands)); e.set_state(F=c["F"].T, hyper=c["hyper"])
    got = e.obs_loglik(cands)
    print(name, "obs", ["%.1e" % (abs(got[i] - r) / abs(r)) if np.isfinite(r) else str(got[i] == r) for i, r in enumerate(c["obs_ll"])])
    for g in range(m.P):
        for which, key in (("sigma", "prior_ll_sigma_%d"), ("lengthscale", "prior_ll_ls_%d")):
            got = e.prior_loglik(g, which, cands)
            line = []
            for i, cand in enumerate(cands):
                ref = c[key % g][i]
                if np.isinf(ref):
                    line.append(str(got[i] == ref)); continue
                o = orc.prior_loglik(m.x, m.F[:, m.groups[g]].T, m.hyper[1 + 2 * g], m.hyper[2 + 2 * g], cand, which)
                line.append("%.0e/%.0e" % (abs(got[i] - o) / max(1, abs(o)), abs(got[i] - ref) / max(1, abs(ref))))
            print(name, g, which, " ".join(line))
    print("status", e.status(raise_on_error=False))
    e.close()
