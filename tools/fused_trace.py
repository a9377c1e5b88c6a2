"""Phase trace of the fused function kernel on the bench workload (development aid).

    nvcc ... -DGPF_FTRACE -o build_variants/libgpf_ftrace.so   (python tools/fused_trace.py --build)
    GPFANOVA_B200_LIB=build_variants/libgpf_ftrace.so python tools/fused_trace.py [--chains 4096] [--n 256]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANT = os.path.join(ROOT, "build_variants", "libgpf_ftrace.so")


def build(extra=()):
    extra = list(extra) + [a for a in sys.argv[1:] if a.startswith("-D")]
    from gpfanova_b200 import build as b
    os.makedirs(os.path.dirname(VARIANT), exist_ok=True)
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-DGPF_FTRACE"] + list(extra) + ["-o", VARIANT, "gpf_api.cu"]
    subprocess.run(cmd, cwd=b.CSRC, check=True)
    print(VARIANT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=5)
    a, _ = ap.parse_known_args()
    if a.build:
        return build()
    import numpy as np
    from gpfanova_b200 import engine, fanova, synthetic
    levels, reps, n_, inter, hc, chains = synthetic.CONFIGS["cfg5_twoway_n256"]
    eff = synthetic.effects(levels, reps)
    x = np.linspace(-1, 1, a.n)[:, None]
    model = fanova.FANOVA(x, np.zeros((a.n, eff.shape[0])), eff, interactions=inter, helmertConvert=hc, chains=1, seed=1)
    y, _ = synthetic.prior_draw(x, model.design_matrix, model.priorGroups())
    groups = [list(g) for g in model.priorGroups()]
    eng = engine.ChainEngine(x, y, model.design_matrix, groups, chains=a.chains, seed=20261019)
    eng.set_state(hyper=synthetic.chain_starts(a.chains, len(groups)))
    eng.sweep(20)
    eng.sync()
    t = (C.c_ulonglong * 24)()
    eng._ck(eng._lib.gpf_debug_trace(eng._h, t))
    eng.set_profile(1)
    eng.sweep(a.sweeps)
    eng.sync()
    prof, _ = eng.profile()
    eng._ck(eng._lib.gpf_debug_trace(eng._h, t))
    v = [int(q) for q in t]
    items = max(1, v[7])
    names = ["table", "factor K_o", "right-hand sides", "factor C", "back substitution", "triangular products", "write-out"]
    tot = sum(v[:7])
    print("ms per sweep:", {k: round(q / a.sweeps, 3) for k, q in prof.items()})
    print("traced items %d, cycles per item %.0f" % (items, tot / items))
    for nm, q in zip(names, v[:7]):
        print("  %-20s %9.0f cycles/item  %5.1f %%" % (nm, q / items, 100.0 * q / max(1, tot)))
    calls = max(1, v[16])
    print("small engine: %d traced calls" % calls)
    for nm, q in zip(["A potrf", "A panel solve", "A tile(1,1)+rhs", "A waiting", "B panel/stores", "B updates", "B rhs", "B waiting"], v[8:16]):
        print("  %-20s %9.0f cycles/call" % (nm, q / calls))


if __name__ == "__main__":
    main()
