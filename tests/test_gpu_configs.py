"""GPU parity at the shapes of BASELINE.json's configs, through the kernels those shapes launch: whole Gibbs sweeps
against the oracle driven by the same Philox stream, at n = 100 / 200 / 256 (several tile rows, both halves of the
centrosymmetric split), with classes of 1, 2, 4, 11 and 64 functions sharing one factorisation, on odd and non-uniform
grids (the non-split variants of the fused kernel), and the fused launch against the sequential sampler order."""
import itertools

import numpy as np
import pytest

import gpfanova_oracle as orc
from helpers import load_case, oracle_model, relerr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpfanova_b200 import engine
    return engine


def make_problem(levels, reps, n, hc, seed, grid="uniform"):
    """synthetic FANOVA data drawn from the model's own prior (test_convergence.py:56-62 style)"""
    rng = np.random.default_rng(seed)
    if grid == "uniform":
        x = np.linspace(-1, 1, n)[:, None]
    else:
        x = np.sort(rng.uniform(-1, 1, size=(n, 1)), axis=0)
    if len(levels) == 1:
        effect = np.array(list(range(levels[0])) * reps)[:, None]
    else:
        effect = np.array(list(itertools.product(*[range(k) for k in levels])) * reps)
    d = orc.Design(effect, interactions=len(levels) > 1, helmertConvert=hc)
    groups = d.prior_groups()
    F = np.zeros((n, d.f))
    L = np.linalg.cholesky(orc.rbf_K(x, 1.0, 0.7) + 1e-8 * np.eye(n))
    for i in range(d.f):
        F[:, i] = L @ rng.standard_normal(n)
    y = F @ d.X.T + rng.normal(0, 0.1, size=(n, d.m))
    return x, y, d, groups, F


#        name                 levels   reps  n    hc     grid          chains  sweeps  oracle chains
SHAPES = [
    ("cfg5_n256",            (3, 3),  5,    256, False, "uniform",    4,      2,      (0, 3)),      # h = 128: 4 + 4 tile rows
    ("cfg2_n100_64chains",   (3, 3),  5,    100, False, "uniform",    64,     2,      (0, 31, 63)),  # h = 50: 2 tile rows
    ("hc3x3_n200",           (3, 3),  2,    200, True,  "uniform",    3,      2,      (0, 2)),      # classes of 2, 2, 4; h = 100
    ("oneway12_hc_n100",     (12,),   2,    100, True,  "uniform",    3,      2,      (1,)),        # class of 11: 32-row passes
    ("odd_n101",             (3, 3),  2,    101, False, "uniform",    3,      2,      (0,)),        # no split: full-size fused kernel
    ("nonuniform_n72",       (3, 3),  2,    72,  False, "random",     3,      2,      (2,)),        # element-wise RBF source
    ("cokol9x9_hc_n200",     (9, 9),  1,    200, True,  "uniform",    2,      1,      (1,)),        # cfg4 shape: 64-function class
    ("threeway_n40",         (2, 2, 2), 2,  40,  False, "uniform",    4,      2,      (0, 3)),      # seven prior groups on four streams
    ("oneway4_n60",          (4,),    3,    60,  False, "uniform",    4,      2,      (0, 2)),      # a group of THREE functions in the four-function kernels
    ("oneway4_hc_n130",      (4,),    3,    130, True,  "uniform",    3,      2,      (1,)),        # ... sharing one factorisation: passes of 2 + 1
] + [   # every boundary of the staged register layout (slots per lane 1 / 2 / 4 / 8 at n <= 32 / 64 / 128 / 256; padding to 8)
    ("bnd_n%d" % nn,         (3,),    2,    nn,  False, "uniform",    2,      1,      (1,)) for nn in (8, 9, 32, 33, 64, 65, 127, 128, 129, 255)
]


# which function-step kernel a shape takes by default, and whether the dense split variant is worth a second run:
#   uniform grid, n <= 256, groups of <= 4 functions -> O(n^2) Schur kernel (option 5); option 3 = 1 -> dense fused kernel on
#   the two half-size blocks
PATHS = [(s, "default") for s in SHAPES] + [(s, "dense_split") for s in SHAPES if s[0] in ("cfg5_n256", "cfg2_n100_64chains", "hc3x3_n200")]


@pytest.mark.parametrize("shape,path", PATHS, ids=["%s-%s" % (s[0], p) for s, p in PATHS])
def test_sweeps_at_config_shapes_follow_oracle(eng, shape, path):
    name, levels, reps, n, hc, grid, C, S, och = shape
    x, y, d, groups, F = make_problem(levels, reps, n, hc, seed=len(name) + n, grid=grid)
    seed = 4242
    e = eng.ChainEngine(x, y, d.X, groups, chains=C, seed=seed, chain_offset=5)
    if path == "dense_split":
        e.set_option(3, 1)
        assert e.get_option(3) == 1 and e.get_option(5) == 0
    elif name in ("cfg5_n256", "cfg2_n100_64chains", "hc3x3_n200"):
        assert e.get_option(5) == 1 and e.get_option(3) == 0
    rng = np.random.default_rng(3)
    hyper = np.concatenate([np.full((C, 1), -2.0), rng.uniform(-0.7, 0.7, size=(C, e.H - 1))], axis=1)
    e.set_state(F=np.zeros((C, e.f, n)), hyper=hyper)
    hist = torch.zeros((S, C, e.H + e.f * n), dtype=torch.float64, device="cuda")
    assert e.sweep(S, thin=0, history=hist) == S
    e.sync()
    e.status()
    hist = hist.cpu().numpy()
    assert np.all(np.isfinite(hist))
    cnt = e.counters()
    assert cnt["draws"] == C * S * e.f and cnt["slice_steps"] == C * S * e.H
    st = orc.Stream(seed)
    for ch in och:
        m = orc.Model(x, y, d.X, groups)
        m.split = bool(e.get_option(3))
        m.hyper = hyper[ch].copy()
        for s in range(S):
            m.sweep(st, ch + 5, s)
            assert np.allclose(hist[s, ch, :e.H], m.hyper, rtol=0, atol=1e-6), (name, ch, s, hist[s, ch, :e.H], m.hyper)
            assert relerr(hist[s, ch, e.H:].reshape(e.f, n), m.F.T) < 1e-6, (name, ch, s)
    e.close()


SEQ = [(s, p) for s, p in PATHS if s[0] != "cfg2_n100_64chains"]


@pytest.mark.parametrize("shape,path", SEQ, ids=["%s-%s" % (s[0], p) for s, p in SEQ])
def test_fused_launch_equals_sequential_order_at_size(eng, shape, path):
    """gpf_set_option(1, 0): one launch per sampler with the general kernel (own factorisation of C per function, factor of
    K_o from HBM); default: one fused launch per sweep (the O(n^2) Schur kernel where it applies, else the dense fused
    kernel).  Same Philox stream, same square root of K_o, same draws to rounding."""
    name, levels, reps, n, hc, grid, C, S, och = shape
    x, y, d, groups, F = make_problem(levels, reps, n, hc, seed=7 + n, grid=grid)
    C = min(C, 3)
    rng = np.random.default_rng(8)
    outs, fchol = [], []
    for batch, share in ((1, 1), (1, 0), (0, 0)):
        e = eng.ChainEngine(x, y, d.X, groups, chains=C, seed=99)
        if path == "dense_split":
            e.set_option(3, 1)
        schur_path = bool(e.get_option(5))
        e.set_option(1, batch)
        e.set_option(2, share)
        hyper = np.concatenate([np.full((C, 1), -2.0), np.random.default_rng(8).uniform(-0.7, 0.7, size=(C, e.H - 1))], axis=1)
        e.set_state(F=F.T, hyper=hyper)
        hist = torch.zeros((1, C, e.H + e.f * n), dtype=torch.float64, device="cuda")
        e.sweep(1, history=hist)
        e.sync()
        e.status()
        outs.append(hist.cpu().numpy())
        fchol.append(e.counters()["function_chol"])
        e.close()
    scale = np.max(np.abs(outs[2]))
    for o in outs[:2]:
        assert np.max(np.abs(o[..., :1 + 2 * len(groups)] - outs[2][..., :1 + 2 * len(groups)])) < 1e-7
        assert np.max(np.abs(o - outs[2])) < 1e-7 * scale, (name, np.max(np.abs(o - outs[2])))
    assert fchol[2] == C * d.f and fchol[1] == C * d.f and fchol[0] <= fchol[1]
    if hc:
        per = 2 if schur_path else 8       # functions per pass: Schur kernel 2 (registers), dense fused kernel 8 (rows riding along)
        assert fchol[0] == C * sum(-(-len(g) // per) for g in groups)      # one factorisation of C per class and pass


@pytest.mark.parametrize("name", ["cokol5x5_hc_n20", "oneway_n64"])
def test_new_goldens_whole_sweeps(eng, name):
    """the cfg4-shaped (5 x 5 levels, helmertConvert) and the n = 64 (split active) reference goldens: sweeps from the golden
    state follow the oracle"""
    c = load_case(name)
    C, S, seed = 3, 3, 77
    e = eng.ChainEngine(c["x"], c["y"], c["design_matrix"], c["groups"], chains=C, seed=seed)
    e.set_state(F=c["F"].T, hyper=c["hyper"])
    hist = torch.zeros((S, C, e.H + e.f * e.n), dtype=torch.float64, device="cuda")
    e.sweep(S, history=hist)
    e.sync()
    e.status()
    hist = hist.cpu().numpy()
    stream = orc.Stream(seed)
    for ch in range(C):
        m = oracle_model(c)
        m.split = bool(e.get_option(3))
        for s in range(S):
            m.sweep(stream, ch, s)
            assert np.allclose(hist[s, ch, :e.H], m.hyper, rtol=0, atol=1e-6), (name, ch, s)
            assert relerr(hist[s, ch, e.H:].reshape(e.f, e.n), m.F.T) < 1e-6, (name, ch, s)
    e.close()


def test_dense_slice_path_with_many_functions(eng):
    """cfg4 shape: the 64-function interaction group rides along the dense slice factorisation as two right-hand-side tile
    rows (gpf_set_option(0, 0): dense engine instead of the Toeplitz/Schur path) -- log-likelihoods and a sweep"""
    x, y, d, groups, F = make_problem((9, 9), 1, 200, True, seed=21)
    C = 4
    rng = np.random.default_rng(4)
    hyper = np.concatenate([np.full((C, 1), -2.0), rng.uniform(-0.6, 0.6, size=(C, 1 + 2 * len(groups) - 1))], axis=1)
    cand = rng.uniform(-1.0, 1.0, size=C)
    res = {}
    for opt in (1, 0):
        e = eng.ChainEngine(x, y, d.X, groups, chains=C, seed=6)
        e.set_option(0, opt)
        e.set_state(F=F.T, hyper=hyper)
        vals = {}
        for g in (1, 3):
            for which in ("sigma", "lengthscale"):
                got = e.prior_loglik(g, which, cand)
                for ch in range(C):
                    o = orc.prior_loglik(x, F[:, groups[g]].T, hyper[ch, 1 + 2 * g], hyper[ch, 2 + 2 * g], cand[ch], which)
                    # cov = K + mean(K) 1e-6 I has cond ~ 1e6 n: two FP64 evaluations of the 64 quadratic forms differ by
                    # ~ eps * cond (the dense path solves with explicit 8 x 8 block inverses: a few times that)
                    ls_, ll_ = (cand[ch], hyper[ch, 2 + 2 * g]) if which == "sigma" else (hyper[ch, 1 + 2 * g], cand[ch])
                    Kc = orc.rbf_K(x, 10.0 ** ls_, 10.0 ** ll_)
                    condc = np.linalg.cond(Kc + Kc.mean() * 1e-6 * np.eye(200))
                    tol = max(5e-9, 16 * np.finfo(float).eps * condc)
                    assert abs(got[ch] - o) <= tol * max(1.0, abs(o)), (opt, g, which, ch, got[ch], o, condc)
                vals[(g, which)] = got
        hist = torch.zeros((1, C, e.H + e.f * 200), dtype=torch.float64, device="cuda")
        e.sweep(1, history=hist)
        e.sync()
        e.status()
        res[opt] = hist.cpu().numpy()
        e.close()
    # same Philox stream, two evaluations of the same log densities: the slice decisions and hence the states agree
    assert np.max(np.abs(res[0][..., :9] - res[1][..., :9])) < 1e-6


@pytest.mark.parametrize("levels,reps,n,hc,C,S,thin", [
    ((3, 3), 5, 256, False, 37, 5, 2),        # cfg5 shape: groups of 1, 2, 2, 4 functions (one cohort: fewer than 64 chains)
    ((3, 3), 2, 100, False, 700, 4, 0),       # more units per sweep than resident warps: chains drift apart; two cohorts
    ((3,), 5, 50, False, 1, 9, 3),            # cfg1 shape: one chain, two groups -- units wait for their chain's y_sigma step
    ((3, 3), 2, 200, True, 66, 3, 1),         # helmertConvert: classes of 2, 2, 4 share a factorisation; cohorts of 36 + 30
    ((3, 3), 2, 31, False, 9, 4, 1),          # n <= 32: one slot per lane from the start
    ((2, 2, 2), 2, 40, False, 70, 3, 1),      # seven prior groups of one function: groups share the four auxiliary streams
])
def test_sweep_schedules_are_bit_identical(eng, levels, reps, n, hc, C, S, thin):
    """Four ways to schedule the same sweeps (sample/container.py:29-78) on the Schur kernels:
      per-stage kernels (gpf_set_option 8 = 0): y_sigma | all functions | the groups' slice kernels;
      group pipelines (default, 8 = 1): every group's functions -> slices on its own stream;
      ... with the fold (8 = 2): the slice kernel that finishes a chain's last group stores the row and runs the next
      y_sigma step -- no k_obs launch and no copy kernels after the first sweep of a call;
      group pipelines with the chains in two independent cohorts of streams (9 = 2);
      one launch (7 = 1, csrc/gpf_sweepk.cuh): units handed to warps by ticket, a chain's y_sigma step and history row done by
      the warp that finishes the chain's sweep -- no device-wide barrier anywhere.
    Same Philox counters, same arithmetic, same summation order: the stored rows, the final state and the counters must be
    EQUAL bit for bit."""
    x, y, d, groups, F = make_problem(levels, reps, n, hc, seed=n + C)
    outs = []
    variants = [("stages", {8: 0}), ("pipelines", {}), ("pipelines + fold", {8: 2}), ("two cohorts", {9: 2}), ("one launch", {7: 1})]
    for name, opts in variants:
        e = eng.ChainEngine(x, y, d.X, groups, chains=C, seed=77, chain_offset=3)
        for k, v in opts.items():
            e.set_option(k, v)
        assert e.get_option(5) == 1 and e.get_option(7) == opts.get(7, 0) and e.get_option(8) == opts.get(8, 1)
        hyper = np.concatenate([np.full((C, 1), -2.0), np.random.default_rng(3).uniform(-0.7, 0.7, size=(C, e.H - 1))], axis=1)
        e.set_state(F=np.zeros((C, e.f, n)), hyper=hyper)
        rows = S if thin <= 0 else S // thin
        hist = torch.full((max(rows, 1), C, e.H + e.f * n), float("nan"), dtype=torch.float64, device="cuda")
        got = e.sweep(S, thin=thin, history=hist)
        e.sync()
        e.status()
        assert got == rows
        launches = e.profile()[1]
        # a second call continues the same chains (sweep counter, state)
        e.sweep(2, thin=0, history=hist[:min(2, rows)].clone())
        e.sync()
        e.status()
        F1, h1 = e.get_state()
        outs.append((hist[:rows].cpu().numpy(), F1, h1, e.counters(), launches))
        e.close()
    ref = outs[0]
    assert np.isfinite(ref[0]).all()
    for (name, _), o in zip(variants[1:], outs[1:]):
        assert np.array_equal(o[0], ref[0]) and np.array_equal(o[1], ref[1]) and np.array_equal(o[2], ref[2]), name
        for k in ("chol", "logp", "draws", "slice_steps", "function_chol", "ko_factor", "jitter"):
            assert o[3][k] == ref[3][k], (name, k)
    assert outs[-1][4] <= 2 and outs[-1][4] < ref[4]       # one launch for the whole call
