"""GPU tests of the operators beside the sweep (SURVEY.md 8f rows N1, N3, N4 and the prior module): derivative kernels
and the FunctionDerivative sampler, prior draws / log-likelihoods, batched samplePrior, transforms for every chain,
per-chain history and CSV, intervals and cross-chain diagnostics -- each against the oracle's restatement of the reference
code (pinned to reference goldens / the live reference in tests/test_oracle.py) or against the goldens directly."""
import itertools
import os

import numpy as np
import pytest

import gpfanova_oracle as orc
from helpers import GOLDEN, load_case, relerr

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpfanova_b200 import engine
    return engine


# ------------------------------------------------------------------------------------------------ kernels
def test_rbf_dist_and_derivative_kernels(eng):
    from gpfanova_b200 import kernel
    c = load_case("oneway_n20_deriv")
    x = c["x"]
    for g in range(2):
        ls, ll = c["hyper"][1 + 2 * g], c["hyper"][2 + 2 * g]
        assert relerr(eng.rbf_dK(x, ls, ll)[0].cpu().numpy(), c["dK_%d" % g]) < 1e-13
        assert relerr(eng.rbf_dK(x, ls, ll, cross=True)[0].cpu().numpy(), c["dKcross_%d" % g]) < 1e-13
    rng = np.random.default_rng(0)
    for n, p in ((9, 1), (40, 3), (256, 1)):
        xx = rng.uniform(-1, 1, size=(n, p))
        for ls in (0.01, 0.5, 30.0):
            ref = orc.rbf_dist(xx, ls)
            got = kernel.RBF.dist(xx, ls)
            # the -2XX^T + |x|^2 + |x'|^2 expansion cancels: absolute rounding of r2 ~ eps |x / l|^2, hence of the distance
            assert np.max(np.abs(got - ref)) <= 1e-7 * max(1.0, np.max(ref))
            assert np.isfinite(got).all()          # (distances stay finite where K underflows: 200 scaled units at l = 0.01)


def test_function_derivative_matches_reference_goldens(eng):
    """mu = K_ba K_a^-1 beta and cov = K_b - K_ba K_a^-1 K_ba^T (function.py:51-60) against the reference's own values;
    both are limited by its K_inv (cond 1e11): agreement to 2e-5 of scale, like the oracle's own restatement.
    At lengthscale > 1 the reference's derivative kernels do not form a covariance (cov is indefinite; the reference draws
    through numpy's SVD map, i.e. |eigenvalues|): the moments are still the reference's, the Cholesky draw is refused."""
    for name in ("oneway_n20_deriv", "oneway_n20_deriv_l1"):
        c = load_case(name)
        x, n = c["x"], c["x"].shape[0]
        for g, members in enumerate(c["groups"]):
            F = c["F"][:, members].T[None]
            dF, mu, cov, st = eng.function_derivative(x, [c["hyper"][1 + 2 * g]], [c["hyper"][2 + 2 * g]], F,
                                                      z=np.zeros((1, len(members), n)), want_moments=True, raise_on_error=False)
            dF, mu, cov, st = dF.cpu().numpy(), mu.cpu().numpy(), cov.cpu().numpy(), int(st[0])
            scale = np.max(np.abs(c["dK_%d" % g]))
            for j, f in enumerate(members):
                assert relerr(mu[0, j], c["dmu_%d" % f]) < 2e-5
                assert np.max(np.abs(cov[0] - c["dcov_%d" % f])) < 2e-5 * scale
            # with the reference's scaling cov = (l - l^2) K_b,true + l^2 cov_true: indefinite exactly when lengthscale > 1
            if c["hyper"][2 + 2 * g] > 0:
                assert np.linalg.eigvalsh(c["dcov_%d" % members[0]]).min() < -1e-3 * scale     # the reference's cov IS indefinite
                assert st < 0 and np.isnan(dF).all()
                with pytest.raises(np.linalg.LinAlgError):
                    eng.function_derivative(x, [c["hyper"][1 + 2 * g]], [c["hyper"][2 + 2 * g]], F)
            elif st >= 0:
                assert np.array_equal(dF, mu)          # z = 0: the draw is the mean
            if c["hyper"][2 + 2 * g] < 0:
                assert st >= 0                         # a proper covariance: the Cholesky draw exists


def test_model_with_derivatives_has_the_reference_schema_and_samples():
    from gpfanova_b200 import fanova
    c = load_case("oneway_n20_deriv_l1")
    m = fanova.FANOVA(c["x"], c["y"], c["effect"], derivatives=True, chains=2, seed=3)
    assert [s.name for s in m.samplers] == [str(s) for s in c["sampler_names"]]
    assert [s.type for s in m.samplers] == [str(s) for s in c["sampler_types"]]
    assert list(m.parameter_cache.index) == [str(s) for s in c["index"]]
    for h, nm in enumerate(m.hyperNames()):
        m.parameter_cache[nm] = c["hyper"][h]
    for f in range(m.f):
        m.parameter_cache[m.functionIndex(f)] = c["F"][:, f]
    ds = m.sampler_dict["dalpha*_0"]
    mu, cov = ds._params()
    assert relerr(mu[:, 0], c["dmu_1"]) < 2e-5
    # sampling: the derivative columns of the stored rows are filled (NaN only where the kernel at that state has a
    # lengthscale that makes the reference's covariance indefinite -- see DESIGN.md)
    m.sample(2)
    cols = m.functionIndex(1, derivative=True)
    assert m.parameter_history.shape == (2, len(c["index"])) and set(cols) <= set(m.parameter_history.columns)
    assert m.chain_history(1).shape == (2, len(c["index"]))


# ------------------------------------------------------------------------------------------------ prior module
def test_gp_loglik_and_prior_class(eng):
    from gpfanova_b200 import prior, kernel
    import pandas as pd
    rng = np.random.default_rng(1)
    for n, nf in ((30, 3), (100, 40), (256, 2)):
        x = np.sort(rng.uniform(-1, 1, size=(n, 1)), axis=0)
        B = 5
        ls, ll = rng.uniform(-0.5, 0.5, B), rng.uniform(-0.8, 0.2, B)
        L0 = np.linalg.cholesky(orc.rbf_K(x, 1.0, 0.5) + 1e-6 * np.eye(n))
        F = np.stack([(L0 @ rng.standard_normal((n, nf))).T for _ in range(B)])
        got = eng.gp_loglik(x, ls, ll, F, add_rel=1e-6).cpu().numpy()
        for b in range(B):
            K = orc.rbf_K(x, 10.0 ** ls[b], 10.0 ** ll[b])
            cov = K + K.mean() * 1e-6 * np.eye(n)
            ref = orc.mvn_logpdf_chol(F[b], cov).sum()
            cond = np.linalg.cond(cov)
            assert abs(got[b] - ref) <= max(5e-9, 16 * np.finfo(float).eps * cond) * max(1.0, abs(ref)), (n, nf, b, got[b], ref)
    # Prior.loglikelihood: cov = L L^T of jitchol(K) (prior.py:86-99) -- K alone needs the jitter ladder
    n = 24
    x = np.linspace(-1, 1, n)[:, None]

    class M(object):
        pass
    mod = M()
    mod.parameter_cache = pd.Series({"s": 0.2, "l": -0.3})
    pr = prior.Prior(x, kernel.RBF(mod, ["s", "l"], logspace=True), [0, 1])
    F = (L0[:n, :n] @ rng.standard_normal((n, 2)))
    K = orc.rbf_K(x, 10 ** 0.2, 10 ** -0.3)
    Lj, tries = orc.jitchol(K)
    ref = orc.mvn_logpdf_chol(F.T, Lj @ Lj.T).sum()
    got = pr.loglikelihood(F)
    assert tries >= 1 and abs(got - ref) <= 1e-6 * abs(ref)        # cond(L L^T) ~ 1e6 / jitter
    draws = pr.sample(size=4000, seed=11)
    assert draws.shape == (4000, n)
    emp = np.cov(draws.T)
    assert np.max(np.abs(emp - K)) < 0.12 * K.max()                 # Monte-Carlo agreement of the covariance
    assert pr.sample(seed=3).shape == (n, 2)


def test_gp_sample_injected_normals(eng):
    rng = np.random.default_rng(2)
    n = 70
    x = np.linspace(-1, 1, n)[:, None]
    ls, ll = np.array([0.1, -0.2, 0.3]), np.array([-0.4, -0.1, 0.2])
    z = rng.standard_normal((3, 11, n))
    got = eng.gp_sample(x, ls, ll, nsamp=11, z=z, add_abs=1e-6).cpu().numpy()
    for b in range(3):
        L = np.linalg.cholesky(orc.rbf_K(x, 10.0 ** ls[b], 10.0 ** ll[b]) + 1e-6 * np.eye(n))
        assert relerr(got[b], (L @ z[b].T).T) < 1e-8
    # with the ladder (K alone): z = identity returns the factor itself; L L^T = K + the ladder's jitter
    zI = np.eye(n)[None]
    L = eng.gp_sample(x, ls[:1], ll[:1], nsamp=n, z=zI).cpu().numpy()[0].T
    K = orc.rbf_K(x, 10.0 ** ls[0], 10.0 ** ll[0])
    resid = L @ L.T - K
    jit = np.diag(resid).mean()
    assert 0 < jit < 1e-2 * K.max() and np.max(np.abs(resid - jit * np.eye(n))) < 1e-9 * K.max()


def test_sample_prior_on_device():
    from gpfanova_b200 import fanova, synthetic
    n = 48
    x = np.linspace(-1, 1, n)[:, None]
    eff = synthetic.effects((3, 2), 4)
    m = fanova.FANOVA(x, np.zeros((n, eff.shape[0])), eff, interactions=True, y_sigma=-2.0, seed=1)
    y, fs = m.samplePrior(seed=5)
    assert y.shape == (n, m.m) and fs.shape == (n, m.f)
    resid = y - fs @ m.design_matrix.T
    assert abs(resid.std() / 0.1 - 1.0) < 0.1                      # noise sd sqrt(10^-2)
    y2, fs2 = m.samplePrior(seed=5)
    assert np.array_equal(y, y2) and np.array_equal(fs, fs2)        # keyed stream
    Y, Fs = m.samplePrior(size=300, seed=9)
    assert Y.shape == (300, n, m.m) and Fs.shape == (300, n, m.f)
    K = orc.rbf_K(x, 1.0, 1.0)
    emp = np.einsum('bi,bj->ij', Fs[:, :, 2], Fs[:, :, 2]) / 300
    assert np.max(np.abs(emp - K)) < 0.3                            # prior covariance of one function over the batch


# ------------------------------------------------------------------------------------------------ N1: transforms at scale
def test_transforms_for_every_chain_and_per_chain_csv(tmp_path):
    from gpfanova_b200 import fanova
    import pandas as pd
    c = load_case("twoway_n24")
    C, S = 5, 4
    m = fanova.FANOVA(c["x"], c["y"], c["effect"], interactions=True, chains=C, seed=8)
    m.sample(S)
    raw = m.chain_history()
    d = orc.Design(c["effect"], interactions=True)
    for ch in range(C):
        frame = m.chain_history(ch)
        assert list(frame.columns) == [str(s) for s in c["index"]] and frame.shape[0] == S
        for r in range(S):
            F = raw[r, ch, m.H:].reshape(m.f, m.n).T
            for nm, v in orc.fanova_transforms(d, F):
                assert relerr(frame.loc[r, m.sampler_dict[nm].parameters].values.astype(float), v) < 1e-12, (ch, r, nm)
            assert np.array_equal(frame.loc[r, m.hyperNames()].values.astype(float), raw[r, ch, :m.H])
    assert np.allclose(m.chain_history(0).values.astype(float), m.parameter_history.values.astype(float), rtol=0, atol=0)
    # per-chain CSV in the reference's schema (container.py:94-95), readable by the reference's load()
    f = tmp_path / "chain3.csv"
    m.save(str(f), chain=3)
    back = pd.read_csv(str(f), index_col=None, float_precision='round_trip')
    assert list(back.columns) == [str(s) for s in c["index"]]
    assert np.allclose(back.values, m.chain_history(3).values.astype(float), rtol=1e-15, atol=0)
    m.save(str(tmp_path / "all_%d.csv"), chain='all')
    assert sorted(os.listdir(str(tmp_path)))[:5] == ["all_%d.csv" % i for i in range(5)]


def test_transform_operator_at_cokol_size(eng):
    """cfg4: 20 x 20 levels with interaction, n = 200: 441 transform rows of 200 time points (88 200 columns) per state,
    against the host einsum of the same contrast matrices"""
    from gpfanova_b200 import fanova
    n = 200
    x = np.linspace(-1, 1, n)[:, None]
    eff = np.array(list(itertools.product(range(20), range(20))))
    m = fanova.FANOVA(x, np.zeros((n, 400)), eff, interactions=True, helmertConvert=True)
    assert len(m.parameter_cache.index) == 9 + 400 * n + (20 + 1 + 20 + 1 + 400) * n
    rng = np.random.default_rng(3)
    F = rng.standard_normal((3, m.f, n))
    got = m._device_transforms(F)
    ref = m._transform_block(F)
    assert got.shape == ref.shape == (3, 442 * n)
    assert relerr(got, ref) < 1e-12


# ------------------------------------------------------------------------------------------------ N4: intervals, diagnostics
def test_function_and_scalar_intervals_match_reference_algorithm(eng):
    from gpfanova_b200 import interval
    rng = np.random.default_rng(6)
    sets = []
    for S, n, alpha in ((200, 12, .95), (333, 5, .9), (64, 30, .5), (1000, 50, .95)):
        X = np.cumsum(rng.standard_normal((S, n)), axis=1) * 0.3 + rng.standard_normal((S, 1))
        ref = orc.function_interval(X, alpha)
        iv = interval.FunctionInterval(X, alpha)
        assert iv.epsilon == ref["epsilon"] and iv.alpha == ref["alpha"] and iv.evaluations == ref["evaluations"]
        assert relerr(iv.mean, ref["mean"]) < 1e-14 and relerr(iv.std, ref["std"]) < 1e-13
        assert np.allclose(iv.lb, ref["lb"], rtol=0, atol=1e-13) and np.allclose(iv.ub, ref["ub"], rtol=0, atol=1e-13)
        inside = X[0] * 0 + ref["mean"]
        assert iv.contains(inside) and not iv.contains(inside + 100)
        assert interval.functionInterval(X, alpha=alpha)[0] == ref["epsilon"]
        sets.append(X)
    # batched: many sample sets in one launch
    Xb = np.stack([sets[0] + k for k in range(7)])
    r = interval.function_intervals(Xb, .95)
    for k in range(7):
        assert r["epsilon"][k] == pytest.approx(orc.function_interval(Xb[k], .95)["epsilon"], rel=1e-12)
    # Chen-Shao HPD threshold
    x = rng.standard_normal(777)
    pi = lambda t: -0.5 * t * t
    si = interval.ScalarInterval(x, .9, pi)
    thr, eps = orc.scalar_interval_threshold(x, .9, pi)
    assert si.epj == thr and np.array_equal(si.epsilon, np.array(eps))
    assert si.contains(0.0) and not si.contains(5.0)
    with pytest.raises(ValueError):
        interval.ScalarInterval(x, 1.5, pi)
    dens = rng.standard_normal((9, 500))
    assert np.array_equal(interval.hpd_thresholds(dens, .8), np.sort(dens, axis=1)[:, 400])


def test_chain_diagnostics(eng):
    rng = np.random.default_rng(7)
    S, C, cols = 41, 13, 300
    hist = rng.standard_normal((S, C, cols)) + rng.standard_normal((1, C, cols)) * np.linspace(0, 2, cols)[None, None]
    for split in (True, False):
        r, mu, var = (t.cpu().numpy() for t in eng.chain_diagnostics(hist, 3, 250, split=split))
        rr, rm, rv = orc.split_rhat(hist[:, :, 3:253], split=split)
        assert relerr(r, rr) < 1e-12 and relerr(mu, rm) < 1e-12 and relerr(var, rv) < 1e-12
    assert r[0] < 1.2 < r[-1]                    # well mixed columns vs columns with chain offsets


def test_model_diagnostics_and_large_history_pipeline():
    """Base.chain_diagnostics on a real run, and the pinned-staging pipeline with a result above 64 MB (page toucher
    active) against a small-chunk run of the same chains"""
    from gpfanova_b200 import fanova, synthetic
    n = 64
    x = np.linspace(-1, 1, n)[:, None]
    eff = synthetic.effects((3, 3), 2)
    m = fanova.FANOVA(x, np.zeros((n, eff.shape[0])), eff, interactions=True, chains=512, seed=4, y_sigma=-2.0)
    m.y, _ = m.samplePrior(seed=1)
    m.sample(30, thin=1)
    diag = m.chain_diagnostics()
    assert list(diag.index) == m.hyperNames() and np.isfinite(diag.values).all() and (diag["rhat"] < 3).all()
    raw = m.chain_history()
    assert raw.nbytes > (64 << 20) and np.isfinite(raw).all()
    assert not (raw[:, :, m.H:] == 0.0).any()                      # no page-toucher zeros in the result
    m2 = fanova.FANOVA(x, m.y, eff, interactions=True, chains=512, seed=4, y_sigma=-2.0)
    eng2 = m2._push_state()
    out2 = eng2.sweep_host(30, thin=1, stage_bytes=4 << 20)
    assert np.array_equal(out2, raw)


# ------------------------------------------------------------------------------------------------ statistical parity, cfg2 size
@pytest.mark.parametrize("case", ["twoway_n100", "twoway_n256"])
def test_posterior_summaries_match_reference_long_chains_cfg2(case):
    """two-way 3 x 3 with interaction, 5 replicates, n = 100 (BASELINE config 2) and n = 256 (config 5, the bench workload:
    everything on the Schur kernels): pooled posterior means of many GPU chains against 8 long chains of the unmodified
    reference (tests/golden/posterior_<case>.npz, oracle/make_posterior_summary.py)"""
    path = os.path.join(GOLDEN, "posterior_%s.npz" % case)
    if not os.path.exists(path):
        pytest.skip("golden not minted")
    from gpfanova_b200 import fanova
    g = dict(np.load(path))
    x, y, eff = g["x"], g["y"], g["effect"]
    C, burn, S = 128, 150, 300
    m = fanova.FANOVA(x, y, eff, interactions=True, chains=C, seed=2027, y_sigma=-2.0)
    m.sample(burn, thin=burn)
    m.sample(S, thin=2)
    raw = m.chain_history()[1:]
    H = m.H
    hyper = raw[:, :, :H]
    F = raw[:, :, H:].reshape(raw.shape[0], C, m.f, m.n)
    ref_h, ref_F = g["hyper_chain_means"], g["F_chain_means"]
    R = ref_h.shape[0]
    ref_h_mean, ref_h_se = ref_h.mean(0), ref_h.std(0, ddof=1) / np.sqrt(R)
    ref_F_mean, ref_F_se = ref_F.mean(0), ref_F.std(0, ddof=1) / np.sqrt(R)
    post_sd_h = np.sqrt(g["hyper_chain_vars"].mean(0))
    post_sd_F = np.sqrt(g["F_chain_vars"].mean(0))
    gpu_h_chain = hyper.mean(0)
    gpu_h_mean, gpu_h_se = gpu_h_chain.mean(0), gpu_h_chain.std(0, ddof=1) / np.sqrt(C)
    gpu_F_chain = F.mean(0)
    gpu_F_mean, gpu_F_se = gpu_F_chain.mean(0), gpu_F_chain.std(0, ddof=1) / np.sqrt(C)
    # the reference chains are short (500 sweeps, autocorrelated): its standard errors are themselves noisy, hence the
    # 10%-of-posterior-sd floor
    zh = np.abs(gpu_h_mean - ref_h_mean) / np.sqrt(ref_h_se ** 2 + gpu_h_se ** 2 + (0.1 * post_sd_h) ** 2)
    assert zh.max() < 5.0, (zh, gpu_h_mean, ref_h_mean)
    zF = np.abs(gpu_F_mean - ref_F_mean) / np.sqrt(ref_F_se ** 2 + gpu_F_se ** 2 + (0.1 * post_sd_F) ** 2)
    assert zF.max() < 5.0, (zF.max(), np.unravel_index(zF.argmax(), zF.shape))
    assert np.sqrt(np.mean((gpu_F_mean - g["F_true"].T) ** 2)) < np.sqrt(np.mean((ref_F_mean - g["F_true"].T) ** 2)) + 0.02


def test_prepare_keeps_results_identical_and_is_optional():
    """FANOVA.prepare(n, thin) (additive API): page-locked result memory for the next sample() call.  Same seed, same data:
    the stored histories with and without it are equal bit for bit; the second sample() call (nothing prepared) falls back to
    the pageable path."""
    import itertools
    from gpfanova_b200 import fanova
    n = 64
    x = np.linspace(-1, 1, n)[:, None]
    eff = np.array(list(itertools.product(range(2), range(2))) * 2)
    rng = np.random.default_rng(4)
    y = np.sin(3 * x) + 0.1 * rng.standard_normal((n, eff.shape[0]))
    hist = []
    for prepared in (False, True):
        m = fanova.FANOVA(x, y, eff, interactions=True, chains=6, seed=12, device=0, store_chains=True)
        if prepared:
            assert m.prepare(8, thin=2) == 4 * 6 * (m.H + m.f * n) * 8
        m.sample(8, thin=2)
        m.sample(3, thin=1)
        hist.append((m.chain_history(), m.parameter_history.to_numpy(dtype=float)))
    assert hist[0][0].shape[0] == 7 and np.array_equal(hist[0][0], hist[1][0]) and np.array_equal(hist[0][1], hist[1][1])
