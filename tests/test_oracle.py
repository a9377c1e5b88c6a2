"""The oracle (oracle/gpfanova_oracle.py) against outputs of the unmodified reference
(tests/golden/*.npz minted by oracle/make_golden.py).  CPU only."""
import math
import os
import sys

import numpy as np
import pytest

import gpfanova_oracle as orc
from helpers import CASES, load_case, oracle_model, relerr


def test_philox_kat():
    # Random123 kat_vectors, philox4x32-10
    assert orc.philox4x32_10(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert orc.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff) == (
        0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert orc.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == (
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    r = orc.philox4x32_10_np(np.array([0x243f6a88, 0]), np.array([0x85a308d3, 0]), np.array([0x13198a2e, 0]),
                             np.array([0x03707344, 0]), 0xa4093822, 0x299f31d0)
    assert int(r[0][0]) == 0xd16cfe09 and int(r[3][0]) == 0x24126ea1


def test_stream_moments():
    s = orc.Stream(12345)
    z = np.concatenate([s.normals(c, 0, 1, 256) for c in range(64)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    u = np.array([s.uniform(0, 0, 0, k) for k in range(2000)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.03


def test_helmert_known():
    assert np.array_equal(orc.helmert(3), np.array([[-1., -1.], [1., -1.], [0., 2.]]))
    hc = orc.helmert_convert(orc.helmert(3))
    assert np.allclose(hc, [[-.70710678, -.40824829], [.70710678, -.40824829], [0, .81649658]], atol=1e-8)


@pytest.mark.parametrize("name", CASES)
def test_design_matrix(name):
    c = load_case(name)
    d = orc.Design(c["effect"], interactions=bool(c["interactions"]), helmertConvert=bool(c["helmertConvert"]))
    assert np.array_equal(d.X, c["design_matrix"])
    assert d.prior_groups() == c["groups"]


@pytest.mark.parametrize("name", CASES)
def test_kernel_and_inverse(name):
    c = load_case(name)
    for g in range(len(c["groups"])):
        K = orc.rbf_K(c["x"], 10.0 ** c["hyper"][1 + 2 * g], 10.0 ** c["hyper"][2 + 2 * g])
        assert relerr(K, c["K_%d" % g]) < 1e-14
        Kinv, L, nt = orc.K_inv(K)
        # K_inv is conditioning-limited (cond up to 1e11): SURVEY 7-H1 measured 1e-7..1.5e-6
        assert relerr(Kinv, c["Kinv_%d" % g]) < 5e-5
        assert nt == 0
    assert abs(1.0 / (10.0 ** c["hyper"][0] + 1e-9) - c["yinv00"]) <= 1e-12 * abs(c["yinv00"])


@pytest.mark.parametrize("name", CASES)
def test_function_conditional(name):
    c = load_case(name)
    m = oracle_model(c)
    for f in range(m.f):
        m_t, n_t = orc.function_moments(m.y, m.X, m.F, f)
        assert relerr(m_t, c["m_t_%d" % f]) < 1e-12
        assert relerr(n_t, c["n_t_%d" % f]) < 1e-14
        g = [i for i, grp in enumerate(m.groups) if f in grp][0]
        # use the reference's own K_inv so the stage is compared on identical inputs
        st = orc.function_conditional(m_t, n_t, m.hyper[0], c["Kinv_%d" % g])
        assert relerr(st["mu"], c["mu_%d" % f]) < 1e-9
        assert relerr(st["cov"], c["cov_%d" % f]) < 1e-9
        # and the whole chain (own K_inv): conditioning-limited floor, SURVEY 7-H1
        K = orc.rbf_K(m.x, 10.0 ** m.hyper[1 + 2 * g], 10.0 ** m.hyper[2 + 2 * g])
        st2 = orc.function_conditional(m_t, n_t, m.hyper[0], orc.K_inv(K)[0])
        assert relerr(st2["mu"], c["mu_%d" % f]) < 1e-7
        assert relerr(st2["cov"], c["cov_%d" % f]) < 1e-7


def test_kat0_survey_values():
    """SURVEY.md 8c KAT-0 numbers (minted from the patched reference at survey time)"""
    c = load_case("kat0_zeroF")
    assert np.allclose(c["K_0"][0], [1, 0.882496902585, 0.606530659713, 0.324652467358, 0.135335283237], atol=1e-12)
    assert np.allclose(c["Kinv_0"][0, :3], [34.448711113334, -86.876526628849, 102.810391160516], rtol=1e-6)
    assert np.allclose(c["m_t_0"], [0.653279951641, -4.484969919624, 1.5, 7.484969919624, 2.346720048359], atol=1e-10)
    assert np.allclose(c["n_t_0"], 6) and np.allclose(c["n_t_2"], 12) and np.allclose(c["m_t_2"], 0.6)
    assert np.allclose(c["mu_0"], [0.081236129778, -0.704197066152, 0.248404547476, 1.206685305079, 0.417091377291], atol=1e-9)
    assert abs(c["cov_0"][0, 1] - 1.1034581661780585e-4) < 1e-12
    assert abs(c["obs_ll"][1] - (-706.323560552463)) < 1e-8
    s = load_case("kat0_sinF")
    i0 = list(s["cands"]).index(0.0)
    i5 = list(s["cands"]).index(0.5)
    i25 = list(s["cands"]).index(2.5)
    assert abs(s["prior_ll_ls_0"][i0] - (-24.192676814878)) < 1e-7
    assert abs(s["prior_ll_sigma_0"][i5] - (-10.320841458200)) < 1e-7
    assert s["prior_ll_ls_0"][i25] == -np.inf


@pytest.mark.parametrize("name", CASES)
def test_loglik(name):
    c = load_case(name)
    m = oracle_model(c)
    for i, cand in enumerate(c["cands"]):
        ref = c["obs_ll"][i]
        got = orc.observation_loglik(m.y, m.X, m.F, cand)
        if np.isinf(ref):
            assert got == ref
        else:
            assert abs(got - ref) <= 1e-11 * abs(ref)
        for g in range(m.P):
            Fg = m.F[:, m.groups[g]].T
            for which, key in (("sigma", "prior_ll_sigma_%d"), ("lengthscale", "prior_ll_ls_%d")):
                ref = c[key % g][i]
                for method, tol in (("eigh", 1e-8), ("chol", 1e-8)):
                    got = orc.prior_loglik(m.x, Fg, m.hyper[1 + 2 * g], m.hyper[2 + 2 * g], cand, which, method=method)
                    if np.isinf(ref):
                        assert got == ref
                    else:
                        # the reference's eigh-based logpdf is itself off by up to 2e-9 rel
                        assert abs(got - ref) <= tol * max(1.0, abs(ref)), (name, g, which, cand, method, got, ref)


@pytest.mark.parametrize("name", CASES)
def test_slice_trajectories(name):
    """slice.pyx:24-62 under the same injected uniform stream as the reference run"""
    c = load_case(name)
    m = oracle_model(c)
    us = c["slice_u"]
    for t in range(us.shape[0]):
        for h in range(m.H):
            seq = us[t, h]
            x1, nev = orc.slice_step(m.hyper[h], lambda cand: m.logp_hyper(h, cand), 0.1, 10, lambda k: seq[k])
            assert abs(x1 - c["slice_x1"][t, h]) < 1e-12, (name, t, h)


@pytest.mark.parametrize("name", ["oneway_n50", "twoway_n24", "twoway_n24_hc"])
def test_transforms(name):
    c = load_case(name)
    d = orc.Design(c["effect"], interactions=bool(c["interactions"]), helmertConvert=bool(c["helmertConvert"]))
    tr = orc.fanova_transforms(d, c["F"])
    assert [t[0] for t in tr] == list(c["transform_names"])
    for (nm, v), ref in zip(tr, c["transform_vals"]):
        assert relerr(v, ref) < 1e-12, nm


def test_jitchol_golden():
    d = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "jitchol.npz")))
    L, nt = orc.jitchol(d["A_pd"])
    assert nt == 0 and relerr(L, d["L_pd"]) < 1e-13
    L, nt = orc.jitchol(d["A_psd"])
    assert nt >= 1
    # jittered factor reproduces A + jitter*I; same jitter ladder as the reference
    assert relerr(L @ L.T, d["L_psd"] @ d["L_psd"].T) < 1e-9
    with pytest.raises(np.linalg.LinAlgError):
        orc.jitchol(d["A_neg"])
    assert bool(d["neg_raises"])


def test_draw_contract_covariance():
    """beta = mu + L_A^{-T} z has covariance A^{-1} (what function.py:40 draws from)"""
    c = load_case("oneway_n50")
    m = oracle_model(c)
    m_t, n_t = orc.function_moments(m.y, m.X, m.F, 1)
    st = orc.function_conditional(m_t, n_t, m.hyper[0], c["Kinv_1"])
    import scipy.linalg
    T = scipy.linalg.solve_triangular(st["L_A"], np.eye(m.n), lower=True, trans='T')
    assert relerr(T @ T.T, st["cov"]) < 1e-9


@pytest.mark.parametrize("name", CASES)
def test_observation_form_is_the_reference_conditional(name):
    """function_conditional_obs (the device contract: no inverse of K) against the reference's mu / cov goldens: the
    difference is bounded by the reference's own K_inv rounding (4e-11..1e-9), and against the 60-digit
    arbiter it is 1e-10 or better where the reference's own path is not."""
    c = load_case(name)
    m = oracle_model(c)
    T = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "truth.npz")))
    for f in range(m.f):
        g = [i for i, grp in enumerate(m.groups) if f in grp][0]
        m_t, n_t = orc.function_moments(m.y, m.X, m.F, f)
        st = orc.function_conditional_obs(m_t, n_t, m.hyper[0], c["K_%d" % g])
        assert relerr(st["mu"], c["mu_%d" % f]) < 5e-9
        assert relerr(st["cov"], c["cov_%d" % f]) < 5e-9
        k = "%s/mu_%d" % (name, f)
        if k in T:
            assert relerr(st["mu"], T[k]) < 1e-11 and relerr(st["cov"], T["%s/cov_%d" % (name, f)]) < 1e-10


@pytest.mark.parametrize("name,centro", [("oneway_n50", False), ("oneway_n50_missing", False), ("oneway_n64", True),
                                         ("unbalanced_p2_n12", False)])
def test_observation_draw_map_has_the_conditional_moments(name, centro):
    """function_draw_obs is affine in z = [z1 | z2]: its value at z = 0 is mu and its Jacobian M satisfies
    M M^T = Sigma -- so a draw has exactly the distribution N(A^-1 b, A^-1) of function.py:38-40."""
    c = load_case(name)
    m = oracle_model(c)
    assert orc.uniform_even_grid(m.x) == centro
    f = 1
    g = [i for i, grp in enumerate(m.groups) if f in grp][0]
    K = c["K_%d" % g]
    m_t, n_t = orc.function_moments(m.y, m.X, m.F, f)
    st = orc.function_conditional_obs(m_t, n_t, m.hyper[0], K)
    SB, jit, tries = orc.prior_sqrt(K, centro)
    assert tries == 0 and relerr(SB @ SB.T, st["Ko"]) < 1e-13
    n = m.n
    mu0 = orc.function_draw_obs(st, SB, np.zeros(2 * n))
    assert relerr(mu0, st["mu"]) < 1e-10
    M = np.stack([orc.function_draw_obs(st, SB, e) - mu0 for e in np.eye(2 * n)], axis=1)
    assert relerr(M @ M.T, st["cov"]) < 1e-9
    if centro:      # the two square roots of K_o give the same law
        L = np.linalg.cholesky(st["Ko"])
        M2 = np.stack([orc.function_draw_obs(st, L, e) - mu0 for e in np.eye(2 * n)], axis=1)
        assert relerr(M2 @ M2.T, M @ M.T) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only in the build container")
def test_oracle_vs_live_reference_sweep_pieces():
    """run the live reference on a fresh random state and compare every stage"""
    import build_ref
    build_ref.build(quiet=True)
    gp = build_ref.import_ref()
    rng = np.random.default_rng(17)
    n = 16
    x = np.sort(rng.uniform(-1, 1, size=(n, 1)), axis=0)
    effect = np.array(list(range(3)) * 3)[:, None]
    y = rng.standard_normal((n, 9))
    ref = gp.fanova.FANOVA(x, y, effect)
    d = orc.Design(effect)
    assert np.array_equal(ref.design_matrix, d.X)
    F = rng.standard_normal((n, 3))
    hyper = rng.uniform(-1, 1, size=5)
    for f in range(3):
        ref.parameter_cache[ref.functionIndex(f)] = F[:, f]
    for nm, v in zip(['y_sigma', 'prior0_sigma', 'prior0_lengthscale', 'prior1_sigma', 'prior1_lengthscale'], hyper):
        ref.parameter_cache[nm] = v
    m = orc.Model(x, y, d.X, d.prior_groups())
    m.F, m.hyper = F.copy(), hyper.copy()
    for h, (nm, kw) in enumerate([('y_sigma', None), ('p0s', dict(p=0, sigma=0.3)), ('p0l', dict(p=0, lengthscale=-0.2)),
                                  ('p1s', dict(p=1, sigma=-0.7)), ('p1l', dict(p=1, lengthscale=0.4))]):
        if h == 0:
            a = ref.observationLikelihood(sigma=-0.4, prior_lb=-2, prior_ub=2)
            b = m.logp_hyper(0, -0.4)
        else:
            a = ref.prior_likelihood(prior_lb=-2, prior_ub=2, **kw)
            b = m.logp_hyper(h, list(kw.values())[1])
        assert abs(a - b) <= 1e-8 * max(1, abs(a))


def test_reference_floor_against_arbiter():
    """tests/golden/truth.npz (60-digit mpmath, oracle/make_truth.py): documents the FP64 floor of the reference's own
    path on the conditioning-limited stages -- the allowance the GPU parity tests use."""
    T = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "truth.npz")))
    worst = dict(Kinv=0.0, mu=0.0, cov=0.0)
    for k, v in T.items():
        name, key = k.split("/")
        if key.startswith("prior_ll"):
            continue
        c = load_case(name)
        kind = key.split("_")[0]
        worst[kind] = max(worst[kind], relerr(c[key], v))
    assert 1e-8 < worst["Kinv"] < 5e-5          # K_inv is only good to ~1e-6
    assert 1e-11 < worst["mu"] < 5e-9           # conditional mean: ~1e-10..1e-9
    assert 1e-11 < worst["cov"] < 5e-9
    c = load_case("oneway_n50")
    m = oracle_model(c)
    for which, key in (("sigma", "prior_ll_sigma_0"), ("lengthscale", "prior_ll_ls_0")):
        tru = T["oneway_n50/prior_ll_%s_0" % which]
        for i, cand in enumerate(c["cands"]):
            if np.isfinite(tru[i]):
                assert abs(c[key][i] - tru[i]) <= 5e-9 * max(1, abs(tru[i]))
                o = orc.prior_loglik(m.x, m.F[:, m.groups[0]].T, m.hyper[1], m.hyper[2], cand, which)
                assert abs(o - tru[i]) <= 5e-9 * max(1, abs(tru[i]))


@pytest.mark.parametrize("name", ["oneway_n20_deriv", "oneway_n20_deriv_l1"])
def test_derivative_restatement_matches_reference(name):
    """rbf_dK is the reference's RBF._dK bit for bit; FunctionDerivative._params to the conditioning of K_inv
    (kernel.pyx:51-64 inverts a cond-1e11 matrix: two evaluations of the same formula differ by ~1e-6)"""
    c = load_case(name)
    for g in range(len(c["groups"])):
        s, l = 10.0 ** c["hyper"][1 + 2 * g], 10.0 ** c["hyper"][2 + 2 * g]
        assert np.array_equal(orc.rbf_dK(c["x"], s, l), c["dK_%d" % g])
        assert np.array_equal(orc.rbf_dK(c["x"], s, l, cross=True), c["dKcross_%d" % g])
        for f in c["groups"][g]:
            mu, cov = orc.function_derivative_params(c["x"], c["F"][:, f], s, l)
            assert relerr(mu, c["dmu_%d" % f]) < 2e-5
            assert np.max(np.abs(cov - c["dcov_%d" % f])) < 2e-5 * max(np.max(np.abs(c["dK_%d" % g])), 1e-12)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only in the build container")
def test_interval_restatement_matches_live_reference():
    """oracle.function_interval / scalar_interval_threshold against gpfanova/interval.py of the live reference"""
    import build_ref
    build_ref.build(quiet=True)
    gp = build_ref.import_ref()
    rng = np.random.default_rng(4)
    for S, n, alpha in ((200, 12, .95), (333, 5, .9), (64, 30, .5)):
        X = np.cumsum(rng.standard_normal((S, n)), axis=1) * 0.3 + rng.standard_normal((S, 1))
        ref = gp.interval.FunctionInterval(X, alpha)
        got = orc.function_interval(X, alpha)
        assert got["epsilon"] == ref.epsilon and got["alpha"] == ref.alpha
        assert np.array_equal(got["lb"], ref.lb) and np.array_equal(got["ub"], ref.ub)
        x = rng.standard_normal(S)
        pi = lambda t: -0.5 * t * t
        ref = gp.interval.ScalarInterval(x, alpha, pi)
        thr, eps = orc.scalar_interval_threshold(x, alpha, pi)
        assert thr == ref.epj and eps == ref.epsilon
