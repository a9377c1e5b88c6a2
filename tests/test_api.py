"""The reference-facing Python API (FANOVA / Base / SamplerContainer / kernels / samplers): construction-time
behaviour on CPU, compute behaviour on the GPU, both against golden outputs of the reference."""
import numpy as np
import pytest

import gpfanova_oracle as orc
from helpers import CASES, load_case, relerr

FANOVA_CASES = [c for c in CASES]


def build(c, **kw):
    from gpfanova_b200 import fanova
    return fanova.FANOVA(c["x"], c["y"], c["effect"], interactions=bool(c["interactions"]),
                         helmertConvert=bool(c["helmertConvert"]), **kw)


def install_state(m, c):
    names = m.hyperNames()
    for h, nm in enumerate(names):
        m.parameter_cache[nm] = c["hyper"][h]
    for f in range(m.f):
        m.parameter_cache[m.functionIndex(f)] = c["F"][:, f]


@pytest.mark.parametrize("name", FANOVA_CASES)
def test_construction_matches_reference(name):
    """design matrix, prior groups, sampler list (names, types, order) and parameter index of the reference"""
    c = load_case(name)
    m = build(c)
    assert np.array_equal(m.design_matrix, c["design_matrix"])
    assert [list(g) for g in m.priorGroups()] == c["groups"]
    assert [s.name for s in m.samplers] == [str(s) for s in c["sampler_names"]]
    assert [s.type for s in m.samplers] == [str(s) for s in c["sampler_types"]]
    assert list(m.parameter_cache.index) == [str(s) for s in c["index"]]
    assert list(m.parameter_history.columns) == [str(s) for s in c["index"]]
    assert (m.parameter_cache.values == 0).all()


def test_initial_values_by_sampler_name_and_mean_model():
    from gpfanova_b200 import mean
    c = load_case("kat0_zeroF")
    m = build(c, y_sigma=-2, prior0_lengthscale=0.5)
    assert m.parameter_cache["y_sigma"] == -2 and m.parameter_cache["prior0_lengthscale"] == 0.5
    mm = mean.Mean(c["x"], c["y"])
    assert mm.design_matrix.shape == (6, 1) and mm.priorGroups() == [[0]]
    assert [s.name for s in mm.samplers] == ["y_sigma", "f0", "prior0_sigma", "prior0_lengthscale"]


def test_transform_block_matches_reference_transforms():
    for name in ["oneway_n50", "twoway_n24", "twoway_n24_hc", "unbalanced_p2_n12"]:
        c = load_case(name)
        m = build(c)
        blk = m._transform_block(c["F"].T[None])[0].reshape(-1, m.n)
        assert blk.shape == c["transform_vals"].shape
        for got, ref in zip(blk, c["transform_vals"]):
            assert relerr(got, ref) < 1e-12


def test_csv_roundtrip_and_resume(tmp_path):
    c = load_case("kat0_sinF")
    m = build(c)
    install_state(m, c)
    m.store()
    m.parameter_cache["y_sigma"] = -1.25
    m.store()
    assert m.parameter_history.shape == (2, len(c["index"]))
    f = tmp_path / "hist.csv"
    m.save(str(f))
    m2 = build(c, parameterFile=str(f))
    assert m2.parameter_history.shape == (2, len(c["index"]))
    assert np.allclose(m2.parameter_history.values, m.parameter_history.values)
    m2.resume_from_last()
    assert m2.parameter_cache["y_sigma"] == -1.25


# ---------------------------------------------------------------------------------------------- GPU
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@gpu
@pytest.mark.parametrize("name", ["kat0_sinF", "oneway_n50_missing", "twoway_n24_hc", "unbalanced_p2_n12"])
def test_model_log_densities_and_kernels(cuda, name):
    c = load_case(name)
    m = build(c)
    install_state(m, c)
    for g in range(len(c["groups"])):
        assert relerr(m.kernels[g].K(m.x), c["K_%d" % g]) < 1e-13
        assert relerr(m.kernels[g].K_inv(m.x), c["Kinv_%d" % g]) < 5e-5
    assert abs(m.y_k.K_inv(m.x)[0, 0] - c["yinv00"]) <= 1e-12 * abs(c["yinv00"])
    for i, cand in enumerate(c["cands"]):
        ref = c["obs_ll"][i]
        got = m.observationLikelihood(sigma=cand, prior_lb=-2, prior_ub=2)
        assert got == ref if np.isinf(ref) else abs(got - ref) <= 1e-10 * abs(ref)
    for g in range(len(c["groups"])):
        for i in (1, 4, 7, 10):
            cand = c["cands"][i]
            for kw, key in ((dict(sigma=cand), "prior_ll_sigma_%d"), (dict(lengthscale=cand), "prior_ll_ls_%d")):
                ref = c[key % g][i]
                got = m.prior_likelihood(p=g, prior_lb=-2, prior_ub=2, **kw)
                assert got == ref if np.isinf(ref) else abs(got - ref) <= 5e-9 * max(1, abs(ref))
    with pytest.raises(ValueError):
        m.prior_likelihood(p=0, prior_lb=-2, prior_ub=2)


@gpu
def test_jitchol_and_kernel_objects(cuda):
    from gpfanova_b200 import linalg, kernel
    import os
    from helpers import GOLDEN
    d = dict(np.load(os.path.join(GOLDEN, "jitchol.npz")))
    assert relerr(linalg.jitchol(d["A_pd"]), d["L_pd"]) < 1e-12
    with pytest.raises(np.linalg.LinAlgError):
        linalg.jitchol(d["A_neg"])

    class M(object):
        pass
    mod = M()
    import pandas as pd
    mod.parameter_cache = pd.Series({"s": 0.3, "l": -0.4})
    k = kernel.RBF(mod, ["s", "l"], logspace=True)
    x = np.linspace(-1, 1, 9)[:, None]
    assert relerr(k.K(x), orc.rbf_K(x, 10 ** 0.3, 10 ** -0.4)) < 1e-13
    assert relerr(k.K(x, None, 0.1), orc.rbf_K(x, 10 ** 0.3, 10 ** 0.1)) < 1e-13
    assert relerr(k.K(x, s=-1.0), orc.rbf_K(x, 10 ** -1.0, 10 ** -0.4)) < 1e-13
    assert relerr(kernel.RBF.dist(x, 0.5), orc.rbf_dist(x, 0.5)) < 1e-7
    w = kernel.White(mod, ["s"], logspace=True)
    assert np.allclose(w.K(x), 10 ** 0.3 * np.eye(9), rtol=1e-14, atol=0)


@gpu
def test_sample_fills_history_like_the_reference(cuda):
    c = load_case("twoway_n24")
    m = build(c, chains=3, seed=5)
    install_state(m, c)
    m.sample(6, thin=2)
    cols = [str(s) for s in c["index"]]
    assert list(m.parameter_history.columns) == cols and m.parameter_history.shape == (3, len(cols))
    assert np.isfinite(m.parameter_history.values.astype(float)).all()
    # the Transform columns of every stored row are the reference's transforms of that row's functions
    d = orc.Design(c["effect"], interactions=True)
    for r in range(3):
        F = np.stack([m.parameter_history.loc[r, m.functionIndex(f)].values.astype(float) for f in range(m.f)], axis=1)
        for nm, v in orc.fanova_transforms(d, F):
            labels = m.sampler_dict[nm].parameters
            assert relerr(m.parameter_history.loc[r, labels].values.astype(float), v) < 1e-12, nm
    # parameter_cache holds the last state of chain 0, transforms included
    last = m.parameter_history.iloc[-1]
    assert np.allclose(m.parameter_cache[cols].values.astype(float), last.values.astype(float), rtol=0, atol=0)
    raw = m.chain_history()
    assert raw.shape == (3, 3, m.H + m.f * m.n)
    assert not np.array_equal(raw[:, 0], raw[:, 1])
    assert m.chain_history(1).shape == (3, len(cols))
    hyper, F = m.chain_state()
    assert np.array_equal(hyper[0], raw[-1, 0, :m.H])
    # a second call continues the chains and appends
    m.sample(2, thin=0, random=True)
    assert m.parameter_history.shape[0] == 5
    assert m.parameterSamples("y_sigma").shape == (5, 1)
    assert m.functionSamples(0).shape == (5, m.n)


@gpu
def test_same_seed_same_chain_and_oracle_agreement(cuda):
    """chains=1 through the public API follows the oracle sweep driven by the same Philox stream"""
    c = load_case("oneway_n50")
    m = build(c, chains=1, seed=77)
    install_state(m, c)
    m.sample(3)
    om = orc.Model(c["x"], c["y"], c["design_matrix"], c["groups"])
    om.F, om.hyper = c["F"].copy(), c["hyper"].copy()
    st = orc.Stream(77)
    for s in range(3):
        om.sweep(st, 0, s)
        row = m.parameter_history.iloc[s]
        assert np.allclose(row[m.hyperNames()].values.astype(float), om.hyper, rtol=0, atol=1e-6)
        for f in range(m.f):
            assert relerr(row[m.functionIndex(f)].values.astype(float), om.F[:, f]) < 1e-6


@gpu
def test_individual_samplers_and_replaced_data(cuda):
    c = load_case("oneway_n50")
    m = build(c, seed=1)
    install_state(m, c)
    fs = [s for s in m.samplers if s.type == "Function"][1]
    mu, cov = fs.conditional(z=np.zeros(2 * m.n))
    assert relerr(mu, c["mu_%d" % fs.f]) < 5e-9
    assert relerr(cov, c["cov_%d" % fs.f]) < 5e-9
    beta = fs.sample()
    assert beta.shape == (m.n,) and np.array_equal(beta, m.parameter_cache[fs.parameters].values)
    sl = m.sampler_dict["prior1_lengthscale"]
    x1 = sl.sample(float(m.parameter_cache["prior1_lengthscale"]))
    assert -2 <= x1 <= 2 and m.parameter_cache["prior1_lengthscale"] == x1
    # the reference's convergence tests overwrite m.y after construction (test_convergence.py:60-62)
    y2, f2 = m.samplePrior()
    assert y2.shape == m.y.shape and f2.shape == (m.n, m.f)
    m.y = y2
    m.sample(2)
    assert m.engine().y is not c["y"] and np.array_equal(m.engine().y, y2)
    # the generic container loop over sampler objects gives the same kind of row
    m._sample()
    m.store()
    assert m.parameter_history.shape[0] == 3


@gpu
def test_not_pd_surfaces_as_linalgerror(cuda):
    c = load_case("kat0_zeroF")
    m = build(c, chains=2)
    m.parameter_cache["prior0_sigma"] = 400.0      # sigma = 1e400 = inf -> K is not finite
    with pytest.raises(np.linalg.LinAlgError):
        m.sample(1)


@gpu
def test_chains_sharded_over_devices_match_single_device(cuda):
    """devices=[0, 1]: contiguous blocks of chains per GPU, Philox keyed by the global chain index -> the same samples as
    on one device (needs two GPUs; on a one-GPU box the same device is used twice, which exercises the same code)"""
    c = load_case("twoway_n24")
    devs = [0, 1] if cuda.cuda.device_count() >= 2 else [0, 0]
    outs = []
    for kw in (dict(device=0), dict(devices=devs)):
        m = build(c, chains=5, seed=12, **kw)
        install_state(m, c)
        m.sample(4, thin=2)
        outs.append((m.chain_history().copy(), m.parameter_history.values.astype(float).copy(), m.chain_state()))
        assert m.counters()["draws"] == 5 * 4 * m.f
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2][0], outs[1][2][0]) and np.array_equal(outs[0][2][1], outs[1][2][1])
