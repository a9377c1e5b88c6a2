"""Statistical parity with the reference sampler (BASELINE north_star: "posterior summaries over long chains must
agree statistically").  tests/golden/posterior_oneway_n16.npz holds per-chain posterior means / variances from 8
long chains (3000 sweeps after 300 burn-in) of the UNMODIFIED reference (oracle/make_posterior_summary.py).  The GPU
sampler runs many chains on the same data; its pooled posterior means must agree with the reference's within the
Monte-Carlo error of the reference chains (their between-chain scatter) plus the posterior scale."""
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_posterior_summaries_match_reference_long_chains():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gpfanova_b200 import fanova
    g = dict(np.load(os.path.join(GOLDEN, "posterior_oneway_n16.npz")))
    x, y, eff = g["x"], g["y"], g["effect"]
    C, burn, S = 256, 300, 600
    m = fanova.FANOVA(x, y, eff, chains=C, seed=2026, y_sigma=-2.0)
    m.sample(burn, thin=burn)
    m.sample(S, thin=2)
    raw = m.chain_history()[1:]                       # drop the burn-in row
    H = m.H
    hyper = raw[:, :, :H]                             # (rows, C, H)
    F = raw[:, :, H:].reshape(raw.shape[0], C, m.f, m.n)
    # reference: chain means and their scatter
    ref_h = g["hyper_chain_means"]                    # (8, 5)
    ref_F = g["F_chain_means"]                        # (8, f, n)
    R = ref_h.shape[0]
    ref_h_mean, ref_h_se = ref_h.mean(0), ref_h.std(0, ddof=1) / np.sqrt(R)
    ref_F_mean, ref_F_se = ref_F.mean(0), ref_F.std(0, ddof=1) / np.sqrt(R)
    post_sd_h = np.sqrt(g["hyper_chain_vars"].mean(0))
    post_sd_F = np.sqrt(g["F_chain_vars"].mean(0))
    gpu_h_chain = hyper.mean(0)                       # (C, H) per-chain means
    gpu_h_mean, gpu_h_se = gpu_h_chain.mean(0), gpu_h_chain.std(0, ddof=1) / np.sqrt(C)
    gpu_F_chain = F.mean(0)
    gpu_F_mean, gpu_F_se = gpu_F_chain.mean(0), gpu_F_chain.std(0, ddof=1) / np.sqrt(C)
    # hyper-parameters: difference of means within 5 combined standard errors (and a 5%-of-posterior-sd floor)
    zh = np.abs(gpu_h_mean - ref_h_mean) / np.sqrt(ref_h_se ** 2 + gpu_h_se ** 2 + (0.05 * post_sd_h) ** 2)
    assert zh.max() < 5.0, (zh, gpu_h_mean, ref_h_mean)
    # functions: same criterion point-wise; all 3 x 16 values
    zF = np.abs(gpu_F_mean - ref_F_mean) / np.sqrt(ref_F_se ** 2 + gpu_F_se ** 2 + (0.05 * post_sd_F) ** 2)
    assert zF.max() < 5.0, (zF.max(), np.unravel_index(zF.argmax(), zF.shape))
    # posterior spread agrees too (variance ratio of hyper-parameters within 25%)
    gpu_var_h = hyper.reshape(-1, H).var(0)
    assert np.all(np.abs(gpu_var_h / post_sd_h ** 2 - 1.0) < 0.35), gpu_var_h / post_sd_h ** 2
    # and the sampler recovers the generating functions about as well as the reference does
    assert np.sqrt(np.mean((gpu_F_mean - g["F_true"].T) ** 2)) < np.sqrt(np.mean((ref_F_mean - g["F_true"].T) ** 2)) + 0.02
