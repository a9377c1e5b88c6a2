"""The two algebraic identities the CUDA path uses in place of the reference's literal arithmetic, checked in NumPy
against the oracle's restatement of the reference (no GPU needed; the GPU tests check the kernels that implement them):

* centrosymmetric split of K_inv on uniform grids (k_group_kinv<2>, gpf_set_option 3);
* observation sum of squares about the least-squares fit for balanced designs (k_obs, gpf_set_option 4).
"""
import itertools

import numpy as np
import pytest

from helpers import relerr
from oracle import gpfanova_oracle as orc

EPS = np.finfo(float).eps


@pytest.mark.parametrize("n", [64, 100, 256])
def test_centrosymmetric_split_reproduces_the_inverse(n):
    """K + 1e-9 I symmetric Toeplitz => K = Q^T diag(K11 + K12 J, K11 - K12 J) Q; K_inv = 1/2 [[S, D J], [J D, J S J]]"""
    rng = np.random.default_rng(n)
    x = np.linspace(-1, 1, n)[:, None]
    h = n // 2
    for _ in range(3):
        sigma, ls = 10.0 ** rng.uniform(-1, 1), 10.0 ** rng.uniform(-0.8, 0.3)
        K = orc.rbf_K(x, sigma, ls) + 1e-9 * np.eye(n)
        ref = orc.K_inv(K - 1e-9 * np.eye(n))[0]                 # kernel.pyx:51-64 through the oracle
        J = np.eye(h)[::-1]
        Bp, Bm = K[:h, :h] + K[:h, h:] @ J, K[:h, :h] - K[:h, h:] @ J
        assert relerr(Bp, Bp.T) < 1e-13 and relerr(Bm, Bm.T) < 1e-13     # (the reference K is Toeplitz to rounding)

        def chol_inverse(B):
            L = np.linalg.cholesky(B)
            Li = np.linalg.solve(L, np.eye(h))
            return Li.T @ Li
        Mp, Mm = chol_inverse(Bp), chol_inverse(Bm)
        S, D = Mp + Mm, Mp - Mm
        Kinv = 0.5 * np.block([[S, D @ J], [J @ D, J @ S @ J]])
        cond = np.linalg.cond(K)
        tol = max(1e-10, max(2.0, 0.5 * np.sqrt(n)) * EPS * cond)
        assert relerr(Kinv, ref) < tol, (relerr(Kinv, ref), cond)
        assert np.max(np.abs(K @ Kinv - np.eye(n))) <= max(1e-12, 32 * n * EPS * cond)


@pytest.mark.parametrize("helmert", [False, True])
def test_sum_of_squares_about_the_least_squares_fit(helmert):
    """X^T X diagonal, nothing missing:  ||y - X F||^2 = ||y - X bhat||^2 + sum_f G_ff ||F_f - bhat_f||^2"""
    rng = np.random.default_rng(3)
    n = 40
    x = np.linspace(-1, 1, n)[:, None]
    effect = np.array(list(itertools.product(range(3), range(3))) * 4)
    d = orc.Design(effect, interactions=True, helmertConvert=helmert)
    X = d.X
    G = X.T @ X
    off = G - np.diag(np.diag(G))
    if np.abs(off).max() > 1e-12 * np.abs(np.diag(G)).max():
        pytest.skip("this contrast coding does not give an orthogonal design")
    F = rng.standard_normal((n, d.f))
    y = F @ X.T + 0.1 * rng.standard_normal((n, d.m))
    Ftry = F + 0.05 * rng.standard_normal(F.shape)
    direct = np.sum((y - Ftry @ X.T) ** 2)
    bhat = (y @ X) / np.diag(G)                                  # n x f
    rss0 = np.sum((y - bhat @ X.T) ** 2)
    split = rss0 + np.sum(np.diag(G) * np.sum((Ftry - bhat) ** 2, axis=0))
    assert abs(split - direct) <= 1e-12 * direct
    for cand in (-2.5, -2.0, -1.0):
        o = orc.observation_loglik(y, X, Ftry, cand)
        sd = np.sqrt(10.0 ** cand)
        ll = -0.5 * split / sd ** 2 - y.size * (np.log(sd) + 0.5 * np.log(2 * np.pi)) - np.log(2.0 - (-2.0))
        assert abs(ll - o) <= 1e-10 * abs(o)
