"""CPU oracle for the gpfanova Gibbs-sampling hot path.  TEST INFRASTRUCTURE ONLY.

A staged NumPy restatement of the reference algorithm (ptonner/gpfanova), every
function citing the reference file:line it follows (paths relative to
/root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline legs may import this module; the product (``gpfanova_b200``) must not.

Parity status: the reference holds no golden vectors / KATs for this path
(SURVEY.md 4, 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/build_ref.py`` emits a py3-importable copy of the unmodified
reference into ``oracle/_ref`` and ``oracle/make_golden.py`` runs it to mint
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this restatement against
those vectors (and, when ``oracle/_ref`` is importable, against the live
reference).

Two deliberate, documented departures, both *injection points* and not changes of
the sampled distribution:

  * Gaussian draw map.  The reference draws through numpy's SVD map
    (function.py:40), which no factorisation-based sampler can reproduce for the same
    z.  The contract here (oracle AND device) is the observation form of the same
    conditional N(A^-1 b, A^-1), A = K_o^-1 + D (function.py:27-38):
        f = S_B z1,  q = q0 - S f - z2,  v = C^-1 q,  beta = f + S_B S_B^T S v
    with S = D^1/2, C = I + S K_o S, q0 = S^-1 b and S_B a square root of K_o = K + 1e-9 I
    (``prior_sqrt``: the Cholesky factor kernel.pyx:56 builds; on a uniform grid with an
    even number >= 64 of time points the block factor of the centrosymmetric split).
    mu = K_o S C^-1 q0 and Sigma = K_o - K_o S C^-1 S K_o are the reference's A^-1 b and
    A^-1 (``function_conditional`` keeps the reference's own arithmetic for comparison);
    equality of the draws with the reference is in distribution.
  * Random streams.  The reference uses the global numpy/scipy stream.  Here all
    randomness comes from Philox4x32-10 keyed (seed; chain, sweep, sampler, draw)
    so the device can replay the identical stream.
"""
import math

import numpy as np

OFFSET = 1e-9  # kernel/kernel.pyx:7
LOG2PI = math.log(2.0 * math.pi)

# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011, Random123) -- counter-based RNG
# --------------------------------------------------------------------------
_PH_M0 = 0xD2511F53
_PH_M1 = 0xCD9E8D57
_PH_W0 = 0x9E3779B9
_PH_W1 = 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """one Philox4x32-10 block; pure-python ints.  KAT: Random123 kat_vectors."""
    for _ in range(10):
        p0 = _PH_M0 * c0
        p1 = _PH_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _M32
        hi1, lo1 = p1 >> 32, p1 & _M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _M32, lo1, (hi0 ^ c3 ^ k1) & _M32, lo0
        k0 = (k0 + _PH_W0) & _M32
        k1 = (k1 + _PH_W1) & _M32
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """vectorised Philox (uint64 numpy arrays holding 32-bit values)."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.broadcast_to(np.asarray(c1, dtype=np.uint64), c0.shape).copy()
    c2 = np.broadcast_to(np.asarray(c2, dtype=np.uint64), c0.shape).copy()
    c3 = np.broadcast_to(np.asarray(c3, dtype=np.uint64), c0.shape).copy()
    c0 = c0.copy()
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    m32 = np.uint64(_M32)
    for _ in range(10):
        p0 = np.uint64(_PH_M0) * c0
        p1 = np.uint64(_PH_M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & m32, lo1, (hi0 ^ c3 ^ k1) & m32, lo0
        k0 = (k0 + np.uint64(_PH_W0)) & m32
        k1 = (k1 + np.uint64(_PH_W1)) & m32
    return c0, c1, c2, c3


def u01(hi, lo):
    """53-bit uniform in (0,1): ((hi>>6)*2^27 + (lo>>5) + 0.5) * 2^-53 (exact in f64)."""
    return ((hi >> 6) * 134217728.0 + (lo >> 5) + 0.5) * (1.0 / 9007199254740992.0)


class Stream(object):
    """The keyed stream used by both the oracle and the device.

    counter = (draw, sampler_id, sweep, chain), key = (seed_lo, seed_hi).
    Each counter yields two uniforms (u_a from words 0,1; u_b from words 2,3).
    """

    def __init__(self, seed):
        self.k0 = seed & _M32
        self.k1 = (seed >> 32) & _M32

    def uniform(self, chain, sweep, sampler, draw):
        r = philox4x32_10(draw & _M32, sampler & _M32, sweep & _M32, chain & _M32, self.k0, self.k1)
        return u01(r[0], r[1])

    def normals(self, chain, sweep, sampler, n):
        """n standard normals by Box-Muller over ceil(n/2) counters (draw = pair index)."""
        npair = (n + 1) // 2
        r = philox4x32_10_np(np.arange(npair), sampler, sweep, chain, self.k0, self.k1)
        ua = ((r[0] >> np.uint64(6)).astype(np.float64) * 134217728.0 + (r[1] >> np.uint64(5)).astype(np.float64) + 0.5) / 9007199254740992.0
        ub = ((r[2] >> np.uint64(6)).astype(np.float64) * 134217728.0 + (r[3] >> np.uint64(5)).astype(np.float64) + 0.5) / 9007199254740992.0
        rad = np.sqrt(-2.0 * np.log(ua))
        ang = 2.0 * np.pi * ub
        z = np.empty(2 * npair)
        z[0::2] = rad * np.cos(ang)
        z[1::2] = rad * np.sin(ang)
        return z[:n]


# --------------------------------------------------------------------------
# L1: kernels + dense linear algebra
# --------------------------------------------------------------------------
def rbf_dist(X, lengthscale):
    """kernel/rbf.pyx:8-14  (RBF.dist)"""
    X = X / lengthscale
    Xsq = np.sum(np.square(X), 1)
    r2 = -2.0 * np.dot(X, X.T) + Xsq[:, None] + Xsq[None, :]
    r2 = np.clip(r2, 0, np.inf)
    return np.sqrt(r2)


def rbf_K(X, sigma, lengthscale):
    """kernel/rbf.pyx:16-19  (RBF._K); sigma, lengthscale already 10** (kernel.pyx:35-37)"""
    d = rbf_dist(X, lengthscale)
    return sigma * np.exp(-0.5 * d ** 2)


def white_K(n, sigma):
    """kernel/white.py:7-9"""
    return sigma * np.eye(n)


def _potrf_lower(A):
    """unblocked lower Cholesky with LAPACK dpotf2's failure rule
    (pivot <= 0 or NaN -> info = column index + 1).  Used instead of scipy so the
    oracle has no BLAS-version dependence; linalg.py:37 calls lapack.dpotrf(lower=1)."""
    try:
        L = np.linalg.cholesky(A)  # LAPACK dpotrf; raises on info>0
    except np.linalg.LinAlgError:
        return None, 1
    return L, 0


def jitchol(A, maxtries=5):
    """linalg.py:35-54.  Returns (L, ntries) with ntries = number of jittered retries used.
    Raises np.linalg.LinAlgError exactly where the reference does."""
    A = np.ascontiguousarray(A)
    L, info = _potrf_lower(A)
    if info == 0:
        return L, 0
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        L, info = _potrf_lower(A + np.eye(A.shape[0]) * jitter)
        if info == 0:
            return L, num_tries
        jitter *= 10
        num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def K_inv(K):
    """kernel/kernel.pyx:51-64:  (K + OFFSET*I)^-1 = L^-T L^-1 with L = jitchol(K+OFFSET I).
    The reference gets L^-1 from the general np.linalg.inv (LU); a triangular solve is the
    same matrix to rounding.  Returns (Kinv, L, ntries)."""
    n = K.shape[0]
    Kj = K + np.eye(n) * OFFSET
    L, ntries = jitchol(Kj)
    Linv = np.linalg.inv(L)
    return np.dot(Linv.T, Linv), L, ntries


# --------------------------------------------------------------------------
# L4 helpers: contrasts, design matrix, prior groups  (host-side, run once)
# --------------------------------------------------------------------------
def helmert(k):
    """patsy 0.4.1 Helmert().code_without_intercept(range(k)).matrix  (fanova.py:160)"""
    contr = np.zeros((k, k - 1))
    contr[1:][np.diag_indices(k - 1)] = np.arange(1, k)
    contr[np.triu_indices(k - 1)] = -1
    return contr


def helmert_convert(c):
    """fanova.py:151-158"""
    c = c.copy()
    n = c.shape[1]
    for i in range(n):
        j = n - i - 1
        c[:, j] /= c[j + 1, j]
        s = np.power(c[j + 1, :], 2).sum()
        c[:, j] *= np.sqrt(1 - 1.0 / (n + 1) - s + np.power(c[j + 1, j], 2))
    return c


class Design(object):
    """FANOVA structure: contrasts, design matrix, prior groups, names.
    fanova.py:11-45 (ctor), :266-281 (priorGroups), :283-310 (_buildDesignMatrix)."""

    def __init__(self, effect, interactions=False, helmertConvert=False):
        effect = np.asarray(effect)
        self.effect = effect
        self.m = effect.shape[0]
        self.k = effect.shape[1]
        self.mk = [np.unique(effect[:, i]).shape[0] for i in range(self.k)]
        if isinstance(interactions, bool):
            self.interactions = interactions
            self._inter = {(i, j): True for i in range(self.k) for j in range(i + 1, self.k)}
        else:
            self.interactions = True
            self._inter = {}
            for i, j in interactions:
                if j < i:
                    i, j = j, i
                self._inter[(i, j)] = True
        self.contrasts = []
        for i in range(self.k):
            h = helmert(self.mk[i])
            if helmertConvert:
                h = helmert_convert(h)
            self.contrasts.append(h)
        self.contrasts_interaction = {}
        for i in range(self.k):
            for j in range(i):
                if self.interactions:
                    self.contrasts_interaction[(j, i)] = np.kron(self.contrasts[j], self.contrasts[i])
        self.inter_index = {}
        self.X = self._design()
        self.f = self.X.shape[1]

    def has_interaction(self, i, k):
        if i > k:
            i, k = k, i
        return self.interactions and (i, k) in self._inter

    def _design(self):
        d = 1
        for i in range(self.k):
            d += self.mk[i] - 1
            for j in range(i):
                if self.has_interaction(j, i):
                    d += (self.mk[i] - 1) * (self.mk[j] - 1)
        x = np.zeros((self.m, d))
        x[:, 0] = 1
        for s in range(self.m):
            ind = 1
            for i in range(self.k):
                x[s, ind:ind + self.mk[i] - 1] = self.contrasts[i][self.effect[s, i], :]
                ind += self.mk[i] - 1
            for i in range(self.k):
                for j in range(i + 1, self.k):
                    if self.has_interaction(i, j):
                        z = self.effect[s, i] * self.mk[j] + self.effect[s, j]
                        w = (self.mk[i] - 1) * (self.mk[j] - 1)
                        x[s, ind:ind + w] = self.contrasts_interaction[(i, j)][z, :]
                        self.inter_index[(i, j)] = ind
                        ind += w
        return x

    def prior_groups(self):
        g = [[0]]
        ind = 1
        for i in range(self.k):
            g.append(list(range(ind, ind + self.mk[i] - 1)))
            ind += self.mk[i] - 1
        for i in range(self.k):
            for k in range(i + 1, self.k):
                if self.has_interaction(i, k):
                    c = (self.mk[i] - 1) * (self.mk[k] - 1)
                    g.append(list(range(ind, ind + c)))
                    ind += c
        return g


# --------------------------------------------------------------------------
# L2: residual projection + conditional posterior of one function
# --------------------------------------------------------------------------
def function_moments(y, X, F, f):
    """base.py:180-190 (residual/functionResidual) + sample/function.py:16-25.

    y: n x m (NaN = missing), X: m x f, F: n x f.  Returns (m_t, n_t), each length n:
      m_t = nansum_o  w_o * (y_o(t) - sum_{g!=f} X[o,g] F[t,g]) / X[o,f],  w_o = X[o,f]^2
      n_t = sum_o  w_o * [y_o(t) observed]          (rows with X[o,f]==0 dropped)
    """
    Fm = F.copy()
    Fm[:, f] = 0.0
    resid = y.T - np.dot(X, Fm.T)                      # m x n
    with np.errstate(divide='ignore', invalid='ignore'):
        resid = (resid.T / X[:, f].T).T
    resid = resid[X[:, f] != 0, :]
    w = np.power(X[:, f], 2)
    w = w[w != 0]
    miss = np.isnan(resid)
    m_t = np.nansum((resid.T * w).T, 0)
    n_t = np.sum(((~miss).T * w).T, 0)
    return m_t, n_t


def function_conditional(m_t, n_t, y_sigma_log10, Kinv_f):
    """sample/function.py:27-38.  y_inv = K_inv of White = I/(10**y_sigma + OFFSET)
    (white.py:7-9 through kernel.pyx:51-64).  Returns dict of every intermediate."""
    n = m_t.shape[0]
    yinv_s = 1.0 / (10.0 ** y_sigma_log10 + OFFSET)
    A = np.diag(n_t * yinv_s) + Kinv_f                 # function.py:30 (broadcast on a diagonal y_inv)
    b = yinv_s * m_t                                   # function.py:32
    L_A, ntries = jitchol(A)                           # function.py:34
    Linv = np.linalg.inv(L_A)                          # function.py:35
    A_inv = np.dot(Linv.T, Linv)                       # function.py:36
    mu = np.dot(A_inv, b)                              # function.py:38
    return dict(A=A, b=b, L_A=L_A, mu=mu, cov=A_inv, ntries=ntries)


def function_draw(mu, L_A, z):
    """a draw from N(mu, (L_A L_A^T)^-1) in the reference's own parametrisation: beta = mu + L_A^{-T} z
    (kept for the statistical checks; the device contract is function_draw_obs)."""
    import scipy.linalg
    return mu + scipy.linalg.solve_triangular(L_A, z, lower=True, trans='T')


def uniform_even_grid(x):
    """the grids on which K is symmetric Toeplitz and the centrosymmetric split is used: one input dimension, evenly
    spaced (to 1e-14 of the range), an even number >= 64 of points"""
    x = np.asarray(x, dtype=float)
    if x.ndim == 2:
        if x.shape[1] != 1:
            return False
        x = x[:, 0]
    n = x.shape[0]
    if n % 2 or n < 64:
        return False
    h = (np.longdouble(x[-1]) - np.longdouble(x[0])) / (n - 1)
    tol = 1e-14 * (abs(x[0]) + abs(x[-1]))
    return bool(np.all(np.abs((np.longdouble(x[0]) + np.arange(n) * h - x).astype(float)) <= tol))


def prior_sqrt(K, centro=False):
    """S_B with S_B S_B^T = K_o + jitter I,  K_o = K + OFFSET I  (kernel.pyx:53-56: jitchol of K + 1e-9 I, ladder of
    linalg.py:41-54 included).  centro: K_o = Q^T diag(B+, B-) Q with B+- = K11 +- K12 J, Q = [[I, J], [I, -J]]/sqrt 2, and
    S_B = Q^T diag(chol B+, chol B-) (the ladder adds the same jitter to both blocks).  Returns (S_B, jitter, tries)."""
    n = K.shape[0]
    Ko = K + np.eye(n) * OFFSET
    if not centro:
        L, tries = jitchol(Ko)
        jit = 0.0 if tries == 0 else Ko.diagonal().mean() * 1e-6 * 10.0 ** (tries - 1)
        return L, jit, tries
    h = n // 2
    K11, K12J = Ko[:h, :h], Ko[:h, h:][:, ::-1]
    jit, tries = 0.0, 0
    while True:
        try:
            Lp = np.linalg.cholesky(K11 + K12J + jit * np.eye(h))
            Lm = np.linalg.cholesky(K11 - K12J + jit * np.eye(h))
            break
        except np.linalg.LinAlgError:
            jit = Ko.diagonal().mean() * 1e-6 if tries == 0 else jit * 10.0
            tries += 1
            if tries > 5:
                raise np.linalg.LinAlgError("not positive definite, even with jitter.")
    r = math.sqrt(0.5)
    SB = np.zeros((n, n))
    SB[:h, :h] = r * Lp
    SB[:h, h:] = r * Lm
    SB[h:, :h] = r * Lp[::-1, :]
    SB[h:, h:] = -r * Lm[::-1, :]
    return SB, jit, tries


def function_conditional_obs(m_t, n_t, y_sigma_log10, K, jitter=0.0):
    """the conditional of sample/function.py:27-38 in observation form (no inverse of K):
       D = diag(n_t) y_inv, S = D^1/2, C = I + S K_o S, q0 = S^-1 (y_inv m_t) (0 where n_t = 0),
       mu = K_o S C^-1 q0 (= A^-1 b),  cov = K_o - K_o S C^-1 S K_o (= A^-1),  K_o = K + (OFFSET + jitter) I."""
    n = m_t.shape[0]
    yinv_s = 1.0 / (10.0 ** y_sigma_log10 + OFFSET)
    Ko = K + np.eye(n) * (OFFSET + jitter)
    d = n_t * yinv_s
    sd = np.sqrt(d)
    q0 = np.where(d > 0, yinv_s * m_t / np.where(sd > 0, sd, 1.0), 0.0)
    C = np.eye(n) + sd[:, None] * Ko * sd[None, :]
    L_C = np.linalg.cholesky(C)
    import scipy.linalg
    v0 = scipy.linalg.cho_solve((L_C, True), q0)
    mu = Ko.dot(sd * v0)
    Wm = scipy.linalg.solve_triangular(L_C, sd[:, None] * Ko, lower=True)
    cov = Ko - Wm.T.dot(Wm)
    return dict(C=C, L_C=L_C, q0=q0, sd=sd, mu=mu, cov=cov, Ko=Ko)


def function_draw_obs(st, SB, z):
    """the draw contract: z = [z1 | z2] (2n standard normals),
       f = S_B z1,  q = q0 - S f - z2,  v = C^-1 q,  beta = S_B (z1 + S_B^T S v)."""
    import scipy.linalg
    n = st['q0'].shape[0]
    z1, z2 = z[:n], z[n:2 * n]
    f = SB.dot(z1)
    q = st['q0'] - st['sd'] * f - z2
    v = scipy.linalg.cho_solve((st['L_C'], True), q)
    return SB.dot(z1 + SB.T.dot(st['sd'] * v))


# --------------------------------------------------------------------------
# log densities handed to the slice sampler
# --------------------------------------------------------------------------
def observation_loglik(y, X, F, x, prior_lb=-2.0, prior_ub=2.0):
    """base.py:196-214.  NB the uniform prior is on sd = sqrt(10**x), support [lb, ub]."""
    yv = np.ravel(y.T)
    mu = np.dot(X, F.T).ravel()
    sd = math.sqrt(10.0 ** x)
    ok = ~np.isnan(yv)
    r = yv[ok] - mu[ok]
    if sd < prior_lb or sd > prior_ub:
        priorll = -np.inf
    else:
        priorll = -math.log(prior_ub - prior_lb)
    return float(np.sum(-0.5 * (r / sd) ** 2 - math.log(sd) - 0.5 * LOG2PI) + priorll)


def mvn_logpdf_eigh(xs, cov):
    """scipy.stats.multivariate_normal(0,cov).logpdf as used at base.py:233-238:
    eigh, eps = 1e6*eps_f64*max|s| cut, pseudo-det / pinv (scipy _PSD)."""
    s, u = np.linalg.eigh(cov)
    eps = 1e6 * np.finfo('d').eps * np.max(np.abs(s))
    if np.min(s) < -eps:
        raise ValueError("not PSD")
    d = s[s > eps]
    if len(d) < len(s):
        raise np.linalg.LinAlgError("singular matrix")
    U = u / np.sqrt(s)
    logpdet = np.sum(np.log(s))
    out = []
    for x in xs:
        maha = np.sum(np.square(np.dot(x, U)))
        out.append(-0.5 * (len(s) * LOG2PI + logpdet + maha))
    return np.array(out)


def mvn_logpdf_chol(xs, cov):
    """same quantity through Cholesky (what the device computes)."""
    import scipy.linalg
    L = np.linalg.cholesky(cov)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    out = []
    for x in xs:
        v = scipy.linalg.solve_triangular(L, x, lower=True)
        out.append(-0.5 * (len(x) * LOG2PI + logdet + np.dot(v, v)))
    return np.array(out)


def prior_loglik(xgrid, Fg, sigma_log10, ls_log10, cand, which, prior_lb=-2.0, prior_ub=2.0, method='chol'):
    """base.py:216-243.  Fg: list/array of the group's functions (|g| x n);
    which in {'sigma','lengthscale'} says which hyper-parameter ``cand`` replaces."""
    if which == 'sigma':
        sigma_log10 = cand
    elif which == 'lengthscale':
        ls_log10 = cand
    else:
        raise ValueError("must provide lengthscale or sigma in likelihood fxn")
    n = xgrid.shape[0]
    cov = rbf_K(xgrid, 10.0 ** sigma_log10, 10.0 ** ls_log10)
    cov = cov + cov.mean() * np.eye(n) * 1e-6          # base.py:222
    if cand < prior_lb or cand > prior_ub:             # scipy uniform(lb, ub-lb).logpdf
        priorll = -np.inf
    else:
        priorll = -math.log(prior_ub - prior_lb)
    if not np.isfinite(priorll):
        return -np.inf
    fn = mvn_logpdf_chol if method == 'chol' else mvn_logpdf_eigh
    ll = 1.0 + float(np.sum(fn(np.asarray(Fg), cov)))  # base.py:235-238
    return ll + priorll


# --------------------------------------------------------------------------
# slice sampler
# --------------------------------------------------------------------------
def slice_step(x, logp, w, m, uniforms):
    """sample/slice.pyx:24-62 (Neal 2003 step-out + shrinkage).

    ``uniforms`` is a callable k -> u_k in (0,1).  Use of the stream:
      u_0 -> level: z = f0 - (1 + Exp(1)) with Exp(1) = -log(u_0)   (slice.pyx:30: expon.rvs(1) is loc=1)
      u_1 -> interval placement, u_2 -> step budget split, u_3.. -> proposals.
    Returns (x1, n_logp_evaluations)."""
    nev = 1
    f0 = logp(x)
    z = f0 - (1.0 - math.log(uniforms(0)))
    u = uniforms(1)
    l = x - w * u
    r = l + w
    v = uniforms(2)
    j = int(m * v)
    k = m - 1 - j
    while j > 0:
        nev += 1
        if not (z < logp(l)):
            break
        j -= 1
        l -= w
    while k > 0:
        nev += 1
        if not (z < logp(r)):
            break
        k -= 1
        r += w
    d = 3
    u = uniforms(d)
    d += 1
    x1 = l + u * (r - l)
    while True:
        nev += 1
        if not (z > logp(x1)):
            break
        if x1 < x:
            l = x1
        else:
            r = x1
        u = uniforms(d)
        d += 1
        x1 = l + u * (r - l)
    return x1, nev


# --------------------------------------------------------------------------
# whole model: sampler order + sweep   (base.py:57-86, container.py:29-53)
# --------------------------------------------------------------------------
class Model(object):
    """Restatement of Base/FANOVA + SamplerContainer state for ONE chain.

    hyper layout: [y_sigma, prior0_sigma, prior0_lengthscale, prior1_sigma, ...] (log10).
    sampler ids (= position in the reference's sampler list, base.py:57-84):
      0: y_sigma; then per group g: its functions, prior{g}_sigma, prior{g}_lengthscale.
    """

    def __init__(self, x, y, X, groups, w=0.1, m=10, hyper_wm=None):
        self.x = np.asarray(x, dtype=float)
        self.y = np.asarray(y, dtype=float)
        self.X = np.asarray(X, dtype=float)
        self.n = self.x.shape[0]
        self.m_obs = self.y.shape[1]
        self.f = self.X.shape[1]
        self.groups = [list(g) for g in groups]
        self.P = len(self.groups)
        self.H = 1 + 2 * self.P
        self.F = np.zeros((self.n, self.f))
        self.hyper = np.zeros(self.H)
        # slice (w, m) per hyper-parameter  (base.py:52-56, 75-82)
        self.wm = [(w, m)] * self.H if hyper_wm is None else list(hyper_wm)
        self.order = []          # list of (kind, arg)
        self.order.append(('slice', 0))
        for g, grp in enumerate(self.groups):
            for f in grp:
                self.order.append(('function', f))
            self.order.append(('slice', 1 + 2 * g))
            self.order.append(('slice', 2 + 2 * g))
        self.counters = dict(chol=0, logp=0, jitter=0)
        # square root of K_o used by the prior draw: True = block factor of the centrosymmetric split (the device's dense
        # kernels on uniform even grids with option 3 on), False = full Cholesky factor; None = split wherever it applies
        self.split = None

    # -- log densities on current state ---------------------------------
    def logp_hyper(self, h, cand):
        self.counters['logp'] += 1
        if h == 0:
            return observation_loglik(self.y, self.X, self.F, cand)
        g = (h - 1) // 2
        which = 'sigma' if (h - 1) % 2 == 0 else 'lengthscale'
        Fg = self.F[:, self.groups[g]].T
        return prior_loglik(self.x, Fg, self.hyper[1 + 2 * g], self.hyper[2 + 2 * g], cand, which)

    def function_step(self, f, z):
        """z: 2n standard normals [z1 | z2] (function_draw_obs)"""
        g = [i for i, grp in enumerate(self.groups) if f in grp][0]
        K = rbf_K(self.x, 10.0 ** self.hyper[1 + 2 * g], 10.0 ** self.hyper[2 + 2 * g])
        split = uniform_even_grid(self.x) if self.split is None else (bool(self.split) and uniform_even_grid(self.x))
        SB, jit, nt = prior_sqrt(K, split)
        m_t, n_t = function_moments(self.y, self.X, self.F, f)
        st = function_conditional_obs(m_t, n_t, self.hyper[0], K, jit)
        self.counters['jitter'] += nt
        beta = function_draw_obs(st, SB, z)
        self.F[:, f] = beta
        st.update(m_t=m_t, n_t=n_t, SB=SB, beta=beta)
        return st

    def sweep(self, stream, chain, sweep_idx, order=None):
        """container.py:29-53 with the keyed Philox stream.  ``order`` = permutation of
        sampler ids for random=True (container.py:34-35), default = fixed order."""
        ids = range(len(self.order)) if order is None else order
        nev = {}
        for sid in ids:
            kind, arg = self.order[sid]
            if kind == 'function':
                z = stream.normals(chain, sweep_idx, sid, 2 * self.n)
                self.function_step(arg, z)
            else:
                h = arg
                w, m = self.wm[h]
                x1, ne = slice_step(self.hyper[h], lambda c: self.logp_hyper(h, c), w, m,
                                    lambda k: stream.uniform(chain, sweep_idx, sid, k))
                self.hyper[h] = x1
                nev[h] = ne
        return nev


# --------------------------------------------------------------------------
# FANOVA transforms appended to every stored row  (fanova.py:58-124)
# --------------------------------------------------------------------------
def fanova_transforms(design, F):
    """Returns list of (name, values[n]) in the reference's sampler order (fanova.py:58-81):
    per effect k: effect_l(t) = sum_j C_k[l,j] beta*_j(t) for every level l, then the finite
    population variance diag(a P a^T)/(mk-1) (fanova.py:117-122); then interaction effects."""
    SUF = ['alpha', 'beta', 'gamma', 'delta', 'epsilon']
    suf = (lambda k: SUF[k]) if design.k <= len(SUF) else (lambda k: 'effect%d' % k)
    out = []
    start = 1
    for k in range(design.k):
        w = design.mk[k] - 1
        Fk = F[:, start:start + w]
        a = np.zeros((F.shape[0], design.mk[k]))
        for l in range(design.mk[k]):
            a[:, l] = np.dot(Fk, design.contrasts[k][l, :])
            out.append(('%s_%d' % (suf(k), l), a[:, l]))
        c = np.ones((design.mk[k], 1))
        Pm = np.eye(design.mk[k]) - np.dot(c, np.dot(np.linalg.inv(np.dot(c.T, c)), c.T))
        fpv = 1.0 / (design.mk[k] - 1) * np.diag(np.dot(a, np.dot(Pm, a.T)))
        out.append(('%s_finiteSampleVariance' % suf(k), fpv))
        start += w
    for k in range(design.k):
        for i in range(k + 1, design.k):
            if design.has_interaction(k, i):
                s0 = design.inter_index[(k, i)]
                w = (design.mk[k] - 1) * (design.mk[i] - 1)
                Fi = F[:, s0:s0 + w]
                for l in range(design.mk[k]):
                    for j in range(design.mk[i]):
                        v = np.dot(Fi, design.contrasts_interaction[(k, i)][l * design.mk[i] + j, :])
                        out.append(('(%s,%s)_(%d,%d)' % (suf(k), suf(i), l, j), v))
    return out


# --------------------------------------------------------------------------
# "next" rows of SURVEY.md 8f: derivative sampler, intervals, diagnostics
# --------------------------------------------------------------------------
def rbf_dK(X, sigma, lengthscale, cross=False):
    """kernel/rbf.pyx:21-38 (RBF._dK), including its scaling by 1/lengthscale (not 1/lengthscale^2) and its use of the
    UNSCALED first input dimension for diff."""
    k = rbf_K(X, sigma, lengthscale)
    diff = X[:, 0][:, None] - X[:, 0][None, :]                      # rbf.pyx:30-32
    if cross:
        return (-1. / lengthscale) * diff * k                       # rbf.pyx:34-35
    return (1. / lengthscale) * (1 - (1. / lengthscale) * (diff ** 2)) * k     # rbf.pyx:37


def function_derivative_params(X, beta, sigma, lengthscale):
    """sample/function.py:51-60 (FunctionDerivative._params): mu = K_ba K_a^-1 beta, cov = K_b - K_ba K_a^-1 K_ba^T with
    K_a^-1 = Kernel.K_inv (kernel.pyx:51-64)."""
    ka_inv = K_inv(rbf_K(X, sigma, lengthscale))[0]
    kb = rbf_dK(X, sigma, lengthscale)
    kba = rbf_dK(X, sigma, lengthscale, cross=True)
    mu = np.dot(kba, np.dot(ka_inv, beta))
    cov = kb - np.dot(kba, np.dot(ka_inv, kba.T))
    return mu, cov


def scalar_interval_threshold(samples, alpha, pi):
    """interval.py:21-35 (ScalarInterval.__init__): sorted densities, threshold epsilon[int(n alpha)]"""
    eps = sorted(pi(t) for t in samples)
    return eps[int(len(samples) * alpha)], eps


def function_interval(samples, alpha, start=1e-6, tol=1e-6, maxiter=100):
    """interval.py:62-115 (FunctionInterval.__init__), statement by statement.  Returns dict(alpha, mean, std, epsilon,
    lb, ub, evaluations)."""
    samples = np.asarray(samples, dtype=float)
    n = samples.shape[0]
    alphaPotential = 1. * np.arange(n) / n                                                      # :65
    alpha_eff = [v for v in alphaPotential if abs(v - alpha) == abs(alpha - alphaPotential).min()][0]   # :66
    mean = samples.mean(0)                                                                      # :68
    std = samples.std(0)                                                                        # :69
    lb, ub = mean - 2 * std, mean + 2 * std                                                     # :70
    evals = [0]

    def check(x):                                                                               # :75
        evals[0] += 1
        return 1. * sum((samples > lb - x).all(1) & (samples < ub + x).all(1)) / n
    epsilon = start
    while check(epsilon) < alpha_eff:                                                           # :80-82
        epsilon *= 2
    eLb, eUb = epsilon / 2, epsilon                                                             # :87
    i = 0
    while True:                                                                                 # :89-110
        if abs(check(eLb) - alpha) < tol:
            epsilon = eLb
            break
        elif abs(check(eUb) - alpha) < tol:
            epsilon = eUb
            break
        nb = (eLb + eUb) / 2
        if check(nb) < alpha:
            eLb = nb
        else:
            eUb = nb
        i += 1
        if i > maxiter:
            epsilon = eLb
            break
    return dict(alpha=alpha_eff, mean=mean, std=std, epsilon=epsilon, lb=mean - std - epsilon, ub=mean + std + epsilon,
                evaluations=evals[0])


def split_rhat(hist, split=True):
    """split-R^ (Gelman et al., BDA3 11.4) per column of hist (S, C, cols): every chain cut into its first and last
    S // 2 rows.  Not in the reference (one chain per process); restated here as the checker of gpf_chain_diagnostics.
    Returns (rhat, pooled mean, posterior variance estimate)."""
    S, C, cols = hist.shape
    if split and S >= 4:
        L = S // 2
        seqs = np.concatenate([hist[:L], hist[S - L:]], axis=1)       # (L, 2C, cols)
    else:
        L, seqs = S, hist
    means = seqs.mean(0)
    W = seqs.var(0, ddof=1).mean(0)
    Bn = means.var(0, ddof=1)
    vhat = (L - 1.0) / L * W + Bn
    return np.sqrt(vhat / W), means.mean(0), vhat
