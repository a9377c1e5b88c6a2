"""Device engine: C chains of one model on one GPU, driven through the C ABI (include/gpfanova_b200.h).

Host-side counterpart of Base.__init__'s state + SamplerContainer.parameter_cache (base.py:16-86,
container.py:8-20) with a leading chain axis.  Device memory is held in torch tensors; all arithmetic
happens in the sm_100a kernels behind the ctypes boundary.
"""
import ctypes as C

import os
import sys
import time

import numpy as np
import torch

from . import _capi

OFFSET = 1e-9          # kernel/kernel.pyx:7
PRIOR_JITTER = 1e-6    # base.py:222


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("gpfanova_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", device if device is not None else torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ----------------------------------------------------------------------------------------------
# stateless batched operators
# ----------------------------------------------------------------------------------------------
def rbf_K(x, log10_sigma, log10_lengthscale, device=None):
    """RBF._K (kernel/rbf.pyx:8-19) for a batch of (sigma, lengthscale) pairs given as log10 values
    (kernel.pyx:35-37).  x: (n, p).  Returns a (batch, n, n) float64 CUDA tensor."""
    dev = _require_cuda(device)
    x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    if x.dim() == 1:
        x = x[:, None]
    s = torch.as_tensor(np.atleast_1d(np.asarray(log10_sigma, dtype=np.float64))).to(dev)
    l = torch.as_tensor(np.atleast_1d(np.asarray(log10_lengthscale, dtype=np.float64))).to(dev)
    assert s.shape == l.shape
    n, p = x.shape
    K = torch.empty((s.numel(), n, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().gpf_rbf_K(_ptr(x), n, p, _ptr(s), _ptr(l), s.numel(), _ptr(K), C.c_void_p(st)))
    return K


def _status_raise(status):
    st = status.cpu().numpy()
    if (st == _capi.STATUS_NONPOS_DIAG).any():
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")      # linalg.py:43
    if (st < 0).any():
        raise np.linalg.LinAlgError("not positive definite, even with jitter.")    # linalg.py:54
    return st


def jitchol(A, maxtries=5, device=None, return_info=False):
    """linalg.jitchol (linalg.py:35-54) for a batch of matrices A: (batch, n, n) or (n, n).
    Returns L (lower, upper zeroed) as a CUDA tensor; raises numpy.linalg.LinAlgError like the reference."""
    dev = _require_cuda(device)
    A = torch.as_tensor(A, dtype=torch.float64).to(dev).contiguous()
    single = A.dim() == 2
    if single:
        A = A[None]
    b, n, _ = A.shape
    L = torch.empty_like(A)
    logdet = torch.empty(b, dtype=torch.float64, device=dev)
    status = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().gpf_jitchol(_ptr(A), n, b, maxtries, _ptr(L), _ptr(logdet), _ptr(status), C.c_void_p(st)))
    tries = _status_raise(status)
    if return_info:
        return (L[0] if single else L), (logdet[0] if single else logdet), (tries[0] if single else tries)
    return L[0] if single else L


def kinv(A, offset=OFFSET, device=None):
    """Kernel.K_inv (kernel/kernel.pyx:51-64) on given matrices: (A + offset I)^-1 through jitchol."""
    dev = _require_cuda(device)
    A = torch.as_tensor(A, dtype=torch.float64).to(dev).contiguous()
    single = A.dim() == 2
    if single:
        A = A[None]
    b, n, _ = A.shape
    out = torch.empty_like(A)
    status = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().gpf_kinv(_ptr(A), n, b, float(offset), _ptr(out), _ptr(status), C.c_void_p(st)))
    _status_raise(status)
    return out[0] if single else out


def _dev_f64(a, dev):
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)


def _x_and_params(x, device, *params):
    dev = _require_cuda(device)
    x = _dev_f64(x, dev)
    if x.dim() == 1:
        x = x[:, None]
    ps = [_dev_f64(np.atleast_1d(np.asarray(p_, dtype=np.float64)), dev) for p_ in params]
    for q in ps[1:]:
        assert q.shape == ps[0].shape
    return dev, x, ps


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def rbf_dist(x, log10_lengthscale, device=None):
    """RBF.dist (kernel/rbf.pyx:8-14) for a batch of lengthscales (log10): (batch, n, n) CUDA tensor of scaled distances"""
    dev, x, (l,) = _x_and_params(x, device, log10_lengthscale)
    n, p = x.shape
    out = torch.empty((l.numel(), n, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_rbf_dist(_ptr(x), n, p, _ptr(l), l.numel(), _ptr(out), _stream(dev)))
    return out


def rbf_dK(x, log10_sigma, log10_lengthscale, cross=False, device=None):
    """RBF._dK (kernel/rbf.pyx:21-38) for a batch of (sigma, lengthscale) pairs (log10): (batch, n, n) CUDA tensor"""
    dev, x, (s, l) = _x_and_params(x, device, log10_sigma, log10_lengthscale)
    n, p = x.shape
    out = torch.empty((s.numel(), n, n), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_rbf_dK(_ptr(x), n, p, _ptr(s), _ptr(l), int(bool(cross)), s.numel(), _ptr(out), _stream(dev)))
    return out


def gp_loglik(x, log10_sigma, log10_lengthscale, F, add_rel=0.0, add_abs=0.0, maxtries=5, device=None, return_info=False):
    """sum_f log N(F[b, f]; 0, K_b + (add_rel mean(K_b) + add_abs) I) for a batch of (sigma, lengthscale) pairs.
    F: (batch, nf, n) or (nf, n) shared by the whole batch.  Returns a (batch,) CUDA tensor; raises LinAlgError like
    linalg.jitchol if a covariance is not positive definite even with jitter."""
    dev, x, (s, l) = _x_and_params(x, device, log10_sigma, log10_lengthscale)
    n, p = x.shape
    b = s.numel()
    F = _dev_f64(F, dev)
    if F.dim() == 2:
        F = F[None].expand(b, -1, -1).contiguous()
    assert F.shape[0] == b and F.shape[2] == n
    ll = torch.empty(b, dtype=torch.float64, device=dev)
    ld = torch.empty(b, dtype=torch.float64, device=dev)
    status = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_gp_loglik(_ptr(x), n, p, _ptr(s), _ptr(l), b, _ptr(F), F.shape[1], float(add_rel), float(add_abs),
                                               int(maxtries), _ptr(ll), _ptr(ld), _ptr(status), _stream(dev)))
    tries = _status_raise(status)
    return (ll, ld, tries) if return_info else ll


def gp_sample(x, log10_sigma, log10_lengthscale, nsamp=1, z=None, seed=0, stream_id=0, add_rel=0.0, add_abs=0.0, maxtries=5,
              device=None):
    """nsamp draws from N(0, K_b + add I) for every (sigma, lengthscale) pair of the batch (base.py:249-254):
    L z with L = jitchol(K_b + add I).  z: (batch, nsamp, n) injected normals or None (Philox keyed by seed / stream_id).
    Returns a (batch, nsamp, n) CUDA tensor."""
    dev, x, (s, l) = _x_and_params(x, device, log10_sigma, log10_lengthscale)
    n, p = x.shape
    b = s.numel()
    zt = None
    if z is not None:
        zt = _dev_f64(z, dev)
        assert tuple(zt.shape) == (b, nsamp, n)
    out = torch.empty((b, int(nsamp), n), dtype=torch.float64, device=dev)
    status = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_gp_sample(_ptr(x), n, p, _ptr(s), _ptr(l), b, int(nsamp), _ptr(zt), int(seed) & (2 ** 64 - 1),
                                               int(stream_id) & 0xffffffff, float(add_rel), float(add_abs), int(maxtries), _ptr(out),
                                               _ptr(status), _stream(dev)))
    _status_raise(status)
    return out


def function_derivative(x, log10_sigma, log10_lengthscale, F, z=None, seed=0, stream_id=0, maxtries=5, device=None,
                        want_moments=False, raise_on_error=True):
    """FunctionDerivative (sample/function.py:43-66) for a batch of kernels: F (batch, nf, n) function values ->
    draws of their derivatives (batch, nf, n); want_moments: also (mu (batch, nf, n), cov (batch, n, n)).
    Where cov is not positive definite even with jitter (the reference's derivative kernels only form a covariance for
    lengthscales <= 1, rbf.pyx:34-37) the draws are NaN and, unless raise_on_error=False, LinAlgError is raised; the last
    element of the returned tuple is the per-item status."""
    dev, x, (s, l) = _x_and_params(x, device, log10_sigma, log10_lengthscale)
    n, p = x.shape
    b = s.numel()
    F = _dev_f64(F, dev)
    assert F.dim() == 3 and F.shape[0] == b and F.shape[2] == n
    nf = F.shape[1]
    zt = None if z is None else _dev_f64(z, dev)
    dF = torch.empty_like(F)
    mu = torch.empty_like(F) if want_moments else None
    cov = torch.empty((b, n, n), dtype=torch.float64, device=dev) if want_moments else None
    for f0 in range(0, nf, 32):                       # at most 32 functions ride along one factorisation
        f1 = min(nf, f0 + 32)
        Fc = F[:, f0:f1].contiguous()
        dc = torch.empty_like(Fc)
        mc = torch.empty_like(Fc) if want_moments else None
        zc = None if zt is None else zt[:, f0:f1].contiguous()
        status = torch.zeros(b, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().gpf_function_derivative(
                _ptr(x), n, p, _ptr(s), _ptr(l), b, _ptr(Fc), f1 - f0, _ptr(zc), int(seed) & (2 ** 64 - 1),
                (int(stream_id) + f0) & 0xffffffff, OFFSET, int(maxtries), _ptr(dc), _ptr(mc), _ptr(cov if f0 == 0 else None),
                _ptr(status), _stream(dev)))
        if raise_on_error:
            _status_raise(status)
        dF[:, f0:f1] = dc
        if want_moments:
            mu[:, f0:f1] = mc
    return (dF, mu, cov, status) if want_moments else (dF, status)


def compose_y(F, X, sd, seed=0, stream_id=0, device=None):
    """y[b] (n, m) = F[b]^T X^T + sd[b] * N(0, 1) noise (base.py:256-257) for a batch of function sets F (batch, f, n)"""
    dev = _require_cuda(device)
    F = _dev_f64(F, dev)
    X = _dev_f64(X, dev)
    b, f, n = F.shape
    m = X.shape[0]
    assert X.shape[1] == f
    sdt = _dev_f64(np.broadcast_to(np.asarray(sd, dtype=np.float64), (b,)), dev)
    y = torch.empty((b, n, m), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_compose_y(_ptr(F), _ptr(X), _ptr(sdt), b, n, m, f, int(seed) & (2 ** 64 - 1),
                                               int(stream_id) & 0xffffffff, _ptr(y), _stream(dev)))
    return y


def fanova_transform(F, blocks, device=None):
    """FANOVA's Transform samplers (fanova.py:58-81) on the device for many stored states at once.
    F: (rows, f, n).  blocks: list of (in0, coef (nout, nin), fpv) -- the block's contrast functions in0 .. in0+nin are
    mapped to nout effect rows, followed by the finite-population-variance row when fpv.  Returns (rows, Q, n) CUDA tensor."""
    dev = _require_cuda(device)
    F = _dev_f64(F, dev)
    rows, f, n = F.shape
    in0 = np.array([b[0] for b in blocks], dtype=np.int32)
    nin = np.array([np.asarray(b[1]).shape[1] for b in blocks], dtype=np.int32)
    nout = np.array([np.asarray(b[1]).shape[0] for b in blocks], dtype=np.int32)
    fpv = np.array([1 if b[2] else 0 for b in blocks], dtype=np.int32)
    coef = np.ascontiguousarray(np.concatenate([np.asarray(b[1], dtype=np.float64).ravel() for b in blocks])) if blocks else np.zeros(1)
    Q = int((nout + fpv).sum())
    out = torch.empty((rows, Q, n), dtype=torch.float64, device=dev)
    if rows and Q:
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().gpf_fanova_transform(_ptr(F), rows, f, n, len(blocks), in0.ctypes.data_as(C.c_void_p),
                                                          nin.ctypes.data_as(C.c_void_p), nout.ctypes.data_as(C.c_void_p),
                                                          fpv.ctypes.data_as(C.c_void_p), coef.ctypes.data_as(C.c_void_p), _ptr(out),
                                                          _stream(dev)))
    return out


def chain_diagnostics(hist, col0=0, ncol=None, split=True, device=None):
    """split-R^ over chains, pooled mean and posterior-variance estimate of the columns col0 .. col0+ncol of a stored
    history hist (S, C, cols) (the layout ChainEngine.sweep writes).  Returns three (ncol,) CUDA tensors."""
    dev = _require_cuda(device)
    hist = _dev_f64(hist, dev)
    S, Cn, cols = hist.shape
    ncol = cols - col0 if ncol is None else int(ncol)
    r, mu, var = (torch.empty(ncol, dtype=torch.float64, device=dev) for _ in range(3))
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_chain_diagnostics(_ptr(hist), S, Cn, cols, int(col0), ncol, int(bool(split)), _ptr(r), _ptr(mu),
                                                       _ptr(var), _stream(dev)))
    return r, mu, var


def row_sort(v, device=None):
    """ascending sort of every row of v (rows, S <= 8192) on the device; returns a new CUDA tensor"""
    dev = _require_cuda(device)
    v = _dev_f64(v, dev).clone()
    if v.dim() == 1:
        v = v[None]
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_row_sort(_ptr(v), v.shape[0], v.shape[1], _stream(dev)))
    return v


def function_interval(samples, alpha=.95, start=1e-6, tol=1e-6, maxiter=100, device=None):
    """FunctionInterval (interval.py:62-131) for a batch of sample sets (batch, S, n) or one (S, n).
    Returns (mean, std, epsilon, evaluations) as CUDA tensors."""
    dev = _require_cuda(device)
    X = _dev_f64(samples, dev)
    if X.dim() == 2:
        X = X[None]
    b, S, n = X.shape
    mean = torch.empty((b, n), dtype=torch.float64, device=dev)
    std = torch.empty((b, n), dtype=torch.float64, device=dev)
    eps = torch.empty(b, dtype=torch.float64, device=dev)
    it = torch.zeros(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().gpf_function_interval(_ptr(X), b, S, n, float(alpha), float(start), float(tol), int(maxiter),
                                                       _ptr(mean), _ptr(std), _ptr(eps), _ptr(it), _stream(dev)))
    return mean, std, eps, it


# ----------------------------------------------------------------------------------------------
# sampler context
# ----------------------------------------------------------------------------------------------
def _host_threads():
    """threads this process may use for host copies / page touching: half the cores, shared by the ranks of the box"""
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(8, (os.cpu_count() or 4) // (2 * local_world)))


_COPY_THREADS = max(1, int(os.environ.get("GPFANOVA_B200_COPY_THREADS", str(max(2, _host_threads())))))
_TOUCH_THREADS = max(1, int(os.environ.get("GPFANOVA_B200_TOUCH_THREADS", str(_host_threads()))))
_POOL = None


def _copy_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=_COPY_THREADS)
    return _POOL


_PREFAULT = None


def _prefault_pool():
    """page-touching threads (shared by the engines of an EngineGroup, which run one sweep_host per device)"""
    global _PREFAULT
    if _PREFAULT is None:
        from concurrent.futures import ThreadPoolExecutor
        _PREFAULT = ThreadPoolExecutor(max_workers=max(8, _TOUCH_THREADS))
    return _PREFAULT


def _env_options():
    out = []
    for item in os.environ.get("GPFANOVA_B200_OPTIONS", "").split(","):
        if "=" in item:
            k, v = item.split("=")
            out.append((int(k), int(v)))
    return out


class ChainEngine(object):
    """C chains of  y (n x m) = F (n x f) X^T + noise  on one GPU.

    hyper layout (log10, per chain): [y_sigma, prior0_sigma, prior0_lengthscale, prior1_sigma, ...].
    Sampler ids (Philox counter word, = position in the reference's sampler list, base.py:57-84):
    0 = y_sigma; then per group: its functions, prior{g}_sigma, prior{g}_lengthscale.
    """

    def __init__(self, x, y, X, groups, chains=1, seed=0, device=None, chain_offset=0, slice_w=None, slice_m=None,
                 prior_lb=-2.0, prior_ub=2.0):
        self.dev = _require_cuda(device)
        self.x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        if self.x.ndim == 1:
            self.x = self.x[:, None]
        self.y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
        self.X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        self.n, self.p = self.x.shape
        self.m = self.y.shape[1]
        self.f = self.X.shape[1]
        assert self.y.shape[0] == self.n and self.X.shape[0] == self.m
        self.groups = [list(g) for g in groups]
        self.P = len(self.groups)
        self.H = 1 + 2 * self.P
        self.C = int(chains)
        gof = np.full(self.f, -1, dtype=np.int32)
        for g, grp in enumerate(self.groups):
            for i in grp:
                gof[i] = g
        assert (gof >= 0).all(), "every function needs a prior group"
        self.group_of_fn = gof
        w = np.full(self.H, 0.1) if slice_w is None else np.ascontiguousarray(slice_w, dtype=np.float64)
        mm = np.full(self.H, 10, dtype=np.int32) if slice_m is None else np.ascontiguousarray(slice_m, dtype=np.int32)
        # sampler list (kind, arg) in the reference's order
        self.order = [("slice", 0)]
        for g, grp in enumerate(self.groups):
            self.order += [("function", i) for i in grp]
            self.order += [("slice", 1 + 2 * g), ("slice", 2 + 2 * g)]
        self.sampler_id = {ka: i for i, ka in enumerate(self.order)}
        d = _capi.ModelDesc()
        d.n, d.p, d.m, d.f, d.P = self.n, self.p, self.m, self.f, self.P
        d.chains, d.device, d.chain_offset, d.seed = self.C, self.dev.index, int(chain_offset), int(seed)
        d.h_x = self.x.ctypes.data_as(_capi.c_dp)
        d.h_y = self.y.ctypes.data_as(_capi.c_dp)
        d.h_X = self.X.ctypes.data_as(_capi.c_dp)
        d.h_group_of_fn = gof.ctypes.data_as(_capi.c_ip)
        d.h_slice_w = w.ctypes.data_as(_capi.c_dp)
        d.h_slice_m = mm.ctypes.data_as(_capi.c_ip)
        d.offset, d.prior_jitter, d.prior_lb, d.prior_ub = OFFSET, PRIOR_JITTER, float(prior_lb), float(prior_ub)
        self._lib = _capi.lib()
        h = C.c_void_p()
        _capi.check(self._lib.gpf_create(C.byref(h), C.byref(d)))
        self._h = h
        self.sweeps_done = 0
        self._dev_hist = None
        self._pinned = []                                     # result blocks from prepare_results()
        for opt, val in _env_options():                       # GPFANOVA_B200_OPTIONS="7=1,3=0": A/B runs of the same program
            self.set_option(opt, val)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _capi.check(rc, self._h)

    def sync(self):
        self._ck(self._lib.gpf_sync(self._h))

    def _new(self, *shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _dev_in(self, a, shape):
        if a is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(self.dev)
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
        torch.cuda.synchronize(self.dev)
        return t

    # -- state ---------------------------------------------------------------------------------
    def _pinned_state(self):
        if getattr(self, "_F_pin", None) is None:
            self._F_pin = torch.empty((self.C, self.f, self.n), dtype=torch.float64, pin_memory=True)
            self._h_pin = torch.empty((self.C, self.H), dtype=torch.float64, pin_memory=True)
        return self._F_pin.numpy(), self._h_pin.numpy()

    def set_state(self, F=None, hyper=None):
        """F: (C, f, n) or (f, n) broadcast to all chains; hyper: (C, H) or (H,).  Staged through pinned memory."""
        Fpin, hpin = self._pinned_state()
        Fp = hp = None
        if F is not None:
            F = np.asarray(F, dtype=np.float64)
            if not np.shares_memory(F, Fpin):
                np.copyto(Fpin, np.broadcast_to(F, Fpin.shape))
            Fp = C.c_void_p(self._F_pin.data_ptr())
        if hyper is not None:
            hyper = np.asarray(hyper, dtype=np.float64)
            if not np.shares_memory(hyper, hpin):
                np.copyto(hpin, np.broadcast_to(hyper, hpin.shape))
            hp = C.c_void_p(self._h_pin.data_ptr())
        self._ck(self._lib.gpf_set_state(self._h, Fp, hp))

    def get_state(self):
        """(F (C, f, n), hyper (C, H)) as numpy views of the engine's pinned staging buffers (refreshed by every call)"""
        Fpin, hpin = self._pinned_state()
        self._ck(self._lib.gpf_get_state(self._h, C.c_void_p(self._F_pin.data_ptr()), C.c_void_p(self._h_pin.data_ptr())))
        return Fpin, hpin

    def status(self, clear=True, raise_on_error=True):
        st = np.zeros(self.C, dtype=np.int32)
        rc = self._lib.gpf_get_status(self._h, st.ctypes.data_as(C.c_void_p), int(clear))
        if rc == _capi.ERR_NOT_PD:
            if raise_on_error:
                bad = np.nonzero(st < 0)[0]
                what = ("not pd: non-positive diagonal elements" if (st[bad] == _capi.STATUS_NONPOS_DIAG).any()
                        else "not positive definite, even with jitter.")
                raise np.linalg.LinAlgError("%s (chains %s)" % (what, bad[:8].tolist()))
        else:
            self._ck(rc)
        return st

    def counters(self, clear=False):
        c = np.zeros(8, dtype=np.uint64)
        self._ck(self._lib.gpf_get_counters(self._h, c.ctypes.data_as(C.c_void_p), int(clear)))
        return dict(zip(["chol", "logp", "jitter", "draws", "inverses", "slice_steps", "function_chol", "ko_factor"],
                        [int(v) for v in c[:8]]))

    # -- pieces of the sweep -------------------------------------------------------------------
    def group_kinv(self, g, want=True):
        out = self._new(self.C, self.n, self.n) if (want and g >= 0) else None
        self._ck(self._lib.gpf_group_kinv(self._h, int(g), _ptr(out)))
        self.sync()
        return out

    def function_step(self, fn, sweep=0, z=None, want=("mu", "LC", "mt", "nt")):
        """Function._sample for function fn of every chain; returns dict of requested intermediates (CUDA tensors).
        z: (C, 2n) injected standard normals [z1 | z2] or None (Philox).  "LC": Cholesky factor of
        C = I + S K_o S (Sigma = K_o - K_o S C^-1 S K_o, S = sqrt(n_t / (sigma_y + 1e-9)))."""
        zt = self._dev_in(z, (self.C, 2 * self.n))
        o = {k: None for k in ("mu", "LC", "mt", "nt")}
        for k in want:
            o[k] = self._new(self.C, self.n, self.n) if k == "LC" else self._new(self.C, self.n)
        sid = self.sampler_id[("function", fn)]
        self._ck(self._lib.gpf_function_step(self._h, int(fn), sid, int(sweep), _ptr(zt), _ptr(o["mu"]), _ptr(o["LC"]),
                                             _ptr(o["mt"]), _ptr(o["nt"])))
        self.sync()
        return {k: v for k, v in o.items() if v is not None}

    def obs_loglik(self, cand):
        ct = self._dev_in(np.broadcast_to(np.asarray(cand, dtype=np.float64), (self.C,)), (self.C,))
        out = self._new(self.C)
        self._ck(self._lib.gpf_obs_loglik(self._h, _ptr(ct), _ptr(out)))
        self.sync()
        return out.cpu().numpy()

    def prior_loglik(self, g, which, cand):
        """which: 'sigma' | 'lengthscale' (base.py:216-243)"""
        w = {"sigma": 0, "lengthscale": 1}[which]
        ct = self._dev_in(np.broadcast_to(np.asarray(cand, dtype=np.float64), (self.C,)), (self.C,))
        out = self._new(self.C)
        self._ck(self._lib.gpf_prior_loglik(self._h, int(g), w, _ptr(ct), _ptr(out)))
        self.sync()
        return out.cpu().numpy()

    def slice_step(self, h, sweep=0, u=None):
        """Slice._sample on hyper-parameter h of every chain; u: (C, k) injected uniforms or None (Philox).
        Returns the number of log-density evaluations per chain."""
        ut = None
        nu = 0
        if u is not None:
            u = np.ascontiguousarray(np.asarray(u, dtype=np.float64))
            nu = u.shape[1]
            ut = self._dev_in(u, (self.C, nu))
        nev = torch.zeros(self.C, dtype=torch.int32, device=self.dev)
        torch.cuda.synchronize(self.dev)
        sid = self.sampler_id[("slice", h)]
        self._ck(self._lib.gpf_slice_step(self._h, int(h), sid, int(sweep), _ptr(ut), nu, _ptr(nev)))
        self.sync()
        return nev.cpu().numpy()

    def sweep(self, n_sweeps=1, thin=0, history=None, order=None, sweep0=None):
        """SamplerContainer.sample (container.py:55-78) for every chain.  history: preallocated CUDA tensor
        (rows, C, H + f*n) or None.  Returns the number of rows stored."""
        if sweep0 is None:
            sweep0 = self.sweeps_done
        op = None
        if order is not None:
            order = np.ascontiguousarray(np.asarray(order, dtype=np.int32).reshape(n_sweeps, len(self.order)))
            op = order.ctypes.data_as(C.c_void_p)
        rows = C.c_longlong(0)
        cap = 0
        if history is not None:
            assert history.is_cuda and history.dtype == torch.float64 and history.is_contiguous()
            assert tuple(history.shape[1:]) == (self.C, self.H + self.f * self.n)
            cap = history.shape[0]
        self._ck(self._lib.gpf_sweep(self._h, int(n_sweeps), int(thin), int(sweep0), op, _ptr(history), cap, C.byref(rows)))
        self.sweeps_done = sweep0 + n_sweeps
        return int(rows.value)

    def set_option(self, option, value):
        """option 0: Toeplitz/Schur slice log-likelihood on uniform grids (1, default) or dense tile engine (0);
        1: batched launches (balanced designs: one fused kernel per sweep for all function draws); 2: shared factorisation
        for functions with identical conditional precision; 3: centrosymmetric split on uniform grids (even n >= 64): all
        n^3 work of the function step on the two half-size blocks; 4: observation likelihood about the
        least-squares fit (balanced designs, nothing missing); 5: function step by the O(n^2) Schur algorithm (uniform grid,
        balanced design, n <= 256, prior groups of at most four functions); 7: all sweeps of a call in one launch without
        device-wide barriers (chain-resident sweep kernel) where option 5 applies to every prior group (default 0); 8: a group's
        functions and slices back to back on the group's own stream (default 1); 9: chains in 1 or 2 independent cohorts of
        streams (default 1)"""
        self._ck(self._lib.gpf_set_option(self._h, int(option), int(value)))

    def get_option(self, option):
        """effective value of an option (5: O(n^2) Schur function step active; 3: split factors of K_o; 6: balanced design; 7: one-launch sweeps active)"""
        return int(self._lib.gpf_get_option(self._h, int(option)))

    def prepare_results(self, n_sweeps, thin=0):
        """Page-locked result memory for an upcoming sweep_host(n_sweeps, thin) call, allocated NOW (cudaHostAlloc costs
        ~0.4 s per GB -- too much to pay inside a sampling call, so nothing is pinned behind the caller's back).  A prepared
        block is consumed by the next sweep_host call it is large enough for: the stored rows then go device -> result by
        DMA alone.  Without one, results land in ordinary pageable memory through the staging pipeline."""
        step = 1 if thin <= 0 else int(thin)
        rows = int(n_sweeps) // step
        if rows <= 0:
            return 0
        cols = self.H + self.f * self.n
        nel = rows * self.C * cols
        self._pinned.append(torch.empty(nel, dtype=torch.float64, pin_memory=True))
        self._ensure_dev_hist(max(1, int(self._STAGE_BYTES // (self.C * cols * 8))), 4)     # (cudaMalloc belongs here too)
        return nel * 8

    _STAGE_BYTES = 96 << 20

    def _ensure_dev_hist(self, cap, nbuf):
        cols = self.H + self.f * self.n
        if self._dev_hist is None or self._dev_hist[0].shape[0] < cap or len(self._dev_hist) < nbuf:
            self._dev_hist = [torch.empty((cap, self.C, cols), dtype=torch.float64, device=self.dev) for _ in range(nbuf)]

    def _take_pinned(self, nel):
        best = None
        for k, t in enumerate(self._pinned):
            if t.numel() >= nel and (best is None or t.numel() < self._pinned[best].numel()):
                best = k
        return None if best is None else self._pinned.pop(best)

    def sweep_host(self, n_sweeps=1, thin=0, order=None, stage_bytes=_STAGE_BYTES, max_chunk_sweeps=None, on_chunk=None):
        """sweep() with the stored states delivered to the host as ONE numpy array (rows, C, H + f*n).

        Double-buffered pipeline: the stored rows of a chunk of sweeps are copied device -> PINNED staging buffer on the
        context's copy stream right behind the chunk's kernels (DMA speed, no pageable staging, beside the kernels of the
        next chunk), and drained into the result array by the host while the GPU already runs the next chunk.  The staging buffers (2 x <= stage_bytes) and the
        device history buffers are kept by the engine and reused by later calls."""
        step = 1 if thin <= 0 else int(thin)
        cols = self.H + self.f * self.n
        rows_total = n_sweeps // step
        order = None if order is None else np.asarray(order, dtype=np.int32).reshape(n_sweeps, len(self.order))
        if rows_total == 0:
            self.sweep(n_sweeps, thin=step, order=order)
            self.sync()
            return np.zeros((0, self.C, cols))
        rowbytes = self.C * cols * 8
        # rows per chunk: bounded by the staging size, and at most a quarter of the call so that the drain of one chunk
        # overlaps the kernels of the next (the last chunk's drain is the only one exposed)
        cap = max(1, int(stage_bytes // rowbytes))            # rows the staging buffers hold (allocated once, reused)
        crow = max(1, min(cap, -(-rows_total // 4)))
        if max_chunk_sweeps:
            crow = max(1, min(crow, int(max_chunk_sweeps) // step))
        pinned = self._take_pinned(rows_total * self.C * cols)
        nbuf = 4 if pinned is not None else 2                  # device history buffers (= chunks that may be in flight)
        self._ensure_dev_hist(cap, nbuf)
        t_begin = time.perf_counter()
        if pinned is not None:
            # ---- prepared result memory (prepare_results): every chunk goes device -> result by DMA on the copy stream; no
            # staging copy, no page faults, no host threads -- the path that scales when eight ranks share one host.  Four
            # device buffers: the host queues up to three chunks ahead of the GPU, so a descheduled host thread does not
            # leave the GPU idle between chunks
            out_t = pinned[:rows_total * self.C * cols].view(rows_total, self.C, cols)
            base_ptr = out_t.data_ptr()
            inflight, done, row0, i = [], 0, 0, 0
            while done < n_sweeps:
                k = min(crow * step, n_sweeps - done)
                rows = k // step
                b = i % 4
                if len(inflight) == 4:                         # buffer b is free once its copy (chunk i-4) has landed
                    self._ck(self._lib.gpf_wait_event(self._h, inflight.pop(0)))
                sub = None if order is None else order[done:done + k]
                if rows == 0:
                    self.sweep(k, thin=step, order=sub)
                else:
                    got = self.sweep(k, thin=step, history=self._dev_hist[b][:rows], order=sub)
                    self._ck(self._lib.gpf_memcpy_d2h_event(self._h, C.c_void_p(base_ptr + row0 * rowbytes),
                                                            _ptr(self._dev_hist[b]), got * rowbytes, b))
                    inflight.append(b)
                    row0 += got
                    i += 1
                done += k
                if on_chunk is not None:
                    on_chunk(done)
            for b in inflight:
                self._ck(self._lib.gpf_wait_event(self._h, b))
            self.sync()
            if os.environ.get("GPFANOVA_B200_TIMING"):
                sys.stderr.write("sweep_host: rows %d chunk rows %d  prepared (pinned) result  total %.3f s\n" % (
                    row0, crow, time.perf_counter() - t_begin))
            return out_t.numpy()[:row0]
        if getattr(self, "_stage", None) is None or self._stage[0].shape[0] < cap:
            self._stage = [torch.empty((cap, self.C, cols), dtype=torch.float64, pin_memory=True) for _ in range(2)]
        out = np.empty((rows_total, self.C, cols))
        # Fresh pageable memory: the first write to every 4 KB page faults (~2 GB/s for ONE thread, but the faults of
        # several threads run side by side: 26 GB/s with eight, tools/pagefault_probe.py).  Background threads touch the
        # pages in drain order (thread j takes blocks j, j + T, j + 2T, ...) while the GPU runs the first sweeps;
        # `touched()` is the length of the prefix every thread has passed.  A drain never writes a region the touchers
        # have not finished with, so their zeros can never land on copied data, and they are joined before the array is
        # returned.
        touchers = []
        touched = lambda: out.size
        if out.nbytes >= (64 << 20):
            flat = out.reshape(-1)
            blk = 1 << 20                                      # doubles per block (8 MB), in drain order
            nblk = -(-flat.size // blk)
            nthr = min(_TOUCH_THREADS, nblk)
            prog = [0] * nthr                                  # blocks finished by each thread

            def prefault(j):
                for bidx in range(j, nblk, nthr):
                    flat[bidx * blk:(bidx + 1) * blk:512] = 0.0
                    prog[j] += 1

            def touched():
                return min(flat.size, min(prog[j] * nthr + j for j in range(nthr)) * blk)
            touchers = [_prefault_pool().submit(prefault, j) for j in range(nthr)]
        pending = []                                           # (buffer index, first row, rows)
        done, row0, i = 0, 0, 0

        tm = {"wait": 0.0, "copy": 0.0, "launch": 0.0} if os.environ.get("GPFANOVA_B200_TIMING") else None

        def drain():
            b, r0, nr = pending.pop(0)
            t0 = time.perf_counter()
            self._ck(self._lib.gpf_wait_event(self._h, b))
            t1 = time.perf_counter()
            _drain_copy(b, r0, nr)
            if tm is not None:
                tm["wait"] += t1 - t0
                tm["copy"] += time.perf_counter() - t1

        def _drain_copy(b, r0, nr):
            end = (r0 + nr) * self.C * cols
            while touched() < end:                             # the page touchers are still inside this region
                time.sleep(0.0002)
            dst = out[r0:r0 + nr].reshape(-1)
            src = self._stage[b].numpy()[:nr].reshape(-1)
            # the copy into fresh pageable memory is page-fault bound (~2 GB/s per thread): spread it over a few threads
            # (numpy releases the GIL while copying)
            nt = _COPY_THREADS if dst.size >= (1 << 20) else 1
            if nt <= 1:
                np.copyto(dst, src)
            else:
                cut = [dst.size * j // nt for j in range(nt + 1)]
                list(_copy_pool().map(lambda j: np.copyto(dst[cut[j]:cut[j + 1]], src[cut[j]:cut[j + 1]]), range(nt)))

        while done < n_sweeps:
            k = min(crow * step, n_sweeps - done)
            rows = k // step
            b = i % 2
            if len(pending) == 2:
                drain()                                        # frees buffer b (chunk i-2) while chunk i-1 runs on the GPU
            sub = None if order is None else order[done:done + k]
            if rows == 0:
                self.sweep(k, thin=step, order=sub)
            else:
                got = self.sweep(k, thin=step, history=self._dev_hist[b][:rows], order=sub)
                self._ck(self._lib.gpf_memcpy_d2h_event(self._h, C.c_void_p(self._stage[b].data_ptr()),
                                                        _ptr(self._dev_hist[b]), got * rowbytes, b))
                pending.append((b, row0, got))
                row0 += got
                i += 1
            done += k
            if on_chunk is not None:
                on_chunk(done)
        while pending:
            drain()
        for t in touchers:
            t.result()                                         # never leave a writer behind the returned array
        self.sync()
        if tm is not None:
            sys.stderr.write("sweep_host: rows %d chunk rows %d  event wait %.3f s  copy %.3f s  total %.3f s\n" % (
                row0, crow, tm["wait"], tm["copy"], time.perf_counter() - t_begin))
        return out[:row0]

    def set_profile(self, on=1):
        """0 off; 1 per-kernel + total CUDA-event timing of the next sweep() calls; 2 total only"""
        self._ck(self._lib.gpf_set_profile(self._h, int(on)))

    def profile(self):
        ms = np.zeros(6)
        nl = C.c_ulonglong(0)
        self._ck(self._lib.gpf_get_profile(self._h, ms.ctypes.data_as(C.c_void_p), C.byref(nl)))
        return dict(zip(["total", "kinv", "function", "prior_slice", "y_sigma", "store"], ms.tolist())), int(nl.value)


def split_chains(chains, parts):
    """contiguous blocks of chains for `parts` devices (or ranks): (counts, offsets); the first chains % parts blocks get
    one chain more.  The offset of a block is the `chain_offset` of its engine (Philox is keyed by the global chain index)."""
    chains, parts = int(chains), int(parts)
    if not 1 <= parts <= chains:
        raise ValueError("need 1 <= number of devices <= chains")
    base, rem = divmod(chains, parts)
    counts = [base + (1 if i < rem else 0) for i in range(parts)]
    offsets = [sum(counts[:i]) for i in range(parts)]
    return counts, offsets


class EngineGroup(object):
    """The chains of one model spread over several GPUs of one box: one ChainEngine (context + stream) per device,
    contiguous blocks of chains, no inter-GPU communication on the sampling path; the stored states are gathered on the
    host.  Same interface as ChainEngine for everything Base uses.  Because the Philox stream is keyed by the GLOBAL
    chain index, results do not depend on the number of devices."""

    def __init__(self, x, y, X, groups, chains, seed, devices, **kw):
        devices = list(devices)
        self.counts, self.offsets = split_chains(chains, len(devices))
        self.engines = [ChainEngine(x, y, X, groups, chains=c, seed=seed, device=d, chain_offset=o, **kw)
                        for c, d, o in zip(self.counts, devices, self.offsets)]
        e0 = self.engines[0]
        self.C = int(chains)
        self.x, self.y, self.X = e0.x, e0.y, e0.X
        self.n, self.p, self.m, self.f, self.P, self.H = e0.n, e0.p, e0.m, e0.f, e0.P, e0.H
        self.order, self.sampler_id, self.dev = e0.order, e0.sampler_id, e0.dev

    def _slices(self):
        return [slice(o, o + c) for o, c in zip(self.offsets, self.counts)]

    @property
    def sweeps_done(self):
        return self.engines[0].sweeps_done

    def close(self):
        for e in self.engines:
            e.close()

    def sync(self):
        for e in self.engines:
            e.sync()

    def set_option(self, option, value):
        for e in self.engines:
            e.set_option(option, value)

    def get_option(self, option):
        return self.engines[0].get_option(option)

    def set_state(self, F=None, hyper=None):
        for e, s in zip(self.engines, self._slices()):
            Fe = None if F is None else np.broadcast_to(np.asarray(F, dtype=np.float64), (self.C, self.f, self.n))[s]
            he = None if hyper is None else np.broadcast_to(np.asarray(hyper, dtype=np.float64), (self.C, self.H))[s]
            e.set_state(F=Fe, hyper=he)

    def get_state(self):
        parts = [e.get_state() for e in self.engines]
        return np.concatenate([p[0] for p in parts], axis=0), np.concatenate([p[1] for p in parts], axis=0)

    def status(self, clear=True, raise_on_error=True):
        st = np.concatenate([e.status(clear=clear, raise_on_error=False) for e in self.engines])
        if raise_on_error and (st < 0).any():
            bad = np.nonzero(st < 0)[0]
            what = ("not pd: non-positive diagonal elements" if (st[bad] == _capi.STATUS_NONPOS_DIAG).any()
                    else "not positive definite, even with jitter.")
            raise np.linalg.LinAlgError("%s (chains %s)" % (what, bad[:8].tolist()))
        return st

    def counters(self, clear=False):
        out = {}
        for e in self.engines:
            for k, v in e.counters(clear=clear).items():
                out[k] = out.get(k, 0) + v
        return out

    def prepare_results(self, n_sweeps, thin=0):
        return sum(e.prepare_results(n_sweeps, thin=thin) for e in self.engines)

    def sweep_host(self, n_sweeps=1, thin=0, order=None, on_chunk=None, **kw):
        from concurrent.futures import ThreadPoolExecutor

        def run(i):
            return self.engines[i].sweep_host(n_sweeps, thin=thin, order=order, on_chunk=on_chunk if i == 0 else None, **kw)
        with ThreadPoolExecutor(max_workers=len(self.engines)) as ex:       # ctypes calls release the GIL
            parts = list(ex.map(run, range(len(self.engines))))
        return np.concatenate(parts, axis=1)

    def sweep(self, n_sweeps=1, thin=0, history=None, order=None, sweep0=None):
        assert history is None, "device-side history is per engine: use sweep_host() or the engines directly"
        for e in self.engines:
            e.sweep(n_sweeps, thin=thin, order=order, sweep0=sweep0)
        return 0

    # per-sampler entry points: run on every device, report the engine holding chain 0 / concatenated per-chain values
    def function_step(self, fn, sweep=0, z=None, want=("mu", "LC", "mt", "nt")):
        outs = []
        for e, s in zip(self.engines, self._slices()):
            outs.append(e.function_step(fn, sweep=sweep, z=None if z is None else np.asarray(z)[s], want=want))
        return outs[0]

    def slice_step(self, h, sweep=0, u=None):
        return np.concatenate([e.slice_step(h, sweep=sweep, u=None if u is None else np.asarray(u)[s])
                               for e, s in zip(self.engines, self._slices())])

    def obs_loglik(self, cand):
        cand = np.broadcast_to(np.asarray(cand, dtype=np.float64), (self.C,))
        return np.concatenate([e.obs_loglik(cand[s]) for e, s in zip(self.engines, self._slices())])

    def prior_loglik(self, g, which, cand):
        cand = np.broadcast_to(np.asarray(cand, dtype=np.float64), (self.C,))
        return np.concatenate([e.prior_loglik(g, which, cand[s]) for e, s in zip(self.engines, self._slices())])
