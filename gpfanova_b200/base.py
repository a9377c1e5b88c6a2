"""Base model  y(t) = X b(t) + noise  -- the API of gpfanova/base.py:7-259 with the Gibbs sweep on the GPU.

What the reference fixes and this class keeps: the constructor arguments, the ORDER of the samplers (base.py:57-84:
`y_sigma`; per prior group its functions (each followed by its derivative when derivatives=True), `prior{g}_sigma`,
`prior{g}_lengthscale`; then the subclass's extra samplers), the label scheme `f{i}(x_t)` (base.py:114-138) and with
them the columns of `parameter_cache` / `parameter_history`.  `sample()` runs the fused device sweep (gpf_sweep) for
`chains` independent chains instead of looping over Python samplers; `parameter_cache` / `parameter_history` describe
chain 0, `chain_state()` / `chain_history()` / `save(f, chain=c)` every chain.

Additive keyword arguments: chains=1, seed=None, device=None, devices=None (GPU ordinals: the chains are sharded over
them, no inter-GPU communication), store_chains=True.
"""
import logging
import os
import random as pyrandom
import sys
import time

import numpy as np
import pandas as pd
import torch

from . import engine as _engine
from .kernel import RBF, White
from .sample import Function, FunctionDerivative, SamplerContainer, Slice

LOG = logging.getLogger(__name__)
DEFAULT_SLICE = (.1, 10)          # (w, m) of every hyper-parameter slice sampler, base.py:52,75,79


class Base(SamplerContainer):
    """Subclasses implement _buildDesignMatrix and priorGroups (base.py:7-14)."""

    def __init__(self, x, y, hyperparam_kwargs={}, derivatives=False, chains=1, seed=None, device=None, devices=None,
                 store_chains=True, *args, **kwargs):
        self.x, self.y = x, y
        self.n, self.p = x.shape
        assert y.shape[0] == self.n, 'x and y must have same first dimension shape!'
        self.m = y.shape[1]
        self.chains = int(chains)
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.device = device
        self.devices = None if devices is None else list(devices)
        self.store_chains = store_chains
        self.derivatives = bool(derivatives)
        self._observationIndexBaseList = None

        self.design_matrix = None
        self.buildDesignMatrix()
        self.f = self.design_matrix.shape[1]
        rank = np.linalg.matrix_rank(self.design_matrix)
        if self.f > rank:
            LOG.error("design matrix is of rank %d, but there are %d functions!" % (rank, self.f))
        if np.isnan(self.y).any():
            LOG.error("NaN values in observation matrix, this is not supported!")

        self._groups = [list(g) for g in self.priorGroups()]
        samplers = self._core_samplers(hyperparam_kwargs)
        self._n_core_samplers = len(samplers)
        samplers += self._additionalSamplers()

        self.H = len(self._hyper_names)
        self._flabels = [lab for f in range(self.f) for lab in self.functionIndex(f)]
        self._dlabels = [lab for f in range(self.f) for lab in self.functionIndex(f, derivative=True)] if self.derivatives else []
        self._eng = None
        self._eng_key = None
        self._manual_calls = 0
        self._chain_blocks = []
        self._deriv_blocks = []
        self._F = np.zeros((self.chains, self.f, self.n))
        self._hyper = np.zeros((self.chains, self.H))

        SamplerContainer.__init__(self, samplers, **kwargs)
        # initial values given by sampler name (container.py:16-18) apply to every chain
        self._F[:] = self.parameter_cache[self._flabels].values.astype(float).reshape(self.f, self.n)[None]
        self._hyper[:] = self.parameter_cache[self._hyper_names].values.astype(float)[None]
        self._transform_labels = [lab for s in self.samplers if s.type == 'Transform' for lab in s.parameters]
        pos = {lab: i for i, lab in enumerate(self.parameter_cache.index)}
        self._pos_h = np.array([pos[l] for l in self._hyper_names], dtype=np.int64)
        self._pos_F = np.array([pos[l] for l in self._flabels], dtype=np.int64)
        self._pos_D = np.array([pos[l] for l in self._dlabels], dtype=np.int64)
        self._pos_T = np.array([pos[l] for l in self._transform_labels], dtype=np.int64)

    def _core_samplers(self, hyperparam_kwargs):
        """the reference's sampler list up to the subclass's extras; also fills y_k, kernels, _hyper_names, _slice_wm"""
        def wm(*keys):
            for k in keys:
                if k in hyperparam_kwargs:
                    return tuple(hyperparam_kwargs[k])
            return DEFAULT_SLICE
        names = self.functionNames()
        self.y_k = White(self, ['y_sigma'], logspace=True)
        self.kernels = []
        self._hyper_names = ['y_sigma']
        self._slice_wm = [wm('y_sigma', 'sigma')]
        out = [Slice('y_sigma', 'y_sigma', lambda v: self.observationLikelihood(sigma=v, prior_lb=-2, prior_ub=2),
                     *self._slice_wm[0], base=self, hyper_index=0)]
        for g, members in enumerate(self._groups):
            ps, pl = 'prior%d_sigma' % g, 'prior%d_lengthscale' % g
            kern = RBF(self, [ps, pl], logspace=True)
            self.kernels.append(kern)
            for f in members:
                nm = names.get(f, "f%d" % f)
                out.append(Function(nm, self.functionIndex(f), self, f, kern))
                if self.derivatives:
                    out.append(FunctionDerivative('d' + nm, self.functionIndex(f, derivative=True), self, f, kern))
            for key, pname, which in (('sigma', ps, 'sigma'), ('lengthscale', pl, 'lengthscale')):
                self._slice_wm.append(wm(key))
                out.append(Slice(pname, pname, lambda v, g=g, which=which: self.prior_likelihood(
                    p=g, prior_lb=-2, prior_ub=2, **{which: v}), *self._slice_wm[-1], base=self,
                    hyper_index=len(self._hyper_names)))
                self._hyper_names.append(pname)
        return out

    # ------------------------------------------------------------------ reference protocol (base.py:88-149)
    def _additionalSamplers(self):
        return []

    def functionNames(self):
        return {}

    def buildDesignMatrix(self):
        if self.design_matrix is None:
            self.design_matrix = self._buildDesignMatrix()

    def _buildDesignMatrix(self):
        raise NotImplementedError("Implement a design matrix for your model!")

    def priorGroups(self):
        raise NotImplementedError("Implement a prior grouping function for your model!")

    def _observationIndexBase(self):
        """the time-point part of every label: str() of the row of x (base.py:129-138)"""
        if self._observationIndexBaseList is None:
            self._observationIndexBaseList = [str(z) for z in self.x]
        return self._observationIndexBaseList

    def functionIndex(self, i, derivative=False, *args, **kwargs):
        stem = 'df%d' if derivative else 'f%d'
        return [(stem % i) + '(%s)' % z for z in self._observationIndexBase()]

    def functionPrior(self, f):
        return next((g for g, members in enumerate(self.priorGroups()) if f in members), -1)

    def _priorParameters(self, i):
        if i < 0:
            return ['y_sigma']
        if i >= len(self.priorGroups()):
            return [None]
        return ['prior%d_sigma' % i, 'prior%d_lengthscale' % i]

    def offset(self):
        return 1e-9

    # ------------------------------------------------------------------ state accessors (base.py:151-198), chain 0
    def functionMatrix(self, remove=[], only=[], derivative=False):
        """current function values from parameter_cache, (n, #functions); removed functions are zero columns"""
        which = list(only) if len(only) > 0 else list(range(self.f))
        out = np.zeros((self.n, len(which)))
        for col, f in enumerate(which):
            if len(only) > 0 or f not in remove:
                out[:, col] = self.parameter_cache[self.functionIndex(f, derivative=derivative)]
        return out

    def functionSamples(self, f, *args, **kwargs):
        return self.parameter_history[self.functionIndex(f, *args, **kwargs)]

    def residual(self, remove=[], only=[]):
        return self.y.T - self.design_matrix.dot(self.functionMatrix(remove, only).T)

    def functionResidual(self, f):
        load = self.design_matrix[:, f]
        keep = load != 0
        return self.residual(remove=[f])[keep] / load[keep, None]

    def observationMean(self):
        return self.design_matrix.dot(self.functionMatrix().T).ravel()

    # ------------------------------------------------------------------ device plumbing
    def engine(self):
        """the ChainEngine for the current (x, y, design matrix); rebuilt if the caller replaced self.y / self.x
        (the reference's convergence tests overwrite m.y after construction, test_convergence.py:60-62)"""
        key = (id(self.x), id(self.y), self.x.shape, self.y.shape)
        stale = self._eng is None or key != self._eng_key
        if not stale and not (np.array_equal(self._eng.y, self.y, equal_nan=True) and np.array_equal(self._eng.x, self.x)):
            stale = True
        if stale:
            done = 0
            if self._eng is not None:
                done = self._eng.sweeps_done      # the Philox stream (keyed by the sweep number) continues across rebuilds
                self._eng.close()
            w = [a for a, _ in self._slice_wm]
            m = [b for _, b in self._slice_wm]
            if self.devices is not None and len(self.devices) > 1:
                self._eng = _engine.EngineGroup(self.x, self.y, self.design_matrix, self._groups, self.chains, self.seed,
                                                self.devices, slice_w=w, slice_m=m)
            else:
                dev = self.device if self.devices is None else self.devices[0]
                self._eng = _engine.ChainEngine(self.x, self.y, self.design_matrix, self._groups, chains=self.chains,
                                                seed=self.seed, device=dev, slice_w=w, slice_m=m)
            self._eng_key = key
            for e in getattr(self._eng, "engines", [self._eng]):
                e.sweeps_done = done
        return self._eng

    def _push_state(self):
        """parameter_cache (chain 0) -> host mirror -> device.  Entries of parameter_cache the caller changed since
        the last sync are applied to every chain."""
        eng = self.engine()
        vals = self.parameter_cache.to_numpy(dtype=float)
        ch = vals[self._pos_h]
        cF = vals[self._pos_F].reshape(self.f, self.n)
        dh = ch != self._hyper[0]
        if dh.any():
            self._hyper[:, dh] = ch[dh][None]
        dF = cF != self._F[0]
        if dF.any():
            self._F[:, dF] = cF[dF][None]
        eng.set_state(F=self._F, hyper=self._hyper)
        return eng

    def _pull_state(self):
        eng = self.engine()
        self._F, self._hyper = eng.get_state()
        vals = self.parameter_cache.to_numpy(dtype=float, copy=True)
        vals[self._pos_h] = self._hyper[0]
        vals[self._pos_F] = self._F[0].ravel()
        if len(self._pos_T):
            vals[self._pos_T] = self._device_transforms(np.ascontiguousarray(self._F[0][None]))[0]
        self.parameter_cache.iloc[:] = vals

    def set_chain_state(self, hyper=None, F=None):
        """per-chain state: hyper (chains, H) in the order of self.hyperNames(); F (chains, f, n)"""
        vals = self.parameter_cache.to_numpy(dtype=float, copy=True)
        if hyper is not None:
            self._hyper = np.array(np.broadcast_to(np.asarray(hyper, dtype=float), (self.chains, self.H)))
            vals[self._pos_h] = self._hyper[0]
        if F is not None:
            self._F = np.array(np.broadcast_to(np.asarray(F, dtype=float), (self.chains, self.f, self.n)))
            vals[self._pos_F] = self._F[0].ravel()
        self.parameter_cache.iloc[:] = vals

    def chain_state(self):
        """(hyper (chains, H), F (chains, f, n)) of every chain after the last sample() call"""
        return np.array(self._hyper), np.array(self._F)

    def hyperNames(self):
        return list(self._hyper_names)

    def _next_manual_sweep(self):
        self._manual_calls += 1
        return 0x80000000 + self._manual_calls

    def _device_function_step(self, f, z=None, want_conditional=False):
        eng = self._push_state()
        zz = None if z is None else np.broadcast_to(np.asarray(z, dtype=float), (self.chains, 2 * self.n))
        out = eng.function_step(f, sweep=self._next_manual_sweep(), z=zz, want=("mu", "LC", "nt") if want_conditional else ())
        eng.status()
        self._pull_state()
        if want_conditional:
            return self._conditional_from_device(f, out)
        return self._F[0, f].copy()

    def _conditional_from_device(self, f, out):
        """(mu, Sigma) of function f's conditional (chain 0) from the device quantities: Sigma = K_o - W^T W with
        W = L_C^-1 S K_o (C = I + S K_o S, S = sqrt(n_t / (sigma_y + 1e-9)), K_o = K + 1e-9 I)"""
        mu = out["mu"][0].cpu().numpy()
        LC = out["LC"][0]
        nt = out["nt"][0]
        yinv = 1.0 / (pow(10.0, float(self._hyper[0, 0])) + self.offset())
        Ko = torch.as_tensor(self.kernels[self.functionPrior(f)].K(self.x) + self.offset() * np.eye(self.n)).to(LC.device)
        sd = torch.sqrt(nt * yinv)
        Wm = torch.linalg.solve_triangular(LC, sd[:, None] * Ko, upper=False)
        return mu, (Ko - Wm.T @ Wm).cpu().numpy()

    def _device_slice_step(self, h, x):
        self.parameter_cache[self._hyper_names[h]] = x
        eng = self._push_state()
        eng.slice_step(h, sweep=self._next_manual_sweep())
        eng.status()
        self._pull_state()
        return float(self._hyper[0, h])

    def _group_log10(self, hyper, g):
        return hyper[..., 1 + 2 * g], hyper[..., 2 + 2 * g]

    def _device_derivatives(self, hyper, F, stream_id, z=None, want_moments=False):
        """FunctionDerivative draws (sample/function.py:43-66) for states hyper (R, H), F (R, f, n): (R, f, n) array of
        beta'_f ~ N(K_ba K_a^-1 beta_f, K_b - K_ba K_a^-1 K_ba^T) with the group's kernel at that state"""
        R = hyper.shape[0]
        dF = np.zeros((R, self.f, self.n))
        moments = {}
        for g, members in enumerate(self._groups):
            ls, ll = self._group_log10(hyper, g)
            zz = None if z is None else np.ascontiguousarray(z[:, members, :])
            out = _engine.function_derivative(self.x, ls, ll, F[:, members, :], z=zz, seed=self.seed, stream_id=stream_id + g,
                                              device=self._dev(), want_moments=want_moments, raise_on_error=False)
            dF[:, members, :] = out[0].cpu().numpy()
            if (out[-1] < 0).any() and not getattr(self, "_deriv_warned", False):
                self._deriv_warned = True
                LOG.error("FunctionDerivative: covariance not positive definite (prior %d); the reference's derivative kernels "
                          "(rbf.pyx:34-37) only form a covariance for lengthscales <= 1 -- derivative draws are NaN there" % g)
            if want_moments:
                moments[g] = (out[1].cpu().numpy(), out[2].cpu().numpy())
        return (dF, moments) if want_moments else dF

    def _dev(self):
        return self.device if self.devices is None else self.devices[0]

    # ------------------------------------------------------------------ log densities (base.py:200-243)
    def observationLikelihood(self, sigma=None, prior_ub=None, prior_lb=None):
        """log p(y | functions, sigma) + log U(prior_lb, prior_ub)(sd), sd = sqrt(10**sigma) (base.py:200-214); chain 0"""
        eng = self._push_state()
        self._check_prior_bounds(eng, prior_lb, prior_ub)
        return float(eng.obs_loglik(float(sigma))[0])

    def prior_likelihood(self, p, sigma=None, lengthscale=None, prior_ub=None, prior_lb=None):
        """log-likelihood of the functions of prior group p under candidate sigma or lengthscale (base.py:216-243)"""
        if lengthscale is None and sigma is None:
            raise ValueError("must provide lengthscale or sigma in likelihood fxn")
        eng = self._push_state()
        self._check_prior_bounds(eng, prior_lb, prior_ub)
        if lengthscale is not None and sigma is not None:
            # both given: evaluate with sigma installed as the current value (the prior term is on the lengthscale)
            orig = np.array(self._hyper)          # (self._hyper may alias the engine's staging buffer)
            hyper = orig.copy()
            hyper[:, 1 + 2 * p] = sigma
            eng.set_state(hyper=hyper)
            val = eng.prior_loglik(p, "lengthscale", float(lengthscale))[0]
            eng.set_state(hyper=orig)
            self._hyper = orig
        elif lengthscale is not None:
            val = eng.prior_loglik(p, "lengthscale", float(lengthscale))[0]
        else:
            val = eng.prior_loglik(p, "sigma", float(sigma))[0]
        st = eng.status(raise_on_error=False)
        if (st < 0).any():
            LOG.error("prior likelihood LinAlgError (%d)" % p)
        return float(val)

    @staticmethod
    def _check_prior_bounds(eng, lb, ub):
        if (lb is not None and float(lb) != -2.0) or (ub is not None and float(ub) != 2.0):
            raise NotImplementedError("the device log-densities are built with the reference's U(-2, 2) hyper-priors "
                                      "(base.py:57,78,83)")

    def samplePrior(self, size=None, seed=None):
        """draw functions from their GP priors at the current hyper-parameters and data from the model (base.py:245-259),
        on the device: one batched launch factors the f kernel matrices (jitchol, i.e. with the jitter ladder: K alone
        is numerically singular) and draws L z, a second one forms y = X f + noise.  Returns (y (n, m), functions (n, f))
        like the reference; with size=B, B independent data sets at once: (B, n, m), (B, n, f).
        Normals are Philox4x32-10 keyed by `seed` (default: drawn from numpy's global stream, which the reference uses)."""
        B = 1 if size is None else int(size)
        seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        grp = np.array([self.functionPrior(i) for i in range(self.f)])
        cur = self.parameter_cache
        ls = np.array([float(cur['prior%d_sigma' % g]) for g in grp])
        ll = np.array([float(cur['prior%d_lengthscale' % g]) for g in grp])
        F = _engine.gp_sample(self.x, ls, ll, nsamp=B, seed=seed, stream_id=1, device=self._dev())          # (f, B, n)
        F = F.permute(1, 0, 2).contiguous()                                                              # (B, f, n)
        sd = np.sqrt(pow(10.0, float(cur['y_sigma'])))
        y = _engine.compose_y(F, self.design_matrix, sd, seed=seed, stream_id=2, device=self._dev())      # (B, n, m)
        y, F = y.cpu().numpy(), F.permute(0, 2, 1).cpu().numpy()
        return (y[0], F[0]) if size is None else (y, F)

    # ------------------------------------------------------------------ the sweep (container.py:29-92)
    def _transform_plan(self):
        return []

    def _transform_block(self, F):
        """values of the Transform samplers for stored states F (S, f, n) -> (S, #transform labels)"""
        return np.zeros((F.shape[0], 0))

    def _device_transforms(self, F):
        """Transform values for states F (R, f, n) on the device (gpf_fanova_transform) -> (R, #transform labels)"""
        plan = self._transform_plan()
        if not plan or F.shape[0] == 0:
            return np.zeros((F.shape[0], 0))
        return _engine.fanova_transform(F, plan, device=self._dev()).reshape(F.shape[0], -1).cpu().numpy()

    def _core_order(self, n_sweeps):
        """random=True: shuffle the reference's full sampler list per sweep (container.py:34-35) and keep the relative
        order of the samplers the device runs (slices and functions)"""
        dev_ids = {i: k for k, i in enumerate(i for i, s in enumerate(self.samplers[:self._n_core_samplers])
                                              if s.type != 'Function' or not getattr(s, 'derivative', False))}
        order = np.empty((n_sweeps, len(dev_ids)), dtype=np.int32)
        ids = list(range(len(self.samplers)))
        for s in range(n_sweeps):
            pyrandom.shuffle(ids)
            order[s] = [dev_ids[i] for i in ids if i in dev_ids]
        return order

    def prepare(self, n=1, thin=0):
        """Optional: allocate page-locked result memory for an upcoming sample(n, thin) call (additive; the reference has
        no counterpart).  The stored states of that call then reach the host by DMA alone -- no staging copy, no page
        faults -- which is what keeps several GPUs of one box from queueing on the host's memory system.  Returns the
        bytes pinned."""
        return self.engine().prepare_results(int(n), thin=int(thin))

    def sample(self, n=1, thin=0, verbose=False, random=False):
        """n sweeps of every chain on the device; every `thin`-th state is stored (thin=0: every sweep)."""
        eng = self._push_state()
        start_time = time.time()
        step = 1 if thin == 0 else int(thin)
        order = self._core_order(n) if random else None
        err = None
        progress = None
        if verbose:
            def progress(done):
                LOG.info("%d/%d iterations (%.2lf%s) finished in %.2lf minutes" % (
                    done, n, 100. * done / n, '%', (time.time() - start_time) / 60))
        t_sw = time.time()
        sweep0 = eng.sweeps_done
        blocks = eng.sweep_host(n, thin=step, order=order, on_chunk=progress,
                                max_chunk_sweeps=max(step, n // 10) if verbose else None)
        t_sw1 = time.time()
        try:
            eng.status()
        except np.linalg.LinAlgError as e:        # keep what was sampled, then propagate like the reference
            err = e
        self._pull_state()
        if blocks.shape[0]:
            self._store_blocks(blocks, sweep0)
        if os.environ.get("GPFANOVA_B200_TIMING"):
            sys.stderr.write("sample: push %.3f s  sweep_host %.3f s  pull+store %.3f s\n" % (
                t_sw - start_time, t_sw1 - t_sw, time.time() - t_sw1))
        if verbose:
            LOG.info("%d samples finished in %.2lf minutes" % (n, (time.time() - start_time) / 60))
        if err is not None:
            raise err

    def _store_blocks(self, raw, sweep0=0):
        """raw: (S, chains, H + f n) device history -> parameter_history rows (chain 0) in the reference's column order;
        with derivatives=True the derivative functions of the stored states are drawn here (they feed nothing back)"""
        dblock = None
        if self.derivatives:
            keep = slice(None) if (self.store_chains and self.chains > 1) else slice(0, 1)
            sub = raw[:, keep, :]
            S, Cn = sub.shape[:2]
            flat = sub.reshape(S * Cn, -1)
            dblock = self._device_derivatives(flat[:, :self.H], flat[:, self.H:].reshape(S * Cn, self.f, self.n),
                                              stream_id=0x40000000 + sweep0).reshape(S, Cn, self.f * self.n)
        if self.store_chains and self.chains > 1:
            self._chain_blocks.append(raw)
            if dblock is not None:
                self._deriv_blocks.append(dblock)
        frame = self._frame_from_raw(raw[:, 0, :], None if dblock is None else dblock[:, 0, :])
        if self.parameter_history.shape[0] == 0:
            self.parameter_history = frame
        else:
            self.parameter_history = pd.concat([self.parameter_history, frame], ignore_index=True)
        self.parameter_history.index = range(self.parameter_history.shape[0])
        if dblock is not None:
            vals = self.parameter_cache.to_numpy(dtype=float, copy=True)
            vals[self._pos_D] = dblock[-1, 0]
            self.parameter_cache.iloc[:] = vals

    def _rows_from_raw(self, raw, dvals=None):
        """(S, H + f n) -> (S, #labels) array in the reference's column order (sampler order; labels base.py:114-138);
        the Transform columns come from the device operator"""
        S = raw.shape[0]
        F = raw[:, self.H:].reshape(S, self.f, self.n)
        out = np.zeros((S, len(self.parameter_cache.index)))
        out[:, self._pos_h] = raw[:, :self.H]
        out[:, self._pos_F] = raw[:, self.H:]
        if len(self._pos_T):
            out[:, self._pos_T] = self._device_transforms(F)
        if dvals is not None and len(self._pos_D):
            out[:, self._pos_D] = dvals
        return out

    def _frame_from_raw(self, raw, dvals=None):
        return pd.DataFrame(self._rows_from_raw(raw, dvals), columns=self.parameter_cache.index)

    def store(self):
        """append the current parameter_cache (chain 0) to parameter_history (container.py:90-92)"""
        SamplerContainer.store(self)

    def chain_history(self, chain=None, as_frame=True):
        """stored states of every chain: array (S, chains, H + f n) laid out [hyper | F (f, n)]; or, for one chain, its
        history in the reference's schema (all columns of parameter_history, Transforms included) as a DataFrame
        (as_frame=False: a plain (S, #labels) array)"""
        if not self._chain_blocks:
            raw = np.zeros((0, self.chains, self.H + self.f * self.n))
        else:
            raw = np.concatenate(self._chain_blocks, axis=0)
            self._chain_blocks = [raw]
        if chain is None:
            return raw
        if self.chains == 1:
            return self.parameter_history if as_frame else self.parameter_history.to_numpy(dtype=float)
        dv = None
        if self._deriv_blocks:
            dv = np.concatenate(self._deriv_blocks, axis=0)
            self._deriv_blocks = [dv]
            dv = dv[:, chain, :]
        rows = self._rows_from_raw(raw[:, chain, :], dv)
        return pd.DataFrame(rows, columns=self.parameter_cache.index) if as_frame else rows

    def save(self, f, chain=None):
        """CSV of parameter_history (container.py:94-95); chain=c writes that chain's history in the same schema,
        chain='all' one file per chain (f must contain a %d for the chain number)"""
        if chain is None:
            return SamplerContainer.save(self, f)
        if chain == 'all':
            for c in range(self.chains):
                self.save(f % c, chain=c)
            return
        rows = self.chain_history(chain, as_frame=False)
        with open(f, "w") as fh:
            fh.write(",".join('"%s"' % c.replace('"', '""') if ("," in c or '"' in c) else c for c in self.parameter_cache.index) + "\n")
            np.savetxt(fh, rows, delimiter=",", fmt="%.17g")

    def chain_diagnostics(self, columns='hyper', split=True):
        """split-R^ over the chains, pooled posterior mean and variance of the stored states (device operator
        gpf_chain_diagnostics).  columns: 'hyper' (the 1 + 2P hyper-parameters) or 'all'.  Returns a DataFrame indexed by
        the column labels of the raw state ([hyper | functions])."""
        raw = self.chain_history()
        ncol = self.H if columns == 'hyper' else raw.shape[2]
        r, mu, var = _engine.chain_diagnostics(raw, 0, ncol, split=split, device=self._dev())
        labels = (self._hyper_names + self._flabels)[:ncol]
        return pd.DataFrame({"rhat": r.cpu().numpy(), "mean": mu.cpu().numpy(), "var": var.cpu().numpy()}, index=labels)

    def counters(self):
        return self.engine().counters()
