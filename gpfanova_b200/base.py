"""Base model y(t) = X b(t) (gpfanova/base.py:7-259) with the Gibbs sweep running on the GPU.

Construction, sampler order and label scheme follow the reference (base.py:16-86, 114-138); `sample()` drives the
fused device sweep (gpf_sweep) for `chains` independent chains instead of looping over Python samplers.
`parameter_cache` / `parameter_history` keep the reference's schema and always describe chain 0; all chains are
available through `chain_state()` / `chain_history()`.

New, additive keyword arguments: chains=1, seed=None, device=None, devices=None (list of GPU ordinals: the chains are
sharded over them, no inter-GPU communication), store_chains=True.
"""
import os
import sys
import logging
import random as pyrandom
import time

import numpy as np
import pandas as pd
import torch

from . import engine as _engine
from .kernel import RBF, White
from .sample import Function, SamplerContainer, Slice


class Base(SamplerContainer):
    """Subclasses implement _buildDesignMatrix and priorGroups (base.py:7-14)."""

    def __init__(self, x, y, hyperparam_kwargs={}, derivatives=False, chains=1, seed=None, device=None, devices=None,
                 store_chains=True, *args, **kwargs):
        self.x = x
        self.y = y
        self._observationIndexBaseList = None
        self.n = self.x.shape[0]
        assert self.y.shape[0] == self.n, 'x and y must have same first dimension shape!'
        self.p = self.x.shape[1]
        self.m = self.y.shape[1]
        self.chains = int(chains)
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
        self.device = device
        self.devices = None if devices is None else list(devices)     # several GPUs of one box: chains are sharded
        self.store_chains = store_chains
        if derivatives:
            raise NotImplementedError("derivatives=True (FunctionDerivative, sample/function.py:43-66) is not part of "
                                      "the GPU hot path yet")

        self.design_matrix = None
        self.buildDesignMatrix()
        self.f = self.design_matrix.shape[1]

        if self.f > np.linalg.matrix_rank(self.design_matrix):
            logging.getLogger(__name__).error("design matrix is of rank %d, but there are %d functions!" % (
                np.linalg.matrix_rank(self.design_matrix), self.f))
        if np.any(np.isnan(self.y)):
            logging.getLogger(__name__).error("NaN values in observation matrix, this is not supported!")

        # kernels and samplers, in the reference's order (base.py:51-84)
        self.y_k = White(self, ['y_sigma'], logspace=True)
        wm = []
        w, m = .1, 10
        if 'y_sigma' in hyperparam_kwargs:
            w, m = hyperparam_kwargs['y_sigma']
        elif 'sigma' in hyperparam_kwargs:
            w, m = hyperparam_kwargs['sigma']
        wm.append((w, m))
        samplers = [Slice('y_sigma', 'y_sigma', lambda x: self.observationLikelihood(sigma=x, prior_lb=-2, prior_ub=2),
                          w, m, base=self, hyper_index=0)]
        self.kernels = []
        fxn_names = self.functionNames()
        self._groups = [list(g) for g in self.priorGroups()]
        self._hyper_names = ['y_sigma']
        for i, p in enumerate(self._groups):
            self.kernels.append(RBF(self, ['prior%d_sigma' % i, 'prior%d_lengthscale' % i], logspace=True))
            for f in p:
                s = fxn_names[f] if f in fxn_names else "f%d" % f
                samplers.append(Function('%s' % s, self.functionIndex(f), self, f, self.kernels[-1]))
            w, m = .1, 10
            if 'sigma' in hyperparam_kwargs:
                w, m = hyperparam_kwargs['sigma']
            wm.append((w, m))
            samplers.append(Slice('prior%d_sigma' % i, 'prior%d_sigma' % i,
                                  lambda x, p=i: self.prior_likelihood(p=p, sigma=x, prior_lb=-2, prior_ub=2), w, m,
                                  base=self, hyper_index=1 + 2 * i))
            w, m = .1, 10
            if 'lengthscale' in hyperparam_kwargs:
                w, m = hyperparam_kwargs['lengthscale']
            wm.append((w, m))
            samplers.append(Slice('prior%d_lengthscale' % i, 'prior%d_lengthscale' % i,
                                  lambda x, p=i: self.prior_likelihood(p=p, lengthscale=x, prior_lb=-2, prior_ub=2), w, m,
                                  base=self, hyper_index=2 + 2 * i))
            self._hyper_names += ['prior%d_sigma' % i, 'prior%d_lengthscale' % i]
        self._slice_wm = wm
        self._n_core_samplers = len(samplers)
        samplers.extend(self._additionalSamplers())

        self.H = len(self._hyper_names)
        self._flabels = [lab for f in range(self.f) for lab in self.functionIndex(f)]
        self._eng = None
        self._eng_key = None
        self._manual_calls = 0
        self._chain_blocks = []
        self._F = np.zeros((self.chains, self.f, self.n))
        self._hyper = np.zeros((self.chains, self.H))

        SamplerContainer.__init__(self, samplers, **kwargs)
        # initial values given by sampler name (container.py:16-18) apply to every chain
        self._F[:] = self.parameter_cache[self._flabels].values.astype(float).reshape(self.f, self.n)[None]
        self._hyper[:] = self.parameter_cache[self._hyper_names].values.astype(float)[None]
        self._transform_labels = [lab for s in self.samplers[self._n_core_samplers:] for lab in s.parameters]
        pos = {lab: i for i, lab in enumerate(self.parameter_cache.index)}
        self._pos_h = np.array([pos[l] for l in self._hyper_names], dtype=np.int64)
        self._pos_F = np.array([pos[l] for l in self._flabels], dtype=np.int64)
        self._pos_T = np.array([pos[l] for l in self._transform_labels], dtype=np.int64)

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

    def functionIndex(self, i, derivative=False, *args, **kwargs):
        if derivative:
            return ['df%d(%s)' % (i, z) for z in self._observationIndexBase()]
        return ['f%d(%s)' % (i, z) for z in self._observationIndexBase()]

    def functionPrior(self, f):
        priors = self.priorGroups()
        for i in range(len(priors)):
            if f in priors[i]:
                return i
        return -1

    def _observationIndexBase(self):
        if self._observationIndexBaseList is None:
            self._observationIndexBaseList = ['%s' % str(z) for z in self.x]
        return self._observationIndexBaseList

    def priorGroups(self):
        raise NotImplementedError("Implement a prior grouping function for your model!")

    def _priorParameters(self, i):
        if i < 0:
            return ['y_sigma']
        if i >= len(self.priorGroups()):
            return [None]
        return ['prior%d_sigma' % i, 'prior%d_lengthscale' % i]

    def offset(self):
        return 1e-9

    # ------------------------------------------------------------------ state accessors (base.py:151-198)
    def functionMatrix(self, remove=[], only=[], derivative=False):
        """current function values of chain 0 from parameter_cache, (n, #functions)"""
        functions = []
        if len(only) > 0:
            for o in only:
                functions.append(self.functionIndex(o, derivative=derivative))
        else:
            for f in range(self.f):
                functions.append(None if f in remove else self.functionIndex(f, derivative=derivative))
        out = np.zeros((self.n, len(functions)))
        for i, z in enumerate(functions):
            if z is not None:
                out[:, i] = self.parameter_cache[z]
        return out

    def functionSamples(self, f, *args, **kwargs):
        return self.parameter_history[self.functionIndex(f, *args, **kwargs)]

    def residual(self, remove=[], only=[]):
        return self.y.T - np.dot(self.design_matrix, self.functionMatrix(remove, only).T)

    def functionResidual(self, f):
        resid = self.residual(remove=[f])
        resid = (resid.T / self.design_matrix[:, f].T).T
        return resid[self.design_matrix[:, f] != 0, :]

    def observationMean(self):
        return np.dot(self.design_matrix, self.functionMatrix().T).ravel()

    # ------------------------------------------------------------------ device plumbing
    def engine(self):
        """the ChainEngine for the current (x, y, design matrix); rebuilt if the caller replaced self.y / self.x
        (the reference's convergence tests overwrite m.y after construction, test_convergence.py:60-62)"""
        key = (id(self.x), id(self.y), self.x.shape, self.y.shape)
        stale = self._eng is None or key != self._eng_key
        if not stale and not (np.array_equal(self._eng.y, self.y, equal_nan=True) and np.array_equal(self._eng.x, self.x)):
            stale = True
        if stale:
            if self._eng is not None:
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
            vals[self._pos_T] = self._transform_block(self._F[0][None])[0]
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
            logging.getLogger(__name__).error("prior likelihood LinAlgError (%d)" % p)
        return float(val)

    @staticmethod
    def _check_prior_bounds(eng, lb, ub):
        if (lb is not None and float(lb) != -2.0) or (ub is not None and float(ub) != 2.0):
            raise NotImplementedError("the device log-densities are built with the reference's U(-2, 2) hyper-priors "
                                      "(base.py:57,78,83)")

    def samplePrior(self):
        """draw functions from their GP priors at the current hyper-parameters and data from the model
        (base.py:245-259).  K and its factor come from the device; normals from numpy's global stream."""
        samples = np.zeros((self.f, self.n))
        for i in range(self.f):
            K = self.kernels[self.functionPrior(i)].K(self.x)
            L = _engine.jitchol(K)      # K alone is numerically singular: the jitter ladder (linalg.py:41-54) applies
            z = torch.as_tensor(np.random.standard_normal(self.n)).to(L.device)
            samples[i, :] = (L @ z).cpu().numpy()
        y = np.dot(self.design_matrix, samples) + np.random.normal(
            0, np.sqrt(pow(10, self.parameter_cache['y_sigma'])), size=(self.m, self.n))
        return y.T, samples.T

    # ------------------------------------------------------------------ the sweep (container.py:29-92)
    def _transform_block(self, F):
        """values of the Transform samplers for stored states F (S, f, n) -> (S, #transform labels)"""
        return np.zeros((F.shape[0], 0))

    def _core_order(self, n_sweeps):
        """random=True: shuffle the reference's full sampler list per sweep (container.py:34-35) and keep the relative
        order of the non-Transform samplers"""
        order = np.empty((n_sweeps, self._n_core_samplers), dtype=np.int32)
        ids = list(range(len(self.samplers)))
        for s in range(n_sweeps):
            pyrandom.shuffle(ids)
            order[s] = [i for i in ids if i < self._n_core_samplers]
        return order

    def sample(self, n=1, thin=0, verbose=False, random=False):
        """n sweeps of every chain on the device; every `thin`-th state is stored (thin=0: every sweep)."""
        logger = logging.getLogger(__name__)
        eng = self._push_state()
        start_time = time.time()
        step = 1 if thin == 0 else int(thin)
        order = self._core_order(n) if random else None
        err = None
        progress = None
        if verbose:
            def progress(done):
                logger.info("%d/%d iterations (%.2lf%s) finished in %.2lf minutes" % (
                    done, n, 100. * done / n, '%', (time.time() - start_time) / 60))
        t_sw = time.time()
        blocks = eng.sweep_host(n, thin=step, order=order, on_chunk=progress,
                                max_chunk_sweeps=max(step, n // 10) if verbose else None)
        t_sw1 = time.time()
        try:
            eng.status()
        except np.linalg.LinAlgError as e:        # keep what was sampled, then propagate like the reference
            err = e
        self._pull_state()
        if blocks.shape[0]:
            self._store_blocks(blocks)
        if os.environ.get("GPFANOVA_B200_TIMING"):
            sys.stderr.write("sample: push %.3f s  sweep_host %.3f s  pull+store %.3f s\n" % (
                t_sw - start_time, t_sw1 - t_sw, time.time() - t_sw1))
        if verbose:
            logger.info("%d samples finished in %.2lf minutes" % (n, (time.time() - start_time) / 60))
        if err is not None:
            raise err

    def _store_blocks(self, raw):
        """raw: (S, chains, H + f n) device history -> parameter_history rows (chain 0) in the reference's column order"""
        if self.store_chains and self.chains > 1:
            self._chain_blocks.append(raw)
        frame = self._frame_from_raw(raw[:, 0, :])
        if self.parameter_history.shape[0] == 0:
            self.parameter_history = frame
        else:
            self.parameter_history = pd.concat([self.parameter_history, frame], ignore_index=True)
        self.parameter_history.index = range(self.parameter_history.shape[0])

    def _frame_from_raw(self, raw):
        """(S, H + f n) -> DataFrame with the reference's columns (sampler order; labels base.py:114-138)"""
        S = raw.shape[0]
        hyper = raw[:, :self.H]
        F = raw[:, self.H:].reshape(S, self.f, self.n)
        cols = self.parameter_cache.index
        out = np.empty((S, len(cols)))
        out[:, self._pos_h] = hyper
        out[:, self._pos_F] = F.reshape(S, self.f * self.n)
        if len(self._pos_T):
            out[:, self._pos_T] = self._transform_block(F)
        return pd.DataFrame(out, columns=cols)

    def store(self):
        """append the current parameter_cache (chain 0) to parameter_history (container.py:90-92)"""
        SamplerContainer.store(self)

    def chain_history(self, chain=None):
        """stored states of every chain: array (S, chains, H + f n) laid out [hyper | F (f, n)], or a DataFrame with
        the reference's columns for one chain"""
        if not self._chain_blocks:
            raw = np.zeros((0, self.chains, self.H + self.f * self.n))
        else:
            raw = np.concatenate(self._chain_blocks, axis=0)
            self._chain_blocks = [raw]
        if chain is None:
            return raw
        if self.chains == 1:
            return self.parameter_history
        return self._frame_from_raw(raw[:, chain, :])

    def counters(self):
        return self.engine().counters()
