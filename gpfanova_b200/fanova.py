"""Functional ANOVA model -- the API of gpfanova/fanova.py:7-325 (constructor arguments, method names and results,
function / effect naming scheme) on top of the GPU sweep of Base.

The model is  y_o(t) = sum_i X[o, i] beta*_i(t) + noise  with one latent function per column of the design matrix:
column 0 the mean, then for every effect k its mk - 1 Helmert contrast functions, then for every interacting pair
(i, j), i < j, the (mk_i - 1)(mk_j - 1) functions of the Kronecker contrast.  Everything in this file runs once on the
host (layout tables, labels) except the Transform values of stored states, which come from the device operator
gpf_fanova_transform for all chains at once."""
import logging

import numpy as np

from .base import Base
from .sample import Transform

SUFFIXES = ['alpha', 'beta', 'gamma', 'delta', 'epsilon']


def helmert_contrast(k):
    """Helmert coding of k levels without intercept (k x (k-1)), the matrix patsy 0.4.1 returns for
    Helmert().code_without_intercept(range(k)) (fanova.py:160): column j is -1 on rows 0..j, j+1 on row j+1."""
    rows, cols = np.indices((k, k - 1))
    return np.where(rows <= cols, -1.0, np.where(rows == cols + 1, rows.astype(float), 0.0))


def orthonormalise_helmert(c):
    """helmertConvert=True (fanova.py:151-158): rescale the Helmert columns, last one first, so that every ROW of the
    contrast matrix has squared length 1 - 1/levels (all levels get the same prior variance)."""
    q = c.shape[1]
    for j in range(q - 1, -1, -1):
        c[:, j] /= c[j + 1, j]
        s = np.power(c[j + 1, :], 2).sum()
        c[:, j] *= np.sqrt(1 - 1. / (q + 1) - s + np.power(c[j + 1, j], 2))      # the reference's own expression (bit-equal X)
    return c


class FANOVA(Base):

    EFFECT_SUFFIXES = SUFFIXES

    def __init__(self, x, y, effect, interactions=False, effectTransforms=True, interactionTransforms=True,
                 helmertConvert=False, *args, **kwargs):
        self.effect = effect
        self.k = effect.shape[1]
        self.mk = [len(np.unique(effect[:, i])) for i in range(self.k)]
        self.effectTransforms = effectTransforms
        self.interactionTransforms = interactionTransforms
        self.helmertConvert = helmertConvert
        # interacting pairs (i < j); a bool switches all pairs on or off, a list names them (fanova.py:22-36)
        all_pairs = [(i, j) for i in range(self.k) for j in range(i + 1, self.k)]
        if isinstance(interactions, bool):
            self.interactions = interactions
            self._interactions = {pr: True for pr in all_pairs}
        elif isinstance(interactions, list):
            self.interactions = True
            self._interactions = {(min(i, j), max(i, j)): True for i, j in interactions}
        self._pairs = [pr for pr in all_pairs if self.hasInteraction(*pr)]
        self.contrasts = [self.buildEffectContrastMatrix(i) for i in range(self.k)]
        self.projectionMatrices = {i: self._projectionMatrix(i) for i in range(self.k)}
        self.contrasts_interaction = {}
        if self.interactions:
            for i, j in all_pairs:
                self.contrasts_interaction[(i, j)] = np.kron(self.contrasts[i], self.contrasts[j])
                self.projectionMatrices[(i, j)] = self._projectionMatrix(i, j)
        # column layout of the design matrix: mean | main effects | interacting pairs
        self._main_start = np.concatenate([[1], 1 + np.cumsum([m - 1 for m in self.mk])]).astype(int)
        self._effectInteractionIndex = {}
        col = int(self._main_start[-1])
        for i, j in self._pairs:
            self._effectInteractionIndex[(i, j)] = col
            col += (self.mk[i] - 1) * (self.mk[j] - 1)
        self._ncol = col
        Base.__init__(self, x, y, *args, **kwargs)

    # ------------------------------------------------------------------ layout
    def hasInteraction(self, i, k):
        return bool(self.interactions) and (min(i, k), max(i, k)) in self._interactions

    def effectSuffix(self, k):
        return SUFFIXES[k] if self.k <= len(SUFFIXES) else 'effect%d' % k

    def effectIndex(self, i, j):
        """column of the j-th contrast function of effect i (fanova.py:229-230)"""
        return 1 + sum(m - 1 for m in self.mk[:i]) + j

    def effectInteractionIndex(self, i, j, k, l):
        """column of an interaction contrast function (fanova.py:232-264, including its fall-through to the start of the
        next block for pairs without an interaction)"""
        if i > k:
            i, j, k, l = k, l, i, j
        if k >= self.k:
            return self.f
        if not self.hasInteraction(i, k):
            return self.effectInteractionIndex(i, 0, k + 1, 0) if i == k - 1 else self.effectInteractionIndex(i + 1, 0, k, 0)
        for lev, eff in ((j, i), (l, k)):
            if lev >= self.mk[eff]:
                logging.getLogger(__name__).warning("%d is out of range of effect %d with size %d" % (lev, eff, self.mk[eff]))
                return -1
        return self._effectInteractionIndex[(i, k)] + j * (self.mk[i] - 1) + l

    def _pair_columns(self, i, k):
        """the function columns of the interaction block (i, k) as the reference's index arithmetic delimits them"""
        first = self.effectInteractionIndex(i, 0, k, 0)
        last = self.effectInteractionIndex(i, 0, k + 1, 0) if i == k - 1 else self.effectInteractionIndex(i + 1, 0, k, 0)
        return list(range(first, last))

    def priorGroups(self):
        """one prior (kernel hyper-parameters) for the mean, one per effect, one per interacting pair (fanova.py:266-281)"""
        groups = [[0]]
        groups += [list(range(self._main_start[i], self._main_start[i + 1])) for i in range(self.k)]
        for i, j in self._pairs:
            s = self._effectInteractionIndex[(i, j)]
            groups.append(list(range(s, s + (self.mk[i] - 1) * (self.mk[j] - 1))))
        return groups

    def _buildDesignMatrix(self):
        """row o = [1 | Helmert row of level effect[o, k] for every effect | Kronecker row for every interacting pair]
        (fanova.py:283-310), built block-wise by fancy indexing"""
        lev = np.asarray(self.effect).astype(int)
        X = np.zeros((self.m, self._ncol))
        X[:, 0] = 1
        for i in range(self.k):
            X[:, self._main_start[i]:self._main_start[i + 1]] = self.contrasts[i][lev[:, i], :]
        for i, j in self._pairs:
            s = self._effectInteractionIndex[(i, j)]
            block = self.contrasts_interaction[(i, j)][lev[:, i] * self.mk[j] + lev[:, j], :]
            X[:, s:s + block.shape[1]] = block
        return X

    # ------------------------------------------------------------------ contrasts
    def buildEffectContrastMatrix(self, i):
        h = helmert_contrast(self.mk[i])
        return orthonormalise_helmert(h) if self.helmertConvert else h

    def _projectionMatrix(self, i, k=None):
        """centering projection I - 11^T/mk used by the finite-population variance (fanova.py:166-171)"""
        if k is not None:
            return None
        return np.eye(self.mk[i]) - np.full((self.mk[i], self.mk[i]), 1.0 / self.mk[i])

    def effectContrastMatrix(self, i, k=None, full=False):
        """contrast matrix of effect i or of the pair (i, k); full=True embeds it in all f function columns"""
        if k is not None and k < i:
            return self.effectContrastMatrix(k, i, full)
        c = self.contrasts[i] if k is None else self.contrasts_interaction[(i, k)]
        if not full:
            return c
        cols = list(range(self._main_start[i], self._main_start[i + 1])) if k is None else self._pair_columns(i, k)
        out = np.zeros((c.shape[0], self.f))
        out[:, cols] = c
        return out

    # ------------------------------------------------------------------ Transform samplers (fanova.py:58-124)
    def _transform_plan(self):
        """blocks (first function column, contrast matrix, finite-population variance?) in the order of the Transform
        samplers: per effect its levels then the variance; then the interacting pairs"""
        blocks = []
        if self.effectTransforms:
            blocks += [(int(self._main_start[i]), self.contrasts[i], True) for i in range(self.k)]
        if self.interactionTransforms:
            blocks += [(self._effectInteractionIndex[pr], self.contrasts_interaction[pr], False) for pr in self._pairs]
        return blocks

    def _additionalSamplers(self):
        out = []
        if self.effectTransforms:
            for i in range(self.k):
                sfx = self.effectSuffix(i)
                out += [Transform('%s_%d' % (sfx, l), self.effectIndexToCache(i, l), lambda i=i, l=l: self.effectSample(i, l))
                        for l in range(self.mk[i])]
                out.append(Transform("%s_finiteSampleVariance" % sfx, self.fsvIndexToCache(i),
                                     lambda i=i: self.finitePopulationVarianceSample(i)))
        if self.interactionTransforms:
            for i, j in self._pairs:
                out += [Transform('(%s,%s)_(%d,%d)' % (self.effectSuffix(i), self.effectSuffix(j), l, q),
                                  self.effectIndexToCache(i, l, j, q), lambda i=i, l=l, j=j, q=q: self.effectSample(i, l, j, q))
                        for l in range(self.mk[i]) for q in range(self.mk[j])]
        return out

    def _transform_block(self, F):
        """Transform values of states F (S, f, n) on the host -> (S, #transform labels); the device counterpart for many
        chains is engine.fanova_transform with the same plan"""
        out = []
        for start, coef, fpv in self._transform_plan():
            a = np.einsum('lj,sjn->sln', coef, F[:, start:start + coef.shape[1], :])       # effect_l = sum_j C[l, j] beta*_j
            out.append(a.reshape(F.shape[0], -1))
            if fpv:
                out.append(np.einsum('sln,sln->sn', a, a - a.mean(axis=1, keepdims=True)) / (coef.shape[0] - 1))
        return np.concatenate(out, axis=1) if out else np.zeros((F.shape[0], 0))

    def effectContrastArray(self, i, k=None, deriv=False):
        cols = range(self._main_start[i], self._main_start[i + 1]) if k is None else self._pair_columns(i, k)
        return self.functionMatrix(only=list(cols))

    def effectArray(self, i, k=None):
        if k is not None:
            return None
        return np.column_stack([self.parameterCurrent("%s_%d" % (self.effectSuffix(i), j)) for j in range(self.mk[i])])

    def effectSample(self, i, j, k=None, l=None):
        """current value of effect level j of effect i, or of level (j, l) of the pair (i, k), chain 0"""
        if k is None:
            return self.effectContrastArray(i).dot(self.contrasts[i][j, :])
        return self.effectContrastArray(i, k).dot(self.contrasts_interaction[(i, k)][j * self.mk[k] + l, :])

    def effectSamples(self, i, j, k=None, l=None, contrast=False):
        return self.parameterSamples(self.effectName(i, j, k, l, contrast))

    def finitePopulationVarianceSample(self, i, k=None):
        if k is not None:
            return None
        a = self.effectArray(i)
        return np.einsum('tl,lm,tm->t', a, self.projectionMatrices[i], a) / (self.mk[i] - 1)

    def relativeInteraction(self, i, j, k, l, i0=None, k0=None):
        """interaction of levels (j, l) relative to the reference levels (i0, k0) over the stored samples"""
        if not self.hasInteraction(i, k):
            return None
        i0, k0 = (i0 or 0), (k0 or 0)
        name = "(%s,%s)_(%%d,%%d)" % (self.effectSuffix(i), self.effectSuffix(k))
        get = lambda a, b: self.parameterSamples(name % (a, b)).values
        return get(j, l) - get(j, k0) - get(i0, l) + get(i0, k0)

    # ------------------------------------------------------------------ names (fanova.py:173-227)
    def functionNames(self):
        names = {0: 'mean'}
        for i in range(self.k):
            for j in range(self.mk[i] - 1):
                names[len(names)] = "%s*_%d" % (self.effectSuffix(i), j)
        # the reference enumerates (i, j, k, l) with the second effect's contrast innermost, for EVERY pair
        for i in range(self.k):
            for j in range(self.mk[i] - 1):
                for k in range(i + 1, self.k):
                    for l in range(self.mk[k] - 1):
                        names[len(names)] = "(%s,%s)*_(%d,%d)" % (self.effectSuffix(i), self.effectSuffix(k), j, l)
        return names

    def effectName(self, k, l, i=None, j=None, contrast=False, finiteVariance=False):
        star = "*" if contrast else ""
        if i is None:
            return "%s%s_%d" % (self.effectSuffix(k), star, l)
        if i < k:
            return self.effectName(i, k, k, l, contrast, finiteVariance)       # (argument order of the reference)
        return "(%s,%s)%s_(%d,%d)" % (self.effectSuffix(k), self.effectSuffix(i), star, l, j)

    def effectIndexToCache(self, k, l, i=None, j=None, contrast=False):
        star = '*' if contrast else ''
        if i is not None and k > i:
            k, l, i, j = i, j, k, l
        stem = "%s%s_%d" % (self.effectSuffix(k), star, l) if i is None else \
            "(%s,%s)%s_(%d,%d)" % (self.effectSuffix(k), self.effectSuffix(i), star, l, j)
        return ["%s(%s)" % (stem, z) for z in self._observationIndexBase()]

    def fsvIndexToCache(self, i, k=None):
        if k is not None:
            return None
        return ["%s_finitePopulationVariance(%s)" % (self.effectSuffix(i), z) for z in self._observationIndexBase()]
