"""shared helpers for the parity tests (oracle side).  Imports oracle/ -- tests only."""
import os

import numpy as np

import gpfanova_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["kat0_zeroF", "kat0_sinF", "oneway_n50", "oneway_n50_missing", "twoway_n24", "twoway_n24_hc",
         "unbalanced_p2_n12", "cokol5x5_hc_n20", "oneway_n64"]


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    gl = d["groups_len"]
    gf = d["groups_flat"]
    groups, o = [], 0
    for L in gl:
        groups.append([int(v) for v in gf[o:o + L]])
        o += L
    d["groups"] = groups
    return d


def relerr(a, b):
    """max-norm relative error  max|a-b| / max|b|"""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    den = np.max(np.abs(b))
    if den == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / den)


def oracle_model(case):
    m = orc.Model(case["x"], case["y"], case["design_matrix"], case["groups"])
    m.F = case["F"].copy()
    m.hyper = case["hyper"].copy()
    return m
