#!/usr/bin/env python
"""Recipe: make the reference (ptonner/gpfanova, Python-2, needs patsy/matplotlib)
importable under Python 3 as ``oracle/_ref/gpfanova``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpfanova_b200/`` may import this or
its output.  Output goes to ``oracle/_ref/`` (git-ignored, NOT gpurun-ignored, so
it travels to the GPU box where /root/reference does not exist).

The reference sources are read where they lie (default /root/reference) and
re-emitted with a minimal, purely mechanical edit list:

  * implicit relative imports  -> explicit relative imports
  * ``except X,e``              -> ``except X as e``
  * ``dict.iteritems``          -> ``dict.items``; ``range`` objects -> lists where shuffled/concatenated
  * mixed tab/space indentation -> tabs (three lines)
  * the three ``.pyx`` files are plain Python plus two ``cdef`` lines and one
    ``double`` annotation: those are removed and the files emitted as ``.py``
  * ``patsy.contrasts.Helmert`` (absent here) -> 6-line shim with patsy 0.4.1's
    ``_helmert_contrast`` formula; unused ``patsy...Sum`` import dropped
  * ``gpfanova.plot`` (needs matplotlib, absent) is not emitted
  * pandas API removed since 0.18: ``DataFrame.append`` -> ``pd.concat``;
    positional ``Series[0]`` -> ``.iloc[0]``
  * ``filter(...)[0]`` -> ``list(filter(...))[0]``

No arithmetic is touched.  Run:  python oracle/build_ref.py [--src /root/reference]
"""
import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

HELMERT_SHIM = '''
class Helmert(object):
	"""shim for patsy.contrasts.Helmert (patsy 0.4.1 _helmert_contrast)"""
	class _M(object):
		def __init__(self, matrix):
			self.matrix = matrix
	def code_without_intercept(self, levels):
		k = len(levels)
		contr = np.zeros((k, k - 1))
		contr[1:][np.diag_indices(k - 1)] = np.arange(1, k)
		contr[np.triu_indices(k - 1)] = -1
		return Helmert._M(contr)
'''


def sub(text, pattern, repl, count=1, flags=0, must=True):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    if must and n == 0:
        raise RuntimeError("edit did not apply: %r" % pattern)
    return new


def tabify(text):
    """normalise leading whitespace of tab-indented files: a space inside the
    leading tabs (py2 tolerated it) is dropped."""
    out = []
    for line in text.split("\n"):
        m = re.match(r"^([\t ]*)(.*)$", line)
        lead, rest = m.group(1), m.group(2)
        if "\t" in lead and " " in lead:
            lead = lead.replace(" ", "")
        out.append(lead + rest)
    return "\n".join(out)


def patch(rel, text):
    if rel == "gpfanova/__init__.py":
        text = re.sub(r"^import (base|mean|fanova|kernel|sample|interval)$", r"from . import \1", text, flags=re.M)
        text = sub(text, r"^import plot\n", "", flags=re.M)
    elif rel == "gpfanova/base.py":
        text = sub(text, r"^from patsy\.contrasts import Sum\n", "", flags=re.M)
        text = sub(text, r"^from sample import", "from .sample import", flags=re.M)
        text = sub(text, r"^from kernel import", "from .kernel import", flags=re.M)
    elif rel == "gpfanova/mean.py":
        text = sub(text, r"^from base import", "from .base import", flags=re.M)
    elif rel == "gpfanova/fanova.py":
        text = tabify(text)
        text = sub(text, r"^from base import", "from .base import", flags=re.M)
        text = sub(text, r"^from sample import", "from .sample import", flags=re.M)
        text = sub(text, r"^from patsy\.contrasts import Helmert\n", HELMERT_SHIM, flags=re.M)
        # py3: range objects are not lists (priorGroups are compared / indexed as lists)
        text = sub(text, r"g\.append\(range\(ind,ind\+self\.mk\[i\]-1\)\)", "g.append(list(range(ind,ind+self.mk[i]-1)))")
        text = sub(text, r"g\.append\(range\(ind,ind\+c\)\)", "g.append(list(range(ind,ind+c)))")
    elif rel == "gpfanova/sample/__init__.py":
        text = re.sub(r"^from (sampler|container|gibbs|slice|transform|fixed|function) import", r"from .\1 import", text, flags=re.M)
    elif rel == "gpfanova/kernel/__init__.py":
        text = re.sub(r"^from (rbf|white) import", r"from .\1 import", text, flags=re.M)
    elif rel == "gpfanova/kernel/white.py":
        text = sub(text, r"^from kernel import", "from .kernel import", flags=re.M)
    elif rel == "gpfanova/kernel/rbf.pyx":
        text = sub(text, r"^from kernel import", "from .kernel import", flags=re.M)
    elif rel == "gpfanova/kernel/kernel.pyx":
        text = sub(text, r"except np\.linalg\.linalg\.LinAlgError,e:", "except np.linalg.LinAlgError as e:")
    elif rel == "gpfanova/sample/slice.pyx":
        text = sub(text, r"def _sample\(self,double x\):", "def _sample(self,x):")
        text = sub(text, r"^\s*cdef double [^\n]*\n", "", flags=re.M)
        text = sub(text, r"^\s*cdef int [^\n]*\n", "", flags=re.M)
    elif rel == "gpfanova/sample/function.py":
        text = tabify(text)
    elif rel == "gpfanova/sample/container.py":
        text = sub(text, r"kwargs\.iteritems\(\)", "kwargs.items()")
        text = sub(text, r"ind = range\(len\(self\.samplers\)\)", "ind = list(range(len(self.samplers)))")
        text = sub(text, r"param = param\[0\]", "param = param.iloc[0]")
        text = sub(text, r"self\.parameter_history = self\.parameter_history\.append\(self\.parameter_cache,ignore_index=True\)",
                   "self.parameter_history = pd.concat([self.parameter_history,self.parameter_cache.to_frame().T],ignore_index=True)")
    elif rel == "gpfanova/interval.py":
        text = re.sub(r"= filter\((lambda x: [^\n]*),alphaPotential\)\[0\]", r"= list(filter(\1,alphaPotential))[0]", text)
    elif rel == "gpfanova/linalg.py":
        pass
    return text


FILES = [
    "gpfanova/__init__.py", "gpfanova/base.py", "gpfanova/mean.py", "gpfanova/fanova.py",
    "gpfanova/linalg.py", "gpfanova/interval.py",
    "gpfanova/sample/__init__.py", "gpfanova/sample/sampler.py", "gpfanova/sample/container.py",
    "gpfanova/sample/gibbs.py", "gpfanova/sample/fixed.py", "gpfanova/sample/transform.py",
    "gpfanova/sample/function.py", "gpfanova/sample/slice.pyx",
    "gpfanova/kernel/__init__.py", "gpfanova/kernel/kernel.pyx", "gpfanova/kernel/rbf.pyx",
    "gpfanova/kernel/white.py",
]


def build(src="/root/reference", out=OUT, quiet=False):
    if not os.path.isdir(os.path.join(src, "gpfanova")):
        raise FileNotFoundError("reference not found at %s" % src)
    for rel in FILES:
        with open(os.path.join(src, rel)) as fh:
            text = fh.read()
        text = patch(rel, text)
        dst = os.path.join(out, rel[:-4] + ".py" if rel.endswith(".pyx") else rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as fh:
            fh.write(text)
    with open(os.path.join(out, "README"), "w") as fh:
        fh.write("generated by oracle/build_ref.py from the unmodified reference; do not commit\n")
    if not quiet:
        print("reference emitted to", out)
    return out


def import_ref(out=OUT):
    """import the patched reference package (as module name ``gpfanova``)."""
    if out not in sys.path:
        sys.path.insert(0, out)
    import gpfanova  # noqa
    return gpfanova


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    build(a.src)
    g = import_ref()
    print("import ok:", g.__file__)
