"""linalg.jitchol of the reference (gpfanova/linalg.py:35-54) on the GPU."""
import numpy as np

from . import engine


def jitchol(A, maxtries=5):
    """Lower Cholesky factor with the reference's jitter-retry ladder.  A: (n, n) array-like.
    Returns a numpy array like the reference; raises numpy.linalg.LinAlgError where the reference does."""
    A = np.ascontiguousarray(np.asarray(A, dtype=np.float64))
    return engine.jitchol(A, maxtries=maxtries).cpu().numpy()
