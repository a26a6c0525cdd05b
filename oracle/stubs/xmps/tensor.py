import numpy as np


def rank(S, tol=None):
    """Number of numerically non-zero singular values of a diagonal matrix or
    vector (numpy.linalg.matrix_rank default tolerance)."""
    S = np.asarray(S)
    s = np.abs(np.diag(S)) if S.ndim == 2 else np.abs(S)
    if s.size == 0 or s.max() == 0:
        return 0
    if tol is None:
        tol = s.max() * max(s.size, 2) * np.finfo(np.float64).eps
    return int(np.sum(s > tol))
