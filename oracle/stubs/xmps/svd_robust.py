import numpy as np
import scipy.linalg


def svd(M, full_matrices=False):
    """gesdd, falling back to gesvd when the divide-and-conquer driver does
    not converge (what xmps.svd_robust exists for)."""
    try:
        return np.linalg.svd(M, full_matrices=full_matrices)
    except np.linalg.LinAlgError:
        return scipy.linalg.svd(M, full_matrices=full_matrices, lapack_driver="gesvd")
