"""numpy model: how many one-sided Jacobi sweeps do steps 4 / 5 need with a SECOND QR preconditioning
(Drmac-Veselic: QR, then QR of the transposed factor, then Jacobi)?  Development aid only."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from orthogonal_classifiers_b200.synthetic import synthetic_images  # noqa: E402
from gram_proto import jacobi_cols_f32  # noqa: E402

f32 = np.float32


def chain_to(x1024, n):
    """Psi_n of the exact SVD chain (fp64), as the (2 b_n) x Q matrix M_n"""
    psi = x1024.reshape(1, -1)
    for k in range(n):
        b = psi.shape[0]
        M = np.concatenate([psi[:, : psi.shape[1] // 2], psi[:, psi.shape[1] // 2:]], axis=0)  # rows (s,l)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        r = min(len(S), 32)
        psi = (S[:r, None] * Vt[:r])
    b = psi.shape[0]
    return np.concatenate([psi[:, : psi.shape[1] // 2], psi[:, psi.shape[1] // 2:]], axis=0)


def mgs_rows(M, pivot=False):
    """M = L Q (rows orthonormalised in order); returns L (lower triangular in the processing order)"""
    M = M.astype(f32).copy()
    m = M.shape[0]
    order = list(range(m))
    L = np.zeros((m, m), f32)
    Q = np.zeros_like(M)
    fro = float((M.astype(np.float64) ** 2).sum())
    live = 0
    for k in range(m):
        if pivot:
            nr = (M[k:] ** 2).sum(1)
            j = k + int(np.argmax(nr))
            M[[k, j]] = M[[j, k]]
            L[[k, j]] = L[[j, k]]
        nrm2 = float(M[k] @ M[k])
        if nrm2 <= 1e-12 * fro:
            continue
        q = M[k] / f32(np.sqrt(nrm2))
        L[k, live] = f32(np.sqrt(nrm2))
        for i in range(k + 1, m):
            c = f32(M[i] @ q)
            L[i, live] = c
            M[i] -= c * q
        live += 1
    return L[:, :live], live


def run(n_img, seed):
    imgs = synthetic_images(n_img, seed=seed)
    res = {k: [] for k in ("s4_plain", "s4_second", "s4_piv", "s4_piv_second", "s5_plain", "s5_second")}
    for img in imgs:
        x = np.zeros(1024); x[:784] = img
        M4 = chain_to(x, 4)            # 32 x 32 (rows beyond 2k are ~0)
        keep = (M4 ** 2).sum(1) > 1e-12 * (M4 ** 2).sum()
        M4 = M4[keep]
        for piv in (False, True):
            L, r = mgs_rows(M4, piv)
            _, _, _, sw = jacobi_cols_f32(L.astype(np.float64), r)
            res["s4_piv" if piv else "s4_plain"].append(sw)
            # second: rows of L orthogonalised -> L = L2 Q2, Jacobi on columns of L2 (left vectors of L)
            L2, r2 = mgs_rows(L, False)
            _, _, _, sw2 = jacobi_cols_f32(L2.astype(np.float64), r2)
            res["s4_piv_second" if piv else "s4_second"].append(sw2)
        M5 = chain_to(x, 5)            # 64 x 16
        # column-pivoted MGS on columns == row-pivoted on M5^T; R = L^T; Jacobi on rows of R == columns of L... of R^T
        L, r = mgs_rows(M5.T.copy(), True)     # M5^T = L Q  -> M5 = Q^T L^T, R = L^T
        _, _, _, sw = jacobi_cols_f32(L.astype(np.float64).T.copy()[:r].T if False else L.astype(np.float64), r)
        res["s5_plain"].append(sw)
        L2, r2 = mgs_rows(L, False)
        _, _, _, sw2 = jacobi_cols_f32(L2.astype(np.float64), r2)
        res["s5_second"].append(sw2)
    for k, v in res.items():
        print(f"{k:16s} mean sweeps {np.mean(v):.2f}  max {max(v)}")


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 48, 3)
