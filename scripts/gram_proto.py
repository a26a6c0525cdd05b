"""numpy model of the Gram-matrix front end of the encoder (D_encode >= 32):

  G_n = X_n X_n^T (fp64) for the row unfoldings X_n (2^(n+1) x 2^(9-n)) of the
  padded image, G_{n-1} = partial trace of G_n; pivoted Cholesky G_n = L L^T
  (fp64, rows in original order, columns in pivot order); one-sided Jacobi on
  the COLUMNS of L in fp32 (odd-even ordering) -> U^(n) Sigma^(n);
  A_n[(s,l), r] = sum_i U^(n-1)[i, l] U^(n)[2i+s, r];  Psi_{N+1} = U^(N)^T X_N.

Checks amplitudes / isometry / singular values against the fp64 SVD chain and
counts sweeps.  Development aid only (not shipped, not a test).
"""
import sys
import numpy as np

sys.path.insert(0, ".")
from orthogonal_classifiers_b200.synthetic import synthetic_images  # noqa: E402

f32 = np.float32


def pivoted_cholesky(G, tol_rel=1e-13):
    p = G.shape[0]
    G = G.copy()
    L = np.zeros((p, p))
    d = np.diag(G).copy()
    dmax = d.max()
    done = np.zeros(p, bool)
    r = 0
    for k in range(p):
        dd = np.where(done, -1.0, d)
        j = int(np.argmax(dd))
        if dd[j] <= tol_rel * dmax:
            break
        piv = np.sqrt(dd[j])
        col = G[:, j] / piv
        col[done] = 0.0
        L[:, k] = col
        L[j, k] = piv
        done[j] = True
        G -= np.outer(col, col)
        d = np.diag(G).copy()
        r += 1
    return L, r


def jacobi_cols_f32(L, r, tol2=1e-12, qstop2=1e-7, floor2=1e-12, max_sweeps=14):
    """odd-even one-sided Jacobi on the first r columns of L (fp32)."""
    V = [L[:, k].astype(f32).copy() for k in range(r)]
    if r & 1:
        V.append(np.zeros_like(V[0]))
    n = len(V)
    fro = sum(float(v @ v) for v in V)
    thr2 = f32(floor2 * fro)
    sweeps = 0
    for sw in range(max_sweeps):
        again = False
        sweeps += 1
        for rr in range(n // 2):
            for phase in (0, 1):
                for k in range(phase, n - 1, 2):
                    a, b = V[k], V[k + 1]
                    aa, bb, ab = f32(a @ a), f32(b @ b), f32(a @ b)
                    if ab * ab > f32(tol2) * aa * bb and aa > thr2 and bb > thr2:
                        if ab * ab > f32(qstop2) * aa * bb:
                            again = True
                        delta = bb - aa
                        gamma = f32(2) * ab
                        w = np.sqrt(delta * delta + gamma * gamma)
                        den = abs(delta) + w
                        sg = np.copysign(gamma, delta * gamma) if delta != 0 else gamma
                        X = den * den + gamma * gamma
                        q = f32(1) / np.sqrt(X)
                        c, s = f32(den * q), f32(sg * q)
                        na = (c * a - s * b).astype(f32)
                        nb = (s * a + c * b).astype(f32)
                        a, b = na, nb
                    V[k], V[k + 1] = b, a  # swap positions
        if not again:
            break
    nrm = np.array([float(v.astype(np.float64) @ v.astype(np.float64)) for v in V])
    order = np.argsort(-nrm, kind="stable")
    sig = np.sqrt(nrm[order])
    U = np.zeros((L.shape[0], L.shape[0]), f32)
    live = 0
    for j, k in enumerate(order):
        if nrm[k] > thr2:
            U[:, j] = V[k] / f32(np.sqrt(nrm[k]))
            live += 1
    return U, sig, live, sweeps


def encode_left(x1024, NL=4):
    """steps 0..NL via Gram matrices. Returns sites A_0..A_NL, Psi_{NL+1}, svals, sweeps."""
    p = 2 ** (NL + 1)
    X = x1024.reshape(p, -1)
    G = X @ X.T
    Gs = {NL: G}
    for n in range(NL - 1, -1, -1):
        g = Gs[n + 1]
        Gs[n] = g[0::2, 0::2] + g[1::2, 1::2]
    Us, sv, sweeps = {}, {}, {}
    for n in range(NL + 1):
        L, r = pivoted_cholesky(Gs[n])
        Us[n], sv[n], _, sweeps[n] = jacobi_cols_f32(L, r)
    sites = []
    Uprev = np.ones((1, 1), f32)
    for n in range(NL + 1):
        U = Us[n]
        b = Uprev.shape[0]
        A = np.zeros((2, b, 2 * b), f32)
        for s in range(2):
            A[s] = Uprev.T @ U[s::2, :]
        sites.append(A)
        Uprev = U
    Psi = (Us[NL].T @ X.astype(f32)).astype(f32)
    return sites, Psi, sv, sweeps


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    NL = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    imgs = synthetic_images(n_img, seed=3)
    errs_amp, errs_iso, errs_sv, sw = [], [], [], {n: [] for n in range(NL + 1)}
    for img in imgs:
        x = np.zeros(1024)
        x[:784] = img
        sites, Psi, sv, sweeps = encode_left(x, NL)
        # reconstruct: PL (2^(NL+1) x b) = product of sites; state = PL @ Psi
        v = np.ones((1, 1))
        for A in sites:
            v = np.einsum("xa,dab->xdb", v, A.astype(np.float64)).reshape(-1, A.shape[2])
        rec = (v @ Psi.astype(np.float64)).reshape(-1)
        errs_amp.append(np.abs(rec - x).max() / np.linalg.norm(x))
        iso = 0.0
        for n, A in enumerate(sites):
            M = A.reshape(-1, A.shape[2]).astype(np.float64)
            g = M.T @ M
            live = np.diag(g) > 0.5
            iso = max(iso, np.abs(g[np.ix_(live, live)] - np.eye(live.sum())).max())
            ref = np.linalg.svd(x.reshape(2 ** (n + 1), -1), compute_uv=False)
            k = min(len(ref), len(sv[n]))
            errs_sv.append(np.abs(ref[:k] - sv[n][:k]).max() / ref[0])
            sw[n].append(sweeps[n])
        errs_iso.append(iso)
    print("amp err max", max(errs_amp), "iso err max", max(errs_iso), "sv err max", max(errs_sv))
    for n in range(NL + 1):
        print("step", n, "sweeps mean", np.mean(sw[n]), "max", max(sw[n]))


if __name__ == "__main__":
    main()
