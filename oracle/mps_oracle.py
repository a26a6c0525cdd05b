"""numpy fp64 restatement of the reference's image -> MPS encode and MPS.MPO
overlap.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference call sites restated here (paths under /root/reference):

* ``tools.py:217-221``  image_to_mps: L = ceil(log2 n), tail zero-pad to 2^L,
  reshape (2,)*L, ``fMPS().left_from_state(.).left_canonicalise(D=D)``.
* ``variational_mpo_classifiers.py:102-104``  mps_encoding (per-image loop).
* ``experiments.py:332-337`` (and 278, 390, 439, 633, 769)  the open-label
  overlap ``np.abs((mps.H.squeeze() @ classifier.squeeze()).data)``.
* ``variational_mpo_classifiers.py:309-318,340-345``  per-label predictions and
  top-k accuracy.

The arithmetic of ``left_from_state`` / ``left_canonicalise`` lives in the
un-vendored dependency ``xmps`` (requirements.txt:69, pinned at commit
0fe1c7ec93425fcbfcd2bd411599c645b1edce62) and the contraction in ``quimb``
(requirements.txt:50, commit 72ae4494); both are restated from their published
behaviour:

* ``left_from_state``: Psi <- state; for n = 0..L-1: reshape Psi to
  (d * D_{n-1}, d^(L-n-1)) with the physical index of site n as the slow row
  index, thin SVD, A_n = U split into d blocks of rows -> (d, D_{n-1}, D_n),
  Psi <- diag(S) V.  The last 1x1 ``S V`` is dropped, so the state is unit norm.
* ``left_canonicalise(D)``: right-to-left SVD sweep truncating every bond to
  min(D, rank), then a left-to-right sweep back (variant "sweep").
"""
from __future__ import annotations

import math

import numpy as np


def pad_image(f_image: np.ndarray) -> tuple[np.ndarray, int]:
    """tools.py:218-220 — flat tail pad to the next power of two."""
    n = f_image.shape[0]
    L = max(1, int(np.ceil(np.log2(n))))
    return np.pad(np.asarray(f_image, dtype=np.float64), (0, 2 ** L - n)), L


def _rank(S: np.ndarray) -> int:
    """Number of numerically non-zero singular values (xmps ``rank`` /
    ``minD=True`` trims these).  numpy.linalg.matrix_rank's default tolerance."""
    if S.size == 0 or S[0] == 0.0:
        return 0
    tol = S[0] * max(S.size, 2) * np.finfo(np.float64).eps
    return int(np.sum(S > tol))


def _unfold(psi: np.ndarray) -> np.ndarray:
    """psi (b, 2q) -> M ((s, l), c) = psi[l, s*q + c]   (site index slow)."""
    b, cols = psi.shape
    q = cols // 2
    return psi.reshape(b, 2, q).transpose(1, 0, 2).reshape(2 * b, q)


def _tt_svd(x: np.ndarray, L: int, D: int | None, trim: bool):
    """Successive (optionally truncated) SVD chain, left to right."""
    psi = x.reshape(1, 2 ** L)
    sites, svals = [], []
    for n in range(L):
        b = psi.shape[0]
        M = _unfold(psi)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        r = S.size
        if trim:
            r = max(1, _rank(S))
        if D is not None:
            r = min(r, D)
        sites.append(U[:, :r].reshape(2, b, r))
        svals.append(S.copy())
        psi = S[:r, None] * Vt[:r]
    return sites, svals, psi  # psi is (r_L, 1): the dropped norm factor


def _right_sweep_truncate(sites, D: int | None, trim: bool):
    """Right-to-left SVD sweep: site m (2, Dl, Dr) -> matrix (Dl, 2*Dr);
    M = U S V; B_m = V[:r]; site m-1 <- site m-1 . U[:, :r] S[:r].  On a
    left-canonical input the S seen at bond m are the true Schmidt values."""
    sites = [s.copy() for s in sites]
    L = len(sites)
    schmidt = [None] * L
    for m in range(L - 1, 0, -1):
        d, Dl, Dr = sites[m].shape
        M = sites[m].transpose(1, 0, 2).reshape(Dl, d * Dr)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        r = S.size
        if trim:
            r = max(1, _rank(S))
        if D is not None:
            r = min(r, D)
        sites[m] = Vt[:r].reshape(r, d, Dr).transpose(1, 0, 2)
        sites[m - 1] = np.einsum("dab,bc->dac", sites[m - 1], U[:, :r] * S[None, :r])
        schmidt[m] = S.copy()
    return sites, schmidt


def _left_sweep(sites, trim: bool):
    """Left-to-right QR-like SVD sweep without truncation; drops the final
    scalar so the result is a unit-norm left-canonical MPS."""
    sites = [s.copy() for s in sites]
    L = len(sites)
    carry = np.ones((1, 1))
    for m in range(L):
        A = np.einsum("ab,dbc->dac", carry, sites[m])
        d, Dl, Dr = A.shape
        M = A.reshape(d * Dl, Dr)
        U, S, Vt = np.linalg.svd(M, full_matrices=False)
        r = S.size
        if trim:
            r = max(1, _rank(S))
        sites[m] = U[:, :r].reshape(d, Dl, r)
        carry = S[:r, None] * Vt[:r]
    return sites


def encode_image(f_image: np.ndarray, D: int | None, variant: str = "inline",
                 trim: bool = True):
    """image (n,) -> (sites, svals).

    sites: list of L arrays (2, b_n, b_{n+1}), left-canonical, unit norm.
    svals: list of L arrays — singular values of the step-n unfolding
           (bond n+1) of the *unnormalised* padded image, before truncation
           (for variant "inline" they are those of the already-truncated chain).

    variant "inline": truncate to D inside the left-to-right chain
                      (BASELINE.json north_star wording).
    variant "sweep" : exact chain, then right-to-left truncating sweep, then
                      left sweep back (restated xmps ``left_canonicalise``).
    Both coincide (up to gauge) when D >= 2^(L/2).
    """
    x, L = pad_image(f_image)
    if not np.any(x):
        # The reference would fail (rank 0); the B200 path defines the output
        # of an all-zero image as an all-zero MPS (overlaps 0) — DESIGN.md §3.
        return [np.zeros((2, 1, 1)) for _ in range(L)], [np.zeros(1) for _ in range(L)]
    if variant == "inline":
        sites, svals, _ = _tt_svd(x, L, D, trim)
        return sites, svals
    if variant == "sweep":
        sites, svals, _ = _tt_svd(x, L, None, trim)
        if D is not None:
            sites, _ = _right_sweep_truncate(sites, D, trim)
            sites = _left_sweep(sites, trim)
        return sites, svals
    raise ValueError(f"unknown variant {variant!r}")


def sweep_singular_values(f_image: np.ndarray, D: int | None, trim: bool = True) -> list[np.ndarray]:
    """What the right-to-left truncating sweep of variant "sweep" sees, in the
    convention of ``oc_encode_batch(..., OC_TRUNC_SWEEP, svals_out)``: entry n
    (bond n+1, n = 0..L-2) = the singular values the sweep finds at site n+1 —
    those of the image already truncated on the bonds to the right — scaled to
    the UNNORMALISED image; entry L-1 = the norm of the truncated image."""
    x, L = pad_image(f_image)
    nrm = np.linalg.norm(x)
    sites, _, _ = _tt_svd(x, L, None, trim)
    swept, schmidt = _right_sweep_truncate(sites, D, trim)
    out = [nrm * schmidt[n + 1] for n in range(L - 1)]
    out.append(np.array([nrm * np.linalg.norm(swept[0])]))
    return out


def mps_to_state(sites) -> np.ndarray:
    """Contract an MPS (list of (2, Dl, Dr)) to its 2^L amplitude vector,
    site 0 = most significant bit of the flat index (C-order reshape,
    tools.py:220)."""
    v = np.ones((1, 1))
    for A in sites:
        d, Dl, Dr = A.shape
        v = np.einsum("xa,dab->xdb", v, A).reshape(-1, Dr)
    return v.reshape(-1)


def bond_singular_values(f_image: np.ndarray) -> list[np.ndarray]:
    """Gauge-free reference for the exact chain: singular values of the
    (2^(n+1) x 2^(L-n-1)) unfolding of the padded image, n = 0..L-1."""
    x, L = pad_image(f_image)
    return [np.linalg.svd(x.reshape(2 ** (n + 1), -1), compute_uv=False) for n in range(L)]


# --------------------------------------------------------------------------
# overlap  (experiments.py:332-337)
# --------------------------------------------------------------------------

def _overlaps_open_legs(mps_sites, clf_sites) -> np.ndarray:
    """Left-to-right contraction carrying every open label leg (any number of
    hairy sites; slow)."""
    env = np.ones((1, 1, 1), dtype=np.result_type(*[w.dtype for w in clf_sites], np.float64))
    for A, W in zip(mps_sites, clf_sites):
        # env[a, B, S] * conj(A)[d, a, a'] * W[d, s, B, B'] -> env'[a', B', S*s]
        t = np.einsum("aBS,dax->dxBS", env, np.conj(A))
        t = np.einsum("dxBS,dsBY->xYSs", t, W)
        env = t.reshape(t.shape[0], t.shape[1], -1)
    return np.abs(env.reshape(-1))


def overlaps_chain(mps_sites, clf_sites) -> np.ndarray:
    """|<image| classifier>| with every label leg left open.

    mps_sites: L arrays (2, a_n, a_{n+1}); clf_sites: L arrays
    (2, s_n, B_n, B_{n+1}) (fMPO.py:21-51 / tools.py:244-254 layout).
    Returns the abs of the contraction over physical legs and bonds, shape
    (prod s_n,) — (16,) for a one-hairy-site classifier.  quimb's ``.H`` is a
    complex conjugate of the image (real here).  For one hairy site the
    contraction order is the one an optimal path search finds: left
    environment, right environment, centre (matmul-shaped steps)."""
    hairy = [n for n, W in enumerate(clf_sites) if W.shape[1] > 1]
    if len(hairy) > 1:
        return _overlaps_open_legs(mps_sites, clf_sites)
    L = len(clf_sites)
    c = hairy[0] if hairy else L - 1
    EL = np.ones((1, 1))
    for n in range(c):                      # EL[a, B]
        A, W = np.conj(mps_sites[n]), clf_sites[n][:, 0]
        EL = sum(A[d].T @ (EL @ W[d]) for d in range(2))
    ER = np.ones((1, 1))
    for n in range(L - 1, c, -1):           # ER[a, B]
        A, W = np.conj(mps_sites[n]), clf_sites[n][:, 0]
        ER = sum((A[d] @ ER) @ W[d].T for d in range(2))
    A, W = np.conj(mps_sites[c]), clf_sites[c]
    out = 0
    for d in range(2):
        H = (EL.T @ A[d]) @ ER              # (B, B')
        out = out + np.tensordot(W[d], H, axes=([1, 2], [0, 1]))
    return np.abs(np.atleast_1d(out))


def dense_classifier(clf_sites) -> np.ndarray:
    """Contract the classifier MPO to a dense (2^L, prod s_n) matrix."""
    v = np.ones((1, 1, 1), dtype=clf_sites[0].dtype)  # (pix, S, B)
    for W in clf_sites:
        d, s, Bl, Br = W.shape
        v = np.einsum("pSB,dsBY->pdSsY", v, W)
        v = v.reshape(v.shape[0] * d, -1, Br)
    return v.reshape(v.shape[0], -1)


def overlaps_dense(f_images: np.ndarray, W: np.ndarray) -> np.ndarray:
    """|x_hat^T W| — valid whenever the encoding is exact (D_encode >= 32 for
    L = 10).  f_images (N, n) unpadded."""
    f_images = np.atleast_2d(np.asarray(f_images, dtype=np.float64))
    n = f_images.shape[1]
    x = np.zeros((f_images.shape[0], W.shape[0]))
    x[:, :n] = f_images
    nrm = np.linalg.norm(x, axis=1, keepdims=True)
    nrm[nrm == 0] = 1.0
    return np.abs((x / nrm) @ W)


def top_k_accuracy(predictions, y_test, k: int) -> float:
    """variational_mpo_classifiers.py:340-345, verbatim semantics
    (np.argpartition top-k, membership test, mean)."""
    top = [np.argpartition(np.asarray(p), -k)[-k:] for p in predictions]
    return float(np.mean([int(i in j) for i, j in zip(y_test, top)]))


def classify_reference_loop(images: np.ndarray, clf_sites, D: int | None,
                            variant: str = "inline"):
    """The reference's execution shape: one Python iteration per image for the
    encode (variational_mpo_classifiers.py:102-104) and one per image for the
    overlap (experiments.py:332-337).  Returns (overlaps (N, S), argmax (N,)).
    This is what bench.py times as the CPU baseline ("port")."""
    clf = [np.asarray(w) for w in clf_sites]
    mps = [encode_image(img, D, variant)[0] for img in images]
    ov = np.array([overlaps_chain(m, clf) for m in mps])
    return ov, np.argmax(ov, axis=1)
