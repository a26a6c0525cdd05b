"""numpy fp64 restatement of the reference's classifier construction ("batch
adding").  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The construction is NOT on the accelerated path (BASELINE.json: "the sum-state
construction stay[s] on the reference host code"); it is restated only to
produce realistic orthogonal / non-orthogonal classifiers for the parity tests.
It is pinned against the reference's own ``fMPO.py`` executed here
(``tests/golden/generate_golden.py``).

Reference code followed (paths under /root/reference):

* ``variational_mpo_classifiers.py:33-51``  create_hairy_bitstrings_data
* ``tools.py:149-180``                       bitstring_data_to_QTN (truncated)
* ``experiments.py:457-461``                 label_last_site_to_centre
* ``deterministic_mpo_classifier.py:21-42``  class_encode_mps_to_mpo / mpo_encoding
* ``fMPO.py:666-689``                        fMPO.add
* ``fMPO.py:338-453``                        compress_centre_one_site
* ``fMPO.py:240-336``                        compress_one_site
* ``experiments.py:482-523``                 prepare_centred_batched_classifier,
                                             adding_centre_batches, add_centre_sublist
* ``deterministic_mpo_classifier.py:50-86``  add_sublist / adding_batches
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import polar


def hairy_bitstrings(n_labels: int, n_sites: int, centre: int | None):
    """Truncated one-site-hairy label product states.

    variational_mpo_classifiers.py:39-51: the hairy site carries
    ``2**(int(log2(n_labels))+1)`` = 16 levels, one-hot at the label; the other
    sites carry level 0.  tools.py:160-166 truncates the non-hairy sites to
    dimension 1.  experiments.py:457-461 swaps the last site with the centre
    one (``len // 2``) for centre-hairy classifiers.
    Returns ``bits[label][site]`` = 1-D vector (dim 1, or 16 on the hairy site).
    """
    s = 2 ** (int(np.log2(n_labels)) + 1)
    hairy = n_sites - 1 if centre is None else centre
    out = []
    for label in range(n_labels):
        sites = []
        for n in range(n_sites):
            if n == hairy:
                v = np.zeros(s)
                v[label] = 1.0
            else:
                v = np.ones(1)
            sites.append(v)
        out.append(sites)
    return out


def class_encode(mps_sites, label: int, bits):
    """deterministic_mpo_classifier.py:21-31 — site (d, a, b) times the label
    product state -> (d, s, a, b)."""
    return [np.einsum("dab,s->dsab", A, bits[label][n]) for n, A in enumerate(mps_sites)]


def mpo_add(x, y):
    """fMPO.py:666-689 — block-diagonal sum; first site concatenates along the
    right bond, last site along the left bond.  The reference allocates
    ``1j * zeros`` so the middle sites come out complex128 with zero imaginary
    part; that dtype is kept."""
    L = len(x)
    out = []
    for i in range(L):
        if i == 0:
            out.append(np.concatenate([x[i], y[i]], 3))
        elif i == L - 1:
            out.append(np.concatenate([x[i], y[i]], 2))
        else:
            d, s, a, b = x[i].shape
            _, _, a2, b2 = y[i].shape
            z = 1j * np.zeros((d, s, a + a2, b + b2))
            z[:, :, :a, :b] = x[i]
            z[:, :, a:, b:] = y[i]
            out.append(z)
    return out


def _rank_diag(S: np.ndarray) -> int:
    if S.size == 0:
        return 0
    tol = S.max() * max(S.size, 2) * np.finfo(np.float64).eps
    return int(np.sum(S > tol))


def _split_l2r(datum):
    """fMPO.py:346-364."""
    d, s, i, j = datum.shape
    M = datum.transpose(0, 2, 1, 3).reshape(d * i, s * j)
    u, S, v = np.linalg.svd(M, full_matrices=False)
    u = np.expand_dims(u.reshape(d, i, -1), 1)
    v = np.expand_dims(v.reshape(-1, s, j), 0).transpose(0, 2, 1, 3)
    return u, S, v


def _split_r2l(datum):
    """fMPO.py:366-380."""
    d, s, i, j = datum.shape
    M = datum.transpose(2, 1, 3, 0).reshape(i * s, j * d)
    u, S, v = np.linalg.svd(M, full_matrices=False)
    u = np.expand_dims(np.expand_dims(u.reshape(i, -1), 0), 0)
    v = v.reshape(-1, s, j, d).transpose(3, 1, 0, 2)
    return u, S, v


def _truncate(A, S, V, D):
    """fMPO.py:382-399."""
    if D is None:
        D = _rank_diag(S)
    return A[:, :, :, :D], S[:D], V[:, :, :D, :]


def mpo_norm(sites) -> float:
    """sqrt(<O|O>) of the MPO (the ``qtn.H @ qtn`` of fMPO.py:448-450)."""
    env = np.ones((1, 1), dtype=complex)
    for W in sites:
        env = np.einsum("ab,dsax,dsby->xy", env, np.conj(W), W)
    return float(np.sqrt(np.abs(env.reshape(-1)[0])))


def compress_centre_one_site(sites, D=None, orthogonalise=False):
    """fMPO.py:338-453."""
    data = [np.array(w) for w in sites]
    L = len(data)
    centre = int(np.argmax([w.shape[1] for w in data]))
    for m in range(centre):  # left -> middle sweep
        A, S, V = _split_l2r(data[m])
        data[m], S, V = _truncate(A, S, V, D)
        SV = S[None, None, :, None] * V
        d1, s1, i1, _ = SV.shape
        d2, s2, _, j2 = data[m + 1].shape
        data[m + 1] = np.einsum("asik,bukj->absuij", SV, data[m + 1]).reshape(
            d1 * d2, s1 * s2, i1, j2)
    for m in range(L - 1, centre, -1):  # right -> middle sweep
        A, S, V = _split_r2l(data[m])
        U, S, data[m] = _truncate(A, S, V, D)
        US = U * S[None, None, None, :]
        d1, s1, i1, _ = data[m - 1].shape
        d2, s2, _, j2 = US.shape
        data[m - 1] = np.einsum("asik,bukj->absuij", data[m - 1], US).reshape(
            d1 * d2, s1 * s2, i1, j2)
    if orthogonalise:
        d, s, i, j = data[centre].shape
        U, _ = polar(data[centre].transpose(0, 2, 1, 3).reshape(d * i, s * j))
        data[centre] = U.reshape(d, i, s, j).transpose(0, 2, 1, 3)
    else:
        data[centre] = data[centre] / mpo_norm(data)
    return data


def compress_one_site(sites, D=None, orthogonalise=False):
    """fMPO.py:240-336 — label leg on the LAST site, left sweep only, polar on
    the last site when ``orthogonalise``; the result is always normalised
    (fMPO.py:330-334 runs unconditionally)."""
    data = [np.array(w) for w in sites]
    L = len(data)
    for m in range(L):
        if m + 1 < L:
            A, S, V = _split_l2r(data[m])
            data[m], S, V = _truncate(A, S, V, D)
            SV = S[None, None, :, None] * V
            d1, s1, i1, _ = SV.shape
            d2, s2, _, j2 = data[m + 1].shape
            data[m + 1] = np.einsum("asik,bukj->absuij", SV, data[m + 1]).reshape(
                d1 * d2, s1 * s2, i1, j2)
        elif orthogonalise:
            d, s, i, j = data[m].shape
            U, _ = polar(data[m].transpose(0, 2, 1, 3).reshape(d * i, s * j))
            data[m] = U.reshape(d, i, s, j).transpose(0, 2, 1, 3)
    data[-1] = data[-1] / mpo_norm(data)
    return data


def _add_all(sub):
    c = sub[0]
    for o in sub[1:]:
        c = mpo_add(c, o)
    return c


def adding_centre_batches(mpos, D, batch_num=2, orthogonalise=False):
    """experiments.py:497-523."""
    out = []
    for i in range(int(len(mpos) / batch_num) + 1):
        sub = mpos[batch_num * i: batch_num * (i + 1)]
        if len(sub) > 0:
            out.append(compress_centre_one_site(_add_all(sub), D, orthogonalise))
    return out


def adding_batches_one_site(mpos, D, batch_num=2, truncate=True, orthogonalise=False):
    """deterministic_mpo_classifier.py:50-86 for one-site-hairy MPOs."""
    if len(mpos) % batch_num != 0:
        if not truncate:
            raise Exception("Batches are not of equal size!")
        expo = int(np.log(len(mpos)) / np.log(batch_num))
        mpos = mpos[: batch_num ** expo]
    out = []
    for i in range(int(len(mpos) / batch_num) + 1):
        sub = mpos[batch_num * i: batch_num * (i + 1)]
        if len(sub) > 0:
            out.append(compress_one_site(_add_all(sub), D, orthogonalise))
    return out


def prepare_centred_sum_states(mps_list, labels, D_batch, batch_nums, n_labels=10):
    """experiments.py:482-495: class-encode, compress each (D=None), then add in
    batches until <= 10 MPOs (one sum state per class when the data is in the
    "one class" arrangement, tools.py:87-100)."""
    L = len(mps_list[0])
    bits = hairy_bitstrings(n_labels, L, centre=L // 2)
    mpos = [compress_centre_one_site(class_encode(m, int(y), bits), None, False)
            for m, y in zip(mps_list, labels)]
    i = 0
    while len(mpos) > 10:
        mpos = adding_centre_batches(mpos, D_batch, batch_nums[i])
        i += 1
    return mpos


def build_centred_classifier(mps_list, labels, D_batch, D_final, batch_nums,
                             orthogonalise, n_labels=10):
    """experiments.py:66-84 with centre_site=True, prep_sum_states=True:
    ``batch_nums[-1]`` is the final batch; returns the classifier site list
    ``(2, s_n, B_n, B_{n+1})`` (complex128 with zero imaginary part, as the
    reference's)."""
    batch_nums = list(batch_nums)
    batch_final = batch_nums.pop(-1)
    sums = prepare_centred_sum_states(mps_list, labels, D_batch, batch_nums, n_labels)
    return adding_centre_batches(sums, D_final, batch_final, orthogonalise)[0]


def real_sites(sites, atol=0.0):
    """The GPU path is real: take ``.real`` after checking the imaginary part
    is exactly (or within atol) zero."""
    out = []
    for w in sites:
        w = np.asarray(w)
        if np.iscomplexobj(w):
            if np.max(np.abs(w.imag), initial=0.0) > atol:
                raise ValueError("classifier has a non-zero imaginary part")
            w = w.real
        out.append(np.ascontiguousarray(w, dtype=np.float64))
    return out
