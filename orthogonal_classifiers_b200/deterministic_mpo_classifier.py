"""Mirror of the classifier-construction entry points of
/root/reference/deterministic_mpo_classifier.py and experiments.py (centre-hairy form), on the
device: ``mpo_encoding`` (:33-42), ``prepare_centred_batched_classifier`` and
``adding_centre_batches`` (experiments.py:482-523).  fp64 throughout."""
from __future__ import annotations

import numpy as np

from . import batched
from .variational_mpo_classifiers import _batch_of


def _hairy_site(q_hairy_bitstrings, L):
    """(centre site, label-leg size) of the reference's label bitstrings
    (variational_mpo_classifiers.py:39-51 + experiments.py:457-461); default: centre L // 2, 16."""
    if q_hairy_bitstrings is None:
        return L // 2, 16
    b = q_hairy_bitstrings[0]
    sizes = [np.asarray(getattr(t, "data", t)).size for t in getattr(b, "tensors", b)]
    if len(sizes) != L:
        raise ValueError("bitstrings and images have different numbers of sites")
    c = int(np.argmax(sizes))
    for lab, bs in enumerate(q_hairy_bitstrings):
        v = np.asarray(getattr(getattr(bs, "tensors", bs)[c], "data", getattr(bs, "tensors", bs)[c])).reshape(-1)
        nz = np.nonzero(v)[0]
        if nz.size != 1 or nz[0] != lab or v[nz[0]] != 1:
            raise ValueError("the device path encodes one-hot label bitstrings (label l -> entry l)")
    return c, int(sizes[c])


def mpo_encoding(mps_train, y_train, q_hairy_bitstrings=None) -> batched.MPOBatch:
    """deterministic_mpo_classifier.py:33-42 — class-encode every image MPS into an MPO."""
    batch = mps_train if isinstance(mps_train, batched.MPSBatch) else _batch_of(mps_train)
    c, S = _hairy_site(q_hairy_bitstrings, batch.L)
    return batched.class_encode(batch, y_train, c, S)


def adding_centre_batches(list_to_add, D, batch_num=2, orthogonalise=False) -> batched.MPOBatch:
    """experiments.py:497-523.  ``list_to_add``: an ``MPOBatch`` (what the functions here
    return).  As in the reference a shorter remainder group is added at the end."""
    if not isinstance(list_to_add, batched.MPOBatch):
        raise TypeError("adding_centre_batches takes the MPOBatch that mpo_encoding / "
                        "prepare_centred_batched_classifier return")
    return batched.add_compress_centre(list_to_add, D, batch_num, orthogonalise)


def prepare_centred_batched_classifier(mps_train, labels, q_hairy_bitstrings, D_total, batch_nums) -> batched.MPOBatch:
    """experiments.py:482-495: class-encode, bring every MPO to its centre-canonical form
    (``compress_centre_one_site(None, False)``), then add in batches of ``batch_nums[i]`` at bond
    ``D_total`` until at most 10 MPOs are left (the per-class sum states)."""
    mpos = mpo_encoding(mps_train, labels, q_hairy_bitstrings)
    mpos = batched.add_compress_centre(mpos, None, 1, False)
    i = 0
    while len(mpos) > 10:
        mpos = adding_centre_batches(mpos, D_total, batch_nums[i])
        i += 1
    return mpos
