"""Mirror of the hot-path entry points of
/root/reference/variational_mpo_classifiers.py (same names and argument
meaning), backed by the CUDA library."""
from __future__ import annotations

import numpy as np

from . import batched
from .tools import MPONetwork, MPSNetwork


class MPSList(list):
    """``mps_encoding``'s return value: a list of per-image networks (as in
    the reference) that also keeps the device-resident batch so that
    ``classifier_predictions`` / ``label_overlaps`` run batched."""

    batch: batched.MPSBatch | None = None


"""
MPS Encoding
"""


def mps_encoding(images, D=2, dtype="f32"):
    """variational_mpo_classifiers.py:102-104 —
    ``[fMPS_to_QTN(image_to_mps(image, D)) for image in images]``."""
    batch = batched.encode_images(np.asarray(images), D, dtype=dtype)
    out = MPSList(MPSNetwork(batch=batch, index=i) for i in range(len(batch)))
    out.batch = batch
    return out


"""
Evaluate classifier
"""


def _batch_of(mps_test) -> batched.MPSBatch:
    b = getattr(mps_test, "batch", None)
    if isinstance(b, batched.MPSBatch) and len(b) == len(mps_test):
        return b
    if isinstance(mps_test, batched.MPSBatch):
        return mps_test
    # a plain list of networks (e.g. a slice): gather rows of the shared batch
    first = mps_test[0]
    if first.batch is not None and all(m.batch is first.batch for m in mps_test):
        import torch
        idx = torch.as_tensor([m.index for m in mps_test], device=first.batch.data.device)
        b = first.batch
        return batched.MPSBatch(b.data.index_select(0, idx), b.bonds, b.offsets, b.L, b.D, b.dtype)
    raise TypeError("mps_test must come from mps_encoding")


def _label_indices(q_hairy_bitstrings, S):
    """Which label-leg entry each bitstring selects.  The reference's
    bitstrings are one-hot on the hairy site (variational_mpo_classifiers.py:
    39-51, tools.py:160-166), so ``clf @ b`` picks entry ``label``."""
    if q_hairy_bitstrings is None:
        return list(range(S))
    if isinstance(q_hairy_bitstrings, int):
        return list(range(q_hairy_bitstrings))
    idx = []
    for b in q_hairy_bitstrings:
        sites = [np.asarray(getattr(t, "data", t)).reshape(-1) for t in getattr(b, "tensors", b)]
        hairy = max(sites, key=lambda v: v.size)
        nz = np.nonzero(hairy)[0]
        if nz.size != 1 or hairy[nz[0]] != 1:
            raise ValueError("only one-hot label bitstrings are supported on the GPU path")
        idx.append(int(nz[0]))
    return idx


def classifier_predictions(mpo_classifier, mps_test, q_hairy_bitstrings=10):
    """variational_mpo_classifiers.py:309-318 — N x n_labels
    ``abs(<image| classifier |bitstring>)``; one batched contraction instead
    of N x n_labels quimb contractions."""
    batch = _batch_of(mps_test)
    ov, _ = batched.overlaps_from_mps(batch, mpo_classifier, return_argmax=False)
    ov = ov.double().cpu().numpy()
    cols = _label_indices(q_hairy_bitstrings, ov.shape[1])
    return [list(row) for row in ov[:, cols]]


def evaluate_classifier_top_k_accuracy(predictions, y_test, k):
    """variational_mpo_classifiers.py:340-345 (host integer work, verbatim
    semantics: argpartition top-k, membership, mean)."""
    top_k_predictions = [np.argpartition(np.asarray(p), -k)[-k:] for p in predictions]
    return np.mean([int(i in j) for i, j in zip(y_test, top_k_predictions)])
