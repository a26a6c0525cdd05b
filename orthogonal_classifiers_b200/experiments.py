"""Mirror of the hot-path pieces of /root/reference/experiments.py, batched on the GPU:

* the inline prediction loops ``[np.abs((mps.H.squeeze() @ clf.squeeze()).data) for mps in ...]``
  over a list of classifiers and the ``(n_D, N, 16)`` arrays they are saved as
  (experiments.py:278-296, 332-352, 390-391, 439-440);
* MPS stacking: ``generate_copy_state`` and the stacked predictions
  (experiments.py:528-546, 631-636).

Classifier construction (``adding_centre_batches`` ...) is in ``deterministic_mpo_classifier.py``;
``mps_stacking`` here runs the whole stacking experiment on the device.
"""
from __future__ import annotations

import os

import numpy as np

from . import batched
from .variational_mpo_classifiers import _batch_of, evaluate_classifier_top_k_accuracy


"""
Prediction arrays (experiments.py:332-352)
"""


def prediction_array(mps_or_images, classifiers, D_encode=32, dtype="f32", trunc=None) -> np.ndarray:
    """``[[np.abs((mps.H.squeeze() @ clf.squeeze()).data) for mps in mps_images] for clf in
    classifiers]`` -> float64 array ``(len(classifiers), N, 16)``: the layout of the
    ``*_d_final_vs_{training,test}_predictions`` files the reference's plotting scripts read
    (experiments.py:340-352; one classifier per D_final / D_batch value).  The images are
    encoded once; every classifier is one overlap launch over the whole batch."""
    if isinstance(mps_or_images, batched.MPSBatch):
        batch = mps_or_images
    elif hasattr(mps_or_images, "batch") or (isinstance(mps_or_images, list) and mps_or_images
                                             and hasattr(mps_or_images[0], "batch")):
        batch = _batch_of(mps_or_images)
    else:
        batch = batched.encode_images(np.asarray(mps_or_images), D_encode, dtype=dtype, trunc=trunc)
    out = []
    for clf in classifiers:
        ov, _ = batched.overlaps_from_mps(batch, clf, return_argmax=False)
        out.append(ov.double().cpu().numpy())
    return np.stack(out) if out else np.zeros((0, len(batch), 0))


def save_prediction_arrays(directory, ortho=None, non_ortho=None, kind="test", sweep="d_final",
                           compressed=None):
    """Write prediction arrays under the reference's file names (experiments.py:349-352):
    ``{ortho,non_ortho}_{sweep}_vs_{kind}_predictions[.npy | _compressed.npz]``.  The reference
    saves test predictions with ``np.save`` and training predictions with
    ``np.savez_compressed`` (``arr_0``); ``compressed`` overrides that choice."""
    os.makedirs(directory, exist_ok=True)
    if compressed is None:
        compressed = kind == "training"
    paths = []
    for name, arr in (("ortho", ortho), ("non_ortho", non_ortho)):
        if arr is None:
            continue
        arr = np.asarray(arr)
        if arr.ndim != 3:
            raise ValueError("prediction arrays are (n_D, N, 16)")
        stem = os.path.join(directory, f"{name}_{sweep}_vs_{kind}_predictions")
        if compressed:
            np.savez_compressed(stem + "_compressed", arr)
            paths.append(stem + "_compressed.npz")
        else:
            np.save(stem, arr)
            paths.append(stem + ".npy")
    return paths


def load_prediction_array(path):
    """Read back either form (the reference: ``np.load(...)['arr_0']`` for the compressed one,
    experiments.py:657)."""
    if path.endswith(".npz"):
        return np.load(path, allow_pickle=False)["arr_0"]
    return np.load(path, allow_pickle=False)


"""
Stacking (experiments.py:528-640)
"""


def generate_copy_state(mps, n_copies: int) -> batched.MPSBatch:
    """experiments.py:530-546: the product of ``n_copies + 1`` copies of every MPS, sites
    relabelled k0..k{L'-1}.  The edge bonds of an MPS are 1, so the product state's MPS is the
    copies' sites side by side: on the batched layout, the per-image block tiled.
    ``mps``: what ``mps_encoding`` returned (or an ``MPSBatch``)."""
    if n_copies < 0:
        raise ValueError("n_copies must be >= 0")
    b = mps if isinstance(mps, batched.MPSBatch) else _batch_of(mps)
    if b.bonds[0] != 1 or b.bonds[-1] != 1:
        raise ValueError("edge bonds must be 1")
    k = n_copies + 1
    L = b.L * k
    if L > batched._lib.OC_MAX_SITES:
        raise ValueError(f"{L} sites > {batched._lib.OC_MAX_SITES}")
    bonds = list(b.bonds[:-1]) * k + [1]
    per = b.offsets[-1]
    offsets = [c * per + o for c in range(k) for o in b.offsets[:-1]] + [k * per]
    return batched.MPSBatch(b.data.repeat(1, k), bonds, offsets, L, b.D, b.dtype)


def stacked_predictions(test_mps_predictions, stacking_unitary, n_copies: int) -> np.ndarray:
    """experiments.py:631-633: ``np.abs((copy_state.H.squeeze() @ stacking_unitary.squeeze()).data)``
    for every test prediction state -> ``(N, 16)``."""
    copies = generate_copy_state(test_mps_predictions, n_copies)
    return batched.label_overlaps(copies, stacking_unitary)


def evaluate_stacking(test_mps_predictions, stacking_unitary, n_copies: int, y_test) -> float:
    """experiments.py:631-643: accuracy of the stacked predictions."""
    return evaluate_classifier_top_k_accuracy(stacked_predictions(test_mps_predictions, stacking_unitary, n_copies),
                                              y_test, 1)


def mps_stacking(training_mps_predictions, test_mps_predictions, n_copies, y_train, y_test,
                 D_batch=32, batch_nums=(4, 8, 13, 13), D_final=32, batch_final=10, ortho_at_end=True,
                 return_details=False):
    """experiments.py:528-643, end to end on the device: copy states of the training label
    vectors -> class-encoded MPOs -> batch-added sum states (``prepare_centred_batched_classifier``,
    D_batch, batch_nums) -> stacking unitary (``adding_centre_batches`` at D_final, polar factor
    when ``ortho_at_end``) -> overlaps of the test copy states with it -> top-1 accuracy.
    The defaults are the reference's hard-coded parameters (experiments.py:596-610)."""
    from .deterministic_mpo_classifier import adding_centre_batches, prepare_centred_batched_classifier
    train_copies = generate_copy_state(training_mps_predictions, n_copies)
    sum_states = prepare_centred_batched_classifier(train_copies, y_train, None, D_batch, list(batch_nums))
    stacking_unitary = adding_centre_batches(sum_states, D_final, batch_final, orthogonalise=ortho_at_end)[0]
    preds = stacked_predictions(test_mps_predictions, stacking_unitary.data, n_copies)
    result = evaluate_classifier_top_k_accuracy(preds, y_test, 1)
    if return_details:
        return result, preds, stacking_unitary
    return result
