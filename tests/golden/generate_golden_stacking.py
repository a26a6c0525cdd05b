"""Golden fixtures for the MPS-stacking path (experiments.py:528-640), produced by the
REFERENCE'S OWN ``mps_stacking`` run unmodified on small synthetic label vectors.
Authoring container only (needs /root/reference); third-party imports resolve to
oracle/stubs.  ``mps_stacking`` returns only the accuracy, so the stacking unitary and the
stacked predictions are captured where the reference's code hands them on
(``data_to_QTN`` / ``evaluate_classifier_top_k_accuracy`` in the experiments namespace).

    python tests/golden/generate_golden_stacking.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import variational_mpo_classifiers as vmc  # noqa: E402  (reference)
import experiments as ex  # noqa: E402  (reference)


def label_vectors(n, y, seed):
    """Normalised, non-negative 16-vectors peaked at the label (what
    tensor_network_stacking_experiment feeds in: rows of a prediction array, normalised)."""
    rng = np.random.default_rng(seed)
    v = 0.3 * rng.random((n, 16))
    v[:, 10:] *= 0.05
    v[np.arange(n), y] += 1.0 + 0.5 * rng.random(n)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def main():
    n_per_class, n_test = 4, 24
    y_train = np.repeat(np.arange(10), n_per_class)      # "one class" arrangement, 4 per class:
    y_test = np.arange(n_test) % 10                      # batch_nums[0] = 4 leaves 10 sum states
    v_train = label_vectors(len(y_train), y_train, 5)
    v_test = label_vectors(n_test, y_test, 6)
    out = dict(v_train=v_train, y_train=y_train, v_test=v_test, y_test=y_test)
    train_mps = vmc.mps_encoding(v_train, None)          # experiments.py:679
    test_mps = vmc.mps_encoding(v_test, None)            # experiments.py:689
    captured = {}
    real_eval, real_d2q = ex.evaluate_classifier_top_k_accuracy, ex.data_to_QTN

    def cap_eval(pred, y, k):
        captured["pred"] = np.array(pred)
        return real_eval(pred, y, k)

    def cap_d2q(data):
        captured["unitary"] = [np.array(w) for w in data]
        return real_d2q(data)

    ex.evaluate_classifier_top_k_accuracy, ex.data_to_QTN = cap_eval, cap_d2q
    try:
        for n_copies in range(4):                        # L = 4, 8, 12, 16
            acc = ex.mps_stacking(train_mps, test_mps, n_copies, y_train, y_test)
            out[f"acc_{n_copies}"] = np.array(acc)
            out[f"pred_{n_copies}"] = captured["pred"]
            for i, w in enumerate(captured["unitary"]):
                assert np.max(np.abs(np.imag(w)), initial=0.0) == 0.0
                out[f"unitary_{n_copies}_site_{i}"] = np.ascontiguousarray(np.real(w))
            print(n_copies, "acc", acc, "L", len(captured["unitary"]),
                  "bonds", [w.shape[3] for w in captured["unitary"]])
    finally:
        ex.evaluate_classifier_top_k_accuracy, ex.data_to_QTN = real_eval, real_d2q
    np.savez_compressed(os.path.join(HERE, "reference_golden_stacking.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(HERE, "reference_golden_stacking.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
