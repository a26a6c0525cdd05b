"""Golden fixtures for the ensemble / padded prediction variants, produced by
the REFERENCE'S OWN functions (variational_mpo_classifiers.py:321-363) on the
classifiers stored in reference_golden.npz.  Authoring container only (needs
/root/reference); third-party imports resolve to oracle/stubs.

    python tests/golden/generate_golden_ensemble.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tools  # noqa: E402  (reference)
import variational_mpo_classifiers as vmc  # noqa: E402  (reference)
import experiments as ex  # noqa: E402  (reference)


def main():
    g = np.load(os.path.join(HERE, "reference_golden.npz"))
    x_test, y_test = g["x_test"], g["y_test"]
    n_sites = 10
    hairy = vmc.create_hairy_bitstrings_data(list(range(10)), n_sites)
    q_last = tools.bitstring_data_to_QTN(hairy, n_sites)
    q_centre = ex.centred_bitstring_to_qtn([ex.label_last_site_to_centre(b) for b in q_last])
    mps_test = vmc.mps_encoding(x_test, 32)
    names = ["e32_b32_f32_nonortho", "e32_b32_f32_ortho", "e32_b8_f8_ortho"]
    ensemble = []
    for name in names:
        sites = [g[f"clf_{name}_site_{i}"] for i in range(n_sites)]
        ensemble.append(tools.data_to_QTN(sites).squeeze())
    e_pred = np.array(vmc.ensemble_predictions(ensemble, mps_test, q_centre))
    out = dict(names=np.array(names), e_pred=e_pred)
    for k in (1, 3):
        out[f"soft_k{k}"] = np.array(vmc.evaluate_soft_ensemble_top_k_accuracy(e_pred, y_test, k))
        out[f"hard_k{k}"] = np.array(vmc.evaluate_hard_ensemble_top_k_accuracy(e_pred, y_test, k))
    # padded predictions: two "paddings", the second with every bitstring scaled by 0.5
    # on its hairy site, so the result is 1.5 x the plain predictions
    import quimb.tensor as qtn
    half = [qtn.TensorNetwork([qtn.Tensor((0.5 if t.data.size > 2 else 1.0) * t.data, t.inds, t.tags)
                               for t in b.tensors]) for b in q_centre]
    out["padded"] = np.array(vmc.padded_classifier_predictions(ensemble[0], mps_test, [q_centre, half]))
    # the evaluators on predictions that are not all correct (random members)
    rng = np.random.default_rng(7)
    rand_y = rng.integers(0, 10, 64)
    rand_e = rng.random((5, 64, 10))
    rand_e[:, np.arange(64), rand_y] += 0.35
    rand_e /= rand_e.sum(2, keepdims=True)
    out["rand_e_pred"], out["rand_y"] = rand_e, rand_y
    for k in (1, 2, 3):
        out[f"rand_soft_k{k}"] = np.array(vmc.evaluate_soft_ensemble_top_k_accuracy(rand_e, rand_y, k))
        out[f"rand_hard_k{k}"] = np.array(vmc.evaluate_hard_ensemble_top_k_accuracy(rand_e, rand_y, k))
        out[f"rand_top_k{k}"] = np.array(vmc.evaluate_classifier_top_k_accuracy(rand_e[0], rand_y, k))
    np.savez_compressed(os.path.join(HERE, "reference_golden_ensemble.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    print("soft", out["soft_k1"], out["soft_k3"], "hard", out["hard_k1"], out["hard_k3"])


if __name__ == "__main__":
    main()
