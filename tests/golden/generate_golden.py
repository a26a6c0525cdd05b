"""Generate the golden fixtures by running the REFERENCE'S OWN code.

Runs only in the authoring container (needs /root/reference, read-only).  The
reference's modules are imported unmodified; their un-vendored third-party
imports resolve to ``oracle/stubs`` (see its README for what is restated).
Nothing in the test-suite or on the GPU box imports /root/reference: the tests
read the ``.npz`` files this script writes next to itself.

    python tests/golden/generate_golden.py
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
import deterministic_mpo_classifier as dmc  # noqa: E402  (reference)
import experiments as ex  # noqa: E402  (reference)
from fMPO import fMPO  # noqa: E402  (reference)

from orthogonal_classifiers_b200.synthetic import synthetic_images  # noqa: E402


def real(sites):
    out = []
    for w in sites:
        w = np.asarray(w)
        assert np.max(np.abs(np.imag(w)), initial=0.0) == 0.0
        out.append(np.ascontiguousarray(np.real(w), dtype=np.float64))
    return out


def site_dict(prefix, sites):
    return {f"{prefix}_site_{i}": w for i, w in enumerate(sites)}


def main():
    n_per_class, n_test = 8, 16
    n_train = 10 * n_per_class
    # "one of each" order 0..9,0..9,... as load_data(equal_numbers=True) returns
    y_train = np.array(list(range(10)) * n_per_class)
    x_train = synthetic_images(n_train, seed=11, labels=y_train)
    y_test = np.arange(n_test) % 10
    x_test = synthetic_images(n_test, seed=12, labels=y_test)
    # tools.py:77-100
    x_arr, y_arr = tools.arrange_data(x_train, y_train, arrangement="one class")

    n_sites = int(np.ceil(np.log2(x_arr.shape[-1])))
    possible_labels = list(range(10))
    hairy = vmc.create_hairy_bitstrings_data(possible_labels, n_sites)
    q_last = tools.bitstring_data_to_QTN(hairy, n_sites)
    q_centre = ex.centred_bitstring_to_qtn([ex.label_last_site_to_centre(b) for b in q_last])

    out = dict(x_train=x_arr, y_train=y_arr, x_test=x_test, y_test=y_test)

    for D_encode, D_batch, D_final in [(32, 32, 32), (10, 10, 32), (32, 8, 8)]:
        tag = f"e{D_encode}_b{D_batch}_f{D_final}"
        mps_train = vmc.mps_encoding(x_arr, D_encode)
        mps_test = vmc.mps_encoding(x_test, D_encode)
        if D_encode == 32:
            # the image MPS sites themselves (gauge-dependent; used only for
            # gauge-invariant checks) for the first 3 test images
            for k in range(3):
                out.update(site_dict(f"mps_e32_img{k}", [t.data for t in mps_test[k].tensors]))

        # --- centre-hairy classifier: experiments.py:482-523 --------------
        sums = ex.prepare_centred_batched_classifier(mps_train, y_arr, q_centre, D_batch, [n_per_class])
        sums_data = [fMPO([np.array(s) for s in m.data]) for m in sums]
        for ortho in (False, True):
            clf = ex.adding_centre_batches(
                [fMPO([np.array(s) for s in m.data]) for m in sums_data], D_final, 10,
                orthogonalise=ortho)[0]
            q_clf = tools.data_to_QTN(clf.data).squeeze()
            # experiments.py:332-337 inline overlap
            preds = np.array([np.abs((m.H.squeeze() @ q_clf.squeeze()).data) for m in mps_test])
            acc = vmc.evaluate_classifier_top_k_accuracy(preds, y_test, 1)
            acc3 = vmc.evaluate_classifier_top_k_accuracy(preds, y_test, 3)
            # variational_mpo_classifiers.py:309-318 per-label form
            preds10 = np.array(vmc.classifier_predictions(q_clf, mps_test, q_centre))
            name = f"{tag}_{'ortho' if ortho else 'nonortho'}"
            out.update(site_dict(f"clf_{name}", real(clf.data)))
            out[f"pred_{name}"] = preds
            out[f"pred10_{name}"] = preds10
            out[f"acc_{name}"] = np.array([acc, acc3])
            print(name, "acc", acc, acc3, "shapes", [w.shape for w in clf.data][4:7])

        # --- last-site-hairy classifier: deterministic_mpo_classifier.py ----
        if D_encode == 32 and D_final == 32:
            sums = dmc.prepare_batched_classifier(mps_train, y_arr, q_last, D_batch, [n_per_class], True)
            for ortho in (False, True):
                clf = dmc.adding_batches(
                    [fMPO([np.array(s) for s in m.data]) for m in sums], D_final, 10,
                    orthogonalise=ortho)[0]
                q_clf = tools.data_to_QTN(clf.data).squeeze()
                preds = np.array([np.abs((m.H.squeeze() @ q_clf.squeeze()).data) for m in mps_test])
                name = f"{tag}_last_{'ortho' if ortho else 'nonortho'}"
                out.update(site_dict(f"clf_{name}", real(clf.data)))
                out[f"pred_{name}"] = preds
                print(name, "shapes", [w.shape for w in clf.data][-3:])

    # one class-encoded image as its own classifier: the reference's
    # known-answer test (test_variational_mpo_classifiers.py:446-457) —
    # overlap of an image with its own class-encoded MPO is 1.
    mps_train = vmc.mps_encoding(x_arr[:3], 32)
    mpo = dmc.mpo_encoding(mps_train, y_arr[:3], q_centre)
    self_ov = [np.abs((m.H.squeeze() @ o.squeeze()).data) for m, o in zip(mps_train, mpo)]
    out["self_overlap"] = np.array(self_ov)
    print("self overlaps", [float(v[int(y)]) for v, y in zip(self_ov, y_arr[:3])])

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"),
          os.path.getsize(os.path.join(HERE, "reference_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
