"""CPU: the ensemble evaluators (host integer work mirrored verbatim from
variational_mpo_classifiers.py:340-363) against values the reference's own
functions produced (tests/golden/generate_golden_ensemble.py)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_ensemble.npz"))


@pytest.mark.parametrize("k", [1, 2, 3])
def test_evaluators_match_reference(g, k):
    import orthogonal_classifiers_b200 as oc
    e, y = g["rand_e_pred"], g["rand_y"]
    assert oc.evaluate_soft_ensemble_top_k_accuracy(e, y, k) == float(g[f"rand_soft_k{k}"])
    assert oc.evaluate_hard_ensemble_top_k_accuracy(e, y, k) == float(g[f"rand_hard_k{k}"])
    assert oc.evaluate_classifier_top_k_accuracy(e[0], y, k) == float(g[f"rand_top_k{k}"])


def test_project_classifier_fixes_the_label_leg():
    from orthogonal_classifiers_b200 import project_classifier
    from orthogonal_classifiers_b200.synthetic import random_classifier
    clf = random_classifier(6, 4, centre=3, seed=5)
    L = len(clf)
    for label in (0, 3, 9):
        bits = [np.array([1.0]) if n != 3 else np.eye(16)[label] for n in range(L)]
        proj = project_classifier(clf, bits)
        assert [w.shape for w in proj] == [(2, 1) + w.shape[2:] for w in clf]
        np.testing.assert_array_equal(proj[3][:, 0], clf[3][:, label])
        np.testing.assert_array_equal(proj[0][:, 0], clf[0][:, 0])
