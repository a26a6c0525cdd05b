"""CPU: pin the numpy oracle against fixtures produced by the reference's own
code (tests/golden/generate_golden.py ran /root/reference/{tools,fMPO,
variational_mpo_classifiers,deterministic_mpo_classifier,experiments}.py)."""
import numpy as np
import pytest

import oracle
from conftest import golden_sites

CENTRE_TAGS = ["e32_b32_f32", "e10_b10_f32", "e32_b8_f8"]


def _dense_match(a, b, tol=1e-9):
    """Dense operators agree up to one global sign."""
    return min(np.abs(a - b).max(), np.abs(a + b).max()) < tol


@pytest.mark.parametrize("tag", CENTRE_TAGS)
@pytest.mark.parametrize("ortho", [False, True])
def test_classifier_builder_matches_reference(golden, tag, ortho):
    """fMPO.add + compress_centre_one_site (+ polar) restated == reference."""
    D_encode, D_batch, D_final = [int(t[1:]) for t in tag.split("_")]
    x, y = golden["x_train"], golden["y_train"]
    # the reference run encoded with the xmps stand-in = oracle variant "sweep"
    mps = [oracle.encode_image(im, D_encode, "sweep")[0] for im in x]
    clf = oracle.build_centred_classifier(mps, y, D_batch, D_final, [8, 10], ortho)
    ref = golden_sites(golden, f"clf_{tag}_{'ortho' if ortho else 'nonortho'}")
    assert [w.shape for w in clf] == [w.shape for w in ref]
    Wo = oracle.dense_classifier(oracle.real_sites(clf))
    Wr = oracle.dense_classifier(ref)
    if not ortho:
        assert _dense_match(Wo, Wr)
    # The training images never excite pixels >= 784 (and other dead chunks), so
    # the centre matrix handed to scipy.linalg.polar (fMPO.py:438-444) is rank
    # deficient and its polar factor is only unique on the training span; the
    # completion depends on LAPACK/gauge, in the reference too.  The unique part:
    a = oracle.overlaps_dense(x, Wo)
    b = oracle.overlaps_dense(x, Wr)
    np.testing.assert_allclose(a, b, atol=1e-9)


@pytest.mark.parametrize("tag", CENTRE_TAGS)
@pytest.mark.parametrize("ortho", [False, True])
def test_overlaps_match_reference(golden, tag, ortho):
    """experiments.py:332-337 inline overlap, through quimb-style contraction
    in the reference run vs the oracle's chain contraction."""
    D_encode = int(tag.split("_")[0][1:])
    name = f"{tag}_{'ortho' if ortho else 'nonortho'}"
    clf = golden_sites(golden, f"clf_{name}")
    pred = golden[f"pred_{name}"]
    for k, im in enumerate(golden["x_test"]):
        mps, _ = oracle.encode_image(im, D_encode, "sweep")
        ov = oracle.overlaps_chain(mps, clf)
        assert ov.shape == (16,)
        np.testing.assert_allclose(ov, pred[k], rtol=0, atol=1e-11)
    # per-label form (variational_mpo_classifiers.py:309-318) == entries 0..9
    np.testing.assert_allclose(golden[f"pred10_{name}"], pred[:, :10], atol=1e-12)
    acc = oracle.top_k_accuracy(pred, golden["y_test"], 1)
    acc3 = oracle.top_k_accuracy(pred, golden["y_test"], 3)
    np.testing.assert_allclose([acc, acc3], golden[f"acc_{name}"])


@pytest.mark.parametrize("ortho", [False, True])
def test_last_site_hairy_matches_reference(golden, ortho):
    name = f"e32_b32_f32_last_{'ortho' if ortho else 'nonortho'}"
    clf = golden_sites(golden, f"clf_{name}")
    assert clf[-1].shape[1] == 16 and clf[-1].shape[3] == 1
    pred = golden[f"pred_{name}"]
    for k, im in enumerate(golden["x_test"][:6]):
        mps, _ = oracle.encode_image(im, 32)
        np.testing.assert_allclose(oracle.overlaps_chain(mps, clf), pred[k], atol=1e-11)
    # restated compress_one_site reproduces the reference operator
    x, y = golden["x_train"], golden["y_train"]
    mps = [oracle.encode_image(im, 32, "sweep")[0] for im in x]
    bits = oracle.hairy_bitstrings(10, 10, centre=None)
    mpos = [oracle.class_encode(m, int(l), bits) for m, l in zip(mps, y)]
    sums = oracle.adding_batches_one_site(mpos, 32, 8)
    mine = oracle.adding_batches_one_site(sums, 32, 10, orthogonalise=ortho)[0]
    Wo, Wr = oracle.dense_classifier(oracle.real_sites(mine)), oracle.dense_classifier(clf)
    if not ortho:
        assert _dense_match(Wo, Wr)
    np.testing.assert_allclose(oracle.overlaps_dense(x, Wo), oracle.overlaps_dense(x, Wr), atol=1e-9)


def test_exact_encoding_is_dense_product(golden):
    """D_encode >= 32: overlaps == |x_hat^T W|; both
    truncation variants coincide."""
    clf = golden_sites(golden, "clf_e32_b32_f32_ortho")
    W = oracle.dense_classifier(clf)
    ovd = oracle.overlaps_dense(golden["x_test"], W)
    np.testing.assert_allclose(ovd, golden["pred_e32_b32_f32_ortho"], atol=1e-11)
    for im, ref in zip(golden["x_test"][:4], ovd):
        for variant in ("inline", "sweep"):
            mps, _ = oracle.encode_image(im, 32, variant)
            np.testing.assert_allclose(oracle.overlaps_chain(mps, clf), ref, atol=1e-11)


def test_reference_mps_sites_gauge_invariants(golden):
    """Image MPS produced in the reference run: unit norm, left-canonical,
    reproduces x/|x| with site 0 = MSB of the flat index (tools.py:217-221)."""
    for k in range(3):
        sites = golden_sites(golden, f"mps_e32_img{k}")
        assert len(sites) == 10 and all(s.shape[0] == 2 for s in sites)
        x, _ = oracle.pad_image(golden["x_test"][k])
        st = oracle.mps_to_state(sites)
        xh = x / np.linalg.norm(x)
        assert min(np.abs(st - xh).max(), np.abs(st + xh).max()) < 1e-12
        mine, _ = oracle.encode_image(golden["x_test"][k], 32)
        assert [a.shape for a in mine] == [a.shape for a in sites]
        for A in mine:
            d, l, r = A.shape
            M = A.reshape(d * l, r)
            np.testing.assert_allclose(M.T @ M, np.eye(r), atol=1e-12)


def test_self_overlap_is_one(golden):
    """test_variational_mpo_classifiers.py:446-457 known answer."""
    so = golden["self_overlap"]
    y = golden["y_train"][:3]
    for v, l in zip(so, y):
        assert abs(v[int(l)] - 1.0) < 1e-12
    bits = oracle.hairy_bitstrings(10, 10, centre=5)
    for im, l in zip(golden["x_train"][:3], y):
        mps, _ = oracle.encode_image(im, 32)
        ov = oracle.overlaps_chain(mps, oracle.class_encode(mps, int(l), bits))
        assert abs(ov[int(l)] - 1.0) < 1e-12 and np.sum(ov > 1e-12) == 1


def test_bond_singular_values_and_structure():
    from orthogonal_classifiers_b200.synthetic import synthetic_images
    im = synthetic_images(3, seed=5)
    for x in im:
        sites, sv = oracle.encode_image(x, 32)
        ref = oracle.bond_singular_values(x)
        for a, b in zip(sv, ref):
            n = min(len(a), len(b))
            np.testing.assert_allclose(a[:n], b[:n], atol=1e-12)
        # test_tools.py:238-248: L == ceil(log2(784)) == 10, d == 2
        assert len(sites) == 10 and all(s.shape[0] == 2 for s in sites)
        s10, _ = oracle.encode_image(x, 10)
        assert max(max(s.shape[1:]) for s in s10) == 10


def test_truncation_variants_differ_only_below_32():
    from orthogonal_classifiers_b200.synthetic import synthetic_images
    x = synthetic_images(1, seed=9)[0]
    a = oracle.mps_to_state(oracle.encode_image(x, 4, "inline")[0])
    b = oracle.mps_to_state(oracle.encode_image(x, 4, "sweep")[0])
    assert abs(abs(a @ b) - 1) > 1e-6          # parity unpinned below 32
    assert abs(np.linalg.norm(a) - 1) < 1e-12 and abs(np.linalg.norm(b) - 1) < 1e-12


def test_zero_image_defined_as_zero():
    sites, sv = oracle.encode_image(np.zeros(784), 32)
    assert all(not np.any(s) for s in sites)
