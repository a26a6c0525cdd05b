"""GPU parity tests: the CUDA path (through the C ABI) against the numpy fp64
oracle and the fixtures produced by the reference's own code.

Tolerances (BASELINE.json north_star): gauge-invariant quantities within 1e-5
(fp32, relative to sigma_max / the largest overlap) or 1e-10 (fp64); labels
equal except declared near-ties (top-2 gap < 1e-6)."""
import numpy as np
import pytest

import oracle
from conftest import golden_sites
from orthogonal_classifiers_b200.synthetic import (adversarial_images, random_classifier,
                                                    synthetic_images)

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-10}


def _state_err(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def _ref_encoding(im, D, variant):
    """Oracle sites + the singular values in the convention of svals_out for that variant."""
    sites, sv = oracle.encode_image(im, D, variant)
    if variant == "sweep" and D is not None and D < 2 ** (len(sites) // 2) and im.any():
        sv = oracle.sweep_singular_values(im, D)
    return sites, sv


def _check_encoding(oc, imgs, D, dtype, check_canonical=True, variant="inline"):
    batch = oc.encode_images(imgs, D, dtype=dtype, return_svals=True, trunc=variant)
    sv = batch.svals.cpu().numpy().astype(np.float64)
    tol = TOL[dtype]
    for i, im in enumerate(imgs):
        sites = [s.astype(np.float64) for s in batch.sites_numpy(i)]
        assert [s.shape for s in sites] == [(2, a, b) for a, b in zip(batch.bonds[:-1], batch.bonds[1:])]
        if not im.any():
            assert all(not s.any() for s in sites)
            continue
        ref_sites, ref_sv = _ref_encoding(im, D, variant)
        st = oracle.mps_to_state(sites)
        # unit norm
        assert abs(np.linalg.norm(st) - 1) < 10 * tol
        smax = max(s[0] for s in ref_sv)
        for n in range(batch.L):
            m = min(sv.shape[2], len(ref_sv[n]))
            assert np.abs(sv[i, n, :m] - ref_sv[n][:m]).max() <= tol * max(smax, ref_sv[n][0]), (i, n)
            assert np.all(sv[i, n, m:] <= tol * smax)
        if check_canonical:
            for n, A in enumerate(sites):
                d, l, r = A.shape
                M = A.reshape(d * l, r)
                G = M.T @ M
                nz = np.diag(G) > 0.5
                assert np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max() < 20 * tol, (i, n)
                assert np.abs(G[~nz]).max(initial=0) < 20 * tol
        yield i, st, ref_sites


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_exact_encoding_reproduces_image(gpu_lib, dtype):
    """D >= 32: the MPS is x/|x| exactly."""
    imgs = np.concatenate([synthetic_images(48, seed=3), adversarial_images()])
    for i, st, _ in _check_encoding(gpu_lib, imgs, 32, dtype):
        x, _ = oracle.pad_image(imgs[i])
        assert _state_err(st, x / np.linalg.norm(x)) < TOL[dtype], i


@pytest.mark.parametrize("variant", ["sweep", "inline"])
@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("D", [2, 4, 8, 10, 16, 20])
def test_truncated_encoding_matches_oracle(gpu_lib, dtype, D, variant):
    """D < 32, both truncation orders (include/oc_b200.h): "sweep" = the order of
    the reference's call ``left_from_state(..).left_canonicalise(D)`` (tools.py:221)
    as restated in oracle/stubs/xmps — exact chain, right-to-left truncating sweep,
    sweep back; "inline" = truncation inside the left-to-right chain.  State
    amplitudes, singular values, unit norm and left-canonical form against the
    oracle of the same variant.  Images whose cut falls in a near-degenerate
    pair of singular values are declared ties and skipped."""
    imgs = np.concatenate([synthetic_images(32, seed=4), adversarial_images()])
    checked = 0
    for i, st, ref_sites in _check_encoding(gpu_lib, imgs, D, dtype, variant=variant):
        _, ref_sv = _ref_encoding(imgs[i], D, variant)
        # the kept subspace is determined to ~eps/gap, gap = relative distance of
        # the singular values either side of the cut
        gaps = [(s[D - 1] - s[D]) / s[0] for s in ref_sv if len(s) > D]
        gap = min(gaps) if gaps else 1.0
        eps = 2e-6 if dtype == "f32" else 1e-13
        if gap < 50 * eps:
            continue
        ref = oracle.mps_to_state(ref_sites)
        assert _state_err(st, ref) < eps / gap + 5 * TOL[dtype], (i, D, gap)
        checked += 1
    assert checked >= 20


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("name", ["e10_b10_f32_ortho", "e10_b10_f32_nonortho"])
def test_truncated_predictions_match_reference_fixtures(gpu_lib, golden, dtype, name):
    """D_encode = 10 (BASELINE config 1's encoder): the arrays the reference's own
    ``mps_encoding`` + inline overlap produced (xmps stand-in: sweep truncation) vs
    the CUDA path with the default truncation order, through the batched API and the
    reference-named entry points."""
    clf = golden_sites(golden, f"clf_{name}")
    ref = golden[f"pred_{name}"]
    ov, am = gpu_lib.classify(golden["x_test"], clf, 10, dtype=dtype)
    assert np.abs(ov - ref).max() <= TOL[dtype] * ref.max()
    srt = np.sort(ref, axis=1)
    tie = (srt[:, -1] - srt[:, -2]) < 1e-6
    assert np.all((am == np.argmax(ref, axis=1)) | tie)
    mps = gpu_lib.mps_encoding(golden["x_test"], 10, dtype=dtype)
    p10 = np.array(gpu_lib.classifier_predictions(gpu_lib.data_to_QTN(clf), mps, 10))
    assert np.abs(p10 - golden[f"pred10_{name}"]).max() <= TOL[dtype] * ref.max()
    acc = gpu_lib.evaluate_classifier_top_k_accuracy(ov, golden["y_test"], 1)
    assert acc == golden[f"acc_{name}"][0]
    # the other order is a different state: the fixtures tell the two apart
    ov_inline, _ = gpu_lib.classify(golden["x_test"], clf, 10, dtype=dtype, trunc="inline")
    assert np.abs(ov_inline - ref).max() > 10 * TOL[dtype] * ref.max()


@pytest.mark.parametrize("variant", ["sweep", "inline"])
def test_config1_end_to_end(gpu_lib, variant):
    """BASELINE.json config 1 at full size: 1000 train / 1000 test, D_encode = D_batch = 10,
    D_total = 32.  The classifier is batch-added on the host (oracle restating
    experiments.py:482-523); the 1000 test images go through the CUDA path and through
    the oracle's per-image loop: overlaps within 1e-5, labels equal except near-ties."""
    n = 1000
    y = np.repeat(np.arange(10), n // 10)
    x = synthetic_images(n, seed=41, labels=y)
    mps = [oracle.encode_image(im, 10, variant)[0] for im in x]
    clf = oracle.real_sites(oracle.build_centred_classifier(mps, y, 10, 32, [10, 10, 10], False))
    yt = np.arange(n) % 10
    xt = synthetic_images(n, seed=42, labels=yt)
    ref, ref_am = oracle.classify_reference_loop(xt, clf, 10, variant)
    for dtype in ("f32", "f64"):
        ov, am = gpu_lib.classify(xt, clf, 10, dtype=dtype, trunc=variant)
        err = np.abs(ov - ref).max() / ref.max()
        assert err <= TOL[dtype], (dtype, err)
        srt = np.sort(ref, axis=1)
        gap = srt[:, -1] - srt[:, -2]
        bad = am != ref_am
        assert np.all(gap[bad] < max(1e-6, 4 * err * ref.max())), (dtype, int(bad.sum()))
        if dtype == "f64":
            assert not bad.any()
        acc = gpu_lib.evaluate_classifier_top_k_accuracy(ov, yt, 1)
        assert abs(acc - oracle.top_k_accuracy(ref, yt, 1)) <= bad.sum() / n
    assert oracle.top_k_accuracy(ref, yt, 1) > 0.5   # the synthetic classes are learnable


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("name", ["e32_b32_f32_ortho", "e32_b32_f32_nonortho", "e32_b8_f8_ortho",
                                  "e32_b32_f32_last_ortho", "e32_b32_f32_last_nonortho"])
def test_predictions_match_reference_fixtures(gpu_lib, golden, dtype, name):
    """encode + overlap vs the arrays the reference's own code produced
    (experiments.py:332-337 through tests/golden/generate_golden.py)."""
    clf = golden_sites(golden, f"clf_{name}")
    ref = golden[f"pred_{name}"]
    ov, am = gpu_lib.classify(golden["x_test"], clf, 32, dtype=dtype)
    assert ov.shape == ref.shape == (16, 16)
    assert np.abs(ov - ref).max() <= TOL[dtype] * ref.max()
    srt = np.sort(ref, axis=1)
    tie = (srt[:, -1] - srt[:, -2]) < 1e-6
    assert np.all((am == np.argmax(ref, axis=1)) | tie)
    # mirrored entry points give the same numbers
    mps = gpu_lib.mps_encoding(golden["x_test"], 32, dtype=dtype)
    p10 = np.array(gpu_lib.classifier_predictions(gpu_lib.data_to_QTN(clf), mps, 10))
    assert p10.shape == (16, 10)
    assert np.abs(p10 - ref[:, :10]).max() <= TOL[dtype] * ref.max()
    if "last" not in name:
        acc = gpu_lib.evaluate_classifier_top_k_accuracy(ov, golden["y_test"], 1)
        assert acc == golden[f"acc_{name}"][0]
        inline = np.abs((mps[3].H.squeeze() @ gpu_lib.data_to_QTN(clf).squeeze()).data)
        assert np.abs(inline - ref[3]).max() <= TOL[dtype] * ref.max()


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("D_enc,D_clf,centre", [(32, 32, 5), (10, 32, 5), (4, 8, 5), (32, 36, 9), (16, 20, 9), (2, 2, 5), (32, 32, 9), (10, 32, 9), (20, 32, 5), (3, 32, 5), (7, 32, 9), (12, 16, 5), (32, 16, 9)])
def test_overlap_kernel_matches_chain_oracle(gpu_lib, dtype, D_enc, D_clf, centre):
    """Overlap kernel alone: feed it the GPU's own MPS and compare with the
    oracle contraction of the same site tensors (gauge-proof)."""
    imgs = np.concatenate([synthetic_images(24, seed=6), adversarial_images()])
    clf = random_classifier(10, D_clf, centre=centre, seed=D_clf)
    batch = gpu_lib.encode_images(imgs, D_enc, dtype=dtype)
    ov = gpu_lib.label_overlaps(batch, clf)
    ref = np.array([oracle.overlaps_chain([s.astype(np.float64) for s in batch.sites_numpy(i)], clf)
                    for i in range(len(imgs))])
    assert np.abs(ov - ref).max() <= TOL[dtype] * ref.max()


@pytest.mark.parametrize("n_pix", [520, 832, 840, 896, 900, 1000])
def test_pixel_counts_around_the_kernel_thresholds(gpu_lib, n_pix):
    """L = 10 images whose pixel count sits on either side of the encoder's structural shortcuts (rows 13-15 of
    X_3 empty: n_pix <= 832 / 896), float32 and uint8 pixels (16-byte aligned rows take the integer Gram kernel,
    the others the fp64 one): amplitudes against the normalised image."""
    rng = np.random.default_rng(n_pix)
    u8 = (rng.random((40, n_pix)) * 255 * (rng.random((40, n_pix)) < 0.6)).astype(np.uint8)
    u8[0] = 0
    for imgs in (u8, (u8.astype(np.float32) / 255.0)):
        batch = gpu_lib.encode_images(imgs, 32)
        assert batch.L == 10
        for i in range(40):
            x = np.zeros(1024)
            x[:n_pix] = u8[i].astype(np.float64) / 255.0
            st = oracle.mps_to_state([t.astype(np.float64) for t in batch.sites_numpy(i)])
            if not x.any():
                assert not st.any()
                continue
            assert _state_err(st, x / np.linalg.norm(x)) < 1e-5, (n_pix, imgs.dtype, i)


def test_small_chains_and_untruncated_vectors(gpu_lib):
    """mps_encoding(label_vectors_16, None) (experiments.py:679,689): L = 4."""
    rng = np.random.default_rng(0)
    for n_pix in (2, 3, 16, 100, 512, 1000):
        v = np.abs(rng.standard_normal((5, n_pix)))
        for dtype in ("f32", "f64"):
            batch = gpu_lib.encode_images(v, None, dtype=dtype)
            for i in range(5):
                st = oracle.mps_to_state([s.astype(np.float64) for s in batch.sites_numpy(i)])
                x, L = oracle.pad_image(v[i])
                assert batch.L == L
                assert _state_err(st, x / np.linalg.norm(x)) < TOL[dtype]


def test_empty_and_ragged_batches(gpu_lib):
    imgs = synthetic_images(7, seed=1)
    clf = random_classifier()
    ov0, am0 = gpu_lib.classify(imgs[:0], clf, 32)
    assert ov0.shape == (0, 16) and am0.shape == (0,)
    full, _ = gpu_lib.classify(imgs, clf, 32)
    for n in (1, 3, 5):
        part, _ = gpu_lib.classify(imgs[:n], clf, 32)
        np.testing.assert_array_equal(part, full[:n])     # batch-size independent, bit-exact
    # strided rows / uint8 input
    wide = np.zeros((7, 800), np.float32)
    wide[:, :784] = imgs
    import torch
    ovw, _ = gpu_lib.classify(torch.from_numpy(wide)[:, :784], clf, 32)
    np.testing.assert_array_equal(ovw, gpu_lib.classify(imgs.astype(np.float32), clf, 32)[0])
    u8 = np.round(imgs * 255).astype(np.uint8)
    ovu, _ = gpu_lib.classify(u8, clf, 32)
    assert np.abs(ovu - full).max() < 1e-5 * full.max()


@pytest.mark.parametrize("n_pix", [784, 1024])
def test_uint8_integer_gram_matches_fp64_gram(gpu_lib, n_pix):
    """uint8 pixels take the exact integer Gram matrix G_3 = X_3 X_3^T from the tensor cores
    (mma.sync u8 x u8 -> s32) instead of fp64 FMAs over converted pixels: same singular values,
    states and overlaps as the fp64 route and as the dense fp64 oracle; ragged batch sizes."""
    from orthogonal_classifiers_b200 import batched
    base = np.concatenate([synthetic_images(500, seed=41), adversarial_images()])
    rng = np.random.default_rng(3)
    if n_pix == 1024:
        full = rng.random((len(base), 1024))
        full[:, :784] = np.maximum(full[:, :784] * 0.2, base)
        base = full
    u8 = np.rint(base * 255).astype(np.uint8)
    clf = random_classifier(10, 32, centre=5, seed=6)
    ref = oracle.overlaps_dense(u8.astype(np.float64) / 255.0, oracle.dense_classifier(clf))
    for n in (len(u8), 33, 5, 1):
        new = gpu_lib.encode_images(u8[:n], 32, return_svals=True)
        ov_new = gpu_lib.label_overlaps(new, clf)
        batched.set_option("enc_u8mma", 0)
        try:
            old = gpu_lib.encode_images(u8[:n], 32, return_svals=True)
            ov_old = gpu_lib.label_overlaps(old, clf)
        finally:
            batched.set_option("enc_u8mma", 1)
        sn, so = new.svals.cpu().numpy(), old.svals.cpu().numpy()
        assert np.abs(sn - so).max() <= 2e-6 * so.max()
        assert np.abs(ov_new - ov_old).max() <= 1e-5 * ov_old.max()
        assert np.abs(ov_new - ref[:n]).max() <= 1e-5 * ref.max()
        live = np.abs(u8[:n].astype(np.float64)).sum(1) > 0
        for i in (0, n // 2, n - 1):
            if not live[i]:
                continue
            a = oracle.mps_to_state([x.astype(np.float64) for x in new.sites_numpy(i)])
            x = np.zeros(1024)
            x[:n_pix] = u8[i].astype(np.float64) / 255.0
            x /= np.linalg.norm(x)
            assert min(np.abs(a - x).max(), np.abs(a + x).max()) < 1e-5


def test_count_correct_and_evaluate(gpu_lib):
    import torch
    rng = np.random.default_rng(5)
    am = rng.integers(0, 16, 100_003).astype(np.int32)
    lab = rng.integers(0, 10, 100_003).astype(np.uint8)
    out = gpu_lib.count_correct(torch.from_numpy(am).cuda(), torch.from_numpy(lab).cuda()).cpu().numpy()
    assert out.tolist() == [int((am == lab).sum()), am.size]
    imgs = synthetic_images(64, seed=2, labels=np.arange(64) % 10)
    clf = random_classifier()
    ov, pred = gpu_lib.classify(imgs, clf, 32)
    acc = gpu_lib.evaluate(imgs, np.arange(64) % 10, clf, 32)
    assert acc == np.mean(pred == np.arange(64) % 10)
    assert acc == gpu_lib.evaluate_classifier_top_k_accuracy(ov, np.arange(64) % 10, 1)


def test_full_split_labels_match_dense_oracle(gpu_lib, golden, capsys):
    """10k test split at D_encode = D_total = 32: overlaps == |x_hat^T W|
    (size-independent property).  fp64: labels bit-exact (no tie allowance beyond
    gaps of 1e-10).  fp32: every label that differs from the fp64 reference's sits
    on a top-2 gap below max(1e-6, 4 x the measured overlap error) — the count and
    the error are printed."""
    n = 10_000
    labels = np.arange(n) % 10
    imgs = synthetic_images(n, seed=21, labels=labels)
    clf = golden_sites(golden, "clf_e32_b32_f32_ortho")
    ref = oracle.overlaps_dense(imgs, oracle.dense_classifier(clf))
    ref_am = np.argmax(ref, axis=1)
    srt = np.sort(ref, axis=1)
    gap = srt[:, -1] - srt[:, -2]
    ov64, am64 = gpu_lib.classify(imgs, clf, 32, dtype="f64")
    assert np.abs(ov64 - ref).max() <= 1e-10 * ref.max()
    assert np.all((am64 == ref_am) | (gap < 1e-10))
    ov, am = gpu_lib.classify(imgs.astype(np.float32), clf, 32)
    err = np.abs(ov - ref).max()
    assert err <= 1e-5 * ref.max()
    bad = am != ref_am
    with capsys.disabled():
        print(f"\n[10k split, f32] labels differing from the fp64 reference: {int(bad.sum())} of {n}; "
              f"max overlap error {err:.2e} (largest overlap {ref.max():.3f}); "
              f"largest top-2 gap among them {gap[bad].max() if bad.any() else 0.0:.2e}")
    assert np.all(gap[bad] < max(1e-6, 4 * err))


def test_self_overlap_known_answer(gpu_lib):
    """test_variational_mpo_classifiers.py:446-457: an image against its own
    class-encoded MPO overlaps 1.000."""
    imgs = synthetic_images(4, seed=8)
    bits = oracle.hairy_bitstrings(10, 10, centre=5)
    for k, im in enumerate(imgs):
        sites, _ = oracle.encode_image(im, 32)
        full = [np.zeros((2, a, b)) for a, b in zip([1, 2, 4, 8, 16, 32, 16, 8, 4, 2], [2, 4, 8, 16, 32, 16, 8, 4, 2, 1])]
        for f, s in zip(full, sites):
            f[:, :s.shape[1], :s.shape[2]] = s
        mpo = oracle.class_encode(full, k, bits)
        ov, am = gpu_lib.classify(im[None], mpo, 32, dtype="f64")
        assert abs(ov[0, k] - 1) < 1e-10 and am[0] == k
        ov32, _ = gpu_lib.classify(im[None], mpo, 32)
        assert np.round(ov32[0, k], 3) == 1.0


def test_generic_and_specialised_kernels_agree(gpu_lib):
    """The runtime-shaped kernels (force_generic) and the specialised ones are
    two CUDA implementations of the same path: same gauge-invariants."""
    imgs = np.concatenate([synthetic_images(40, seed=14), adversarial_images()]).astype(np.float32)
    clf = random_classifier(10, 32, centre=5, seed=3)
    for D in (32, 10, 3):
        fast = gpu_lib.encode_images(imgs, D, return_svals=True)
        ov_fast = gpu_lib.label_overlaps(fast, clf)
        gpu_lib.set_force_generic(True)
        try:
            gen = gpu_lib.encode_images(imgs, D, return_svals=True)
            ov_gen = gpu_lib.label_overlaps(gen, clf)
        finally:
            gpu_lib.set_force_generic(False)
        sf, sg = fast.svals.cpu().numpy(), gen.svals.cpu().numpy()
        assert np.abs(sf - sg).max() <= 1e-5 * sg.max()
        if D == 32:
            assert np.abs(ov_fast - ov_gen).max() <= 1e-5 * ov_gen.max()


@pytest.mark.parametrize("option", ["enc_gram", "enc_pack", "enc_mgs4", "enc_qr5", "enc_qr6", "enc_split5", "ov_static", "ov_mma", "ov_pad"])
def test_alternative_code_paths_agree(gpu_lib, option):
    """Every fast path has its predecessor behind an option: the Gram-matrix front
    end (steps 0-3) vs. the Jacobi head kernels, the packed step 4 vs. the paired
    one, the compile-time overlap shapes vs. the run-time ones.  Same
    gauge-invariants from both."""
    from orthogonal_classifiers_b200 import batched
    imgs = np.concatenate([synthetic_images(300, seed=21), adversarial_images()]).astype(np.float32)
    clf = random_classifier(10, 32, centre=5, seed=4)
    for D in (32, 20):
        both = D == 32 or option == "ov_pad"   # (truncated encodings: only the overlap options compare overlaps)
        new = gpu_lib.encode_images(imgs, D, return_svals=True)
        ov_new = gpu_lib.label_overlaps(new, clf) if both else None
        batched.set_option(option, 0)
        try:
            old = gpu_lib.encode_images(imgs, D, return_svals=True)
            ov_old = gpu_lib.label_overlaps(old, clf) if both else None
        finally:
            batched.set_option(option, 1)
        sn, so = new.svals.cpu().numpy(), old.svals.cpu().numpy()
        assert np.abs(sn - so).max() <= 1e-5 * so.max()
        for i in (0, 7, 150, 299, len(imgs) - 1):
            a = oracle.mps_to_state([x.astype(np.float64) for x in new.sites_numpy(i)])
            b = oracle.mps_to_state([x.astype(np.float64) for x in old.sites_numpy(i)])
            assert _state_err(a, b) < 1e-5
        if both:
            assert np.abs(ov_new - ov_old).max() <= 1e-5 * ov_old.max()
    if option.startswith("ov_"):
        # the last-site-hairy classifier (fMPO.compress_one_site) has its own static / tensor-core modes
        clf9 = random_classifier(10, 32, centre=9, seed=5)
        for D in (32, 10):
            enc = gpu_lib.encode_images(imgs, D)
            ov_new = gpu_lib.label_overlaps(enc, clf9)
            batched.set_option(option, 0)
            try:
                ov_old = gpu_lib.label_overlaps(enc, clf9)
            finally:
                batched.set_option(option, 1)
            assert np.abs(ov_new - ov_old).max() <= 1e-5 * ov_old.max()


@pytest.mark.gpu
@pytest.mark.parametrize("n_img", [1, 5, 8 * 148 + 3, 4099])
def test_tcgen05_overlap_mode_matches_mma_sync(gpu_lib, n_img):
    """``ov_tc5`` (opt-in, MODE 6 of the overlap kernel): site 4 of the natural 32/32 profile as one
    shared-operand tcgen05.mma per CTA (A = the images' Y = A4^T E4 in tensor memory, B = the classifier
    site in shared memory, 3xTF32, accumulator in TMEM).  Same overlaps as the mma.sync kernel (MODE 2)
    and as the dense fp64 oracle, for full CTA iterations, a ragged last one and fewer images than a tile."""
    from orthogonal_classifiers_b200 import batched
    imgs = synthetic_images(n_img, seed=31).astype(np.float32)
    clf = random_classifier(10, 32, centre=5, seed=6)
    enc = gpu_lib.encode_images(imgs, 32)
    ov_mma = gpu_lib.label_overlaps(enc, clf)
    batched.set_option("ov_tc5", 1)
    try:
        ov_tc5 = gpu_lib.label_overlaps(enc, clf)
        ov_tc5_again = gpu_lib.label_overlaps(enc, clf)   # a second launch: fresh TMEM allocation, same bits
    finally:
        batched.set_option("ov_tc5", 0)
    ref = oracle.overlaps_dense(imgs, oracle.dense_classifier(clf))
    assert np.abs(ov_tc5 - ov_mma).max() <= 1e-5 * ref.max()
    assert np.abs(ov_tc5 - ref).max() <= 1e-5 * ref.max()
    assert np.array_equal(ov_tc5, ov_tc5_again)


@pytest.mark.gpu
@pytest.mark.parametrize("D", [2, 4, 8, 10, 16, 20])
def test_fused_sweep_classify_is_gauge_free(gpu_lib, D):
    """encode + classify with the reference's truncation order (OC_TRUNC_SWEEP): the fused entry points contract
    the RIGHT-canonical MPS the right-to-left truncating sweep leaves (block positions and bond indices mirrored,
    `OverlapParams::mirrored`) instead of first sweeping back to the left gauge as `oc_encode_batch` must.
    Same state, so the same overlaps - against the two-step path, for every kernel family (zero-padded natural-32
    classifier, last-site-hairy, run-time shapes) and through the host-buffer call."""
    from orthogonal_classifiers_b200 import batched
    imgs = np.concatenate([synthetic_images(700, seed=41), adversarial_images()]).astype(np.float32)
    for clf in (random_classifier(10, 32, centre=5, seed=7), random_classifier(10, 32, centre=9, seed=8),
                random_classifier(10, 20, centre=5, seed=9), random_classifier(10, 12, centre=3, seed=10)):
        ov_fused, am_fused = gpu_lib.classify(imgs, clf, D, trunc="sweep")
        enc = gpu_lib.encode_images(imgs, D, trunc="sweep")            # left-canonical MPS, then the overlap
        ov_two = gpu_lib.label_overlaps(enc, clf)
        batched.set_option("fused_sweep", 0)
        try:
            ov_off, am_off = gpu_lib.classify(imgs, clf, D, trunc="sweep")
        finally:
            batched.set_option("fused_sweep", 1)
        scale = ov_two.max()
        assert np.abs(ov_fused - ov_two).max() <= 1e-5 * scale
        assert np.abs(ov_off - ov_two).max() <= 1e-5 * scale
        srt = np.sort(ov_two, axis=1)
        tie = (srt[:, -1] - srt[:, -2]) < 1e-5 * scale
        assert np.all((am_fused == ov_two.argmax(1)) | tie)
        ov_host, am_host = gpu_lib.classify_host(imgs, clf, D, trunc="sweep")
        assert np.abs(ov_host - ov_two).max() <= 1e-5 * scale


@pytest.mark.gpu
def test_two_stream_chunk_split_is_bit_identical(gpu_lib):
    """``enc_split2`` (opt-in): the two halves of an encoder chunk run on two
    streams and rejoin the caller's stream.  An image's MPS does not depend on
    which other images share its kernel launch, so MPS, singular values,
    overlaps and labels must be bit-identical to the single-stream chain."""
    import torch
    from orthogonal_classifiers_b200 import batched
    base = np.concatenate([synthetic_images(2041, seed=33), adversarial_images()]).astype(np.float32)
    imgs = torch.from_numpy(np.tile(base, (20, 1))[:40001]).cuda()   # >= 32,768: the split engages; odd size
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=6))
    ref = gpu_lib.encode_images(imgs, 32, return_svals=True)
    ov_ref, am_ref = batched.overlaps_from_mps(ref, clf)
    batched.set_option("enc_split2", 1)
    try:
        new = gpu_lib.encode_images(imgs, 32, return_svals=True)
        ov_new, am_new = batched.overlaps_from_mps(new, clf)
        torch.cuda.synchronize()
    finally:
        batched.set_option("enc_split2", 0)
    assert torch.equal(new.data, ref.data)
    assert torch.equal(new.svals, ref.svals)
    assert torch.equal(ov_new, ov_ref) and torch.equal(am_new, am_ref)


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))


def test_encoding_does_not_depend_on_batch_composition(gpu_lib):
    """Images share warps (two per warp in most kernels, up to four of equal
    live-row count in the packed step 4, chosen by a device-side counting sort
    with atomics): an image's MPS must be bit-identical whatever else is in the
    batch and in whatever order."""
    imgs = synthetic_images(257, seed=23).astype(np.float32)
    full = gpu_lib.encode_images(imgs, 32)
    perm = np.random.default_rng(0).permutation(len(imgs))
    shuffled = gpu_lib.encode_images(imgs[perm], 32)
    alone = gpu_lib.encode_images(imgs[5:6], 32)
    for i in (0, 5, 100, 256):
        j = int(np.where(perm == i)[0][0])
        for x, y in zip(full.sites_numpy(i), shuffled.sites_numpy(j)):
            np.testing.assert_array_equal(x, y)
    for x, y in zip(full.sites_numpy(5), alone.sites_numpy(0)):
        np.testing.assert_array_equal(x, y)
    # guests in the packed step 4 (a warp's idle lanes take an image with a shorter line, which then
    # runs in the longer-vector instantiation of its host): same bits as equal-length warps only
    from orthogonal_classifiers_b200 import batched
    many = synthetic_images(3000, seed=29).astype(np.float32)
    mixed = gpu_lib.encode_images(many, 32, return_svals=True)
    batched.set_option("enc_mix", 0)
    try:
        plain = gpu_lib.encode_images(many, 32, return_svals=True)
    finally:
        batched.set_option("enc_mix", 1)
    assert torch_equal(mixed.data, plain.data) and torch_equal(mixed.svals, plain.svals)


def test_packed_step4_guest_plans(gpu_lib):
    """The guest plan of the packed step 4 for line lengths the stroke images do not produce: images with
    exactly k non-zero 64-pixel rows, k = 1..16, in lopsided proportions (more guests than hosts, hosts
    without guests, 14- to 16-lane hosts whose free lanes only fit 2- to 4-lane guests, an empty image).
    Same bits with and without guests; amplitudes against the normalised image."""
    from orthogonal_classifiers_b200 import batched
    rng = np.random.default_rng(17)
    counts = {1: 7, 2: 40, 3: 9, 4: 33, 5: 5, 6: 61, 7: 3, 8: 90, 9: 11, 10: 150, 11: 64, 12: 37, 13: 201, 14: 23, 15: 18, 16: 12}
    imgs = [np.zeros(1024)]
    for k, n in counts.items():
        for _ in range(n):
            x = np.zeros((16, 64))
            x[:k] = rng.random((k, 64)) * (rng.random((k, 1)) + 0.05)
            imgs.append(x.reshape(-1))
    imgs = np.stack(imgs)[rng.permutation(len(imgs))].astype(np.float32)
    mixed = gpu_lib.encode_images(imgs, 32, return_svals=True)
    batched.set_option("enc_mix", 0)
    try:
        plain = gpu_lib.encode_images(imgs, 32, return_svals=True)
    finally:
        batched.set_option("enc_mix", 1)
    assert torch_equal(mixed.data, plain.data) and torch_equal(mixed.svals, plain.svals)
    for i in range(0, len(imgs), 37):
        x = imgs[i].astype(np.float64)
        if not x.any():
            continue
        x /= np.linalg.norm(x)
        a = oracle.mps_to_state([t.astype(np.float64) for t in mixed.sites_numpy(i)])
        assert min(np.abs(a - x).max(), np.abs(a + x).max()) < 1e-5, i


def test_scale_invariance_and_low_contrast(gpu_lib):
    """The MPS is unit norm (the final 1x1 S.V is dropped), so the overlaps must
    not depend on the pixel scale: every threshold in the encoder is relative.
    Also images with almost-dependent rows (low contrast on a constant offset)."""
    imgs = synthetic_images(200, seed=37).astype(np.float64)
    clf = random_classifier(10, 32, centre=5, seed=8)
    W = oracle.dense_classifier(clf)
    ref = oracle.overlaps_dense(imgs, W)
    for scale in (1e-3, 1.0, 255.0, 1e3):
        ov, _ = gpu_lib.classify((imgs * scale).astype(np.float32), clf, 32)
        assert np.abs(ov - ref).max() <= 1e-5 * ref.max(), scale
    # graded spectra: sigma_2 / sigma_1 ~ 1e-3 .. 1e-4 and singular values down to the 1e-6 floor.
    # (Caught a single fp32 rotation at step 8 leaving cos ~ 1e-4 between graded columns, and a
    # one-pass polish leaving 1e-3 between columns with sigma < 3e-5 sigma_max.)  At contrast 1e-3
    # fp32 pixels themselves carry the strokes to only ~6e-5, hence the looser bound there.
    for contrast, tol in ((1e-1, 1e-5), (1e-2, 1e-5), (1e-3, 2e-5)):
        low = (0.5 + contrast * imgs).astype(np.float32)
        ref_low = oracle.overlaps_dense(low.astype(np.float64), W)
        ov, _ = gpu_lib.classify(low, clf, 32)
        assert np.abs(ov - ref_low).max() <= tol * ref_low.max(), contrast
        batch = gpu_lib.encode_images(low[:16], 32)
        for i in range(16):
            for n, A in enumerate(batch.sites_numpy(i)):
                M = A.reshape(-1, A.shape[2]).astype(np.float64)
                G = M.T @ M
                nz = np.diag(G) > 0.5
                assert np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max() < 20 * TOL["f32"], (contrast, i, n)


def test_image_families(gpu_lib):
    """Inputs the MNIST-shaped synthetic strokes do not cover: flat spectra (white
    noise), exactly rank-deficient patterns (ramps, checkerboards, blocks), almost
    rank-1 images (constant + 1e-4 noise), twelve decades of dynamic range."""
    rng = np.random.default_rng(5)
    n = 96
    yy, xx = np.mgrid[0:28, 0:28]
    fam = {
        "white noise": rng.random((n, 784)),
        "sparse pixels": (rng.random((n, 784)) < 0.01) * rng.random((n, 784)),
        "ramps": np.stack([((a * xx + b * yy) % 28 / 27.0).reshape(-1) for a, b in rng.integers(0, 5, (n, 2))]),
        "checker": np.stack([(((xx // p) + (yy // q)) % 2).reshape(-1).astype(float)
                             for p, q in rng.integers(1, 8, (n, 2))]),
        "blocks": np.stack([((xx > a) & (xx < a + 9) & (yy > b) & (yy < b + 6)).reshape(-1).astype(float)
                            for a, b in rng.integers(0, 18, (n, 2))]),
        "const + noise": 0.7 + 1e-4 * rng.standard_normal((n, 784)),
        "huge range": rng.random((n, 784)) ** 12,
    }
    clf = random_classifier(10, 32, centre=5, seed=9)
    W = oracle.dense_classifier(clf)
    for name, imgs in fam.items():
        x32 = imgs.astype(np.float32)
        ref = oracle.overlaps_dense(x32.astype(np.float64), W)
        ov, am = gpu_lib.classify(x32, clf, 32)
        assert np.abs(ov - ref).max() <= 1e-5 * ref.max(), name
        srt = np.sort(ref, axis=1)
        tie = (srt[:, -1] - srt[:, -2]) < 1e-5 * ref.max()
        assert np.all((am == ref.argmax(1)) | tie), name
        batch = gpu_lib.encode_images(x32[:8], 32)
        for i in range(8):
            for A in batch.sites_numpy(i):
                M = A.reshape(-1, A.shape[2]).astype(np.float64)
                G = M.T @ M
                nz = np.diag(G) > 0.5
                assert np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max() < 20 * TOL["f32"], name


def test_full_32x32_images(gpu_lib):
    """BASELINE.json config 5 uses 32 x 32 = 1024-pixel images: all 16 rows of the 16 x 64
    unfolding and all 32 of the 32 x 32 one are live (784-pixel images leave 3 / 7 of them
    zero), which selects the 16-lane step-4 and the full-length step-5 kernels."""
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:32, 0:32]
    blobs = np.stack([np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * w ** 2)).reshape(-1)
                      for cx, cy, w in zip(rng.uniform(4, 28, 150), rng.uniform(4, 28, 150), rng.uniform(1, 6, 150))])
    imgs = np.concatenate([blobs, rng.random((150, 1024))]).astype(np.float32)
    clf = random_classifier(10, 32, centre=5, seed=12)
    ref = oracle.overlaps_dense(imgs.astype(np.float64), oracle.dense_classifier(clf))
    ov, am = gpu_lib.classify(imgs, clf, 32)
    assert np.abs(ov - ref).max() <= 1e-5 * ref.max()
    srt = np.sort(ref, axis=1)
    assert np.all((am == ref.argmax(1)) | ((srt[:, -1] - srt[:, -2]) < 1e-5 * ref.max()))
    batch = gpu_lib.encode_images(imgs[::25], 32, return_svals=True)
    sv = batch.svals.cpu().numpy().astype(np.float64)
    for j, i in enumerate(range(0, len(imgs), 25)):
        ref_sv = oracle.bond_singular_values(imgs[i].astype(np.float64))
        for n in range(10):
            m = min(sv.shape[2], len(ref_sv[n]))
            assert np.abs(sv[j, n, :m] - ref_sv[n][:m]).max() <= 1e-5 * ref_sv[0][0], (i, n)


def test_chunk_boundary_of_the_encoder(gpu_lib):
    """The encoder works in chunks of 131,072 images (workspace and sort lists are
    per chunk): images on both sides of the boundary must come out as they do in
    a small batch."""
    import torch
    base = synthetic_images(512, seed=29).astype(np.float32)
    n = 131072 + 700
    reps = (n + len(base) - 1) // len(base)
    big = torch.from_numpy(np.tile(base, (reps, 1))[:n]).cuda()
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=6))
    from orthogonal_classifiers_b200 import batched
    ov_big, am_big = batched.classify_device(big, clf, 32)
    ov_small, am_small = batched.classify_device(big[:512].contiguous(), clf, 32)
    ov_big, am_big = ov_big.cpu().numpy(), am_big.cpu().numpy()
    ov_small, am_small = ov_small.cpu().numpy(), am_small.cpu().numpy()
    for i in (0, 131071, 131072, 131073, n - 1):
        np.testing.assert_array_equal(ov_big[i], ov_small[i % 512])
        assert am_big[i] == am_small[i % 512]


def test_host_pipeline_matches_single_shot(gpu_lib):
    """Chunked, copy/compute-overlapped host API == one-shot classify, bit-exact."""
    import torch
    imgs = synthetic_images(1000, seed=31).astype(np.float32)
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=2))
    ref_ov, ref_am = gpu_lib.classify(imgs, clf, 32)
    pipe = gpu_lib.HostPipeline(clf, 784, 32, chunk=300)
    pin = torch.from_numpy(imgs).pin_memory()
    ov = torch.empty((1000, 16), dtype=torch.float32).pin_memory()
    am = torch.empty((1000,), dtype=torch.int32).pin_memory()
    for _ in range(2):
        ov.zero_(); am.zero_()
        pipe.run(pin, ov, am)
        np.testing.assert_array_equal(ov.numpy().astype(np.float64), ref_ov)
        np.testing.assert_array_equal(am.numpy(), ref_am)


def test_ensemble_and_padded_predictions_match_reference(gpu_lib, golden):
    """variational_mpo_classifiers.py:321-337 on the reference's own outputs
    (tests/golden/generate_golden_ensemble.py)."""
    import os
    ge = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_ensemble.npz"))
    oc = gpu_lib
    mps_test = oc.mps_encoding(golden["x_test"], 32)
    ensemble = [oc.data_to_QTN(golden_sites(golden, f"clf_{n}")) for n in ge["names"]]
    e = np.array(oc.ensemble_predictions(ensemble, mps_test, 10))
    ref = ge["e_pred"]
    assert e.shape == ref.shape
    assert np.abs(e - ref).max() < 2e-5 * ref.max()
    e_list = oc.ensemble_predictions(ensemble, mps_test, 10)   # carries the device votes (k = 1)
    assert e_list.soft_argmax is not None
    for k in (1, 3):
        assert oc.evaluate_soft_ensemble_top_k_accuracy(e, golden["y_test"], k) == float(ge[f"soft_k{k}"])
        assert oc.evaluate_hard_ensemble_top_k_accuracy(e, golden["y_test"], k) == float(ge[f"hard_k{k}"])
        assert oc.evaluate_soft_ensemble_top_k_accuracy(e_list, golden["y_test"], k) == float(ge[f"soft_k{k}"])
        assert oc.evaluate_hard_ensemble_top_k_accuracy(e_list, golden["y_test"], k) == float(ge[f"hard_k{k}"])
    # padded: explicit product bitstrings (second padding scaled by 0.5 on the hairy site)
    L, c = 10, 5
    def bits(label, scale):
        return [np.array([1.0]) if n != c else scale * np.eye(16)[label] for n in range(L)]
    padded = [[bits(l, 1.0) for l in range(10)], [bits(l, 0.5) for l in range(10)]]
    p = np.array(oc.padded_classifier_predictions(ensemble[0], mps_test, padded))
    assert p.shape == ge["padded"].shape
    assert np.abs(p - ge["padded"]).max() < 2e-5 * ge["padded"].max()
    # the general-bitstring route of classifier_predictions agrees with the one-hot fast path
    fast = np.array(oc.classifier_predictions(ensemble[0], mps_test, 10))
    slow = np.array(oc.classifier_predictions(ensemble[0], mps_test, [bits(l, 0.5) for l in range(10)]))
    assert np.abs(2 * slow - fast).max() < 2e-5 * fast.max()
