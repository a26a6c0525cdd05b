"""GPU tests of the entry points either side of the kernels: the host-buffer C-ABI
call, the generic-L overlap (MPS stacking shapes), classifier files, trimming,
and the per-(device, stream) library state."""
import os

import numpy as np
import pytest

import oracle
from orthogonal_classifiers_b200.synthetic import (adversarial_images, random_classifier,
                                                    synthetic_images)

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-10}


def _random_mps(rng, bonds, n):
    """n random (not normalised, not canonical) MPSs with the given bond profile, packed
    the way include/oc_b200.h lays a batch out."""
    sites = [[rng.standard_normal((2, a, b)) / np.sqrt(a) for a, b in zip(bonds[:-1], bonds[1:])] for _ in range(n)]
    flat = np.stack([np.concatenate([s.reshape(-1) for s in img]) for img in sites])
    return sites, flat


def _random_mpo(rng, bonds, c, S):
    return [rng.standard_normal((2, S if n == c else 1, a, b)) / np.sqrt(2 * a)
            for n, (a, b) in enumerate(zip(bonds[:-1], bonds[1:]))]


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("L,c,amax,Bmax", [(4, 0, 4, 4), (4, 3, 4, 4), (4, 1, 2, 4), (8, 2, 8, 16), (8, 7, 16, 16),
                                           (16, 8, 8, 32), (16, 0, 4, 8), (16, 15, 32, 32), (12, 5, 16, 12),
                                           (20, 13, 6, 10), (64, 40, 4, 4)])
def test_overlap_any_length_and_hairy_site(gpu_lib, dtype, L, c, amax, Bmax):
    """The contraction of experiments.py:633 (MPS stacking: L = 4 (n + 1) sites, label leg
    wherever the stacking MPO carries it) — chains of 4..64 sites, the hairy site at the
    ends and in the middle, image and classifier bonds unrelated."""
    import torch
    from orthogonal_classifiers_b200 import batched
    rng = np.random.default_rng(L * 100 + c)
    a = [min(2 ** n, 2 ** (L - n), amax) for n in range(L + 1)]
    B = [min(2 ** n, 2 ** (L - n), Bmax) for n in range(L + 1)]
    n_img = 37
    sites, flat = _random_mps(rng, a, n_img)
    clf = _random_mpo(rng, B, c, 16)
    offs = np.concatenate([[0], np.cumsum([2 * x * y for x, y in zip(a[:-1], a[1:])])]).tolist()
    tdt = torch.float32 if dtype == "f32" else torch.float64
    mps = batched.MPSBatch(torch.from_numpy(flat).to(tdt).cuda(), a, offs, L, None, dtype)
    ov, am = batched.overlaps_from_mps(mps, clf)
    ov = ov.double().cpu().numpy()
    ref = np.array([oracle.overlaps_chain(s, clf) for s in sites])
    assert ov.shape == ref.shape == (n_img, 16)
    # a random chain of L sites has entries of very different size: compare relative to the
    # largest overlap of the batch, as everywhere
    assert np.abs(ov - ref).max() <= TOL[dtype] * ref.max()
    srt = np.sort(ref, axis=1)
    tie = (srt[:, -1] - srt[:, -2]) < 4 * TOL[dtype] * ref.max()
    assert np.all((am.cpu().numpy() == ref.argmax(1)) | tie)


@pytest.mark.parametrize("in_dtype", [np.float32, np.float64, np.uint8])
def test_classify_host_c_abi(gpu_lib, in_dtype):
    """oc_classify_host: host buffers in and out through one C-ABI call; the pipelined
    chunks (pinned and pageable memory, ragged last chunk, row stride > n_pix) give
    exactly what the device-pointer entry points give."""
    import torch
    from orthogonal_classifiers_b200 import batched
    base = np.concatenate([synthetic_images(1000, seed=31), adversarial_images()])
    if in_dtype == np.uint8:
        imgs = np.rint(base * 255).astype(np.uint8)
    else:
        imgs = base.astype(in_dtype)
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=2))
    ref_ov, ref_am = gpu_lib.classify(imgs, clf, 32)
    try:
        for chunk in (0, 300, 5000):
            batched.set_option("host_chunk", chunk)
            ov, am = gpu_lib.classify_host(imgs, clf, 32)
            np.testing.assert_array_equal(ov.astype(np.float64), ref_ov)
            np.testing.assert_array_equal(am, ref_am)
        batched.set_option("host_chunk", 256)
        pin = torch.from_numpy(imgs).pin_memory()
        ovp = torch.empty((len(imgs), 16), dtype=torch.float32).pin_memory()
        amp = torch.empty((len(imgs),), dtype=torch.int32).pin_memory()
        for _ in range(2):
            ovp.zero_(); amp.zero_()
            gpu_lib.classify_host(pin, clf, 32, out=(ovp, amp))
            np.testing.assert_array_equal(ovp.numpy().astype(np.float64), ref_ov)
            np.testing.assert_array_equal(amp.numpy(), ref_am)
        # rows with a stride: the last row must not be read past its n_pix pixels
        wide = np.zeros((len(imgs), 800), dtype=in_dtype)
        wide[:, :784] = imgs
        ovw, amw = gpu_lib.classify_host(wide[:, :784], clf, 32)
        np.testing.assert_array_equal(ovw.astype(np.float64), ref_ov)
        # truncated encoders and fp64 through the same call
        for D, trunc in ((10, "sweep"), (10, "inline")):
            r_ov, r_am = gpu_lib.classify(imgs[:300], clf, D, trunc=trunc)
            ov, am = gpu_lib.classify_host(imgs[:300], clf, D, trunc=trunc)
            np.testing.assert_array_equal(ov.astype(np.float64), r_ov)
        # the library keeps its device copy and re-laid form of the classifier while the caller's is unchanged:
        # another classifier of the same shapes, then the first again (all four compute streams: 5 chunks)
        batched.set_option("host_chunk", 250)
        clf_b = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=77))
        rb_ov, _ = gpu_lib.classify(imgs, clf_b, 32)
        for c, r in ((clf_b, rb_ov), (clf, ref_ov), (clf_b, rb_ov)):
            ov, _ = gpu_lib.classify_host(imgs, c, 32)
            np.testing.assert_array_equal(ov.astype(np.float64), r)
        clf64 = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=2), dtype="f64")
        r_ov, _ = gpu_lib.classify(imgs[:200], clf64, 32, dtype="f64")
        ov, _ = gpu_lib.classify_host(imgs[:200], clf64, 32, dtype="f64")
        np.testing.assert_array_equal(ov, r_ov)
    finally:
        batched.set_option("host_chunk", 0)
    # argument errors come back as codes, not crashes
    from orthogonal_classifiers_b200 import _lib
    assert _lib.lib.oc_classify_host(imgs.ctypes.data, 7, 4, 784, 784, 10, 32, 0, 0, clf.packed_host.ctypes.data,
                                     clf.packed_host.size, clf._bonds_c, clf._s_c, ref_ov.ctypes.data, None, None) == 1
    assert _lib.lib.oc_classify_host(imgs.ctypes.data, 0, -3, 784, 784, 10, 32, 0, 0, clf.packed_host.ctypes.data,
                                     clf.packed_host.size, clf._bonds_c, clf._s_c, ref_ov.ctypes.data, None, None) == 1


def test_classify_host_default_schedule(gpu_lib):
    """The default schedule of oc_classify_host on a batch large enough for it - four compute streams, the
    ramp 4,000 / 8,000 / 14,000 / 22,000 for uint8 pixels, eight staging buffers reused (> 8 chunks of float32) -
    returns exactly what one device-resident call returns, for every stream count."""
    from orthogonal_classifiers_b200 import batched
    base = np.concatenate([synthetic_images(3000, seed=5), adversarial_images()])
    reps = 85000 // len(base) + 1
    f32 = np.tile(base, (reps, 1))[:85000].astype(np.float32)
    u8 = np.rint(f32 * 255).astype(np.uint8)
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=9))
    for imgs in (u8, f32):
        ref_ov, ref_am = gpu_lib.classify(imgs, clf, 32)
        try:
            for streams in (4, 2, 1):
                batched.set_option("host_streams", streams)
                ov, am = gpu_lib.classify_host(imgs, clf, 32)
                np.testing.assert_array_equal(ov.astype(np.float64), ref_ov)
                np.testing.assert_array_equal(am, ref_am)
        finally:
            batched.set_option("host_streams", 4)


def test_classifier_files_round_trip(gpu_lib, golden, tmp_path):
    """tools.py:262-291: save one ``site_{i}.npy`` per site, load them back (squeezed
    dimensions re-expanded), classify with the loaded classifier."""
    from conftest import golden_sites
    clf = golden_sites(golden, "clf_e32_b32_f32_ortho")
    q = gpu_lib.data_to_QTN(clf)
    gpu_lib.save_qtn_classifier(q, "round_trip", root=str(tmp_path))
    assert sorted(os.listdir(tmp_path / "round_trip")) == sorted(f"site_{i}.npy" for i in range(10))
    back = gpu_lib.load_qtn_classifier("round_trip", root=str(tmp_path))
    for a, b in zip(back.data, clf):
        np.testing.assert_array_equal(a, b)
    # the reference saves the SQUEEZED network (experiments.py / tools.py:262-266): edge bonds
    # and the size-1 label legs are gone; the loader puts them back
    sq = [np.squeeze(w) if n != 5 else w.reshape(2, 16, 32, 16) for n, w in enumerate(clf)]
    os.makedirs(tmp_path / "squeezed")
    for i, w in enumerate(sq):
        np.save(tmp_path / "squeezed" / f"site_{i}", w)
    back2 = gpu_lib.load_qtn_classifier("squeezed", root=str(tmp_path))
    assert [w.shape for w in back2.data] == [w.shape for w in clf]
    ov, _ = gpu_lib.classify(golden["x_test"], back2, 32)
    ref = golden["pred_e32_b32_f32_ortho"]
    assert np.abs(ov - ref).max() <= 1e-5 * ref.max()


def test_image_to_mps_trim_and_structure(gpu_lib):
    """tools.py:217-221 / test_tools.py:238-248: L = ceil(log2(784)) = 10, d = 2, D as asked;
    ``trim=True`` drops the zero trailing bond columns (xmps' minD) without changing the state."""
    imgs = np.concatenate([synthetic_images(3, seed=7), adversarial_images()[1:3]])
    for k, im in enumerate(imgs):
        for D in (32, 10, 2):
            full = gpu_lib.image_to_mps(im, D, dtype="f64")
            trim = gpu_lib.image_to_mps(im, D, dtype="f64", trim=True)
            assert full.L == trim.L == 10 and full.d == 2 and full.D == D
            a = oracle.mps_to_state(full.data)
            b = oracle.mps_to_state(trim.data)
            assert np.abs(a - b).max() < 1e-12
            assert all(t.shape[1] <= f.shape[1] and t.shape[2] <= f.shape[2] for t, f in zip(trim.data, full.data))
            ref, _ = oracle.encode_image(im, D, "sweep")
            # the oracle trims by numerical rank as xmps' minD does: same bond dimensions.  (Not for
            # the constant image: its exact rank deficiency leaves singular values of ~1e-16 sigma_max
            # in the fp64 LAPACK chain, on either side of numpy's rank tolerance.)
            if k != 4:
                assert [t.shape for t in trim.data] == [r.shape for r in ref]
    single = gpu_lib.image_to_mps(adversarial_images()[1], 32, dtype="f64", trim=True)
    assert max(max(t.shape[1:]) for t in single.data) == 1      # a one-pixel image is a product state


def test_streams_do_not_share_scratch(gpu_lib):
    """Library state is keyed by (device, stream): two streams working at the same time on
    different batches (different sizes, so their workspaces differ) give what each gives alone."""
    import torch
    from orthogonal_classifiers_b200 import batched
    a = torch.from_numpy(synthetic_images(20000, seed=51).astype(np.float32)).cuda()
    b = torch.from_numpy(synthetic_images(9000, seed=52).astype(np.float32)).cuda()
    clf = gpu_lib.DeviceClassifier(random_classifier(10, 32, centre=5, seed=6))
    ref_a = [t.clone() for t in batched.classify_device(a, clf, 32)]
    ref_b = [t.clone() for t in batched.classify_device(b, clf, 10)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            ra = batched.classify_device(a, clf, 32)
        with torch.cuda.stream(s2):
            rb = batched.classify_device(b, clf, 10)
        outs.append((ra, rb))
    torch.cuda.synchronize()
    for ra, rb in outs:
        assert torch.equal(ra[0], ref_a[0]) and torch.equal(ra[1], ref_a[1])
        assert torch.equal(rb[0], ref_b[0]) and torch.equal(rb[1], ref_b[1])
    from orthogonal_classifiers_b200._lib import check, lib
    check(lib.oc_release(int(s1.cuda_stream)))
    check(lib.oc_release(int(s2.cuda_stream)))


def test_classifier_cache_follows_content_changes(gpu_lib):
    """oc_classifier_cache: the re-laid classifier is reused across calls; a DeviceClassifier built
    at the same address with other contents, or other image bond profiles, must not see stale data."""
    imgs = synthetic_images(64, seed=53).astype(np.float32)
    for seed in (1, 2, 3):
        sites = random_classifier(10, 32, centre=5, seed=seed)
        clf = gpu_lib.DeviceClassifier(sites)
        W = oracle.dense_classifier(sites)
        ref = oracle.overlaps_dense(imgs.astype(np.float64), W)
        for _ in range(2):
            ov, _ = gpu_lib.classify(imgs, clf, 32)
            assert np.abs(ov - ref).max() <= 1e-5 * ref.max()
        b10 = gpu_lib.encode_images(imgs, 10)
        ov10 = gpu_lib.label_overlaps(b10, clf)
        ref10 = np.array([oracle.overlaps_chain([s.astype(np.float64) for s in b10.sites_numpy(i)], sites)
                          for i in range(len(imgs))])
        assert np.abs(ov10 - ref10).max() <= 1e-5 * ref10.max()
        del clf


def test_stacking_matches_reference_fixture(gpu_lib):
    """experiments.py:528-640 (BASELINE config 4): label 16-vectors -> L = 4 MPS
    (``mps_encoding(v, None)``) -> product of n + 1 copies -> overlap with the stacking
    unitary, L = 4 (n + 1) <= 16, label leg on the centre site.  Fixture: the reference's own
    ``mps_stacking`` (tests/golden/generate_golden_stacking.py) — its stacking unitary, stacked
    predictions and accuracy for n = 0..3."""
    from orthogonal_classifiers_b200 import experiments as ex
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_stacking.npz"))
    for dtype in ("f32", "f64"):
        mps = gpu_lib.mps_encoding(g["v_test"], None, dtype=dtype)
        assert mps.batch.L == 4 and mps.batch.bonds == [1, 2, 4, 2, 1]
        for n in range(4):
            L = 4 * (n + 1)
            unitary = [g[f"unitary_{n}_site_{i}"] for i in range(L)]
            assert int(np.argmax([w.shape[1] for w in unitary])) == L // 2
            copies = ex.generate_copy_state(mps, n)
            assert copies.L == L and len(copies) == len(g["v_test"])
            # the copy state is the n + 1-fold tensor power of the normalised vector
            st = oracle.mps_to_state([s.astype(np.float64) for s in copies.sites_numpy(3)])
            v = g["v_test"][3] / np.linalg.norm(g["v_test"][3])
            ref_st = v
            for _ in range(n):
                ref_st = np.kron(ref_st, v)
            assert min(np.abs(st - ref_st).max(), np.abs(st + ref_st).max()) < 10 * TOL[dtype]
            pred = ex.stacked_predictions(mps, unitary, n)
            ref = g[f"pred_{n}"]
            assert pred.shape == ref.shape
            assert np.abs(pred - ref).max() <= TOL[dtype] * ref.max(), (dtype, n)
            assert ex.evaluate_stacking(mps, unitary, n, g["y_test"]) == float(g[f"acc_{n}"])


def test_prediction_arrays_and_files(gpu_lib, golden, tmp_path):
    """experiments.py:332-352: the (n_D, N, 16) arrays and the files the plotting scripts read."""
    from conftest import golden_sites
    from orthogonal_classifiers_b200 import experiments as ex
    names = ["e32_b32_f32_nonortho", "e32_b32_f32_ortho", "e32_b8_f8_ortho"]
    clfs = [golden_sites(golden, f"clf_{n}") for n in names]
    mps = gpu_lib.mps_encoding(golden["x_test"], 32)
    arr = ex.prediction_array(mps, clfs)
    assert arr.shape == (3, 16, 16) and arr.dtype == np.float64
    for k, n in enumerate(names):
        assert np.abs(arr[k] - golden[f"pred_{n}"]).max() <= 1e-5 * golden[f"pred_{n}"].max()
    np.testing.assert_array_equal(ex.prediction_array(golden["x_test"], clfs, 32), arr)
    paths = ex.save_prediction_arrays(str(tmp_path), ortho=arr[1:], non_ortho=arr[:1], kind="test")
    assert sorted(os.path.basename(p) for p in paths) == ["non_ortho_d_final_vs_test_predictions.npy",
                                                          "ortho_d_final_vs_test_predictions.npy"]
    np.testing.assert_array_equal(np.load(tmp_path / "ortho_d_final_vs_test_predictions.npy"), arr[1:])
    paths = ex.save_prediction_arrays(str(tmp_path), ortho=arr, kind="training")
    assert os.path.basename(paths[0]) == "ortho_d_final_vs_training_predictions_compressed.npz"
    np.testing.assert_array_equal(np.load(paths[0])["arr_0"], arr)      # experiments.py:657 reads arr_0
    np.testing.assert_array_equal(ex.load_prediction_array(paths[0]), arr)


def test_ensemble_votes_on_device(gpu_lib):
    """oc_ensemble_vote / oc_label_mix against the reference's host arithmetic
    (variational_mpo_classifiers.py:321-363 restated in numpy), ties included."""
    import torch
    from collections import Counter
    from orthogonal_classifiers_b200 import batched
    rng = np.random.default_rng(3)
    K, N, S, nl = 5, 3001, 16, 10
    pred = rng.random((K, N, S))
    y = rng.integers(0, nl, N)
    pred[:, np.arange(N), y] += 0.3
    # exact vote ties: every member of the first 200 images votes for a different label
    for i in range(200):
        for k in range(K):
            pred[k, i, :nl] = 0.01
            pred[k, i, (i + 2 * k) % nl] = 1.0
    for dtype, tdt in (("f64", torch.float64), ("f32", torch.float32)):
        p = pred.astype(np.float64 if dtype == "f64" else np.float32)
        norm, soft, sa, ha = batched.ensemble_vote(torch.from_numpy(p).cuda(), nl)
        p64 = p.astype(np.float64)
        ref_norm = p64[:, :, :nl] / p64[:, :, :nl].sum(2, keepdims=True)
        tol = 1e-12 if dtype == "f64" else 1e-6
        assert np.abs(norm.double().cpu().numpy() - ref_norm).max() < tol
        ref_soft = ref_norm.sum(0)
        assert np.abs(soft.double().cpu().numpy() - ref_soft).max() < K * tol
        srt = np.sort(ref_soft, axis=1)
        clear = (srt[:, -1] - srt[:, -2]) > 1e-5
        assert np.all((sa.cpu().numpy() == ref_soft.argmax(1)) | ~clear)
        votes = p64[:, :, :nl].argmax(2).T                              # (N, K)
        ref_hard = np.array([Counter(v.tolist()).most_common(1)[0][0] for v in votes])
        np.testing.assert_array_equal(ha.cpu().numpy(), ref_hard)
    # label mixing: random product bitstrings, two paddings
    sg = rng.standard_normal((N, S))
    V = rng.standard_normal((S, 2 * nl))
    out = batched.label_mix(torch.from_numpy(sg).cuda(), V, nl, 2).cpu().numpy()
    ref = np.abs(sg @ V[:, :nl]) + np.abs(sg @ V[:, nl:])
    assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()


def _left_cut_singular_values(mpos, D, c=5):
    """Singular values the left sweep of compress_centre_one_site (fMPO.py:400-421) sees at each
    site of the sum of ``mpos`` (oracle restatement), to tell clean cuts from degenerate ones."""
    data = oracle.mpo_oracle._add_all(mpos)
    out = []
    for m in range(c):
        d, s, i, j = data[m].shape
        U, S, Vt = np.linalg.svd(np.asarray(data[m]).real.transpose(0, 2, 1, 3).reshape(d * i, s * j), full_matrices=False)
        r = min(D, int(np.sum(S > S[0] * 1e-13)))
        out.append(S)
        SV = S[:r, None] * Vt[:r]
        data[m + 1] = np.einsum("rk,dskj->dsrj", SV, np.asarray(data[m + 1]).real)   # s = 1 left of the centre
    return out


@pytest.mark.parametrize("tag", ["e32_b32_f32", "e10_b10_f32", "e32_b8_f8"])
@pytest.mark.parametrize("ortho", [False, True])
def test_classifier_construction_on_device(gpu_lib, golden, tag, ortho):
    """fMPO.add + compress_centre_one_site (+ polar) on the device (fMPO.py:338-453, 666-689;
    experiments.py:482-523) against the classifiers the reference's own code built
    (tests/golden/generate_golden.py): the dense operator of the non-orthogonalised ones up to a
    global sign, the overlaps on the training span for the orthogonalised ones (the polar factor
    of the rank-deficient centre matrix is unique only there, in the reference too).

    ``e32_b8_f8`` cannot be compared value by value: its D_batch = 8 cut falls INSIDE a block of
    exactly equal singular values (the summands are isometries, so the summed site matrix has
    singular values sqrt(5), 2, sqrt(3), ... with high multiplicity), where the subspace an SVD
    keeps is whatever LAPACK returns — in the reference too.  The test shows the tie and checks
    what is defined: shapes, isometric sites, unit norm / partial isometry."""
    from conftest import golden_sites
    D_encode, D_batch, D_final = [int(t[1:]) for t in tag.split("_")]
    x, y = golden["x_train"], golden["y_train"]
    mps = gpu_lib.mps_encoding(x, D_encode, dtype="f64")
    sums = gpu_lib.prepare_centred_batched_classifier(mps, y, None, D_batch, [8])
    assert len(sums) == 10
    clf = gpu_lib.adding_centre_batches(sums, D_final, 10, orthogonalise=ortho)[0]
    ref = golden_sites(golden, f"clf_{tag}_{'ortho' if ortho else 'nonortho'}")
    assert [w.shape for w in clf.data] == [w.shape for w in ref]
    Wg, Wr = oracle.dense_classifier(clf.data), oracle.dense_classifier(ref)
    for n, w in enumerate(clf.data):          # canonical around the centre
        if n < 5:
            M = w[:, 0].reshape(-1, w.shape[3])
            G = M.T @ M
        elif n > 5:
            M = w[:, 0].transpose(1, 0, 2).reshape(w.shape[2], -1)
            G = M @ M.T
        else:
            continue
        nz = np.diag(G) > 0.5
        assert np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max() < 1e-10 and np.abs(G[~nz]).max(initial=0) < 1e-10
    c = clf.data[5]
    Mc = c.transpose(0, 2, 1, 3).reshape(c.shape[0] * c.shape[2], -1)
    if ortho:
        G = Mc @ Mc.T                          # partial isometry: M M^T is a projector
        assert np.abs(G @ G - G).max() < 1e-8
    else:
        assert abs(np.linalg.norm(Wg) - 1) < 1e-10 and abs(np.linalg.norm(Mc) - 1) < 1e-10
    if tag == "e32_b8_f8":
        bits = oracle.hairy_bitstrings(10, 10, centre=5)
        enc = [oracle.class_encode(oracle.encode_image(im, 32)[0], int(l), bits) for im, l in zip(x[:8], y[:8])]
        S4 = _left_cut_singular_values(enc, 8)[3]      # bond 4: 16 -> 8
        assert abs(S4[7] - S4[8]) < 1e-12 * S4[0]      # the cut is a tie: no unique answer
        return
    scale = np.abs(Wr).max()
    if not ortho:
        assert min(np.abs(Wg - Wr).max(), np.abs(Wg + Wr).max()) <= 1e-5 * scale
    a, b = oracle.overlaps_dense(x, Wg), oracle.overlaps_dense(x, Wr)
    assert np.abs(a - b).max() <= 1e-5 * b.max()
    # and the classifier classifies: same predictions on the test images as the reference's
    ov, _ = gpu_lib.classify(golden["x_test"], clf.data, D_encode, dtype="f64")
    pred = golden[f"pred_{tag}_{'ortho' if ortho else 'nonortho'}"]
    assert np.abs(ov - pred).max() <= 1e-5 * pred.max()


def test_add_compress_groups_and_remainder(gpu_lib):
    """adding_centre_batches semantics (experiments.py:497-509): consecutive groups, a shorter
    remainder group at the end; each group equals the oracle's add + compress of its members —
    without truncation (D = 32) and with a truncating D on D_encode = 10 images, whose summed site
    matrices have no degenerate singular values at the cut."""
    n = 23
    y = np.arange(n) % 10
    x = synthetic_images(n, seed=61, labels=y)
    bits = oracle.hairy_bitstrings(10, 10, centre=5)
    for D_enc, D in ((32, 32), (10, 8)):
        mps = gpu_lib.mps_encoding(x, D_enc, dtype="f64")
        mpos = gpu_lib.mpo_encoding(mps, y)
        out = gpu_lib.adding_centre_batches(mpos, D, 5, orthogonalise=False)
        assert len(out) == 5                                   # 5 + 5 + 5 + 5 + 3
        enc = [oracle.class_encode(oracle.encode_image(im, D_enc, "sweep")[0], int(l), bits) for im, l in zip(x, y)]
        ref = oracle.adding_centre_batches(enc, D, 5, False)
        for g in range(5):
            if D < D_enc:
                for S in _left_cut_singular_values(enc[5 * g:5 * g + 5], D):
                    assert len(S) <= D or S[D - 1] - S[D] > 1e-6 * S[0]   # clean cuts only
            Wg = oracle.dense_classifier(out[g].data)
            Wr = oracle.dense_classifier(oracle.real_sites(ref[g]))
            assert min(np.abs(Wg - Wr).max(), np.abs(Wg + Wr).max()) <= 1e-8 * np.abs(Wr).max(), (D_enc, D, g)


def test_mps_stacking_end_to_end(gpu_lib):
    """experiments.py:528-643 with the stacking unitary BUILT on the device (class encoding, batch
    adding at D_batch = 32, polar factor) from the fixture's training vectors: same accuracy as the
    reference's run for n = 0..3 copies; for n <= 1 (no bond is truncated) the overlaps on the
    training vectors — where the polar factor is unique — equal those of the reference's unitary."""
    from orthogonal_classifiers_b200 import experiments as ex
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_stacking.npz"))
    train = gpu_lib.mps_encoding(g["v_train"], None, dtype="f64")
    test = gpu_lib.mps_encoding(g["v_test"], None, dtype="f64")
    for n in range(4):
        acc, preds, unitary = ex.mps_stacking(train, test, n, g["y_train"], g["y_test"], return_details=True)
        L = 4 * (n + 1)
        ref_unitary = [g[f"unitary_{n}_site_{i}"] for i in range(L)]
        assert [w.shape for w in unitary.data] == [w.shape for w in ref_unitary]
        assert acc == float(g[f"acc_{n}"])
        if n <= 1:
            mine = ex.stacked_predictions(train, unitary.data, n)
            ref = ex.stacked_predictions(train, ref_unitary, n)
            assert np.abs(mine - ref).max() <= 1e-8 * ref.max(), n
