"""Overlap error of the shipped path on several image families (vs the dense fp64 oracle)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
import oracle
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier, adversarial_images
rng = np.random.default_rng(5)
N = 256
yy, xx = np.mgrid[0:28, 0:28]
fam = {
    "strokes": synthetic_images(N, seed=41),
    "white noise": rng.random((N, 784)),
    "sparse pixels": (rng.random((N, 784)) < 0.01) * rng.random((N, 784)),
    "ramps": np.stack([((a * xx + b * yy) % 28 / 27.0).reshape(-1) for a, b in rng.integers(0, 5, (N, 2))]),
    "checker": np.stack([(((xx // p) + (yy // q)) % 2).reshape(-1).astype(float) for p, q in rng.integers(1, 8, (N, 2))]),
    "blocks": np.stack([((xx > a) & (xx < a + 9) & (yy > b) & (yy < b + 6)).reshape(-1).astype(float) for a, b in rng.integers(0, 18, (N, 2))]),
    "const+noise1e-4": 0.7 + 1e-4 * rng.standard_normal((N, 784)),
    "huge range": rng.random((N, 784)) ** 12,
    "adversarial": adversarial_images(),
}
clf = random_classifier(10, 32, centre=5, seed=9)
W = oracle.dense_classifier(clf)
for name, imgs in fam.items():
    x32 = imgs.astype(np.float32)
    ref = oracle.overlaps_dense(x32.astype(np.float64), W)
    ov, am = oc.classify(x32, clf, 32)
    b = oc.encode_images(x32, 32, return_sweeps=True)
    iso = 0.0
    for i in range(0, len(x32), max(1, len(x32) // 16)):
        for A in b.sites_numpy(i):
            M = A.reshape(-1, A.shape[2]).astype(np.float64); G = M.T @ M; nz = np.diag(G) > 0.5
            if nz.any(): iso = max(iso, np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max())
    srt = np.sort(ref, axis=1); tie = (srt[:, -1] - srt[:, -2]) < 1e-5 * max(ref.max(), 1e-30)
    lab = np.mean((am == ref.argmax(1)) | tie)
    print(f"{name:18s} overlap err/max {np.abs(ov - ref).max() / max(ref.max(), 1e-30):.2e}  iso {iso:.1e}  labels ok {lab:.3f}  max sweeps {b.sweeps.max(0).values.cpu().numpy().tolist()}")
