"""Worst-case accuracy of the encoder over a batch: reconstructed amplitudes vs x/|x| (D = 32),
isometry of every site, and overlaps vs the dense oracle."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
import oracle
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
imgs = synthetic_images(N, seed=5, labels=np.arange(N) % 10, dtype=np.float32)
b = oc.encode_images(torch.from_numpy(imgs).cuda(), 32, return_sweeps=True)
worst = 0.0; iso = 0.0
for i in range(0, N, max(1, N // 256)):
    sites = [s.astype(np.float64) for s in b.sites_numpy(i)]
    st = oracle.mps_to_state(sites)
    x = np.zeros(1024); x[:784] = imgs[i]; x /= np.linalg.norm(x)
    worst = max(worst, min(np.abs(st - x).max(), np.abs(st + x).max()))
    for A in sites:
        M = A.reshape(-1, A.shape[2]); G = M.T @ M; nz = np.diag(G) > 0.5
        iso = max(iso, np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max())
clf = random_classifier(10, 32, centre=5, seed=1)
ov, am = oc.classify(imgs, clf, 32)
ref = oracle.overlaps_dense(imgs.astype(np.float64), oracle.dense_classifier(clf))
print(f"max amplitude err {worst:.2e}  max isometry err {iso:.2e}  max overlap err/max {np.abs(ov-ref).max()/ref.max():.2e}  "
      f"mean sweeps {b.sweeps.float().mean(0).cpu().numpy().round(2)}")
