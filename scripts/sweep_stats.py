"""Per-image Jacobi sweep counts of every encoder step, and what a warp (max over its images) pays."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200.synthetic import synthetic_images
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
b = oc.encode_images(torch.from_numpy(imgs).cuda(), 32, return_sweeps=True, return_svals=True)
sw = b.sweeps.cpu().numpy(); sv = b.svals.cpu().numpy()
k = (sv[:, 3] > 0).sum(1)
rng = np.random.default_rng(0)
for n in range(10):
    s = sw[:, n]
    print("step", n, "mean %.2f" % s.mean(), "hist", np.bincount(s, minlength=10)[:12])
for n, ipw in ((3, 2), (4, 3), (5, 4), (6, 8)):
    s = sw[:, n]
    for kk in (None,) if n != 4 else (9, 10, 11, 12):
        sel = s if kk is None else s[k == kk]
        m = len(sel) // ipw * ipw
        p = rng.permutation(sel)[:m].reshape(-1, ipw)
        srt = np.sort(sel)[:m].reshape(-1, ipw)
        print("step", n, "k", kk, "ipw", ipw, "mean %.3f  E[max random] %.3f  E[max sorted by sweeps] %.3f" % (sel.mean(), p.max(1).mean(), srt.max(1).mean()))
# correlation of step-4 sweeps with simple predictors
s4 = sw[:, 4].astype(float)
r3 = sv[:, 3]; cond = np.log10(r3[:, 0] / np.maximum(r3[np.arange(len(k)), k - 1], 1e-30))
print("corr(step-4 sweeps, k) %.3f  corr(step-4 sweeps, log cond bond 4) %.3f  corr(step-4 sweeps, step-3 sweeps) %.3f" % (
    np.corrcoef(s4, k)[0, 1], np.corrcoef(s4, cond)[0, 1], np.corrcoef(s4, sw[:, 3])[0, 1]))
