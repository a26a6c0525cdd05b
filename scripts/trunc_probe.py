"""Truncated encodes (D < 32, oc_encode_fast path) against the inline fp64 oracle: state error and singular values."""
import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
import oracle
from orthogonal_classifiers_b200.synthetic import synthetic_images
base = synthetic_images(24, seed=43).astype(np.float64)
for name, imgs in (("strokes", base), ("low contrast 1e-2", 0.5 + 1e-2 * base), ("low contrast 1e-3", 0.5 + 1e-3 * base)):
    x32 = imgs.astype(np.float32)
    for D in (2, 4, 8, 10, 16, 20):
        b = oc.encode_images(x32, D, return_svals=True)
        sv = b.svals.cpu().numpy().astype(np.float64)
        worst = 0.0; wsv = 0.0
        for i in range(len(x32)):
            ref_sites, ref_sv = oracle.encode_image(x32[i].astype(np.float64), D, "inline")
            st = oracle.mps_to_state([s.astype(np.float64) for s in b.sites_numpy(i)])
            rs = oracle.mps_to_state(ref_sites)
            worst = max(worst, min(np.abs(st - rs).max(), np.abs(st + rs).max()))
            for n in range(10):
                m = min(sv.shape[2], len(ref_sv[n]))
                wsv = max(wsv, np.abs(sv[i, n, :m] - ref_sv[n][:m]).max() / ref_sv[0][0])
        print(f"{name:18s} D={D:2d}  state err {worst:.2e}  sv err {wsv:.2e}")
