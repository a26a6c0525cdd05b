"""Overlap error on low-contrast images (constant offset + small strokes) for the encoder variants."""
import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
import oracle
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
imgs = synthetic_images(200, seed=37).astype(np.float64)
clf = random_classifier(10, 32, centre=5, seed=8)
W = oracle.dense_classifier(clf)
for eps in (1e-1, 1e-2, 1e-3):
    low = (0.5 + eps * imgs).astype(np.float32)
    ref = oracle.overlaps_dense(low.astype(np.float64), W)
    for opts in ({}, {"enc_gram": 0}, {"enc_mgs4": 0}, {"enc_qr5": 0}, {"ov_mma": 0}, {"enc_gram": 0, "enc_mgs4": 0, "enc_qr5": 0, "ov_mma": 0}, {"enc_hw": 0}, {"force_generic": 1}):
        for k, v in opts.items(): batched.set_option(k, v)
        ov, _ = oc.classify(low, clf, 32)
        for k in opts: batched.set_option(k, 0 if k == "force_generic" else 1)
        b = oc.encode_images(low[:8], 32, return_svals=True)
        print(f"eps {eps:g} {str(opts):70s} err/max {np.abs(ov - ref).max() / ref.max():.2e}")
    sv = b.svals.cpu().numpy()[0]
    print("  svals image 0, steps 3..5 (rel):", [np.round(sv[n][:6] / sv[n][0], 7).tolist() for n in (3, 4, 5)])
