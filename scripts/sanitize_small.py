"""Small encode + classify over every kernel family, for compute-sanitizer."""
import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200.synthetic import adversarial_images, random_classifier, synthetic_images

imgs = np.concatenate([synthetic_images(90, seed=1), adversarial_images()]).astype(np.float32)
for D, Dc, c in ((32, 32, 5), (10, 32, 5), (3, 8, 5), (32, 20, 9)):
    clf = random_classifier(10, Dc, centre=c, seed=D)
    ov, am = oc.classify(imgs, clf, D)
    b = oc.encode_images(imgs, D, return_svals=True, return_sweeps=True)
    print(D, Dc, c, float(ov.max()), int(am.sum()), float(b.svals.max()))
oc.set_force_generic(True)
ov, am = oc.classify(imgs[:8], random_classifier(10, 32, centre=5, seed=0), 32)
ov64, _ = oc.classify(imgs[:8].astype(np.float64), random_classifier(10, 32, centre=5, seed=0), 32, dtype="f64")
print("generic", float(np.abs(ov - ov64).max()))
v = np.abs(np.random.default_rng(0).standard_normal((5, 16)))
print("L=4", oc.encode_images(v, None).data.shape)
