"""Two fused encode + classify passes at D_encode < 32, OC_TRUNC_SWEEP (for an ncu launch list)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
N = 70000
D = int(sys.argv[1]) if len(sys.argv) > 1 else 16
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
for _ in range(2):
    batched.classify_device(imgs, clf, D, trunc="sweep")
torch.cuda.synchronize()
print("ok")
