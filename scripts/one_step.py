"""One warm-up and one measured device-resident encode + classify step of the bench workload (70,000 images,
D_encode = D_total = 32), then the same overlap with `ov_tc5` — the command behind the per-kernel ncu captures."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier

N = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
for _ in range(2):
    batch = oc.encode_images(imgs, 32)
    batched.overlaps_from_mps(batch, clf)
    batched.set_option("ov_tc5", 1)
    batched.overlaps_from_mps(batch, clf)
    batched.set_option("ov_tc5", 0)
torch.cuda.synchronize()
print("ok")
