"""oc_classify_host per-chunk timeline with the default schedule (uint8 host pixels, 70,000 images)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
N = 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = np.tile(imgs, ((N + 8191) // 8192, 1))[:N]
u8 = torch.from_numpy(np.rint(imgs * 255).astype(np.uint8)).pin_memory()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
ov = torch.empty((N, 16), dtype=torch.float32).pin_memory()
am = torch.empty((N,), dtype=torch.int32).pin_memory()
for _ in range(3): batched.classify_host(u8, clf, 32, out=(ov, am))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10): batched.classify_host(u8, clf, 32, out=(ov, am))
b.record(); torch.cuda.synchronize()
print(f"{a.elapsed_time(b)/10:.3f} ms per call", file=sys.stderr)
batched.set_option("host_trace", 1); batched.classify_host(u8, clf, 32, out=(ov, am)); batched.set_option("host_trace", 0)
