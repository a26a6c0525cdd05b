"""Device timing of the two truncation orders (70k images) for D_encode < 32."""
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

def timeit(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for D in (2, 4, 8, 10, 16, 20, 32):
    row = []
    for trunc in ("inline", "sweep"):
        t = timeit(lambda: batched.classify_device(imgs, clf, D, trunc=trunc))
        row.append(f"{trunc} {t:.3f} ms")
    print(f"N={N} D_encode={D}: " + " | ".join(row), flush=True)
