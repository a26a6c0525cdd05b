"""Device-resident encode + classify: float32 vs uint8 pixels, several batch sizes."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
base = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for N in (8750, 17500, 35000, 52500, 70000):
    f = torch.from_numpy(np.tile(base, ((N + 8191) // 8192, 1))[:N]).cuda()
    u = (f * 255).round().to(torch.uint8)
    ws = torch.empty((N, 2728), device="cuda")
    tf = timeit(lambda: batched.classify_device(f, clf, 32, workspace=ws))
    tu = timeit(lambda: batched.classify_device(u, clf, 32, workspace=ws))
    print(f"N={N}: f32 {tf:.3f} ms  u8 {tu:.3f} ms", flush=True)
