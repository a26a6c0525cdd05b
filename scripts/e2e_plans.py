"""oc_classify_host with explicit chunk schedules (OC_HOST_SIZES), uint8 host pixels, 70,000 images."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier
N = 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = np.tile(imgs, ((N + 8191) // 8192, 1))[:N]
u8 = torch.from_numpy(np.rint(imgs * 255).astype(np.uint8)).pin_memory()
f32 = torch.from_numpy(imgs).pin_memory()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
ov = torch.empty((N, 16), dtype=torch.float32).pin_memory()
am = torch.empty((N,), dtype=torch.int32).pin_memory()
def timeit(x, reps=10):
    for _ in range(3): batched.classify_host(x, clf, 32, out=(ov, am))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): batched.classify_host(x, clf, 32, out=(ov, am))
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
plans = ["17500,17500,17500,17500", "17500,17500,35000", "14000,14000,42000", "11667,11667,46666", "8750,8750,17500,35000",
         "8750,17500,43750", "10000,20000,40000", "12000,12000,23000,23000", "17500,52500", "10000,15000,45000", "14000,21000,35000",
         "8750,8750,8750,8750,35000", "20000,20000,30000", "23334,23333,23333"]
for name, x in (("u8", u8), ("f32", f32)):
    for p in plans:
        os.environ["OC_HOST_SIZES"] = p
        print(f"{name} {p:28s}: {timeit(x):.3f} ms", flush=True)
