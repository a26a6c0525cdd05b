"""oc_classify_host with explicit chunk schedules (OC_HOST_SIZES) on 2 / 4 compute streams, uint8 host pixels."""
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
plans = ["4000,8000,14000,22000,22000", "3000,6000,12000,24500,24500", "4000,8000,16000,21000,21000", "5000,10000,15000,20000,20000",
         "4000,8000,12000,16000,30000", "2000,4000,8000,16000,20000,20000", "4000,8000,14000,44000", "6000,12000,26000,26000",
         "4000,6000,10000,14000,18000,18000", "3000,5000,8000,12000,21000,21000", "4000,8000,14000,22000,22000"]
for streams in (4, 4):
    batched.set_option("host_streams", streams)
    for p in plans:
        os.environ["OC_HOST_SIZES"] = p
        print(f"streams {streams} {p:36s}: {timeit(u8):.3f} ms", flush=True)
