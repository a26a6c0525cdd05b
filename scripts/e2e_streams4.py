"""oc_classify_host with 2 / 4 compute streams and chunk sizes (uint8 and float32 host pixels, 70,000 images)."""
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
batched.set_option("host_streams", 1); batched.set_option("host_chunk", 0)
batched.classify_host(u8, clf, 32, out=(ov, am)); ref = ov.clone()
batched.set_option("host_ramp", 0)
for name, x in (("u8", u8), ("f32", f32)):
    for streams in (2, 4):
        for chunk in (5000, 7000, 8750, 10000, 11667, 14000, 17500):
            batched.set_option("host_streams", streams); batched.set_option("host_chunk", chunk)
            t = timeit(x)
            same = bool((ov == ref).all())
            print(f"{name} streams {streams} chunk {chunk:6d}: {t:.3f} ms = {N/t/1e3:.2f} M img/s  identical {same}", flush=True)
