"""Per-chunk timeline of oc_classify_host (option host_trace)."""
import sys, time
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
for streams, chunk, ramp in ((1, 35000, 1), (2, 17500, 0), (2, 35000, 0)):
    batched.set_option("host_streams", streams); batched.set_option("host_chunk", chunk); batched.set_option("host_ramp", ramp)
    for _ in range(3): batched.classify_host(u8, clf, 32, out=(ov, am))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): batched.classify_host(u8, clf, 32, out=(ov, am))
    torch.cuda.synchronize()
    print(f"streams {streams} chunk {chunk} ramp {ramp}: wall {1e3*(time.perf_counter()-t0)/5:.3f} ms per call", file=sys.stderr)
    batched.set_option("host_trace", 1)
    batched.classify_host(u8, clf, 32, out=(ov, am))
    batched.set_option("host_trace", 0)
