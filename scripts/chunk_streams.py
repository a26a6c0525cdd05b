"""Device-resident encode + classify of 70,000 images as chunks on 1 / 2 / 3 streams: do concurrent chunks fill
each other's kernel tails, and does a chunk whose MPS fits the L2 (8,750 x 10.9 KB = 95 MB) cost anything?"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier

N = 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
ov = torch.empty((N, 16), dtype=torch.float32, device="cuda")
am = torch.empty((N,), dtype=torch.int32, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]

def run(chunk, ns):
    cur = torch.cuda.current_stream()
    if ns == 0:
        batched.classify_device(imgs, clf, 32, out=(ov, am))
        return
    ev0 = torch.cuda.Event(); ev0.record(cur)
    k = 0
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        s = streams[k % ns]; k += 1
        s.wait_event(ev0)
        with torch.cuda.stream(s):
            batched.classify_device(imgs[lo:hi], clf, 32, out=(ov[lo:hi], am[lo:hi]))
    for s in streams[:ns]:
        e = torch.cuda.Event(); e.record(s); cur.wait_event(e)

def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

print(f"whole batch, one stream: {timeit(lambda: run(N, 0)):.3f} ms")
ref = ov.clone()
for chunk in (35000, 17500, 8750):
    for ns in (1, 2, 3, 4):
        t = timeit(lambda: run(chunk, ns))
        torch.cuda.synchronize()
        same = bool((ov == ref).all())
        print(f"chunk {chunk:6d} on {ns} stream(s): {t:.3f} ms  (bit-identical: {same})")
