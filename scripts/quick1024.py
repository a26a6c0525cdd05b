"""Device timing for full 32 x 32 (1024-pixel) images, the shape of BASELINE.json config 5."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import random_classifier
N = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
rng = np.random.default_rng(3)
yy, xx = np.mgrid[0:32, 0:32]
base = np.stack([np.clip(sum(np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * w ** 2))
                             for cx, cy, w in zip(rng.uniform(4, 28, 3), rng.uniform(4, 28, 3), rng.uniform(1, 4, 3))), 0, 1).reshape(-1)
                 for _ in range(4096)]).astype(np.float32)
base[base < 0.1] = 0
imgs = torch.from_numpy(np.tile(base, ((N + 4095) // 4096, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))
def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
batch = oc.encode_images(imgs, 32)
t_enc = timeit(lambda: oc.encode_images(imgs, 32))
t_ov = timeit(lambda: batched.overlaps_from_mps(batch, clf))
print(f"1024 px N={N} encode {t_enc:.3f} ms overlap {t_ov:.3f} ms -> {N/(t_enc+t_ov)*1e3/1e6:.2f} M img/s")
b2 = oc.encode_images(imgs[:4096], 32, return_sweeps=True)
print("mean sweeps per step:", b2.sweeps.float().mean(0).cpu().numpy().round(2))
