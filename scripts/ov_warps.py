"""How does the overlap kernel scale with the warps resident on an SM?  (FFMA static mode, `ov_mma` = 0: the
tensor-core mode's CTA-wide label product is written for exactly 8 warps.)"""
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
batch = oc.encode_images(imgs, 32)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
batched.set_option("ov_mma", 0)
for w in (8, 7, 6, 5, 4, 3, 2):
    batched.set_option("ov_warps", w)
    print(f"FFMA static mode, {w} warps per SM: {timeit(lambda: batched.overlaps_from_mps(batch, clf))*1e3:.1f} us")
batched.set_option("ov_warps", 8); batched.set_option("ov_mma", 1)
print(f"tensor-core mode, 8 warps: {timeit(lambda: batched.overlaps_from_mps(batch, clf))*1e3:.1f} us")
