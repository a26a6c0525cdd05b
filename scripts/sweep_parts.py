"""Where the time of OC_TRUNC_SWEEP goes: inline chain on the images, inline chain on bit-reversed
(full 1024-pixel) images, the whole sweep encode."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200.synthetic import synthetic_images

N = 70000
base = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
pad = np.zeros((8192, 1024), np.float32); pad[:, :784] = base
idx = np.arange(1024); rev = np.zeros_like(idx)
for b in range(10): rev |= ((idx >> b) & 1) << (9 - b)
imgs = torch.from_numpy(np.tile(base, ((N + 8191) // 8192, 1))[:N]).cuda()
rimgs = torch.from_numpy(np.tile(pad[:, rev], ((N + 8191) // 8192, 1))[:N]).cuda()

def timeit(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for D in (2, 4, 8, 10, 16, 20):
    a = timeit(lambda: oc.encode_images(imgs, D, trunc="inline"))
    b = timeit(lambda: oc.encode_images(rimgs, D, trunc="inline"))
    c = timeit(lambda: oc.encode_images(imgs, D, trunc="sweep"))
    print(f"D={D}: inline {a:.3f} | inline on bit-reversed {b:.3f} | sweep {c:.3f} -> bitrev + QR pass {c-b:.3f} ms", flush=True)
