"""Quick device timing of oc_encode_batch / oc_overlap_batch (70k images, D=32)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier

N = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 32
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=int(sys.argv[3]) if len(sys.argv) > 3 else 5, seed=1))

def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

batch = oc.encode_images(imgs, D)
t_enc = timeit(lambda: oc.encode_images(imgs, D))
t_ov = timeit(lambda: batched.overlaps_from_mps(batch, clf))
print(f"N={N} D={D} encode {t_enc:.3f} ms  overlap {t_ov:.3f} ms  total {t_enc+t_ov:.3f} ms  -> {N/(t_enc+t_ov)*1e3/1e6:.2f} M img/s")
b2 = oc.encode_images(imgs[:8192], D, return_sweeps=True)
print("mean sweeps per step:", b2.sweeps.float().mean(0).cpu().numpy().round(2))
b3 = oc.encode_images(imgs[:2048], D, return_svals=True)
sv = b3.svals.cpu().numpy()
nzc = (sv > 0).sum(2)
print("nonzero svals per step: mean", nzc.mean(0).round(2), "max", nzc.max(0))
rel = sv / sv[:, :, :1].clip(1e-30)
print("svals in (0,1e-5] rel: per step", ((rel > 0) & (rel < 1e-5)).sum(2).mean(0).round(2))
