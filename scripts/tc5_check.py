"""MODE 6 of the overlap kernel (site 4 on tcgen05, `ov_tc5`) against MODE 2 (mma.sync): same overlaps, timing."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images, random_classifier

N = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1))

def timeit(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

batch = oc.encode_images(imgs, 32)
res = {}
for n_try in (N, 8 * 148 + 3, 5, 1):   # full CTA iterations, a ragged last iteration, fewer images than one tile
    sub = oc.encode_images(imgs[:n_try], 32)
    out = {}
    for opt in (0, 1):
        batched.set_option("ov_tc5", opt)
        ov, am = batched.overlaps_from_mps(sub, clf)
        torch.cuda.synchronize()
        out[opt] = (ov.cpu().numpy(), am.cpu().numpy())
    err = np.abs(out[0][0] - out[1][0]).max() / np.abs(out[0][0]).max()
    print(f"N={n_try}: max overlap diff tc5 vs mma.sync {err:.3e} (of the largest overlap), labels differing {(out[0][1] != out[1][1]).sum()}")
    assert err < 1e-5
for opt in (0, 1, 0, 1):
    batched.set_option("ov_tc5", opt)
    t = timeit(lambda: batched.overlaps_from_mps(batch, clf))
    print(f"ov_tc5={opt}: overlap {t*1e3:.1f} us per {N} images")
batched.set_option("ov_tc5", 0)
