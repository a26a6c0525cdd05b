"""Two halves of an encoder chunk on two streams (oc_set_option("enc_split2", 1)) against the
single-stream chain: device timing on 70,000 images and bit-identity of the results."""
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
_, offs = batched.mps_layout(10, 32)
ws = torch.empty((N, offs[-1]), dtype=torch.float32, device="cuda")
ov = torch.empty((N, 16), dtype=torch.float32, device="cuda")
am = torch.empty((N,), dtype=torch.int32, device="cuda")

def run():
    batched.classify_device(imgs, clf, 32, workspace=ws, out=(ov, am))

def timeit(reps=20):
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): run()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {}
for rep in range(2):
    for flag in (0, 1):
        batched.set_option("enc_split2", flag)
        t = timeit()
        res.setdefault(flag, []).append(t)
        if rep == 0:
            res[("ov", flag)] = ov.clone(); res[("ws", flag)] = ws.clone(); res[("am", flag)] = am.clone()
batched.set_option("enc_split2", 0)
print(f"N={N} encode+classify ms: single-stream {min(res[0]):.3f}  split2 {min(res[1]):.3f}")
print("identical overlaps:", bool((res[('ov', 0)] == res[('ov', 1)]).all()), " labels:", bool((res[('am', 0)] == res[('am', 1)]).all()),
      " max |mps diff|:", float((res[('ws', 0)] - res[('ws', 1)]).abs().max()))
