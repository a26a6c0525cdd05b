"""Time oc_encode_batch / oc_overlap_batch under different launch options."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200._lib import lib, check
from orthogonal_classifiers_b200.synthetic import synthetic_images

N = 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, (9, 1))[:N]).cuda()
def t_encode(reps=5):
    for _ in range(2): oc.encode_images(imgs, 32)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): oc.encode_images(imgs, 32)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for w in (4, 2, 1):
    check(lib.oc_set_option(b"enc_warps", w))
    print("enc_warps", w, "encode ms", round(t_encode(), 3))
