"""Per-image sweep counts and bond ranks of the encoder (to study the cost of pairing two images per warp)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200.synthetic import synthetic_images
N = 8192
imgs = synthetic_images(N, seed=2, labels=np.arange(N) % 10, dtype=np.float32)
b = oc.encode_images(torch.from_numpy(imgs).cuda(), 32, return_sweeps=True, return_svals=True)
sw = b.sweeps.cpu().numpy()
rk = (b.svals.cpu().numpy() > 0).sum(2)
np.savez("gpurun_out/pair_stats.npz", sweeps=sw, ranks=rk)
print(sw.mean(0), rk.mean(0))
