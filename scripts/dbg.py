import sys
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200.synthetic import synthetic_images
imgs = synthetic_images(2048, seed=2, labels=np.arange(2048) % 10, dtype=np.float32)
b = oc.encode_images(torch.from_numpy(imgs).cuda(), 32, return_sweeps=True, return_svals=True)
s = b.sweeps.cpu().numpy()
print("sweeps", (s % 100).mean(0).round(2))
print("mlive ", ((s // 100) % 100).mean(0).round(2))
print("lost  ", (s // 10000).mean(0).round(2))
sv = b.svals.cpu().numpy()
print("sv step3 img0", sv[0,3,:16])
A = b.sites_numpy(0)
print("A3 col norms", np.linalg.norm(A[3].reshape(16,16),axis=0).round(3))
print("A4 row-l norms", np.linalg.norm(A[4].transpose(1,0,2).reshape(16,64),axis=1).round(3))
print("A4 col norms", np.linalg.norm(A[4].reshape(32,32),axis=0).round(3))
