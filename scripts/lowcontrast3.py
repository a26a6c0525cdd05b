import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
from orthogonal_classifiers_b200.synthetic import synthetic_images
imgs = synthetic_images(8, seed=37).astype(np.float64)
low = (0.5 + 1e-3 * imgs).astype(np.float32)
b = oc.encode_images(low, 32, return_svals=True)
A = b.sites_numpy(0)[5].astype(np.float64)
M = A.reshape(-1, A.shape[2]); G = M.T @ M
np.set_printoptions(precision=1, linewidth=250, suppress=False)
print("svals step5 rel", (b.svals[0, 5, :16].cpu().numpy() / b.svals[0, 5, 0].item()))
print("diag-1", np.diag(G) - 1)
off = G - np.diag(np.diag(G))
print("max offdiag", np.abs(off).max(), "at", np.unravel_index(np.abs(off).argmax(), off.shape))
print(off[:6, :6])
