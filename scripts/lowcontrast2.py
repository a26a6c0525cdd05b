import sys
sys.path.insert(0, ".")
import numpy as np
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200 import batched
import oracle
from orthogonal_classifiers_b200.synthetic import synthetic_images
imgs = synthetic_images(8, seed=37).astype(np.float64)
eps = 1e-3
low = (0.5 + eps * imgs).astype(np.float32)
for name, opts in (("default", {}), ("old hw", {"enc_gram": 0, "enc_mgs4": 0, "enc_qr5": 0, "enc_pack": 0}), ("fast", {"enc_hw": 0}), ("generic", {"force_generic": 1})):
    for k, v in opts.items(): batched.set_option(k, v)
    b = oc.encode_images(low, 32, return_svals=True, return_sweeps=True)
    for k in opts: batched.set_option(k, 0 if k == "force_generic" else 1)
    sv = b.svals.cpu().numpy().astype(np.float64)
    i = 0
    x, _ = oracle.pad_image(low[i].astype(np.float64)); xh = x / np.linalg.norm(x)
    sites = [s.astype(np.float64) for s in b.sites_numpy(i)]
    st = oracle.mps_to_state(sites)
    err = min(np.abs(st - xh).max(), np.abs(st + xh).max())
    ref_sv = oracle.bond_singular_values(low[i].astype(np.float64))
    sverr = [np.abs(sv[i, n, :len(ref_sv[n])][:sv.shape[2]] - ref_sv[n][:sv.shape[2]]).max() / ref_sv[n][0] for n in range(10)]
    # partial reconstructions: contract sites 0..n and compare the projector onto the exact row space
    iso = []
    for n, A in enumerate(sites):
        M = A.reshape(-1, A.shape[2]); G = M.T @ M; nz = np.diag(G) > 0.5
        iso.append(np.abs(G[np.ix_(nz, nz)] - np.eye(nz.sum())).max())
    print(f"{name:8s} amp err {err:.2e} (x max {np.abs(xh).max():.2e})  sweeps {b.sweeps[i].cpu().numpy().tolist()}")
    print("     sv err/step", " ".join(f"{e:.1e}" for e in sverr))
    print("     iso/step   ", " ".join(f"{e:.1e}" for e in iso))
