"""Per-launch encoder times (CUDA events inside the library) for a few option settings."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200._lib import lib, check
from orthogonal_classifiers_b200.synthetic import synthetic_images
N = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
for opts in ([], [(b"enc_sort", 0)], [(b"enc_hw", 0)]):
    check(lib.oc_set_option(b"enc_sort", 1)); check(lib.oc_set_option(b"enc_hw", 1))
    for k, v in opts: check(lib.oc_set_option(k, v))
    for _ in range(3): oc.encode_images(imgs, 32)
    check(lib.oc_set_option(b"enc_events", 1))
    for _ in range(8): oc.encode_images(imgs, 32)
    ms = (C.c_float * 5)(); check(lib.oc_encode_step_ms(ms, 5)); check(lib.oc_set_option(b"enc_events", 0))
    print(opts, [round(float(v), 3) for v in ms], "total", round(sum(ms), 3))
