"""Encoder per-step times (library events) for uint8 device pixels, integer Gram (IMMA) on / off."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200._lib import lib, check
from orthogonal_classifiers_b200.synthetic import synthetic_images
N = 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
u8 = torch.from_numpy(np.rint(np.tile(imgs, ((N + 8191) // 8192, 1))[:N] * 255).astype(np.uint8)).cuda()
for v in (1, 0, 1, 0):
    check(lib.oc_set_option(b"enc_u8mma", v))
    for _ in range(3): oc.encode_images(u8, 32)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): oc.encode_images(u8, 32)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 20
    check(lib.oc_set_option(b"enc_events", 1))
    for _ in range(8): oc.encode_images(u8, 32)
    ms = (C.c_float * 5)(); check(lib.oc_encode_step_ms(ms, 5)); check(lib.oc_set_option(b"enc_events", 0))
    print("enc_u8mma", v, "encode %.3f ms" % t, [round(float(x), 3) for x in ms], flush=True)
check(lib.oc_set_option(b"enc_u8mma", 1))
