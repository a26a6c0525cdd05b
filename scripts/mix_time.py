"""Per-step encoder times (library events) with an option on / off, e.g. `mix_time.py enc_mix`."""
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import orthogonal_classifiers_b200 as oc
from orthogonal_classifiers_b200._lib import lib, check
from orthogonal_classifiers_b200.synthetic import synthetic_images
opt = sys.argv[1].encode() if len(sys.argv) > 1 else b"enc_mix"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 70000
imgs = synthetic_images(8192, seed=2, labels=np.arange(8192) % 10, dtype=np.float32)
imgs = torch.from_numpy(np.tile(imgs, ((N + 8191) // 8192, 1))[:N]).cuda()
def plain(reps=20):
    for _ in range(3): oc.encode_images(imgs, 32)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): oc.encode_images(imgs, 32)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for v in (1, 0, 1, 0):
    check(lib.oc_set_option(opt, v))
    t = plain()
    check(lib.oc_set_option(b"enc_events", 1))
    for _ in range(8): oc.encode_images(imgs, 32)
    ms = (C.c_float * 5)(); check(lib.oc_encode_step_ms(ms, 5)); check(lib.oc_set_option(b"enc_events", 0))
    print(opt.decode(), v, "encode %.3f ms" % t, [round(float(x), 3) for x in ms], flush=True)
check(lib.oc_set_option(opt, 1))
