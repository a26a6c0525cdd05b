"""BASELINE.json config 5: 1,000,000 synthetic full 32 x 32 images sharded over the
GPUs of one box (strong scaling: the total is fixed), D_encode = D_total = 32,
classifier replicated, one all-reduce of (correct, total).

    python scripts/config5.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29511 scripts/config5.py

Device-resident shards, CUDA events on the launching stream, barrier + synchronize on
both sides, max over ranks.  Prints one JSON line on rank 0.  Size-independent checks
at the full size: every image is counted exactly once (all-reduced total = 1,000,000)
and the shard is a tiling of 4,096 distinct images, so every repeat must give the
bit-identical overlaps and label of its first occurrence wherever it lands in a chunk.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

N_TOTAL = int(os.environ.get("OC_C5_IMAGES", "1000000"))
STEPS, WARMUP, BASE = 5, 3, 4096


def base_images(seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:32, 0:32]
    out = np.empty((BASE, 1024), dtype=np.float32)
    for i in range(BASE):
        k = int(rng.integers(2, 5))
        cx, cy, w = rng.uniform(4, 28, k), rng.uniform(4, 28, k), rng.uniform(1, 4, k)
        img = sum(np.exp(-((xx - a) ** 2 + (yy - b) ** 2) / (2 * c ** 2)) for a, b, c in zip(cx, cy, w))
        out[i] = np.clip(img, 0, 1).reshape(-1)
    out[out < 0.1] = 0
    return np.rint(out * 255).astype(np.float32) / 255


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import orthogonal_classifiers_b200 as oc
    from orthogonal_classifiers_b200 import batched
    from orthogonal_classifiers_b200.synthetic import random_classifier

    lo, hi = oc.shard_bounds(N_TOTAL, rank, world)
    n = hi - lo
    base = torch.from_numpy(base_images(2 + rank)).to(dev)
    imgs = base.repeat((n + BASE - 1) // BASE, 1)[:n].contiguous()
    labels = torch.from_numpy(np.random.default_rng(100 + rank).integers(0, 10, n).astype(np.uint8)).to(dev)
    clf = oc.DeviceClassifier(random_classifier(10, 32, centre=5, seed=1), dtype="f32", device=dev)
    _, offs = batched.mps_layout(10, 32)
    ws = torch.empty((n, offs[-1]), dtype=torch.float32, device=dev)
    ov = torch.empty((n, 16), dtype=torch.float32, device=dev)
    am = torch.empty((n,), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        batched.classify_device(imgs, clf, 32, workspace=ws, out=(ov, am))
        counts = batched.count_correct(am, labels)
        return oc.all_reduce_counts(counts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(WARMUP):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(STEPS):
        counts = step()
    b.record(stream)
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / STEPS

    # repeats of a base image: identical overlaps and labels wherever they sit in the shard
    reps = n // BASE
    ok = True
    if reps >= 2:
        ov3 = ov[: reps * BASE].view(reps, BASE, 16)
        am2 = am[: reps * BASE].view(reps, BASE)
        ok = bool((ov3 == ov3[0:1]).all().item()) and bool((am2 == am2[0:1]).all().item())
    flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(flag)
    c = counts.cpu()
    if rank == 0:
        print(json.dumps({
            "config": "BASELINE.json configs[4]: 1M synthetic 32x32 images, D_encode=D_total=32, "
                      "classifier replicated, count all-reduce",
            "n_gpus": world, "images_total": N_TOTAL, "images_per_gpu": n, "scaling": "strong",
            "ms_per_pass": ms, "images_per_s": N_TOTAL / (ms * 1e-3), "steps": STEPS, "warmup": WARMUP,
            "counted_total": int(c[1].item()), "counted_once": int(c[1].item()) == N_TOTAL,
            "repeats_bit_identical": int(flag.item()) == 0,
            "accuracy_random_labels": float(c[0].item()) / max(1, int(c[1].item()))}))
    if world > 1:
        dist.destroy_process_group()
    if int(flag.item()) != 0 or int(c[1].item()) != N_TOTAL:
        sys.exit(1)


if __name__ == "__main__":
    main()
