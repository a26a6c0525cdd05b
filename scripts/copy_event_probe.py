"""Does an event recorded BETWEEN two async H2D copies of one stream release a waiting stream when the first
copy ends, or only when the copy engine's queue drains?  (oc_classify_host's first chunk started late.)"""
import torch, time
N = 14_000_000
h0 = torch.empty(N, dtype=torch.uint8).pin_memory(); h1 = torch.empty(N, dtype=torch.uint8).pin_memory()
d0 = torch.empty(N, dtype=torch.uint8, device="cuda"); d1 = torch.empty_like(d0)
x = torch.zeros(1 << 20, device="cuda")
cp, cp2, comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
for variant in ("one copy stream", "two copy streams chained by an event"):
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); e_c0 = torch.cuda.Event(enable_timing=True)
        e_c1 = torch.cuda.Event(enable_timing=True); e_k = torch.cuda.Event(enable_timing=True)
        t0.record(torch.cuda.current_stream()); cp.wait_event(t0); cp2.wait_event(t0); comp.wait_event(t0)
        with torch.cuda.stream(cp):
            d0.copy_(h0, non_blocking=True); e_c0.record(cp)
        with torch.cuda.stream(comp):
            comp.wait_event(e_c0); x.add_(1.0); e_k.record(comp)
        if variant.startswith("one"):
            with torch.cuda.stream(cp):
                d1.copy_(h1, non_blocking=True); e_c1.record(cp)
        else:
            with torch.cuda.stream(cp2):
                cp2.wait_event(e_c0); d1.copy_(h1, non_blocking=True); e_c1.record(cp2)
        torch.cuda.synchronize()
    print(f"{variant}: copy 0 done {t0.elapsed_time(e_c0):.3f} ms, kernel behind it done {t0.elapsed_time(e_k):.3f} ms, "
          f"copy 1 done {t0.elapsed_time(e_c1):.3f} ms")
