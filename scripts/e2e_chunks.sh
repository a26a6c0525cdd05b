#!/bin/bash
# oc_classify_host chunk schedule: e2e (uint8) and e2e_f32 for a few stage sizes
for c in 8750 11667 17500 23334 35000 70000; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config5 --e2e-chunk $c 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print($c, 'u8 %.3f ms  f32 %.3f ms' % (d['e2e']['ms_per_step'], d['e2e_f32']['ms_per_step']))"
done
