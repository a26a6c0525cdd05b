#!/bin/bash
# end-to-end throughput for several H2D/compute pipeline chunk sizes
for c in ${CHUNKS:-8750 17500 23334 35000 70000}; do
  python bench.py --no-cpu-baseline --e2e-chunk $c 2>/dev/null | tail -1 > /tmp/b.json
  python - "$c" <<'P'
import json, sys
d = json.load(open('/tmp/b.json'))
print(sys.argv[1], 'device', round(d['value'] / 1e6, 2), 'e2e f32', round(d['e2e']['value'] / 1e6, 2), 'e2e u8', round(d['e2e_u8']['value'] / 1e6, 2))
P
done
