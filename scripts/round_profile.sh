#!/bin/bash
# Round-end evidence: bench line (plain), ncu launch list of the same command, ncu --set full of the
# dominant kernel (step 4).  Outputs under gpurun_out/.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1.log 2> gpurun_out/bench_r1.err || exit 1
tail -1 gpurun_out/bench_r1.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r1g.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hw_row4q_kernel -s 4 -c 1 -o gpurun_out/prof_step4_r1g -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/prof_step4_r1g.ncu-rep
