#!/bin/bash
# Round-2 final evidence: GPU tests, bench line (plain), ncu launch list of the same command, ncu --set full of
# every kernel of one device step (float32 and uint8 pixels).  Outputs under gpurun_out/.
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r2f.log 2>&1; tail -3 gpurun_out/gputest_r2f.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f.log 2> gpurun_out/bench_r2f.err || exit 1
tail -c 600 gpurun_out/bench_r2f.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2f_ref.log 2> gpurun_out/bench_r2f_ref.err
tail -c 600 gpurun_out/bench_r2f_ref.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r2f.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2f.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python scripts/one_step.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -s 17 -c 17 -o gpurun_out/step_r2f -f python scripts/one_step.py > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/step_r2f.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:gram -s 4 -c 2 -o gpurun_out/gram_u8_r2f -f python scripts/u8_gram_time.py > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/gram_u8_r2f.ncu-rep
