#!/bin/bash
# Round-2 final evidence (outputs under gpurun_out/; one ncu report per gpurun call: 64 MiB come back at most).
#   part 1: GPU tests, bench line (plain), reference arm, ncu launch list of the bench command
#   part 2: ncu --set full of every kernel of one device step (float32 pixels)
#   part 3: ncu --set full of the uint8 Gram / head kernels
set -x
case "${1:-1}" in
1)
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r2f.log 2>&1; tail -3 gpurun_out/gputest_r2f.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2f.log 2> gpurun_out/bench_r2f.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2f_ref.log 2> gpurun_out/bench_r2f_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r2f.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2f.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
python scripts/sweep_time.py > gpurun_out/sweep_time_r2f.log 2>&1
python scripts/u8_gram_time.py > gpurun_out/u8_gram_time_r2f.log 2>&1
;;
2)
python scripts/one_step.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -s 16 -c 16 -o gpurun_out/step_r2f -f python scripts/one_step.py > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/step_r2f.ncu-rep
;;
3)
ncu --set full --clock-control none --import-source on -k regex:gram -s 4 -c 2 -o gpurun_out/gram_u8_r2f -f python scripts/u8_gram_time.py > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/gram_u8_r2f.ncu-rep
;;
esac
