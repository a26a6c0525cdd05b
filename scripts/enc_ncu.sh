#!/bin/bash
# launch list (gpu__time_duration) of a few encode calls: python scripts/quick.py under ncu
python scripts/quick.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_q.csv python scripts/quick.py > gpurun_out/ncu_q.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_q.csv
