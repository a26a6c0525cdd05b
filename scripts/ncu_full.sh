#!/bin/bash
# ncu --set full of kernels matching $1 (regex) in scripts/quick.py; report to gpurun_out/$2.ncu-rep
python scripts/quick.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-4} -c ${4:-1} -o gpurun_out/$2 -f python scripts/quick.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/$2.ncu-rep
