#!/bin/bash
# ncu --set full of the GAP kernel inside bench.py's kernel-level leg (run under gpurun); raw + source pages as csv.
TAG=${1:-r2}
cd /root/repo
ARGS="--kernels-only --steps 2 --warmup 3"
ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:"k_gap_pairs" --launch-skip 3 -c 2 -f -o /tmp/${TAG}_gap \
    python bench.py $ARGS > gpurun_out/${TAG}_gap_ncu.log 2>&1
ncu -i /tmp/${TAG}_gap.ncu-rep --page raw --csv > gpurun_out/${TAG}_gap_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_gap.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_gap_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_gap_*; tail -3 gpurun_out/${TAG}_gap_ncu.log
