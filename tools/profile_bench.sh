#!/bin/bash
# ncu evidence for bench.py's device leg (run under gpurun): plain run first (must exit 0), then the launch list, then one
# --set full capture of the hot kernels (the .ncu-rep stays in /tmp: gpurun_out/ is limited to 64 MiB; the raw page is
# exported as csv).  TAG names the files under gpurun_out/.
W=${1:-C4}; TAG=${2:-r1p}; NFULL=${3:-20}
cd /root/repo
ARGS="--workload $W --no-e2e --no-cpu-baseline --steps 2 --warmup 3"
python bench.py $ARGS > gpurun_out/${TAG}_${W}_plain.json 2> gpurun_out/${TAG}_${W}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_${W}_plain.err; exit 1; }
ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_${W}_ncu1.log 2>&1
ncu --target-processes application-only --set full --clock-control none --import-source on \
    -k regex:"k_gap_pairs|k_borders_packed|k_myers|k_seed|k_lcs|k_warp_per_job" --launch-skip 28 -c $NFULL -f -o /tmp/${TAG}_${W}_full \
    python bench.py $ARGS > gpurun_out/${TAG}_${W}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_${W}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_${W}_full_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_${W}_*
