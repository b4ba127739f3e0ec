#!/bin/bash
# ncu evidence for bench.py's kernel-level leg (run under gpurun).  Plain run first (must exit 0), then the launch list of the
# same command, then one --set full capture of the hot kernels' big launches; the raw page is exported as csv (the .ncu-rep
# stays in /tmp: gpurun_out/ is limited to 64 MiB).  TAG names the files under gpurun_out/.
W=${1:-C3}; TAG=${2:-r2}
cd /root/repo
ARGS="--workload $W --kernels-only --steps 2 --warmup 3"
timeout 300 python bench.py $ARGS > gpurun_out/${TAG}_${W}_plain.json 2> gpurun_out/${TAG}_${W}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_${W}_plain.err; exit 1; }
timeout 600 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_${W}_ncu1.log 2>&1
# the merged batch's launches are the last ones of the run: skip the warm-up's (3 warm-up steps of the one-batch replay + 2 merged)
timeout 900 ncu --target-processes application-only --set full --clock-control none --import-source on \
    -k regex:"k_gap_pairs|k_borders_packed|k_myers|k_seed|k_lcs_len|k_lcs_locate|k_warp_per_job|k_borders_chunked" -f -o /tmp/${TAG}_${W}_full \
    python bench.py $ARGS > gpurun_out/${TAG}_${W}_ncu2.log 2>&1
ncu -i /tmp/${TAG}_${W}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_${W}_full_raw.csv 2>/dev/null
python tools/ncu_kernels_json.py gpurun_out/${TAG}_${W}_full_raw.csv $W > gpurun_out/${TAG}_${W}_ncu_kernels.json
ls -la gpurun_out/${TAG}_${W}_*
