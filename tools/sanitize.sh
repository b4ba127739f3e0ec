#!/bin/bash
# compute-sanitizer over the device path (run under gpurun): memcheck / racecheck / initcheck on (1) the C-ABI parity
# tests (every kernel, golden + fuzz batches) and (2) the shipped est-fact (in-process engine: the sanitizer follows one
# process) on regression cases.  Logs: gpurun_out/sanitizer_<tag>_*.log; one summary line per run.  VERDICT r1 #3.
TAG=${1:-r2}
cd /root/repo
OUT=gpurun_out
S="compute-sanitizer --error-exitcode 99 --print-limit 20"
unpack() { mkdir -p /tmp/san_$1 && xz -dc tests/golden/estfact/$1/genomic.txt.xz > /tmp/san_$1/genomic.txt && xz -dc tests/golden/estfact/$1/ests.txt.xz > /tmp/san_$1/ests.txt; }
prog() {  # tool case threads
  unpack $2
  ( cd /tmp/san_$2 && timeout 900 $S --tool $1 /root/repo/pintron_b200/bin/est-fact --engine inproc --threads $3 --quiet \
      > /root/repo/$OUT/sanitizer_${TAG}_$1_$2_t$3.log 2>&1
    echo "$1 est-fact $2 threads=$3 rc=$? $(grep 'ERROR SUMMARY' /root/repo/$OUT/sanitizer_${TAG}_$1_$2_t$3.log | tail -1) raw-md5 $(md5sum raw-multifasta-out.txt | cut -c1-8) expected $(python3 -c "import json;print(json.load(open('/root/repo/tests/golden/estfact/$2/expected.json'))['raw-multifasta-out.txt']['md5'][:8])")" )
}
for tool in memcheck racecheck initcheck; do
  timeout 900 $S --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_engine_gpu.py -q -m gpu -x \
      -k "mixed_batch or gap_packed or packed_borders or golden or parts_equal or bit_parallel" > $OUT/sanitizer_${TAG}_${tool}_cabi.log 2>&1
  echo "$tool C-ABI tests rc=$? $(grep -E 'passed|failed' $OUT/sanitizer_${TAG}_${tool}_cabi.log | tail -1) $(grep 'ERROR SUMMARY' $OUT/sanitizer_${TAG}_${tool}_cabi.log | tail -1)"
  prog $tool test-AMBN 6
done
prog memcheck test_gtf7 1
prog memcheck test_gtf7 6
