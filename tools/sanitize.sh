#!/bin/bash
# compute-sanitizer over the device path (run under gpurun): memcheck / racecheck / initcheck on (1) the C-ABI parity
# tests' mixed batch and (2) the shipped est-fact on small regression cases, single- and multi-threaded.  Logs go to
# gpurun_out/sanitizer_<tag>_*.log; a summary line per run is printed.  VERDICT r1 #3.
TAG=${1:-r2}
cd /root/repo
OUT=gpurun_out
S="compute-sanitizer --error-exitcode 99 --print-limit 20"
unpack() { mkdir -p /tmp/san_$1 && xz -dc tests/golden/estfact/$1/genomic.txt.xz > /tmp/san_$1/genomic.txt && xz -dc tests/golden/estfact/$1/ests.txt.xz > /tmp/san_$1/ests.txt; }
for tool in memcheck racecheck initcheck; do
  # (1) every kernel once, through the C ABI
  timeout 1500 $S --tool $tool python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mixed_batch or gap_packed or packed_borders or golden" \
      > $OUT/sanitizer_${TAG}_${tool}_cabi.log 2>&1
  echo "$tool cabi rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $OUT/sanitizer_${TAG}_${tool}_cabi.log) clean-summaries; $(grep 'ERROR SUMMARY' $OUT/sanitizer_${TAG}_${tool}_cabi.log | tail -1)"
  # (2) the program (in-process engine: the sanitizer follows this process only)
  for c in test-AMBN test_gtf7; do
    unpack $c
    for th in 1 6; do
      [ $c = test_gtf7 ] && [ $tool != memcheck ] && continue
      ( cd /tmp/san_$c && timeout 1500 $S --tool $tool /root/repo/pintron_b200/bin/est-fact --engine inproc --threads $th --quiet \
          > /root/repo/$OUT/sanitizer_${TAG}_${tool}_${c}_t$th.log 2>&1; echo "$tool $c threads=$th rc=$? $(grep 'ERROR SUMMARY' /root/repo/$OUT/sanitizer_${TAG}_${tool}_${c}_t$th.log | tail -1) md5 $(md5sum raw-multifasta-out.txt | cut -c1-8)" )
    done
  done
done
