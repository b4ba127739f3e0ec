#!/bin/bash
# run est-fact N times on a regression fixture with the default thread count and print the md5 of the five outputs
C=${1:-test_gtf7}; N=${2:-6}
mkdir -p /tmp/fl_$C && cd /tmp/fl_$C
xz -dc /root/repo/tests/golden/estfact/$C/genomic.txt.xz > genomic.txt
xz -dc /root/repo/tests/golden/estfact/$C/ests.txt.xz > ests.txt
for i in $(seq 1 $N); do
  /root/repo/pintron_b200/bin/est-fact --quiet $3 $4 > /dev/null 2>&1
  echo "run $i rc=$? $(md5sum raw-multifasta-out.txt processed-ests.txt megs.txt processed-megs.txt meg-edges.txt | cut -c1-8 | tr '\n' ' ')"
done
