#!/bin/bash
# Times the shipped est-fact on a synthetic C3 sample for a few thread/fiber settings (run under gpurun).
READS=${1:-20000}
cd /root/repo
python - <<PY
import os
from pintron_b200.synth import Synth
os.makedirs("/tmp/c3", exist_ok=True)
s = Synth(os.environ.get("WORKLOAD", "C3"), reads=$READS)
open("/tmp/c3/genomic.txt","wb").write(s.genome_fasta())
open("/tmp/c3/ests.txt","wb").write(s.ests_fasta(0,$READS))
PY
nproc
cd /tmp/c3
shift
for cfg in "$@"; do
  echo "== $cfg"
  s=$(date +%s.%N)
  /root/repo/pintron_b200/bin/est-fact $cfg 2> err.log
  echo "rc=$?"
  grep -E "Timer (Algorithm|Total)|device batches|thread-seconds|scheduler|by phase" err.log
done
