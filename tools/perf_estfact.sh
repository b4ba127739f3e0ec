#!/bin/bash
# Times the shipped est-fact on a synthetic sample for a few thread/fiber settings (run under gpurun).
#   tools/perf_estfact.sh READS "<est-fact flags>" ...      env: WORKLOAD=C3|C4|C5mini, PC_PROFILE=1 for per-op device times
READS=${1:-20000}
W=${WORKLOAD:-C3}
cd /root/repo
python - <<PY
import os, time
from pintron_b200.synth import Synth
os.makedirs("/tmp/perf_$W", exist_ok=True)
t=time.time()
s = Synth("$W", reads=$READS)
open("/tmp/perf_$W/genomic.txt","wb").write(s.genome_fasta())
open("/tmp/perf_$W/ests.txt","wb").write(s.ests_fasta(0,$READS))
print("synth s", time.time()-t)
PY
nproc
cd /tmp/perf_$W
ls -la ests.txt genomic.txt
shift
for cfg in "$@"; do
  echo "== $W $READS reads: $cfg"
  s=$(date +%s.%N)
  { time /root/repo/pintron_b200/bin/est-fact $cfg 2> err.log ; } 2> time.log
  e=$(date +%s.%N); echo "rc=$? wall $(python -c "print(round($e - $s, 3))") s"
  tr "\n" " " < time.log; echo
  grep -E "Timer (Algorithm|Total)|device batches|thread-seconds|scheduler|by phase|pc profile|pc op" err.log
done
