#!/bin/bash
# Sampling profile (tools/ef_prof.c) of the shipped host code, one worker thread.  Build first: make -C pintron_b200/host prof
#   WORKLOAD=C4 tools/gprof_estfact.sh READS [--lines]
READS=${1:-5000}
W=${WORKLOAD:-C4}
cd /root/repo
python - <<PY
import os
from pintron_b200.synth import Synth
os.makedirs("/tmp/gp_$W", exist_ok=True)
s = Synth("$W", reads=$READS)
open("/tmp/gp_$W/genomic.txt","wb").write(s.genome_fasta())
open("/tmp/gp_$W/ests.txt","wb").write(s.ests_fasta(0,$READS))
PY
cd /tmp/gp_$W
/root/repo/pintron_b200/bin/est-fact-pg --threads 1 2> err.log
grep -E "Timer (Algorithm|Total)|device batches|thread-seconds|scheduler|by phase" err.log
python /root/repo/tools/ef_prof_report.py /root/repo/pintron_b200/bin/est-fact-pg ef_prof.out $2 | tee /root/repo/gpurun_out/prof_$W.txt | head -40
python /root/repo/tools/ef_prof_report.py /root/repo/pintron_b200/bin/est-fact-pg ef_prof.out --lines > /root/repo/gpurun_out/prof_lines_$W.txt
