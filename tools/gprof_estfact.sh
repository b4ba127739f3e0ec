#!/bin/bash
# Sampling profile (tools/ef_prof.c) of the shipped HOST code, one worker thread, as a client of est-factd: the profiled
# process holds no CUDA, so every sample is per-EST host code.  Build first: make -C pintron_b200/host prof
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
export EST_FACTD_SOCKET=/tmp/gp_$W/efd.sock
/root/repo/pintron_b200/bin/est-factd --socket $EST_FACTD_SOCKET --idle-timeout 20 > efd.log 2>&1 &
while [ ! -S $EST_FACTD_SOCKET ]; do sleep 0.1; done
/root/repo/pintron_b200/bin/est-fact-pg --engine daemon --threads 1 2> err.log
grep -E "Timer (Algorithm|Total)|lane batches|thread-seconds|scheduler|by phase|engine \(" err.log
python /root/repo/tools/ef_prof_report.py /root/repo/pintron_b200/bin/est-fact-pg ef_prof.out $2 | tee /root/repo/gpurun_out/prof_$W.txt | head -60
python /root/repo/tools/ef_prof_report.py /root/repo/pintron_b200/bin/est-fact-pg ef_prof.out --lines > /root/repo/gpurun_out/prof_lines_$W.txt
/root/repo/pintron_b200/bin/est-factd --socket $EST_FACTD_SOCKET --stop
