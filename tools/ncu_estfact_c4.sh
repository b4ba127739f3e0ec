#!/bin/bash
# per-launch device times of the shipped est-fact on a small C4 sample (one host thread, so launches are serial)
cd /root/repo
python - <<PY
import os
from pintron_b200.synth import Synth
os.makedirs("/tmp/c4s", exist_ok=True)
s = Synth("C4", reads=1500)
open("/tmp/c4s/genomic.txt","wb").write(s.genome_fasta())
open("/tmp/c4s/ests.txt","wb").write(s.ests_fasta(0,1500))
PY
cd /tmp/c4s
/root/repo/pintron_b200/bin/est-fact --threads 1 --fibers 256 2> plain.log && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /root/repo/gpurun_out/launches_c4.csv /root/repo/pintron_b200/bin/est-fact --threads 1 --fibers 256 > ncu.log 2>&1
tail -3 plain.log
