#!/bin/bash
# est-fact on a regression fixture with PC_CAPTURE, then every captured job against the oracle port (GPU box)
C=${1:-test-788}
cd /root/repo
mkdir -p /tmp/dbg_$C && cd /tmp/dbg_$C
xz -dc /root/repo/tests/golden/estfact/$C/genomic.txt.xz > genomic.txt
xz -dc /root/repo/tests/golden/estfact/$C/ests.txt.xz > ests.txt
PC_CAPTURE=/tmp/dbg_$C/cap.bin /root/repo/pintron_b200/bin/est-fact --quiet --threads 2
md5sum raw-multifasta-out.txt
python /root/repo/tools/check_capture.py /tmp/dbg_$C/cap.bin /tmp/dbg_$C/genomic.txt
