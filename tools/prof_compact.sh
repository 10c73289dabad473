#!/bin/bash
# GPU box: full ncu captures of k_walk and k_emit at config #2 (4 x 2.5 M records), plus the timing line.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python tools/variants.py default > $O/variants.log 2>&1; tail -3 $O/variants.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 2 -c 1 -f -o $O/walk python tools/variants.py default > $O/ncu_walk.log 2>&1; tail -2 $O/ncu_walk.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit -s 2 -c 1 -f -o $O/emit python tools/variants.py default > $O/ncu_emit.log 2>&1; tail -2 $O/ncu_emit.log
ls -la $O | grep ncu-rep
