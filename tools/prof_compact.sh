#!/bin/bash
# GPU box: full ncu captures of k_walk and k_emit at config #2 (4 x 2.5 M records), plus the timing line.  The library that was
# profiled is saved next to the captures (tools/ncu_lines.py maps SASS to source lines through it).
set -u
O=gpurun_out
mkdir -p $O
cp incubator_pegasus_b200/libpegasus_b200.so $O/lib_at_profile.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 2 -c 1 -f -o $O/walk python tools/variants.py default > $O/ncu_walk.log 2>&1; tail -1 $O/ncu_walk.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emit -s 2 -c 1 -f -o $O/emit python tools/variants.py default > $O/ncu_emit.log 2>&1; tail -1 $O/ncu_emit.log
