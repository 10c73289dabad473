#!/bin/bash
# GPU box: full ncu capture of k_walk at config #2 for one group width (PGS_WALK_G, default 4)
set -u
O=gpurun_out; mkdir -p $O
cp incubator_pegasus_b200/libpegasus_b200.so $O/lib_at_profile.so
G=${1:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 2 -c 1 -f -o $O/walk_g$G python tools/variants.py PGS_WALK_G=$G > $O/ncu_walk_g$G.log 2>&1; tail -1 $O/ncu_walk_g$G.log
