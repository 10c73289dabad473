#!/bin/bash
# GPU box: k_walk / k_emit time at config #2 over the walker's diagnostic knobs (group width, segment budget, register cap).
set -u
O=gpurun_out; mkdir -p $O
specs=()
for g in ${SWEEP_G:-2 4 8}; do for w in ${SWEEP_W:-65536 131072}; do for m in ${SWEEP_M:-4 5 6 8}; do specs+=("PGS_WALK_G=$g,PGS_SEG_WEIGHT=$w,PGS_WALK_MINB=$m"); done; done; done
timeout 1500 python tools/variants.py "${specs[@]}" > $O/walk_sweep.log 2>&1
grep "==" $O/walk_sweep.log
