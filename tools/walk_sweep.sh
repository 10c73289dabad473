#!/bin/bash
# GPU box: k_walk / k_emit time at config #2 over the walker's group width (PGS_WALK_G, the one diagnostic knob).
set -u
O=gpurun_out; mkdir -p $O
specs=()
for g in ${SWEEP_G:-1 2 4 8 16}; do specs+=("PGS_WALK_G=$g"); done
timeout 1500 python tools/variants.py "${specs[@]}" > $O/walk_sweep.log 2>&1
grep "==" $O/walk_sweep.log
