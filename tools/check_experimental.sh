#!/bin/bash
# Runs on the GPU box: parity tests and kernel time of the k_merge<NT, EXP=true> instantiation (PGS_EXPERIMENTAL=1:
# packed key-rebuild metadata, per-survivor varint packing, sample-then-refine rank searches).  The default
# instantiation contains none of that code (checked by diffing its SASS), so tests/ never exercise it.
set -u
mkdir -p gpurun_out
# bit mask: 1 = base items (packed key-rebuild metadata, per-survivor varint packing), 2 = + sample-then-refine rank,
# 4 = + staged heads (one-pass chunk writes), 8 = + shuffle-scan key rebuild
for mask in 1 3 5 9 15; do
  echo "== PGS_EXPERIMENTAL=$mask"
  PGS_EXPERIMENTAL=$mask timeout 300 python -m pytest tests/test_compaction_gpu.py tests/test_edge_cases_gpu.py tests/test_engine_variants_gpu.py -x -q -m gpu 2>&1 | tail -3
done
timeout 300 python tools/variants.py default PGS_EXPERIMENTAL=1 PGS_EXPERIMENTAL=3 PGS_EXPERIMENTAL=5 PGS_EXPERIMENTAL=9 PGS_EXPERIMENTAL=15 2>&1 | tail -12
