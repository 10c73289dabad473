#!/bin/bash
# Runs on the GPU box: parity tests and kernel time of the k_merge<NT, EXP=true> instantiation (PGS_EXPERIMENTAL=1:
# packed key-rebuild metadata, per-survivor varint packing, sample-then-refine rank searches).  The default
# instantiation contains none of that code (checked by diffing its SASS), so tests/ never exercise it.
set -u
mkdir -p gpurun_out
PGS_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_compaction_gpu.py tests/test_edge_cases_gpu.py tests/test_engine_variants_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 200 python tools/variants.py default PGS_EXPERIMENTAL=1 2>&1 | tail -6
