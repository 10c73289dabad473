#!/bin/bash
# Runs on the GPU box: memcheck + racecheck of a small compaction, the GPU parity tests, kernel timing at config #2.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_compaction_gpu.py -k "small or default_ttl" -x -q > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_compaction_gpu.py -k "small" -x -q > $O/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazards" $O/racecheck.log; tail -3 $O/racecheck.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/gpu_tests.log 2>&1; echo "tests rc=$?"; tail -15 $O/gpu_tests.log
timeout 600 python tools/variants.py default > $O/variants.log 2>&1; tail -3 $O/variants.log
