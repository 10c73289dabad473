#!/bin/bash
# Runs on the GPU box (under gpurun): tests, bench, ncu launch list, one full ncu capture per hot kernel, phase timing.
# Outputs land in gpurun_out/; tools/summarize_profiles.py turns them into the tracked files under profiles/.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -x -q -m gpu > $O/final_tests.log 2>&1; tail -2 $O/final_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/final_bench.json 2> $O/final_bench.err; head -c 400 $O/final_bench.json; echo
PGS_PHASE_TIMING=1 timeout 300 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-reads > /dev/null 2> $O/final_phases.err; grep "k_merge phases" $O/final_phases.err | tail -1 > $O/final_phases.txt; cat $O/final_phases.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > $O/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_merge -s 1 -c 1 -f -o $O/final_k_merge python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e --skip-reads > $O/final_ncu_merge.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_get -s 1 -c 1 -f -o $O/final_reads python bench.py --steps 1 --warmup 1 --skip-cpu --skip-e2e > $O/final_ncu_reads.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_scan -s 1 -c 1 -f -o $O/final_scan python bench.py --steps 1 --warmup 1 --skip-cpu --skip-e2e > $O/final_ncu_scan.log 2>&1
ls -la $O | grep final
