#!/bin/bash
# diagnostics: k_merge phase timing for a list of PGS_VARIANT values (one bench run each)
out=gpurun_out/variants.log
: > $out
for v in "$@"; do
  echo "== variant $v" >> $out
  PGS_VARIANT=$v PGS_PHASE_TIMING=1 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-reads --skip-e2e 2> gpurun_out/v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms', d['roofline']['kernel_ms'])" >> $out
  grep "k_merge phases" gpurun_out/v.err | tail -1 >> $out
  PGS_VARIANT=$v python bench.py --steps 5 --warmup 3 --skip-cpu --skip-reads --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('untimed kernel_ms', d['roofline']['kernel_ms'])" >> $out
done
cat $out
