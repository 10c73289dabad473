#!/bin/bash
for kb in 0 64 80; do
  PGS_SCAN_POOL_KB=$kb PGS_PHASE_TIMING=1 timeout 200 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/sp.json 2> gpurun_out/sp.err
  echo "== pool_kb=$kb"; grep "k_scan phases" gpurun_out/sp.err | tail -1
  PGS_SCAN_POOL_KB=$kb timeout 200 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print({k:(round(v['value']/1e6,1),round(v['e2e']/1e6,1),round(v['kernel_ms'],3)) for k,v in d['reads'].items() if isinstance(v,dict) and 'value' in v})"
done
