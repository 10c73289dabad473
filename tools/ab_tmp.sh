#!/bin/bash
# temporary: pick the faster of two builds of the read kernels on this box, then run the profile round with it
set -u
O=gpurun_out; mkdir -p $O
M() { timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-sharded --skip-sweep --skip-ycsb 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith(chr(123))][-1]); r=d['reads']; print(r['scan']['kernel_ms'], r['get']['kernel_ms'])"; }
A=$(M); echo "lb8: $A" | tee $O/ab.txt
cp incubator_pegasus_b200/libpegasus_b200.so /tmp/lb8.so
cp incubator_pegasus_b200/libpegasus_b200_nolb.so incubator_pegasus_b200/libpegasus_b200.so
B=$(M); echo "nolb: $B" | tee -a $O/ab.txt
python - <<PY | tee -a $O/ab.txt
a=float("$A".split()[0]); b=float("$B".split()[0])
print("pick", "lb8" if a < b else "nolb")
open("/tmp/pick","w").write("lb8" if a < b else "nolb")
PY
if [ "$(cat /tmp/pick)" = "lb8" ]; then cp /tmp/lb8.so incubator_pegasus_b200/libpegasus_b200.so; fi
bash tools/profile_round.sh
