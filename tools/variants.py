#!/usr/bin/env python
"""Diagnostics: k_merge kernel time + per-phase cycle shares for a list of PGS_VARIANT / env settings.
Usage: python tools/variants.py [--ctas N] SPEC [SPEC ...]   where SPEC = "KEY=VAL,KEY=VAL" (e.g. PGS_VARIANT=1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (device init)
import incubator_pegasus_b200 as pgs
from incubator_pegasus_b200 import synth

args = sys.argv[1:]
ctas = 1
if args and args[0] == "--ctas":
    ctas = int(args[1]); args = args[2:]
n = int(os.environ.get("VAR_RECORDS", "2500000"))
runs = synth.compaction_runs(k=4, n_per_run=n, hk_len=16, sk_len=32, user_len=256, now=synth.NOW, seed=1000)
host_runs = [pgs.build_run(r) for r in runs]
eng = pgs.Engine(device=0, ctas_per_sm=ctas)
part = eng.partition(app_id=1, pidx=0)
ids = [part.upload(h) for h in host_runs]
for spec in args:
    saved = {}
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1); saved[k] = os.environ.get(k); os.environ[k] = v
    os.environ.pop("PGS_PHASE_TIMING", None)
    ms = []
    for i in range(7):
        res = part.compact(ids, out_level=1, bottommost=1, now=synth.NOW, enabled=True, flags=3)
        if i >= 3: ms.append(res.merge_kernel_ms)
    print(f"== {spec}: kernel_ms {sum(ms)/len(ms):.3f} (min {min(ms):.3f}) walk {res.walk_ms:.3f} emit {res.emit_ms:.3f} plan {res.device_ms - res.merge_kernel_ms:.3f} segs {res.n_tiles}", flush=True)
    os.environ["PGS_PHASE_TIMING"] = "1"
    sys.stderr.flush()
    part.compact(ids, out_level=1, bottommost=1, now=synth.NOW, enabled=True, flags=3)
    os.environ.pop("PGS_PHASE_TIMING", None)
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
