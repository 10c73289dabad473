#!/usr/bin/env python
"""Diagnostics: compaction kernel times (k_walk, k_emit, planner) at BASELINE config #2 for a list of environment settings.
Usage: python tools/variants.py SPEC [SPEC ...]   where SPEC = "default" or "KEY=VAL,KEY=VAL" (e.g. PGS_WALK_G=8)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: F401  (device init)
import incubator_pegasus_b200 as pgs
from incubator_pegasus_b200 import synth

args = sys.argv[1:]
n = int(os.environ.get("VAR_RECORDS", "2500000"))
runs = synth.compaction_runs(k=4, n_per_run=n, hk_len=16, sk_len=32, user_len=256, now=synth.NOW, seed=1000)
host_runs = [pgs.build_run(r) for r in runs]
eng = pgs.Engine(device=0)
part = eng.partition(app_id=1, pidx=0)
ids = [part.upload(h) for h in host_runs]
for spec in args:
    saved = {}
    for kv in spec.split(","):
        if "=" in kv:
            k, v = kv.split("=", 1); saved[k] = os.environ.get(k); os.environ[k] = v
    ms = []
    for i in range(7):
        res = part.compact(ids, out_level=1, bottommost=1, now=synth.NOW, enabled=True, flags=3)
        if i >= 3: ms.append(res.merge_kernel_ms)
    print(f"== {spec}: kernel_ms {sum(ms)/len(ms):.3f} (min {min(ms):.3f}) walk {res.walk_ms:.3f} emit {res.emit_ms:.3f} plan {res.device_ms - res.merge_kernel_ms:.3f} segs {res.n_tiles}", flush=True)
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
