"""pgs_batcher_get on the device: many host threads read through one batcher; every answer equals the answer of a direct
pgs_get_batch on the key's partition, and the launches were shared.  (Runs last in the GPU suite: it is the newest path.)"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from incubator_pegasus_b200 import synth

pytestmark = pytest.mark.gpu


def test_batched_gets_match_direct_gets(pgs, engine):
    rng = np.random.default_rng(21)
    parts, keysets = [], []
    try:
        for p, n_runs in enumerate((3, 1, 0, 4)):  # partition 2 holds nothing
            part = engine.partition(app_id=9, pidx=p)
            parts.append(part)
            runs = synth.compaction_runs(k=max(1, n_runs), n_per_run=2000, seed=90 + p)[:n_runs]
            for r in runs:
                part.upload_records(r)
            ks = [r.key(int(i)) for r in runs for i in rng.integers(0, r.n, 60)]
            keysets.append(ks + [b"\x00\x06absent" + bytes([p])])
        work = [(slot, k) for slot in range(len(parts)) for q in range(len(parts)) for k in keysets[q][:50]]
        order = rng.permutation(len(work))
        work = [work[i] for i in order]
        want = {}
        for slot, part in enumerate(parts):
            ks = sorted({k for s, k in work if s == slot})
            flat = np.frombuffer(b"".join(ks), np.uint8).copy()
            off = np.zeros(len(ks) + 1, np.uint32)
            off[1:] = np.cumsum([len(k) for k in ks])
            st, res, arena, _ = part.get_batch(flat, off, synth.NOW)
            assert st == 0
            for i, k in enumerate(ks):
                r = res[i]
                want[(slot, k)] = (r.status, r.expire_ts, r.expired, arena[r.value_off:r.value_off + r.value_len].tobytes() if r.status == pgs.OK else None)
        b = pgs.Batcher(parts, max_batch=64, max_wait_us=2000)
        try:
            def one(item):
                slot, k = item
                st, r, v = b.get(slot, k, synth.NOW, cap=1024)
                return (slot, k), (st, r.status, r.expire_ts, r.expired, v)
            with ThreadPoolExecutor(max_workers=32) as pool:
                got = list(pool.map(one, work))
            hits = 0
            for key, (st, status, ets, expired, v) in got:
                assert st == 0
                assert (status, ets, expired, v) == want[key], key
                hits += status == pgs.OK
            assert hits > 100
            requests, launches = b.stats()
            assert requests == len(work) and 0 < launches < requests  # windows were shared
            # a value that does not fit the caller's buffer: the length comes back
            slot, k = next((s, k) for (s, k), w in want.items() if w[0] == pgs.OK and len(w[3]) > 8)
            st, r, v = b.get(slot, k, synth.NOW, cap=8)
            assert st == 0 and r.status == pgs.INCOMPLETE and r.value_len == len(want[(slot, k)][3]) and v is None
            assert lib_invalid(pgs, b)
        finally:
            b.close()
    finally:
        for part in parts:
            part.close()


def lib_invalid(pgs, b):
    import ctypes as C
    r = pgs.GetResult()
    return pgs.lib().pgs_batcher_get(b.h, 99, b"k", 1, 0, None, 0, C.byref(r)) == pgs.INVALID_ARGUMENT
