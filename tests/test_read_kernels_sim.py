"""CPU check of the READ kernels' logic (k_get, k_scan_fwd of incubator_pegasus_b200/csrc/read_kernels.cuh) inside the host SIMT
interpreter (tools/simt, see test_kernel_sim.py): random multi-version data over several runs, point lookups and forward
range scans with every flag of the request, compared with a plain Python model of RocksDB's visibility rules and of the
reference's iterator loop (src/server/pegasus_server_impl.cpp:617-756, 1266-1320).  Also: the Bloom filters built at upload and
by the compaction walker never reject a present key or hash-key prefix."""
import ctypes as C

import numpy as np
import pytest

from incubator_pegasus_b200 import synth
from test_kernel_sim import sim, sim_compact  # noqa: F401  (fixture + helper)

NOW = synth.NOW


def be(n, w):
    return int(n).to_bytes(w, "big")


def raw_key(hk, sk):
    return be(len(hk), 2) + hk + sk


def value(ets, data):
    return be(ets, 4) + be((1 << 8) | 2, 8) + data


def make_db(pgs, rng, n_runs, hashkeys, sort_per_hk, big=False):
    """n_runs runs, NEWEST FIRST (as Partition.runs): each write gets a global seq; a run holds a random subset of keys."""
    seq = 0
    runs_items = []
    for r in range(n_runs):  # oldest run first while generating
        items = {}
        for hk in hashkeys:
            for s in range(sort_per_hk):
                if rng.random() < 0.45:
                    continue
                sk = b"s%04d" % s
                for _ in range(int(rng.integers(1, 3))):
                    seq += 1
                    u = rng.random()
                    if u < 0.12:
                        items[(raw_key(hk, sk), -seq)] = (raw_key(hk, sk), seq, 0, b"")
                    else:
                        ets = 0 if u < 0.6 else (NOW + int(rng.integers(1, 1000)) if u < 0.85 else NOW - int(rng.integers(0, 1000)))
                        dl = int(rng.integers(0, 30)) if not big or rng.random() < 0.9 else int(rng.integers(600, 1500))
                        items[(raw_key(hk, sk), -seq)] = (raw_key(hk, sk), seq, 1, value(ets, bytes(rng.integers(0, 256, dl, dtype=np.uint8))))
        runs_items.append([items[k] for k in sorted(items)])
    runs_items.reverse()  # newest first
    return [pgs.Records.from_list(it) for it in runs_items if it], runs_items


def visible(runs_items):
    """newest version of every user key; tombstones hide the key.  -> sorted [(key, value)]"""
    best = {}
    for items in runs_items:
        for k, s, t, v in items:
            if k not in best or s > best[k][0]:
                best[k] = (s, t, v)
    return [(k, v) for k, (s, t, v) in sorted(best.items()) if t == 1], best


def run_args(pgs, runs, block_size=4096, ri=16):
    brs = [pgs.build_run(r, block_size, ri) for r in runs]
    k = len(brs)
    return (k, (C.c_void_p * k)(*[b.data.ctypes.data for b in brs]), (C.c_uint64 * k)(*[b.data.shape[0] for b in brs]),
            (C.c_void_p * k)(*[b.blk_off.ctypes.data for b in brs]), (C.c_void_p * k)(*[b.blk_size.ctypes.data for b in brs]),
            (C.c_uint32 * k)(*[b.n_blocks for b in brs])), brs


@pytest.mark.parametrize("use_bloom", [1, 0])
def test_sim_get(pgs, sim, use_bloom):
    rng = np.random.default_rng(5)
    hks = [bytes(rng.integers(0, 256, int(rng.integers(1, 9)), dtype=np.uint8)) for _ in range(12)] + [b""]
    runs, items = make_db(pgs, rng, 4, hks, 40, big=True)
    vis, best = visible(items)
    args, keep = run_args(pgs, runs, block_size=1024)
    present = list(best.keys())
    absent = [raw_key(h, b"s%04d" % s) for h in hks[:4] for s in (41, 77)] + [raw_key(b"zz", b""), b"", b"\x00", b"\x00\x05ab", b"\xff" * 300]
    keys = [present[i] for i in rng.permutation(len(present))[:300]] + absent
    flat = np.frombuffer(b"".join(keys), np.uint8).copy()
    off = np.zeros(len(keys) + 1, np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    res = (pgs.GetResult * len(keys))()
    arena = np.zeros(1 << 20, np.uint8)
    stats = (C.c_uint64 * 3)()
    st = sim.sim_get(*args, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(keys), NOW,
                     arena.ctypes.data_as(C.c_void_p), C.c_uint64(arena.shape[0]), res, stats, use_bloom)
    assert st == 0
    for i, k in enumerate(keys):
        r = res[i]
        if k not in best or best[k][1] == 0:
            assert r.status == pgs.NOT_FOUND and not r.expired, (i, k)
            continue
        v = best[k][2]
        ets = int.from_bytes(v[:4], "big")
        assert r.expire_ts == ets
        if 0 < ets <= NOW:
            assert r.status == pgs.NOT_FOUND and r.expired
        else:
            assert r.status == pgs.OK and arena[r.value_off:r.value_off + r.value_len].tobytes() == v[12:], (i, k)
    if use_bloom:
        assert stats[2] > 0                                     # some run probes were saved ...
    else:
        assert stats[2] == 0
    print("probes", stats[1], "skipped", stats[2])


def test_sim_get_multi_partition(pgs, sim):
    """pgs_get_batch_multi's kernel shape: one launch, every key looked up in the runs of its own partition slot only"""
    rng = np.random.default_rng(15)
    hks = [bytes(rng.integers(0, 256, int(rng.integers(1, 6)), dtype=np.uint8)) for _ in range(8)]
    runs, items = make_db(pgs, rng, 4, hks, 30)
    per_run = {}
    for key, seq, typ, val, run in [(k, s, t, v, r) for r, rr in enumerate(runs) for k, s, t, v in [(rr.key(i), int(rr.seq[i]), int(rr.type[i]), rr.value(i)) for i in range(rr.n)]]:
        per_run.setdefault(run, []).append((key, seq, typ, val))
    def best_of(run_ids):
        best = {}
        for r in run_ids:
            for k, s, t, v in per_run.get(r, []):
                if k not in best or s > best[k][0]:
                    best[k] = (s, t, v)
        return best
    slots = [best_of([0, 1]), best_of([2, 3]), {}]
    allkeys = sorted(set(k for b in slots for k in b))
    keys = [allkeys[i] for i in rng.permutation(len(allkeys))[:240]] + [b"", b"\x00\x09nothing"]
    flat = np.frombuffer(b"".join(keys), np.uint8).copy()
    off = np.zeros(len(keys) + 1, np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    args, keep = run_args(pgs, runs, block_size=1024)
    res = (pgs.GetResult * len(keys))()
    arena = np.zeros(1 << 20, np.uint8)
    stats = (C.c_uint64 * 3)()
    assert sim.sim_get(*args, flat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(keys), NOW,
                       arena.ctypes.data_as(C.c_void_p), C.c_uint64(arena.shape[0]), res, stats, 3) == 0
    hits = 0
    for i, k in enumerate(keys):
        best, r = slots[i % 3], res[i]
        if k not in best or best[k][1] == 0:
            assert r.status == pgs.NOT_FOUND and not r.expired, (i, k)
            continue
        v = best[k][2]
        ets = int.from_bytes(v[:4], "big")
        if 0 < ets <= NOW:
            assert r.status == pgs.NOT_FOUND and r.expired
        else:
            hits += 1
            assert r.status == pgs.OK and arena[r.value_off:r.value_off + r.value_len].tobytes() == v[12:], (i, k)
    assert hits > 30


def model_scan(vis, q, now):
    """the reference loop over the visible records; returns dict like pgs_scan_result + kvs"""
    start, stop = q["start"], q["stop"]
    pos = 0
    while pos < len(vis) and vis[pos][0] < start:
        pos += 1
    prefix = None
    if q.get("prefix") and len(start) >= 2:
        hl = int.from_bytes(start[:2], "big")
        if 2 + hl <= len(start):
            prefix = start[:2 + hl]

    def valid(p):
        if p >= len(vis):
            return False
        k = vis[p][0]
        if prefix is not None and k[:len(prefix)] != prefix:
            return False
        if q.get("has_upper") and k >= stop:
            return False
        return True
    count = it = exp = fil = 0
    size = 0
    kvs = []
    complete = False
    first_excl = not q["start_inclusive"]
    while count < q["max_count"] and it < q["max_iter_count"] and not (q["max_iter_size"] > 0 and size >= q["max_iter_size"]) and valid(pos):
        k, v = vis[pos]
        if k > stop or (k == stop and not q["stop_inclusive"]):
            complete = True
            break
        if first_excl:
            first_excl = False
            if k == start:
                pos += 1
                continue
        it += 1
        ets = int.from_bytes(v[:4], "big")
        hl = int.from_bytes(k[:2], "big")
        hk, sk = k[2:2 + hl], k[2 + hl:]

        def match(ft, pat, s):
            if ft == 0 or not pat:
                return True
            return {1: pat in s, 2: s.startswith(pat), 3: s.endswith(pat)}[ft]
        if 0 < ets <= now:
            exp += 1
        elif not match(q.get("hft", 0), q.get("hpat", b""), hk) or not match(q.get("sft", 0), q.get("spat", b""), sk):
            fil += 1
        else:
            ko = sk if q["key_mode"] == 1 else k
            vo = b"" if q.get("no_value") else v[12:]
            count += 1
            size += len(ko) + len(vo)
            if not q.get("count_only"):
                kvs.append((ko, vo, ets if q.get("return_expire_ts") else 0))
        if k == stop:
            complete = True
            break
        pos += 1
    iv = valid(pos)
    return dict(count=count, iter_count=it, expire_count=exp, filter_count=fil, size=size, complete=complete, iter_valid=iv,
                resume=vis[pos][0] if iv and not complete else None, kvs=kvs)


def do_scans(pgs, sim, args, reqs, lanes=0):
    n = len(reqs)
    keep = []

    def blob(b):
        buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if b else b"\0")
        keep.append(buf)
        return pgs.Blob(C.cast(buf, C.POINTER(C.c_uint8)), len(b))
    arr = (pgs.ScanRequest * n)()
    for i, q in enumerate(reqs):
        r = arr[i]
        r.start, r.stop = blob(q["start"]), blob(q["stop"])
        r.start_inclusive, r.stop_inclusive = int(q["start_inclusive"]), int(q["stop_inclusive"])
        r.no_value, r.key_mode, r.return_expire_ts = int(q.get("no_value", 0)), q["key_mode"], int(q.get("return_expire_ts", 0))
        r.count_only, r.prefix_same_as_start = int(q.get("count_only", 0)), int(q.get("prefix", 0))
        r.reserved[0] = int(q.get("has_upper", 0))
        r.hash_filter_type, r.sort_filter_type = q.get("hft", 0), q.get("sft", 0)
        r.hash_filter, r.sort_filter = blob(q.get("hpat", b"")), blob(q.get("spat", b""))
        r.max_count, r.max_iter_count, r.max_iter_size = q["max_count"], q["max_iter_count"], q["max_iter_size"]
    astride, kstride, rstride = 1 << 16, 512, 512
    arena = np.zeros(astride * n, np.uint8)
    kvs = (pgs.KV * (kstride * n))()
    resume = np.zeros(rstride * n, np.uint8)
    res = (pgs.ScanResult * n)()
    st = sim.sim_scan(*args, arr, n, NOW, C.c_uint64(astride), kstride, arena.ctypes.data_as(C.c_void_p), kvs,
                      resume.ctypes.data_as(C.c_void_p), rstride, res, lanes)
    assert st == 0, st
    out = []
    for i in range(n):
        r = res[i]
        a = arena[i * astride:(i + 1) * astride]
        recs = [(a[kv.key_off:kv.key_off + kv.key_len].tobytes(), a[kv.value_off:kv.value_off + kv.value_len].tobytes(), kv.expire_ts)
                for kv in kvs[i * kstride:i * kstride + r.n_kvs]]
        out.append(dict(count=r.count, iter_count=r.iter_count, expire_count=r.expire_count, filter_count=r.filter_count, size=r.size,
                        complete=bool(r.complete), iter_valid=bool(r.iter_valid),
                        resume=resume[i * rstride:i * rstride + r.resume_len].tobytes() if r.iter_valid and not r.complete else None, kvs=recs))
    return out


@pytest.mark.parametrize("n_runs,lanes", [(4, 0), (1, 0), (6, 16), (3, 32)])
def test_sim_scan_forward(pgs, sim, n_runs, lanes):
    rng = np.random.default_rng(100 + n_runs)
    hks = [b"h%d" % i for i in range(7)] + [b"", b"h1x", bytes([0xff, 0xff])]
    runs, items = make_db(pgs, rng, n_runs, hks, 30)
    vis, best = visible(items)
    args, keep = run_args(pgs, runs, block_size=512, ri=4)
    reqs = []
    for hk in hks + [b"nope"]:
        lo, hi = raw_key(hk, b""), raw_key(hk, b"\xff" * 8)
        nxt = bytearray(raw_key(hk, b""))
        while nxt and nxt[-1] == 0xff:
            nxt.pop()
        nxt[-1] += 1
        nxt = bytes(nxt)
        base = dict(start=lo, stop=nxt, start_inclusive=True, stop_inclusive=False, key_mode=1, prefix=1,
                    max_count=3000, max_iter_count=3000, max_iter_size=0)
        reqs.append(base)                                                               # multi_get: whole hash key
        reqs.append(dict(base, max_count=7))                                            # count limit
        reqs.append(dict(base, max_iter_count=9))                                       # iteration limit
        reqs.append(dict(base, max_iter_size=100))                                      # size limit
        reqs.append(dict(base, start=raw_key(hk, b"s0010"), stop=raw_key(hk, b"s0020"), stop_inclusive=True))
        reqs.append(dict(base, start=raw_key(hk, b"s0010"), start_inclusive=False, stop=raw_key(hk, b"s0010"), stop_inclusive=True))
        reqs.append(dict(base, start=raw_key(hk, b"s0005"), start_inclusive=False, no_value=1))
        reqs.append(dict(base, sft=3, spat=b"7"))                                       # sort-key postfix filter
        reqs.append(dict(base, sft=1, spat=b"01", count_only=1))
        reqs.append(dict(base, has_upper=1, count_only=1, max_count=2**32 - 1, max_iter_count=2**32 - 1))  # sortkey_count
        reqs.append(dict(base, key_mode=0, prefix=0, stop=hi, hft=2, hpat=hk[:1], return_expire_ts=1, max_count=11))  # scanner batch
    reqs.append(dict(start=b"", stop=b"\xff\xff\xff", start_inclusive=True, stop_inclusive=True, key_mode=0, prefix=0,
                     max_count=100000, max_iter_count=100000, max_iter_size=0))         # full table
    reqs.append(dict(start=b"\x00\x02h", stop=b"\x00\x02h5", start_inclusive=True, stop_inclusive=False, key_mode=0, prefix=0,
                     max_count=40, max_iter_count=1000, max_iter_size=0))
    got = do_scans(pgs, sim, args, reqs, lanes)
    for i, (q, g_) in enumerate(zip(reqs, got)):
        want = model_scan(vis, q, NOW)
        assert g_ == want, (i, q, {k: (g_[k], want[k]) for k in want if g_[k] != want[k]})


@pytest.mark.parametrize("n_runs,lanes", [(4, 0), (5, 16)])
def test_sim_scan_multi_partition(pgs, sim, n_runs, lanes):
    """pgs_range_scan_many_multi's kernel shape: one launch, every request merges the runs of its own partition slot only"""
    rng = np.random.default_rng(300 + n_runs)
    hks = [b"h%d" % i for i in range(5)] + [b"", bytes([0xff, 0xff])]
    runs, items = make_db(pgs, rng, n_runs, hks, 30, big=True)  # values of 600..1500 bytes among them: several copy rounds
    assert len(runs) == n_runs
    half = n_runs // 2
    vis_of = [visible(items[:half])[0], visible(items[half:])[0], []]
    args, keep = run_args(pgs, runs, block_size=512, ri=4)
    reqs = []
    for hk in hks + [b"nope"]:
        nxt = bytearray(raw_key(hk, b""))
        while nxt and nxt[-1] == 0xff:
            nxt.pop()
        nxt[-1] += 1
        base = dict(start=raw_key(hk, b""), stop=bytes(nxt), start_inclusive=True, stop_inclusive=False, key_mode=1, prefix=1,
                    max_count=3000, max_iter_count=3000, max_iter_size=0)
        for q in (base, dict(base, max_count=7), dict(base, start=raw_key(hk, b"s0010"), stop=raw_key(hk, b"s0020"), stop_inclusive=True),
                  dict(base, sft=1, spat=b"01", count_only=1),
                  dict(base, key_mode=0, prefix=0, stop=raw_key(hk, b"\xff" * 8), return_expire_ts=1, max_count=11)):
            reqs += [q, q, q]   # the same request against each of the three slots
    got = do_scans(pgs, sim, args, reqs, lanes | 0x100)
    nonempty = 0
    for i, (q, g_) in enumerate(zip(reqs, got)):
        want = model_scan(vis_of[i % 3], q, NOW)
        nonempty += want["count"] > 0
        assert g_ == want, (i, i % 3, q, {k: (g_[k], want[k]) for k in want if g_[k] != want[k]})
    assert nonempty > 20


def test_sim_bloom_no_false_negatives(pgs, sim):
    rng = np.random.default_rng(9)
    runs = synth.compaction_runs(k=3, n_per_run=400, seed=9)
    out, x = sim_compact(pgs, sim, runs, bottommost=True, seg_weight=32 * 1024)
    got = pgs.decode_blocks(out)
    miss = 0
    for i in range(got.n):
        k = got.key(i)
        assert sim.sim_result_bloom_check(k, len(k)) == 1
        pl = 2 + int.from_bytes(k[:2], "big")
        assert sim.sim_result_bloom_check(k[:pl], pl) == 1
    for _ in range(2000):
        k = bytes(rng.integers(0, 256, 50, dtype=np.uint8))
        miss += sim.sim_result_bloom_check(k, len(k))
    assert miss < 200  # ~1 % expected at 10 bits per entry (the filter is sized for twice the entries here)
