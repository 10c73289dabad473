"""The reference's scanner function tests (src/test/function_test/base_api/test_scan.cpp:140-389), restated against a
Python dict model: a hash key with many sort keys, plain hash keys, hash keys written with a TTL; the scans of the
reference test (ALL_SORT_KEY, BOUND_INCLUSIVE, BOUND_EXCLUSIVE, ONE_POINT, HALF_INCLUSIVE, VOID_SPAN, OVERALL,
OVERALL_COUNT_ONLY, REQUEST_EXPIRE_TS) must return exactly the model's slices.  The client side of a scanner
(start/stop raw keys, batches until the context completes) follows src/client_lib/pegasus_client_impl.cpp:1135-1192
and pegasus_scanner_impl.cpp; it lives in rrdb_harness.Backend.scan_all.

Runs on the CPU oracle (pins the oracle to the reference's expectations) and, marked gpu, on the CUDA engine."""
import random

import pytest

from rrdb_harness import Backend, next_blob, raw_key

NOW = 200_000_000
TTL = 24 * 60 * 60


def rand_str(rnd, lo=4, hi=24):
    return bytes(rnd.choice(b"abcdefghijklmnopqrstuvwxyz0123456789") for _ in range(rnd.randint(lo, hi)))


@pytest.fixture(scope="module", params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def filled(request):
    kind = request.param
    engine = request.getfixturevalue("engine") if kind == "gpu" else None
    be = Backend(kind, engine, opts={"l0_compaction_trigger": 4})
    be.set_partition_version(0)  # a table with one partition: partition_version = partition_count - 1
    rnd = random.Random(20260922)
    model, ttl_model = {}, {}
    big = rand_str(rnd)
    model[big] = {}
    while len(model[big]) < 400:
        model[big][rand_str(rnd)] = rand_str(rnd, 1, 60)
    while len(model) < 61:
        hk = rand_str(rnd)
        if hk in model:
            continue
        model[hk] = {rand_str(rnd): rand_str(rnd, 1, 60) for _ in range(10)}
    plain = dict(model)
    while len(model) < 121:
        hk = rand_str(rnd)
        if hk in model:
            continue
        model[hk] = {rand_str(rnd): rand_str(rnd, 1, 60) for _ in range(10)}
        ttl_model[hk] = {sk: (v, NOW + TTL) for sk, v in model[hk].items()}
    # several flushed runs + a live memtable, like a replica that has been running for a while
    order = list(model)
    rnd.shuffle(order)
    for i, hk in enumerate(order):
        items = list(model[hk].items())
        for j in range(0, len(items), 100):
            be.multi_put(hk, dict(items[j:j + 100]), expire_ts=(NOW + TTL) if hk in ttl_model else 0, now=NOW)
        if i % 30 == 29:
            be.flush(NOW)
    yield be, model, ttl_model, big, rnd
    be.close()


def scan_hk(be, hk, start=b"", stop=b"", **kw):
    # the client never sends an empty range (pegasus_client_impl.cpp:1182-1188): the scanner completes at once;
    # the server answers such a request with kOk and no rows (pegasus_server_impl.cpp:1225-1241)
    a, b = raw_key(hk, start), (raw_key(hk, stop) if stop else next_blob(raw_key(hk, b"")))
    si, ti = kw.get("start_inclusive", True), (kw.get("stop_inclusive", False) if stop else False)
    if a > b or (a == b and not (si and ti)):
        r = be.get_scanner(a, b, start_inclusive=si, stop_inclusive=ti, batch_size=37, now=NOW)
        assert r["error"] == 0 and r["kvs"] == []
        return {}
    kvs, batches = be.scan_all(hk, start_sk=start, stop_sk=stop, batch_size=37, now=NOW, **kw)
    assert batches[-1]["error"] == 0 and batches[-1]["context_id"] == -1  # PERR_SCAN_COMPLETE
    out = {}
    for k, v, _ in kvs:
        hl = int.from_bytes(k[:2], "big")
        assert k[2:2 + hl] == hk
        sk = k[2 + hl:]
        assert sk not in out  # check_and_put: no duplicates
        out[sk] = v
    return out


def test_all_sort_key(filled):
    be, model, _, big, _ = filled
    assert scan_hk(be, big) == model[big]


def test_bounds(filled):
    be, model, _, big, rnd = filled
    keys = sorted(model[big])
    for _ in range(6):
        i1 = rnd.randrange(200)
        i2 = i1 + rnd.randrange(150) + 20
        start, stop = keys[i1], keys[i2]
        got = scan_hk(be, big, start, stop, start_inclusive=True, stop_inclusive=True)      # BOUND_INCLUSIVE
        assert got == {k: model[big][k] for k in keys[i1:i2 + 1]}
        got = scan_hk(be, big, start, stop, start_inclusive=False, stop_inclusive=False)    # BOUND_EXCLUSIVE
        assert got == {k: model[big][k] for k in keys[i1 + 1:i2]}
        assert scan_hk(be, big, start, start, start_inclusive=True, stop_inclusive=True) == {start: model[big][start]}  # ONE_POINT
        assert scan_hk(be, big, start, start, start_inclusive=True, stop_inclusive=False) == {}  # HALF_INCLUSIVE
        assert scan_hk(be, big, stop, start, start_inclusive=True, stop_inclusive=True) == {}    # VOID_SPAN


def full_scan(be, **kw):
    """an unordered scanner over the whole partition (pegasus_client_impl.cpp:1221-1237): start "\x00\x00" inclusive, stop
    "\xff\xff" exclusive, full_scan and partition-hash validation on (pegasus_scanner_impl.cpp:60-72,413-416)"""
    r = be.get_scanner(b"\x00\x00", b"\xff\xff", start_inclusive=True, stop_inclusive=False, full_scan=True,
                       validate_partition_hash=True, batch_size=100, now=NOW, **kw)
    out, count, guard = list(r["kvs"]), max(r["kv_count"], 0), 0
    while r["error"] == 0 and r["context_id"] >= 0 and guard < 10000:
        r = be.scan(r["context_id"], now=NOW)
        out += r["kvs"]
        count += max(r["kv_count"], 0)
        guard += 1
    assert r["error"] == 0 and r["context_id"] == -1
    return out, count


def test_overall(filled):
    be, model, _, _, _ = filled
    kvs, _ = full_scan(be)
    data = {}
    for k, v, _ in kvs:
        hl = int.from_bytes(k[:2], "big")
        data.setdefault(k[2:2 + hl], {})
        assert k[2 + hl:] not in data[k[2:2 + hl]]
        data[k[2:2 + hl]][k[2 + hl:]] = v
    assert data == model


def test_overall_count_only(filled):
    be, model, _, _, _ = filled
    kvs, count = full_scan(be, only_return_count=True)
    assert kvs == []
    assert count == sum(len(m) for m in model.values())


def test_request_expire_ts(filled):
    be, model, ttl_model, _, _ = filled
    kvs, _ = full_scan(be, return_expire_ts=True)
    data, ttl_data = {}, {}
    for k, v, ets in kvs:
        hl = int.from_bytes(k[:2], "big")
        hk, sk = k[2:2 + hl], k[2 + hl:]
        data.setdefault(hk, {})[sk] = v
        if ets > 0:
            ttl_data.setdefault(hk, {})[sk] = (v, ets)
    assert data == model
    assert ttl_data == ttl_model
