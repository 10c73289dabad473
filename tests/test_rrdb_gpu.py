"""GPU parity of the rrdb operator surface (on_get / on_multi_get / on_batch_get / on_sortkey_count / on_ttl /
on_get_scanner / on_scan through pgs_rrdb_*, kernels k_get / k_scan_fwd / k_scan / k_walk / k_emit) against the CPU oracle and
the reference's own golden tables."""
import json
import os
import random

import numpy as np
import pytest

from rrdb_harness import Backend, next_blob, raw_key, same_response

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
T = json.load(open(os.path.join(GOLD, "tables.json")))
MG = json.load(open(os.path.join(GOLD, "multi_get_basic.json")))
NOW = 200_000_000


def b(s):
    return s.encode("latin-1")


@pytest.fixture()
def pair(engine):
    g, o = Backend("gpu", engine), Backend("oracle")
    yield g, o
    g.close()
    o.close()


def both(pair, fn, *a, **kw):
    rg = getattr(pair[0], fn)(*a, **kw)
    ro = getattr(pair[1], fn)(*a, **kw)
    return rg, ro


def check(pair, fn, *a, **kw):
    rg, ro = both(pair, fn, *a, **kw)
    ok, detail = same_response(rg, ro)
    assert ok, (fn, a, kw, detail)
    return rg


@pytest.mark.parametrize("direction", ["forward", "reverse"])
def test_multi_get_basic_tables(pair, direction):
    g = MG[direction]
    hk = b"basic_test_multi_get"
    for be in pair:
        assert be.multi_put(hk, {b(k): b(v) for k, v in g["fixture"]}) == 0
    assert check(pair, "sortkey_count", hk)["count"] == 13
    for case in g["cases"]:
        o = case["options"]
        r = check(pair, "multi_get", hk, b(case["start"]), b(case["stop"]), o["start_inclusive"], o["stop_inclusive"],
                  max_kv_count=case["max_count"], max_kv_size=1000000, reverse=o["reverse"],
                  filter_type=o["sort_key_filter_type"], filter_pattern=b(o["sort_key_filter_pattern"]))
        assert r["error"] == case["error"], case["title"]
        assert sorted([k.decode("latin-1"), v.decode("latin-1")] for k, v, _ in r["kvs"]) == case["expect"], case["title"]
    for be in pair:
        be.put(hk, b"", b"expire_value", expire_ts=NOW - 1, now=NOW - 5)
    r = check(pair, "multi_get", hk, max_kv_count=2, now=NOW)
    assert r["error"] == 7 and [k for k, _, _ in r["kvs"]] == [b"1", b"1-abcdefg"]
    for be in pair:
        assert be.multi_remove(hk, [b(k) for k, _ in g["fixture"]]) == (0, 13)
    assert check(pair, "sortkey_count", hk)["count"] == 0


def _prepare_range_read(pair, total, expired):
    hk = b"range_read_hashkey"
    for be in pair:
        if expired:
            be.multi_put(hk, {b"1-%d" % i: b"value" for i in range(expired)}, expire_ts=NOW - 10, now=NOW - 20)
        if total > expired:
            be.multi_put(hk, {b"2-%d" % i: b"value" for i in range(expired, total)})
    return hk


@pytest.mark.parametrize("row", T["range_read_multiget"]["rows"])
def test_range_read_multiget_table(engine, row):
    exp, total, max_count, want_err, want_n = row
    pair = (Backend("gpu", engine), Backend("oracle"))
    try:
        hk = _prepare_range_read(pair, total, exp)
        r = check(pair, "multi_get", hk, max_kv_count=max_count, max_kv_size=1000000, now=NOW)
        assert (r["error"], len(r["kvs"])) == (want_err, want_n)
    finally:
        for be in pair:
            be.close()


@pytest.mark.parametrize("row", T["range_read_scan"]["rows"])
def test_range_read_scan_table(engine, row):
    exp, total, batch, _want = row
    pair = (Backend("gpu", engine), Backend("oracle"))
    try:
        hk = _prepare_range_read(pair, total, exp)
        (kg, bg), (ko, bo) = both(pair, "scan_all", hk, batch_size=batch, now=NOW)
        assert kg == ko and len(kg) == total - exp
        assert len(bg) == len(bo)
        for x, y in zip(bg, bo):
            ok, d = same_response(x, y)
            assert ok, d
        r = check(pair, "sortkey_count", hk, now=NOW)
        assert r["count"] == total - exp
    finally:
        for be in pair:
            be.close()


def test_get_ttl_batch_get(pair):
    for be in pair:
        be.put(b"h", b"s1", b"v1")
        be.put(b"h", b"s2", b"v2", expire_ts=NOW + 100)
        be.put(b"h", b"s3", b"v3", expire_ts=NOW - 100, now=NOW - 200)
        be.put(b"h", b"s4", b"")
        be.put(b"", b"only_sort", b"x")
        be.remove(b"h", b"s1")
        be.put(b"h", b"s1", b"v1-new")
        be.put(b"h2", b"s", b"gone")
        be.flush(NOW)
        be.remove(b"h2", b"s")
    for hk, sk in [(b"h", b"s1"), (b"h", b"s2"), (b"h", b"s3"), (b"h", b"s4"), (b"h", b"nope"), (b"h2", b"s"), (b"", b"only_sort"), (b"zz", b"")]:
        check(pair, "get", hk, sk, now=NOW)
        check(pair, "ttl", hk, sk, now=NOW)
    r = check(pair, "get", b"h", b"s1", now=NOW)
    assert r["error"] == 0 and r["kvs"][0][1] == b"v1-new"
    assert check(pair, "ttl", b"h", b"s2", now=NOW)["ttl"] == 100
    assert check(pair, "ttl", b"h", b"s1", now=NOW)["ttl"] == -1
    assert check(pair, "get", b"h", b"s3", now=NOW)["error"] == 1
    r = check(pair, "batch_get", [(b"h", b"s1"), (b"h", b"s3"), (b"x", b"y"), (b"h", b"s4"), (b"", b"only_sort")], now=NOW)
    assert len(r["kvs"]) == 3
    assert check(pair, "batch_get", [], now=NOW)["error"] == 4
    r = check(pair, "multi_get", b"h", sort_keys=[b"s1", b"s2", b"s3", b"zz", b"s4"], now=NOW)
    assert [k for k, _, _ in r["kvs"]] == [b"s1", b"s2", b"s4"]
    check(pair, "multi_get", b"h", sort_keys=[b"s1", b"s2", b"s4"], max_kv_count=2, now=NOW)
    check(pair, "multi_get", b"h", sort_keys=[b"s1", b"s2", b"s4"], no_value=True, now=NOW)
    assert check(pair, "multi_get", b"h", filter_type=9, now=NOW)["error"] == 4
    assert both(pair, "multi_put", b"h", {})[0] == 4  # empty kvs -> kInvalidArgument, empty record written
    for be in pair:
        be.flush(NOW)
    check(pair, "multi_get", b"h", now=NOW)


def test_scanner_filters_and_flags(pair):
    hks = [b"a", b"ab", b"b", b"user_1", b"user_2", b"zz"]
    for be in pair:
        for hk in hks:
            be.multi_put(hk, {b"k%02d" % i: b"v" * (i + 1) for i in range(12)}, expire_ts=0)
            be.put(hk, b"tmp", b"x", expire_ts=NOW + 77)
        be.flush(NOW)
        be.remove(b"ab", b"k03")
    full_start, full_stop = raw_key(b"", b""), b"\xff\xff"
    # drive both scanners batch by batch
    for kw in [dict(), dict(batch_size=5), dict(batch_size=1), dict(no_value=True, batch_size=7), dict(return_expire_ts=True, batch_size=50),
               dict(only_return_count=True, batch_size=9), dict(hash_filter=(2, b"user"), batch_size=4), dict(hash_filter=(1, b"b")),
               dict(sort_filter=(3, b"1"), batch_size=3), dict(sort_filter=(2, b"k1")), dict(hash_filter=(3, b"_2"), sort_filter=(1, b"0"))]:
        rg, ro = both(pair, "get_scanner", full_start, full_stop, full_scan=True, now=NOW, **kw)
        n = 0
        while True:
            ok, d = same_response(rg, ro)
            assert ok, (kw, n, d)
            if rg["context_id"] < 0 or rg["error"] != 0:
                break
            rg, ro = pair[0].scan(rg["context_id"], now=NOW), pair[1].scan(ro["context_id"], now=NOW)
            n += 1
            assert n < 500
    # hash-key scoped scans use the prefix iterator
    for hk in hks:
        (kg, _), (ko, _) = both(pair, "scan_all", hk, batch_size=4, now=NOW)
        assert kg == ko and len(kg) == (12 if hk == b"ab" else 13)
        (kg, _), (ko, _) = both(pair, "scan_all", hk, start_sk=b"k03", stop_sk=b"k09", batch_size=100, now=NOW)
        assert kg == ko
    # unknown context / cleared context
    assert check(pair, "scan", 12345, now=NOW)["error"] == 1
    rg, ro = both(pair, "get_scanner", full_start, full_stop, full_scan=True, batch_size=2, now=NOW)
    pair[0].clear_scanner(rg["context_id"])
    pair[1].clear_scanner(ro["context_id"])
    assert pair[0].scan(rg["context_id"], now=NOW)["error"] == 1 and pair[1].scan(ro["context_id"], now=NOW)["error"] == 1
    # empty ranges and unsupported filter
    check(pair, "get_scanner", raw_key(b"b", b""), raw_key(b"a", b""), now=NOW)
    check(pair, "get_scanner", raw_key(b"a", b"x"), raw_key(b"a", b"x"), start_inclusive=True, stop_inclusive=False, now=NOW)
    assert check(pair, "get_scanner", full_start, full_stop, hash_filter=(7, b"x"), now=NOW)["error"] == 4


def test_scan_context_is_a_snapshot(pair):
    for be in pair:
        be.multi_put(b"snap", {b"%03d" % i: b"v%d" % i for i in range(30)})
    rg, ro = both(pair, "get_scanner", raw_key(b"snap", b""), next_blob(raw_key(b"snap", b"")), batch_size=10, now=NOW)
    for be in pair:  # mutate after the scanner was opened
        be.multi_remove(b"snap", [b"%03d" % i for i in range(10, 20)])
        be.put(b"snap", b"015x", b"new")
        be.manual_compact(NOW)
    seen_g, seen_o = list(rg["kvs"]), list(ro["kvs"])
    while rg["context_id"] >= 0:
        rg, ro = pair[0].scan(rg["context_id"], now=NOW), pair[1].scan(ro["context_id"], now=NOW)
        seen_g += rg["kvs"]
        seen_o += ro["kvs"]
    assert seen_g == seen_o and len(seen_g) == 30
    check(pair, "multi_get", b"snap", now=NOW)


def test_ttl_default_ttl_and_manual_compact(pair):
    """shape of src/test/function_test/base_api/test_ttl.cpp:79-205: default_ttl rewrites TTL-less records on
    manual compaction, expired ones disappear, explicit TTLs survive."""
    for be in pair:
        be.put(b"ttl", b"no_ttl", b"a")
        be.put(b"ttl", b"with_ttl", b"b", expire_ts=NOW + 1000)
        be.put(b"ttl", b"dead", b"c", expire_ts=NOW - 5, now=NOW - 10)
        be.flush(NOW)
    assert check(pair, "ttl", b"ttl", b"no_ttl", now=NOW)["ttl"] == -1
    for be in pair:
        be.update_envs({"default_ttl": "500"})
        be.manual_compact(NOW)
    assert check(pair, "ttl", b"ttl", b"no_ttl", now=NOW)["ttl"] == 500
    assert check(pair, "ttl", b"ttl", b"with_ttl", now=NOW)["ttl"] == 1000
    assert check(pair, "get", b"ttl", b"dead", now=NOW)["error"] == 1
    for be in pair:
        be.put(b"ttl", b"later", b"d", now=NOW)  # db_expire_ts: default ttl applied at write time
    assert check(pair, "ttl", b"ttl", b"later", now=NOW)["ttl"] == 500
    check(pair, "multi_get", b"ttl", now=NOW + 600)
    (kg, _), (ko, _) = both(pair, "scan_all", b"ttl", now=NOW + 600, return_expire_ts=True)
    assert kg == ko


def test_user_specified_compaction_and_split_validation(pair):
    ops = T["create_operations"]["json"]
    for be in pair:
        for hk in (b"hashkey_1", b"xhashkey", b"other"):
            be.multi_put(hk, {b"a_sortkey": b"1", b"sortkey_b": b"2", b"plain": b"3"}, expire_ts=NOW + 100)
        be.update_envs({"user_specified_compaction": ops})
        be.manual_compact(NOW)
    for hk in (b"hashkey_1", b"xhashkey", b"other"):
        check(pair, "multi_get", hk, now=NOW)
        (kg, _), (ko, _) = both(pair, "scan_all", hk, return_expire_ts=True, now=NOW)
        assert kg == ko
    assert check(pair, "sortkey_count", b"hashkey_1", now=NOW)["count"] == 0  # COT_DELETE on prefix "hashkey"
    # partition split: stale half hidden by scans (validate hash) and removed by compaction
    for be in pair:
        be.update_envs({"user_specified_compaction": "", "replica.split.validate_partition_hash": "true"})
        be.set_partition_version(1)
    full_start, full_stop = raw_key(b"", b""), b"\xff\xff"
    rg, ro = both(pair, "get_scanner", full_start, full_stop, full_scan=True, batch_size=1000, now=NOW)
    ok, d = same_response(rg, ro)
    assert ok, d
    rg2, ro2 = both(pair, "get_scanner", full_start, full_stop, full_scan=True, batch_size=1000, validate_partition_hash=False, now=NOW)
    assert same_response(rg2, ro2)[0] and len(rg2["kvs"]) >= len(rg["kvs"])
    for be in pair:
        be.manual_compact(NOW)
    rg3, ro3 = both(pair, "get_scanner", full_start, full_stop, full_scan=True, batch_size=1000, validate_partition_hash=False, now=NOW)
    assert same_response(rg3, ro3)[0] and len(rg3["kvs"]) == len(rg["kvs"])


def test_randomized_differential(engine):
    rnd = random.Random(20240917)
    pair = (Backend("gpu", engine, opts={"l0_compaction_trigger": 3}), Backend("oracle", opts={"l0_compaction_trigger": 3}))
    hks = [b"", b"a", b"ab", b"abc", b"b\xff", b"b\xff\xff", b"user%d" % 7, b"k" * 40]
    sks = [b"", b"0", b"1", b"10", b"1\xff", b"2", b"zz", b"m" * 60] + [b"s%03d" % i for i in range(40)]
    now = NOW
    try:
        for step in range(400):
            op = rnd.random()
            hk = rnd.choice(hks)
            if op < 0.35:
                sk = rnd.choice(sks)
                if not hk and not sk:
                    continue
                ets = rnd.choice([0, 0, now + rnd.randint(1, 50), now - rnd.randint(1, 50)])
                val = bytes(rnd.getrandbits(8) for _ in range(rnd.randint(0, 300)))
                for be in pair:
                    be.put(hk, sk, val, expire_ts=ets, now=now)
            elif op < 0.45:
                sk = rnd.choice(sks)
                if not hk and not sk:
                    continue
                for be in pair:
                    be.remove(hk, sk)
            elif op < 0.50:
                for be in pair:
                    be.flush(now)
            elif op < 0.52:
                for be in pair:
                    be.manual_compact(now)
            elif op < 0.62:
                check(pair, "get", hk, rnd.choice(sks), now=now)
            elif op < 0.78:
                a, c = sorted([rnd.choice(sks), rnd.choice(sks)])
                check(pair, "multi_get", hk, rnd.choice([b"", a]), rnd.choice([b"", c]), rnd.random() < 0.5, rnd.random() < 0.5,
                      max_kv_count=rnd.choice([0, 1, 3, 100]), max_kv_size=rnd.choice([0, 50, 100000]), no_value=rnd.random() < 0.2,
                      reverse=rnd.random() < 0.4, filter_type=rnd.choice([0, 0, 1, 2, 3]),
                      filter_pattern=rnd.choice([b"", b"s", b"1", b"0", b"s01", b"\xff"]), now=now)
            elif op < 0.84:
                check(pair, "sortkey_count", hk, now=now)
            elif op < 0.92:
                if hk:
                    (kg, bg), (ko, bo) = both(pair, "scan_all", hk, batch_size=rnd.choice([1, 3, 10, 0]), now=now,
                                              no_value=rnd.random() < 0.2, sort_filter=(rnd.choice([0, 1, 2, 3]), rnd.choice([b"", b"s0", b"1"])))
                    assert kg == ko and len(bg) == len(bo)
            else:
                rg, ro = both(pair, "get_scanner", raw_key(b"", b""), b"\xff\xff", full_scan=True, batch_size=rnd.choice([5, 50]), now=now,
                              hash_filter=(rnd.choice([0, 1, 2, 3]), rnd.choice([b"", b"a", b"b"])))
                n = 0
                while True:
                    ok, d = same_response(rg, ro)
                    assert ok, (step, d)
                    if rg["context_id"] < 0 or rg["error"]:
                        break
                    rg, ro = pair[0].scan(rg["context_id"], now=now), pair[1].scan(ro["context_id"], now=now)
                    n += 1
            if step % 50 == 49:
                now += 20
    finally:
        for be in pair:
            be.close()


def test_manual_compact_env_trigger(pair):
    """pegasus_manual_compact_service `once` rule: trigger_time newer than the last finish starts a full compaction
    with the filter; the same env again does nothing; `disabled` wins."""
    unix = lambda now: now + 1451606400
    for be in pair:
        be.put(b"mc", b"keep", b"1")
        be.put(b"mc", b"dead", b"2", expire_ts=NOW - 5, now=NOW - 10)
        be.flush(NOW)
        be.put(b"mc", b"keep2", b"3")
        be.update_envs({"default_ttl": "900"})
    for be in pair:
        be.update_envs({"manual_compact.disabled": "true", "manual_compact.once.trigger_time": str(unix(NOW))}, now=NOW)
    assert check(pair, "ttl", b"mc", b"keep", now=NOW)["ttl"] == -1  # nothing ran
    for be in pair:
        be.update_envs({"manual_compact.once.trigger_time": str(unix(NOW)), "manual_compact.once.bottommost_level_compaction": "force"}, now=NOW)
    assert check(pair, "ttl", b"mc", b"keep", now=NOW)["ttl"] == 900  # default_ttl rewrite happened
    assert check(pair, "get", b"mc", b"dead", now=NOW)["error"] == 1
    for be in pair:
        be.update_envs({"default_ttl": "100"})
        be.update_envs({"manual_compact.once.trigger_time": str(unix(NOW))}, now=NOW + 50)  # not newer than the last finish
        be.put(b"mc", b"later", b"4", now=NOW + 50)
    assert check(pair, "ttl", b"mc", b"keep", now=NOW + 50)["ttl"] == 850
    check(pair, "multi_get", b"mc", now=NOW + 50)
