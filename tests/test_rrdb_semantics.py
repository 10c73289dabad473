"""rrdb-surface semantics that round 1 got wrong or left open, checked on the CPU oracle (always) and on the
CUDA engine (-m gpu) with the same assertions:
  * a remove that fills the memtable compacts with the real clock (default-TTL rewrite = now + ttl, not 0 + ttl)
  * update_app_envs follows the reference's parsers and "absent key = deleted" rule
    (pegasus_server_impl.cpp:2814-2826, 2966-3001)
  * parked scan contexts expire after five minutes (pegasus_server_impl.cpp:1377-1385)
  * last_flushed_decree only moves when the memtable becomes an HBM run
  * point reads see the memtable in place (rocksdb_wrapper.cpp:78-127), no flush"""
import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend, raw_key, same_response, next_blob

NOW = synth.NOW


def backends():
    return [pytest.param("oracle", id="oracle"), pytest.param("gpu", marks=pytest.mark.gpu, id="gpu")]


@pytest.fixture
def make(request):
    made = []

    def _make(kind, **kw):
        eng = request.getfixturevalue("engine") if kind == "gpu" else None
        b = Backend(kind, eng, **kw)
        made.append(b)
        return b
    yield _make
    for b in made:
        b.close()


@pytest.mark.parametrize("kind", backends())
def test_remove_triggered_compaction_uses_the_clock(make, kind):
    be = make(kind, opts={"memtable_bytes": 600, "l0_compaction_trigger": 2}, envs={"default_ttl": "1000"})
    # records without TTL, written with default_ttl already set: put stamps now + ttl itself; an explicit expire_ts=far future
    # record keeps its own.  What matters here are records that reach the compaction filter with expire_ts == 0:
    be.update_envs({}, now=NOW)  # default_ttl untouched (absent key), split validation off
    be.update_envs({"default_ttl": "0"}, now=NOW)
    for i in range(6):
        be.put(b"h", b"s%02d" % i, b"v" * 40, now=NOW)      # expire_ts 0 in the value
    be.flush(NOW)
    be.update_envs({"default_ttl": "1000"}, now=NOW)
    for i in range(40):                                       # removes of other keys fill the tiny memtable -> flush -> L0 compaction
        be.remove(b"x", b"k%03d" % i, now=NOW + 5)
    r = be.ttl(b"h", b"s00", now=NOW + 6)
    assert r["error"] == 0
    assert r["ttl"] == 1000 + 5 - 6, r                       # rewritten as (NOW + 5) + 1000 by the compaction that a remove triggered
    assert be.get(b"h", b"s03", now=NOW + 900)["error"] == 0  # still alive; with a clock of 0 it would read as expired (2016)


@pytest.mark.parametrize("kind", backends())
def test_env_parsers_follow_the_reference(make, kind):
    be = make(kind, envs={"default_ttl": "100"})
    be.put(b"h", b"a", b"1", now=NOW)
    assert be.ttl(b"h", b"a", now=NOW)["ttl"] == 100
    for bad in ("abc", "-1", "", "12x", "99999999999"):
        be.update_envs({"default_ttl": bad}, now=NOW)        # invalid: the old value stays
        be.put(b"h", b"b" + bad.encode(), b"1", now=NOW)
        assert be.ttl(b"h", b"b" + bad.encode(), now=NOW)["ttl"] == 100, bad
    be.update_envs({"default_ttl": "7"}, now=NOW)
    be.put(b"h", b"c", b"1", now=NOW)
    assert be.ttl(b"h", b"c", now=NOW)["ttl"] == 7
    # split validation: buf2bool is case-insensitive; garbage changes nothing; an absent key switches it off
    be2 = make(kind, pidx=1)
    be2.set_partition_version(3)
    for i in range(64):
        be2.put(b"hk%02d" % i, b"s", b"v", now=NOW)
    be2.flush(NOW)
    start, stop = b"", b"\xff\xff\xff"  # the whole table

    def scan_count():
        r = be2.get_scanner(start, stop, batch_size=1000, full_scan=True, now=NOW)
        return len(r["kvs"])
    assert scan_count() == 64
    be2.update_envs({"replica.split.validate_partition_hash": "TRUE"}, now=NOW)
    n_on = scan_count()
    assert 0 < n_on < 64                                      # stale-split keys are skipped silently
    be2.update_envs({"replica.split.validate_partition_hash": "yes"}, now=NOW)
    assert scan_count() == n_on                               # unparsable: unchanged
    be2.update_envs({}, now=NOW)
    assert scan_count() == 64                                 # deleted from the env map: off


@pytest.mark.parametrize("kind", backends())
def test_scan_contexts_expire(make, kind):
    be = make(kind)
    for i in range(30):
        be.put(b"h", b"s%02d" % i, b"v", now=NOW)
    be.flush(NOW)
    start, stop = raw_key(b"h", b""), next_blob(raw_key(b"h", b""))
    r = be.get_scanner(start, stop, batch_size=5, now=NOW)
    assert r["context_id"] >= 0 and len(r["kvs"]) == 5
    r2 = be.scan(r["context_id"], now=NOW + 299)              # still there just before the deadline; re-parked at NOW + 299
    assert r2["error"] == 0 and len(r2["kvs"]) == 5 and r2["context_id"] >= 0
    r3 = be.scan(r2["context_id"], now=NOW + 299 + 300)       # five minutes after it was parked: gone
    assert r3["error"] == 1 and r3["kvs"] == []
    a = be.get_scanner(start, stop, batch_size=5, now=NOW)
    b = be.get_scanner(start, stop, batch_size=5, now=NOW + 100)
    assert be.f("rrdb_gc")(be.h, NOW + 350) == 1              # only the older one
    assert be.scan(a["context_id"], now=NOW + 351)["error"] == 1
    assert be.scan(b["context_id"], now=NOW + 351)["error"] == 0


@pytest.mark.parametrize("kind", backends())
def test_decrees_and_memtable_reads(make, kind):
    be = make(kind, opts={"l0_compaction_trigger": 100})
    be.put(b"h", b"a", b"old", now=NOW)
    be.flush(NOW)
    d_flushed = be.f("rrdb_last_flushed_decree")(be.h)
    assert d_flushed == be.decree
    be.put(b"h", b"a", b"new", now=NOW)                       # newer version in the memtable
    be.put(b"h", b"b", b"only-mem", expire_ts=NOW + 50, now=NOW)
    be.remove(b"h", b"gone", now=NOW)
    assert be.f("rrdb_last_flushed_decree")(be.h) == d_flushed   # nothing reached an HBM run yet
    assert be.f("rrdb_last_committed_decree")(be.h) == be.decree
    assert be.get(b"h", b"a", now=NOW)["kvs"][0][1] == b"new"
    assert be.ttl(b"h", b"b", now=NOW + 10)["ttl"] == 40
    assert be.get(b"h", b"b", now=NOW + 60)["error"] == 1 and be.get(b"h", b"b", now=NOW + 60)["expire_count"] == 1
    r = be.multi_get(b"h", sort_keys=[b"a", b"b", b"gone", b"zz"], now=NOW)
    assert [(k, v) for k, v, _ in r["kvs"]] == [(b"a", b"new"), (b"b", b"only-mem")]
    r = be.batch_get([(b"h", b"b"), (b"h", b"a"), (b"h", b"nope")], now=NOW)
    assert [(k, v) for k, v, _, _ in r["kvs"]] == [(b"hb", b"only-mem"), (b"ha", b"new")]
    assert be.f("rrdb_last_flushed_decree")(be.h) == d_flushed   # the point reads did not flush
    r = be.multi_get(b"h", now=NOW)                           # a range read does
    assert [(k, v) for k, v, _ in r["kvs"]] == [(b"a", b"new"), (b"b", b"only-mem")]
    assert be.f("rrdb_last_flushed_decree")(be.h) == be.decree


@pytest.mark.parametrize("kind", backends())
def test_batched_writes_of_one_decree(make, kind):
    """on_batched_write_requests (pegasus_server_write.cpp:92-222): puts and removes of one decree apply together, an empty
    batch only advances the decree, an operation that may not be batched applies nothing"""
    be = make(kind)
    be.put(b"h", b"gone", b"x", now=NOW)
    rc, errs = be.batched_writes([("put", b"h", b"a", b"1", NOW + 40), ("put", b"h", b"b", b"2"), ("remove", b"h", b"gone"), ("put", b"h", b"a", b"3")], now=NOW)
    assert rc == 0 and errs == [0, 0, 0, 0]
    assert be.get(b"h", b"a", now=NOW)["kvs"][0][1] == b"3" and be.ttl(b"h", b"a", now=NOW)["ttl"] == -1   # the later put of the batch wins
    assert be.get(b"h", b"b", now=NOW)["error"] == 0 and be.get(b"h", b"gone", now=NOW)["error"] == 1
    d = be.decree
    assert be.batched_writes([], now=NOW) == (0, []) and be.f("rrdb_last_committed_decree")(be.h) == d + 1
    assert be.batched_writes([("put", b"h", b"c", b"9"), ("multi_put", b"h", b"d", b"9")], now=NOW)[0] == 4
    assert be.get(b"h", b"c", now=NOW)["error"] == 1
    # pegasus_write_service_test.cpp:170-207 (test_batched_writes): 100 puts, then removes of the same 100 keys, one decree, every
    # response kOk; what the batch leaves behind is nothing
    ops = [("put", b"hash_key", b"sort_key_%d" % i, b"value_%d" % i) for i in range(100)]
    ops += [("remove", b"hash_key", b"sort_key_%d" % i) for i in range(100)]
    rc, errs = be.batched_writes(ops, now=NOW, ts_us=1000)
    assert rc == 0 and errs == [0] * 200
    assert be.multi_get(b"hash_key", now=NOW)["kvs"] == []
    assert be.sortkey_count(b"hash_key", now=NOW)["count"] == 0


@pytest.mark.parametrize("kind", backends())
def test_periodic_manual_compaction(make, kind):
    """pegasus_manual_compact_service.cpp:186-219: a time of the local day between the last finished manual compaction and
    now starts one; the same env afterwards does nothing until the next time of day passes; `once` is looked at first"""
    import time
    unix = NOW + 1451606400
    lt = time.localtime(unix)
    noon = NOW - (lt.tm_hour * 3600 + lt.tm_min * 60 + lt.tm_sec) + 12 * 3600  # 12:00 of NOW's local day, Pegasus seconds
    be = make(kind)
    be.put(b"h", b"keep", b"1", now=noon)                       # expire_ts 0: default_ttl is not set yet
    be.flush(noon)
    be.put(b"h", b"keep2", b"2", now=noon)
    per = "manual_compact.periodic.trigger_time"
    be.update_envs({"default_ttl": "9000", per: "13:00,25:00,xx"}, now=noon)           # 13:00 has not come, the rest is not a time
    assert be.ttl(b"h", b"keep", now=noon)["ttl"] == -1
    be.update_envs({"default_ttl": "9000", per: "11:00,13:00", "manual_compact.disabled": "true"}, now=noon)
    assert be.ttl(b"h", b"keep", now=noon)["ttl"] == -1        # disabled wins
    be.update_envs({"default_ttl": "9000", per: "11:00,13:00", "manual_compact.max_concurrent_running_count": "0"}, now=noon)
    assert be.ttl(b"h", b"keep", now=noon)["ttl"] == -1        # no compaction may run
    be.update_envs({"default_ttl": "9000", per: "11:00,13:00"}, now=noon)              # 11:00 passed since the last finish (never)
    assert be.ttl(b"h", b"keep", now=noon)["ttl"] == 9000      # the filter's default-TTL rewrite happened
    assert be.ttl(b"h", b"keep2", now=noon)["ttl"] == 9000     # the memtable was flushed into it
    be.update_envs({"default_ttl": "0", per: "11:00,13:00"}, now=noon + 60)
    be.put(b"h", b"late", b"3", now=noon + 60)                  # expire_ts 0 again
    be.update_envs({"default_ttl": "100", per: "11:00,13:00"}, now=noon + 120)         # finished at 12:00: 11:00 no longer fires
    assert be.ttl(b"h", b"late", now=noon + 120)["ttl"] == -1
    be.update_envs({"default_ttl": "100", per: "11:00,13:00"}, now=noon + 3600 + 30)   # 13:00 passed
    assert be.ttl(b"h", b"late", now=noon + 3600 + 30)["ttl"] == 100
    assert be.ttl(b"h", b"keep", now=noon + 3600 + 30)["ttl"] == 9000- 3630
    assert be.get(b"h", b"keep2", now=noon + 3600 + 30)["error"] == 0
