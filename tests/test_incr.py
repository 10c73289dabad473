"""incr through the rrdb surface (first operator of the write-side set): the reference's own unit cases
(src/server/test/pegasus_write_service_impl_test.cpp:234-316, NonIdempotentIncrTest) restated against the oracle (CPU) and the
CUDA engine (-m gpu), plus a differential random history.  The read-before-write goes through the memtable overlay and
pgs_get_batch like any point read."""
import random

import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend, same_response

NOW = synth.NOW
I64_MAX, I64_MIN = 2**63 - 1, -2**63
INVALID = 4  # rocksdb::Status::kInvalidArgument


def backends():
    return [pytest.param("oracle", id="oracle"), pytest.param("gpu", marks=pytest.mark.gpu, id="gpu")]


@pytest.fixture
def be(request):
    kind = request.param
    eng = request.getfixturevalue("engine") if kind == "gpu" else None
    b = Backend(kind, eng, pidx=11)
    yield b
    b.close()


def value_of(be, now=NOW):
    r = be.get(b"incr_hash_key", b"incr_sort_key", now=now)
    return None if r["error"] else r["kvs"][0][1]


@pytest.mark.parametrize("be", backends(), indirect=True)
def test_reference_unit_cases(be):
    K = (b"incr_hash_key", b"incr_sort_key")
    # IncrOneOnAbsentRecord / IncrOneOnExistingRecord / IncrBigOnExistingRecord
    assert be.incr(*K, 1, now=NOW) == (0, 0, 1) and value_of(be) == b"1"
    be.put(*K, b"10", now=NOW)
    assert be.incr(*K, 1, now=NOW) == (0, 0, 11) and value_of(be) == b"11"
    be.put(*K, b"10", now=NOW)
    assert be.incr(*K, 100, now=NOW) == (0, 0, 110) and value_of(be) == b"110"
    # IncrNegative
    be.remove(*K, now=NOW)
    assert be.incr(*K, -100, now=NOW) == (0, 0, -100)
    assert be.incr(*K, -1, now=NOW) == (0, 0, -101) and value_of(be) == b"-101"
    # IncrZero
    be.remove(*K, now=NOW)
    assert be.incr(*K, 0, now=NOW) == (0, 0, 0) and value_of(be) == b"0"
    for base in (10, -10):
        be.put(*K, b"%d" % base, now=NOW)
        assert be.incr(*K, 0, now=NOW) == (0, 0, base)
    # IncrOnNonNumericRecord: the call succeeds, the response carries kInvalidArgument, the record is untouched
    be.put(*K, b"abc", now=NOW)
    assert be.incr(*K, 1, now=NOW)[:2] == (0, INVALID) and value_of(be) == b"abc"
    # IncrOverflow / IncrUnderflow: the response returns the base value
    be.put(*K, b"1", now=NOW)
    assert be.incr(*K, I64_MAX, now=NOW) == (0, INVALID, 1) and value_of(be) == b"1"
    be.put(*K, b"-1", now=NOW)
    assert be.incr(*K, I64_MIN, now=NOW) == (0, INVALID, -1) and value_of(be) == b"-1"
    # IncrOnExpireRecord
    be.remove(*K, now=NOW)
    assert be.incr(*K, 10, expire_ts_seconds=1, now=NOW) == (0, 0, 10)   # expire_ts 1: long expired
    assert value_of(be) is None
    assert be.incr(*K, 100, now=NOW) == (0, 0, 100) and value_of(be) == b"100"
    # an empty value counts as 0; strtoll base 0 accepts hex / octal like dsn::buf2int64
    be.put(*K, b"", now=NOW)
    assert be.incr(*K, 7, now=NOW) == (0, 0, 7)
    be.put(*K, b"0x10", now=NOW)
    assert be.incr(*K, 1, now=NOW) == (0, 0, 17)
    be.put(*K, b" 5", now=NOW)   # strtoll skips leading blanks, the whole buffer is still consumed
    assert be.incr(*K, 1, now=NOW) == (0, 0, 6)
    be.put(*K, b"5 ", now=NOW)
    assert be.incr(*K, 1, now=NOW)[:2] == (0, INVALID)


@pytest.mark.parametrize("be", backends(), indirect=True)
def test_expiry_rules(be):
    K = (b"incr_hash_key", b"incr_sort_key")
    be.put(*K, b"5", expire_ts=NOW + 100, now=NOW)
    assert be.incr(*K, 1, expire_ts_seconds=0, now=NOW) == (0, 0, 6) and be.ttl(*K, now=NOW)["ttl"] == 100   # kept
    assert be.incr(*K, 1, expire_ts_seconds=NOW + 50, now=NOW) == (0, 0, 7) and be.ttl(*K, now=NOW)["ttl"] == 50
    assert be.incr(*K, 1, expire_ts_seconds=-1, now=NOW) == (0, 0, 8) and be.ttl(*K, now=NOW)["ttl"] == -1   # cleared
    be.flush(NOW)   # the base now lives in an HBM run, not in the memtable
    assert be.incr(*K, 2, now=NOW) == (0, 0, 10) and value_of(be) == b"10"


@pytest.mark.gpu
def test_random_history_matches_the_oracle(engine):
    opts = {"memtable_bytes": 2 << 10, "l0_compaction_trigger": 3}
    g, o = Backend("gpu", engine, pidx=12, opts=opts), Backend("oracle", pidx=12, opts=opts)
    rnd = random.Random(9)
    try:
        for i in range(1500):
            hk, sk = b"c%02d" % rnd.randrange(40), b"n"
            kind = rnd.random()
            if kind < 0.6:
                a = (rnd.choice([1, -1, 5, 10**12, I64_MAX]), rnd.choice([0, 0, -1, NOW + 30, NOW - 1]))
                assert g.incr(hk, sk, a[0], expire_ts_seconds=a[1], now=NOW) == o.incr(hk, sk, a[0], expire_ts_seconds=a[1], now=NOW), (i, a)
            elif kind < 0.7:
                v = rnd.choice([b"", b"abc", b"42", b"-7", b"9223372036854775807"])
                g.put(hk, sk, v, now=NOW); o.put(hk, sk, v, now=NOW)
            elif kind < 0.75:
                g.remove(hk, sk, now=NOW); o.remove(hk, sk, now=NOW)
            else:
                ok, d = same_response(g.get(hk, sk, now=NOW), o.get(hk, sk, now=NOW))
                assert ok, (i, d)
        assert g.f("rrdb_last_flushed_decree")(g.h) == o.f("rrdb_last_flushed_decree")(o.h) > 0
    finally:
        g.close(); o.close()
