"""check_and_set / check_and_mutate through the rrdb surface: validate_check's table
(src/server/pegasus_write_service_impl.h:1144-1270) and the sequences of the reference's function test
(src/test/function_test/base_api/test_check_and_set.cpp:33-1580; the client turns kTryAgain into set_succeed = false) restated
against the oracle (CPU) and the CUDA engine (-m gpu).  The checked value is read like any point read: memtable overlay, then
pgs_get_batch."""
import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend

NOW = synth.NOW
OK, INVALID, TRY_AGAIN = 0, 4, 13
(NO_CHECK, NOT_EXIST, NOT_EXIST_OR_EMPTY, EXIST, NOT_EMPTY, ANYWHERE, PREFIX, POSTFIX, B_LT, B_LE, B_EQ, B_GE, B_GT,
 I_LT, I_LE, I_EQ, I_GE, I_GT) = range(18)


def backends():
    return [pytest.param("oracle", id="oracle"), pytest.param("gpu", marks=pytest.mark.gpu, id="gpu")]


@pytest.fixture
def be(request):
    kind = request.param
    eng = request.getfixturevalue("engine") if kind == "gpu" else None
    b = Backend(kind, eng, pidx=13)
    yield b
    b.close()


def get(be, hk, sk):
    r = be.get(hk, sk, now=NOW)
    return None if r["error"] else r["kvs"][0][1]


# (stored value or None, check type, operand) -> passes?  ("inv": kInvalidArgument)
TABLE = [
    (None, NO_CHECK, b"", True), (b"x", NO_CHECK, b"zz", True),
    (None, NOT_EXIST, b"", True), (b"", NOT_EXIST, b"", False), (b"v", NOT_EXIST, b"", False),
    (None, NOT_EXIST_OR_EMPTY, b"", True), (b"", NOT_EXIST_OR_EMPTY, b"", True), (b"v", NOT_EXIST_OR_EMPTY, b"", False),
    (None, EXIST, b"", False), (b"", EXIST, b"", True), (b"v", EXIST, b"", True),
    (None, NOT_EMPTY, b"", False), (b"", NOT_EMPTY, b"", False), (b"v", NOT_EMPTY, b"", True),
    (None, ANYWHERE, b"v", False), (b"", ANYWHERE, b"v", False), (b"", ANYWHERE, b"", True), (b"v111v", ANYWHERE, b"", True),
    (b"v111v", ANYWHERE, b"111", True), (b"v111v", ANYWHERE, b"y", False), (b"v111v", ANYWHERE, b"v111v", True), (b"v111v", ANYWHERE, b"v111vv", False),
    (b"v111v", PREFIX, b"v", True), (b"v111v", PREFIX, b"v111", True), (b"v111v", PREFIX, b"111", False), (b"v111v", PREFIX, b"", True), (None, PREFIX, b"", False),
    (b"v111v", POSTFIX, b"v", True), (b"v111v", POSTFIX, b"111v", True), (b"v111v", POSTFIX, b"111", False), (b"v111v", POSTFIX, b"2v111v", False),
    (None, B_EQ, b"", False), (b"", B_EQ, b"", True), (b"v1", B_EQ, b"v1", True), (b"v1", B_EQ, b"v2", False),
    (b"v1", B_LT, b"v2", True), (b"v2", B_LT, b"v2", False), (b"v3", B_LT, b"v2", False),
    (b"v1", B_LE, b"v2", True), (b"v2", B_LE, b"v2", True), (b"v3", B_LE, b"v2", False),
    (b"v1", B_GE, b"v2", False), (b"v2", B_GE, b"v2", True), (b"v3", B_GE, b"v2", True),
    (b"v1", B_GT, b"v2", False), (b"v2", B_GT, b"v2", False), (b"v3", B_GT, b"v2", True), (b"v", B_LT, b"v1", True),
    (None, I_EQ, b"1", False), (b"", I_EQ, b"1", "inv"), (b"1", I_EQ, b"1", True), (b"1", I_EQ, b"", "inv"), (b"1", I_EQ, b"v1", "inv"),
    (b"v1", I_EQ, b"1", "inv"), (b"1", I_EQ, b"88888888888888888888888888888888888888888888888", "inv"), (b"0", I_EQ, b"0x0", True),
    (b"10", I_LT, b"9", False), (b"-10", I_LT, b"9", True), (b"9", I_LT, b"9", False), (b"9", I_LE, b"9", True), (b"10", I_LE, b"9", False),
    (b"9", I_GE, b"9", True), (b"8", I_GE, b"9", False), (b"10", I_GT, b"9", True), (b"9", I_GT, b"9", False),
    (b"9223372036854775807", I_GT, b"-9223372036854775808", True),
]


@pytest.mark.parametrize("be", backends(), indirect=True)
def test_validate_check_table(be):
    hk = b"cas_table"
    for i, (stored, ctype, operand, want) in enumerate(TABLE):
        sk = b"k%03d" % i
        if stored is not None:
            be.put(hk, sk, stored, now=NOW)
        if i % 7 == 3:
            be.flush(NOW)  # some of the checked values live in HBM runs, the others in the memtable
        r = be.check_and_set(hk, sk, ctype, operand, sk, b"new", now=NOW)
        assert r["rc"] == OK, (i, r)
        assert r["returned"] and r["exist"] == (stored is not None) and r["check_value"] == stored, (i, r)
        if want is True:
            assert r["error"] == OK and get(be, hk, sk) == b"new", (i, stored, ctype, operand, r)
        else:
            assert r["error"] == (INVALID if want == "inv" else TRY_AGAIN), (i, stored, ctype, operand, r)
            assert get(be, hk, sk) == stored, (i, r)


@pytest.mark.parametrize("be", backends(), indirect=True)
def test_reference_sequences(be):
    hk = b"check_and_set_test_value_not_exist"
    # value_not_exist, k1 (:36-87)
    r = be.check_and_set(hk, b"k1", NOT_EXIST, b"", b"k1", b"v1", now=NOW)
    assert (r["error"], r["returned"], r["exist"]) == (OK, True, False) and get(be, hk, b"k1") == b"v1"
    r = be.check_and_set(hk, b"k1", NOT_EXIST, b"", b"k1", b"v2", now=NOW)
    assert (r["error"], r["returned"], r["exist"], r["check_value"]) == (TRY_AGAIN, True, True, b"v1") and get(be, hk, b"k1") == b"v1"
    r = be.check_and_set(hk, b"k1", NOT_EXIST, b"", b"k1", b"v1", return_check_value=False, now=NOW)
    assert (r["error"], r["returned"]) == (TRY_AGAIN, False)
    # k3 checked, k4 set (:125-164)
    r = be.check_and_set(hk, b"k3", NOT_EXIST, b"", b"k4", b"v4", now=NOW)
    assert r["error"] == OK and get(be, hk, b"k3") is None and get(be, hk, b"k4") == b"v4"
    # an expired value does not exist for the check
    be.put(hk, b"k5", b"old", expire_ts=NOW - 1, now=NOW)
    r = be.check_and_set(hk, b"k5", NOT_EXIST, b"", b"k5", b"fresh", ttl_ts=NOW + 30, now=NOW)
    assert (r["error"], r["exist"]) == (OK, False) and be.ttl(hk, b"k5", now=NOW)["ttl"] == 30
    # invalid_type (:1555-1578): nothing is read, nothing is returned
    r = be.check_and_set(hk, b"k1", 100, b"v", b"k1", b"v1", now=NOW)
    assert (r["rc"], r["error"], r["returned"]) == (OK, INVALID, False) and get(be, hk, b"k1") == b"v1"
    d0 = be.f("rrdb_last_committed_decree")(be.h)
    assert d0 == be.decree  # failed requests advanced the decree too (empty_put)


@pytest.mark.parametrize("be", backends(), indirect=True)
def test_check_and_mutate(be):
    hk = b"cam"
    be.put(hk, b"guard", b"7", now=NOW)
    be.put(hk, b"old", b"x", now=NOW)
    r = be.check_and_mutate(hk, b"guard", I_GE, b"5", [("put", b"a", b"1", NOW + 60), ("put", b"b", b""), ("del", b"old")], now=NOW)
    assert (r["rc"], r["error"], r["check_value"]) == (OK, OK, b"7")
    assert get(be, hk, b"a") == b"1" and get(be, hk, b"b") == b"" and get(be, hk, b"old") is None and be.ttl(hk, b"a", now=NOW)["ttl"] == 60
    r = be.check_and_mutate(hk, b"guard", I_LT, b"5", [("put", b"a", b"2")], now=NOW)
    assert r["error"] == TRY_AGAIN and get(be, hk, b"a") == b"1"
    assert be.check_and_mutate(hk, b"guard", NO_CHECK, b"", [], now=NOW)["error"] == INVALID          # empty mutate list
    assert be.check_and_mutate(hk, b"guard", NO_CHECK, b"", [(5, b"a", b"z")], now=NOW)["error"] == INVALID  # bad operation
    assert get(be, hk, b"a") == b"1"
    r = be.check_and_mutate(hk, b"missing", EXIST, b"", [("del", b"a")], now=NOW)
    assert (r["error"], r["exist"]) == (TRY_AGAIN, False) and get(be, hk, b"a") == b"1"
