"""src/test/function_test/base_api/test_ttl.cpp:79-205 restated on the CPU oracle with an explicit clock (the reference
sleeps; here `now` is a parameter, so its "within error_allow" windows become exact values).  Pins the oracle's write-time
default TTL (rocksdb_wrapper.cpp:280-288), read-side expiry (pegasus_server_impl.cpp:443-448, on_ttl :1088-1149) and the
compaction filter's default_ttl rewrite / expiry drop (key_ttl_compaction_filter.h:55-92) behind the manual-compaction
env trigger (pegasus_manual_compact_service.cpp:83-121).  The CUDA engine is compared with the oracle on the same shape
in test_rrdb_gpu.py::test_ttl_default_ttl_and_manual_compact."""
from rrdb_harness import Backend

T0 = 200_000_000
HK = b"ttl_test_hash_key"
SK0, SK1, SK2 = b"ttl_test_sort_key_0", b"ttl_test_sort_key_1", b"ttl_test_sort_key_2"
V0, V1, V2 = b"ttl_test_value_0", b"ttl_test_value_1", b"ttl_test_value_2"
DEFAULT_TTL, SPECIFY_TTL, SLEEP, ENV_EFFECT = 3600, 5, 10, 31
unix = lambda now: now + 1451606400  # epoch_now() counts from 2016-01-01 (pegasus_utils.h:39-41)


def get(be, sk, now):
    r = be.get(HK, sk, now=now)
    return r["error"], (r["kvs"][0][1] if r["kvs"] else None)


def ttl(be, sk, now):
    r = be.ttl(HK, sk, now=now)
    return r["error"], r["ttl"]


def test_set_without_default_ttl():
    be = Backend("oracle")
    try:
        be.put(HK, SK1, V1, expire_ts=T0 + SPECIFY_TTL, now=T0)  # client: expire_ts = epoch_now() + ttl_seconds
        assert get(be, SK1, T0) == (0, V1)
        assert ttl(be, SK1, T0) == (0, SPECIFY_TTL)
        be.put(HK, SK2, V2, now=T0)
        assert get(be, SK2, T0) == (0, V2)
        assert ttl(be, SK2, T0) == (0, -1)
        t1 = T0 + SLEEP
        assert ttl(be, SK1, t1)[0] == 1 and get(be, SK1, t1)[0] == 1  # PERR_NOT_FOUND
        assert ttl(be, SK2, t1) == (0, -1) and get(be, SK2, t1) == (0, V2)
        be.update_envs({"manual_compact.once.trigger_time": str(unix(t1))}, now=t1)
        assert ttl(be, SK1, t1)[0] == 1 and get(be, SK1, t1)[0] == 1
        assert ttl(be, SK2, t1) == (0, -1) and get(be, SK2, t1) == (0, V2)
    finally:
        be.close()


def test_set_with_default_ttl():
    be = Backend("oracle")
    try:
        be.put(HK, SK0, V0, now=T0)
        be.update_envs({"default_ttl": str(DEFAULT_TTL)}, now=T0)
        be.put(HK, SK1, V1, expire_ts=T0 + SPECIFY_TTL, now=T0)
        assert get(be, SK1, T0) == (0, V1)
        assert ttl(be, SK1, T0) == (0, SPECIFY_TTL)
        be.put(HK, SK2, V2, now=T0)  # no TTL given: the table default applies at write time
        assert get(be, SK2, T0) == (0, V2)
        assert ttl(be, SK2, T0) == (0, DEFAULT_TTL)
        t1 = T0 + SLEEP
        assert ttl(be, SK0, t1) == (0, -1)  # written before the env: still forever
        assert ttl(be, SK1, t1)[0] == 1 and get(be, SK1, t1)[0] == 1
        assert ttl(be, SK2, t1) == (0, DEFAULT_TTL - SLEEP) and get(be, SK2, t1) == (0, V2)
        t2 = t1 + ENV_EFFECT
        be.update_envs({"manual_compact.once.trigger_time": str(unix(t2))}, now=t2)
        assert ttl(be, SK0, t2) == (0, DEFAULT_TTL)  # the compaction filter gave the TTL-less record now + default_ttl
        assert ttl(be, SK1, t2)[0] == 1 and get(be, SK1, t2)[0] == 1
        assert ttl(be, SK2, t2) == (0, DEFAULT_TTL - SLEEP - ENV_EFFECT)
    finally:
        be.close()
