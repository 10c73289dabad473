"""The manual-compaction rules of the product's host library (pgs_manual_compact_decide / _state_check) against the cases the
reference's own test holds (src/server/test/manual_compact_service_test.cpp:120-330), and the compression setting
(pgs_parse_compression_types) against src/server/test/pegasus_compression_options_test.cpp:115-155.  No device work."""
import incubator_pegasus_b200 as pgs

COMPACTED_TS = 1500000000  # manual_compact_service_test.cpp:45
MIDNIGHT = 1499961600      # some day's 00:00:00; the reference asks localtime, the product takes it as an argument
DAY_NOW = (MIDNIGHT + 12 * 3600) * 1000


def hhmm(s):
    h, m = s.split(":")
    return MIDNIGHT + int(h) * 3600 + int(m) * 60


def decide(envs, now_ms=DAY_NOW, last_s=0, num_levels=7):
    return pgs.manual_compact_decide(envs, now_ms, last_s * 1000, MIDNIGHT, num_levels)


def test_check_compact_disabled():  # :120-143
    for v, want in ((None, 0), ("", 0), ("true", 1), ("false", 0), ("1", 0), ("0", 0), ("abc", 0)):
        envs = {} if v is None else {"manual_compact.disabled": v}
        assert decide(envs).disabled == want, v
    # a disabled table starts nothing even when a rule would fire (:86-89)
    d = decide({"manual_compact.disabled": "true", "manual_compact.once.trigger_time": str(COMPACTED_TS + 1)}, last_s=COMPACTED_TS)
    assert d.disabled == 1 and d.rule == 0


def test_max_concurrent_running_count():  # :147-166, :91-94
    fire = {"manual_compact.once.trigger_time": str(COMPACTED_TS + 1)}
    assert decide(fire).max_concurrent_running_count == 2**31 - 1 and decide(fire).rule == 1
    for v, count, rule in (("3", 3, 1), ("0", 0, 0), ("-1", -1, 0), ("abc", 2**31 - 1, 1), ("", 2**31 - 1, 1)):
        d = decide(dict(fire, **{"manual_compact.max_concurrent_running_count": v}))
        assert (d.max_concurrent_running_count, d.rule) == (count, rule), v


def test_check_once_compact():  # :145-176, compacted at 1500000000
    key = "manual_compact.once.trigger_time"
    for v, want in ((None, 0), ("", 0), ("abc", 0), ("-1", 0), (str(COMPACTED_TS - 1), 0), (str(COMPACTED_TS), 0),
                    (str(COMPACTED_TS + 1), 1), (str(COMPACTED_TS + 10**8), 1)):
        envs = {} if v is None else {key: v}
        assert decide(envs, last_s=COMPACTED_TS).rule == want, v


def test_check_periodic_compact():  # :178-259
    key = "manual_compact.periodic.trigger_time"

    def fires(v, last, now="12:00"):
        envs = {} if v is None else {key: v}
        return decide(envs, now_ms=hhmm(now) * 1000, last_s=last).rule == 2
    # invalid trigger time formats
    for v in (None, "", ",", "12:oo", str(COMPACTED_TS), "24:00", "10:60", "-1:00"):
        assert not fires(v, 0), v
    # compacted at 10:00: has been compacted
    for v in ("9:00", "3:00,9:00", "10:00"):
        assert not fires(v, hhmm("10:00")), v
    # compacted at 09:00, single compact time
    for now, want in (("08:00", False), ("09:30", False), ("10:30", True)):
        assert fires("10:00", hhmm("09:00"), now) == want, now
    # multiple compact times
    for now, want in (("08:00", False), ("09:30", False), ("10:30", True)):
        assert fires("10:00,21:00", hhmm("09:00"), now) == want, now
    # compacted at 11:00
    for now, want in (("11:01", False), ("20:30", False), ("21:01", True)):
        assert fires("10:00,21:00", hhmm("11:00"), now) == want, now
    # compacted at 21:50
    assert not fires("10:00,21:00", hhmm("21:50"), "22:00")
    # `once` wins when both fire (:96-104); the options then come from the once.* keys
    d = decide({key: "10:00", "manual_compact.once.trigger_time": str(hhmm("09:30")), "manual_compact.periodic.target_level": "3",
                "manual_compact.once.target_level": "2"}, now_ms=hhmm("10:30") * 1000, last_s=hhmm("09:00"))
    assert (d.rule, d.target_level) == (1, 2)
    d = decide({key: "10:00", "manual_compact.periodic.target_level": "3", "manual_compact.once.target_level": "2",
                "manual_compact.periodic.bottommost_level_compaction": "force"}, now_ms=hhmm("10:30") * 1000, last_s=hhmm("09:00"))
    assert (d.rule, d.target_level, d.bottommost_force) == (2, 3, 1)


def test_midnight_from_the_clock():
    """today_midnight_s = -1: the library derives the local day from now_ms, whatever the zone of this host is"""
    import time
    now_s = 1700000000
    lt = time.localtime(now_s)
    midnight = now_s - (lt.tm_hour * 3600 + lt.tm_min * 60 + lt.tm_sec)
    t = "%d:%02d" % (lt.tm_hour, lt.tm_min)  # the minute that holds now_s started at or before now_s
    d = pgs.manual_compact_decide({"manual_compact.periodic.trigger_time": t}, (now_s + 61) * 1000, (midnight - 1) * 1000, -1, 7)
    assert d.rule == 2
    d = pgs.manual_compact_decide({"manual_compact.periodic.trigger_time": t}, (now_s - 3600) * 1000, (midnight - 1) * 1000, -1, 7)
    assert d.rule == 0 or time.localtime(now_s - 3600).tm_mday != lt.tm_mday


def test_extract_manual_compact_opts():  # :261-300, num_levels = 7
    fire = {"manual_compact.once.trigger_time": str(COMPACTED_TS + 1)}
    d = decide(fire)
    assert (d.rule, d.target_level, d.bottommost_force) == (1, -1, 0)
    for tl, bl, want_tl, want_force in (("2", "force", 2, 1), ("-1", "skip", -1, 0), ("-2", "nonono", -1, 0), ("8", None, -1, 0),
                                        ("7", None, 7, 0), ("0", None, -1, 0), ("abc", "FORCE", -1, 0)):
        envs = dict(fire, **{"manual_compact.once.target_level": tl})
        if bl is not None:
            envs["manual_compact.once.bottommost_level_compaction"] = bl
        d = decide(envs)
        assert (d.target_level, d.bottommost_force) == (want_tl, want_force), (tl, bl)


def test_check_manual_compact_state_0_interval():  # :302-315
    first = 1500000000 * 1000
    ok, enq = pgs.manual_compact_state_check(first, 0, 0, 0)
    assert ok and enq == first                                     # 1st start ok
    assert pgs.manual_compact_state_check(first, 0, 0, enq) == (False, enq)  # 1st start not ok: one is queued
    last = first + 1000                                            # the compaction took one second; enqueue time reset
    ok, enq = pgs.manual_compact_state_check(last, last, 0, 0)
    assert ok and enq == last                                      # 2nd start ok
    assert not pgs.manual_compact_state_check(last, last, 0, enq)[0]


def test_check_manual_compact_state_1h_interval():  # :317-343
    first = 1500000000
    ok, enq = pgs.manual_compact_state_check(first * 1000, 0, 3600, 0)
    assert ok
    assert not pgs.manual_compact_state_check(first * 1000, 0, 3600, enq)[0]
    last = (first + 10) * 1000                                     # cost 10 seconds
    for past, want in ((1800, False), (3609, False), (3610, False), (3611, True)):
        ok, enq = pgs.manual_compact_state_check((first + past) * 1000, last, 3600, 0)
        assert ok == want, past
    assert not pgs.manual_compact_state_check((first + 3611) * 1000, last, 3600, enq)[0]


def test_parse_compression_types():
    """pegasus_compression_options_test.cpp:115-155 (num_levels = 7): the rocksdb_compression_type setting"""
    import ctypes as C
    L = pgs.lib()
    L.pgs_parse_compression_types.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p]
    none, snappy, lz4, zstd = 0, 1, 4, 7
    H = "per_level:"

    def parse(cfg, before=None):
        buf = (C.c_uint8 * 7)(*(before or [9] * 7))
        return L.pgs_parse_compression_types(cfg.encode(), 7, buf), list(buf)
    ok = [("none", [none] * 7), ("snappy", [none, none] + [snappy] * 5), ("lz4", [none, none] + [lz4] * 5), ("zstd", [none, none] + [zstd] * 5),
          (H + "none", [none] * 7), (H + "none,snappy", [none] + [snappy] * 6), (H + "none,lz4,snappy,zstd", [none, lz4, snappy, zstd, zstd, zstd, zstd]),
          (H + "none,lz4,snappy,zstd,lz4,snappy,zstd", [none, lz4, snappy, zstd, lz4, snappy, zstd]),
          (H + "none,lz4,snappy,zstd,lz4,snappy,zstd,zstd", [none, lz4, snappy, zstd, lz4, snappy, zstd])]
    for cfg, want in ok:
        assert parse(cfg) == (0, want), cfg
    old = [none, lz4, snappy, zstd, lz4, snappy, zstd]
    for cfg in ("none1", "Snappy", ",zstd", H + ":snappy", H + "snappy,snappy1", "per_leve:snappy", "per_levelsnappy", "not_support_zip"):
        assert parse(cfg, old) == (pgs.INVALID_ARGUMENT, old), cfg   # refused, the previous table stays
    # the default the reference starts with (check_rocksdb_compression_types_default, 6 levels shown there)
    assert parse("lz4")[1][:6] == [none, none, lz4, lz4, lz4, lz4]
