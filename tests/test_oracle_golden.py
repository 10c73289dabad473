"""CPU: the oracle against the golden tables of the reference's own tests (tests/golden/*.json) and the
product's host-side pieces (key schema, ops parser) against the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from rrdb_harness import Backend, raw_key

GOLD = os.path.join(os.path.dirname(__file__), "golden")
T = json.load(open(os.path.join(GOLD, "tables.json")))
MG = json.load(open(os.path.join(GOLD, "multi_get_basic.json")))
NOW = 200_000_000


def b(s):
    return s.encode("latin-1") if isinstance(s, str) else s


def test_pattern_rule_tables(oracle):
    L = oracle.lib()
    for name in ("hashkey_pattern_rule", "sortkey_pattern_rule"):
        for value, pattern, mt, want in T[name]["rows"]:
            assert bool(L.orc_string_pattern_match(b(value), len(value), mt, b(pattern), len(pattern))) == want, (name, value, pattern, mt)


def test_ttl_range_rule_table(oracle):
    L = oracle.lib()
    L.orc_ttl_range_rule_match.argtypes = [C.c_uint32] * 4
    for start, stop, rel, want in T["ttl_range_rule"]["rows"]:
        assert bool(L.orc_ttl_range_rule_match(start, stop, (rel + NOW) & 0xFFFFFFFF, NOW)) == want


def test_rule_create(oracle):
    L = oracle.lib()
    for row in T["rule_create"]["rows"]:
        pat = C.create_string_buffer(64)
        mt, st, sp = C.c_int32(), C.c_uint32(), C.c_uint32()
        ok = L.orc_rule_create(row["type"], b(row["params"]), len(row["params"]), pat, 64, C.byref(mt), C.byref(st), C.byref(sp))
        assert bool(ok) == row["ok"], row
        if row["ok"]:
            if "pattern" in row:
                assert pat.value.decode() == row["pattern"] and mt.value == row["match_type"]
            else:
                assert (st.value, sp.value) == (row["start_ttl"], row["stop_ttl"])


def _value(oracle, expire_ts, data=b""):
    buf = (C.c_uint8 * (12 + len(data)))()
    n = oracle.lib().orc_generate_value(1, C.c_uint32(expire_ts), C.c_uint64(0), data, len(data), buf, len(buf))
    return bytes(buf[:n])


def _one_op(oracle, op_type, ttl_type, ttl_value, rules):
    L = oracle.lib()
    n = len(rules)
    rt = (C.c_int32 * n)(*[r[0] for r in rules])
    mt = (C.c_int32 * n)(*[r[1] for r in rules])
    pats = (C.c_char_p * n)(*[b(r[2]) for r in rules])
    st = (C.c_uint32 * n)(*[r[3] for r in rules])
    sp = (C.c_uint32 * n)(*[r[4] for r in rules])
    L.orc_ops_build.argtypes = [C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_uint32]
    return C.c_void_p(L.orc_ops_build(op_type, ttl_type, ttl_value, n, rt, mt, pats, st, sp, 1))


def test_all_rules_match_table(oracle):
    L = oracle.lib()
    L.orc_op_all_rules_match.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p,
                                         C.c_uint32, C.c_uint32]
    for want, hk, sk, ttl, hp, hm, sp_, sm, start, stop in T["all_rules_match"]["rows"]:
        op = _one_op(oracle, 1, 3, 0, [(0, hm, hp, 0, 0), (1, sm, sp_, 0, 0), (2, 3, "", start, stop)])
        v = _value(oracle, ttl + NOW)
        assert bool(L.orc_op_all_rules_match(op, 0, b(hk), len(hk), b(sk), len(sk), v, len(v), NOW)) == want
    op = _one_op(oracle, 0, 0, 0, [])
    assert bool(L.orc_op_all_rules_match(op, 0, b"hash", 4, b"sort", 4, b"", 0, NOW)) == T["all_rules_match"]["empty_rules_match"]


def test_delete_and_update_ttl_tables(oracle):
    L = oracle.lib()
    L.orc_op_filter.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32,
                                C.c_uint32, C.c_void_p, C.POINTER(C.c_int32)]
    for want, hk, pat, mt in T["delete_key_filter"]["rows"]:
        op = _one_op(oracle, 1, 3, 0, [(0, mt, pat, 0, 0)])
        ch = C.c_int32()
        nv = (C.c_uint8 * 16)()
        assert bool(L.orc_op_filter(op, 0, b(hk), len(hk), b"", 0, b"", 0, NOW, nv, C.byref(ch))) == want
    eb = T["update_ttl_filter"]["epoch_begin"]
    for changed, expect_ts, hk, ets, pat, mt, op_type, value in T["update_ttl_filter"]["rows"]:
        op = _one_op(oracle, 0, op_type, value, [(0, mt, pat, 0, 0)])
        v = _value(oracle, ets)
        ch = C.c_int32()
        nv = (C.c_uint8 * len(v))()
        assert L.orc_op_filter(op, 0, b(hk), len(hk), b"", 0, v, len(v), NOW, nv, C.byref(ch)) == 0  # never deletes
        assert bool(ch.value) == changed
        if changed:
            new_ts = int.from_bytes(bytes(nv[:4]), "big")
            if op_type == 2:
                assert new_ts + eb == expect_ts
            elif op_type == 1:
                assert new_ts == expect_ts
            else:
                assert new_ts == expect_ts + NOW


def _describe_ops(oracle, json_text):
    L = oracle.lib()
    ops = oracle.Ops(json_text)
    out = []
    for i in range(len(ops)):
        ot, tt, tv, nr = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32()
        L.orc_ops_describe(C.c_void_p(ops.h), i, C.byref(ot), C.byref(tt), C.byref(tv), C.byref(nr))
        rules = []
        for r in range(nr.value):
            rt, mt, st, sp = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32()
            pat = C.create_string_buffer(64)
            L.orc_ops_describe_rule(C.c_void_p(ops.h), i, r, C.byref(rt), C.byref(mt), pat, 64, C.byref(st), C.byref(sp))
            rules.append({"type": rt.value, "pattern": pat.value.decode(), "match_type": mt.value, "start_ttl": st.value, "stop_ttl": sp.value})
        out.append({"op": ot.value, "ttl_type": tt.value, "ttl_value": tv.value, "rules": rules})
    return out


def test_create_operations_json(oracle, pgs):
    g = T["create_operations"]
    got = _describe_ops(oracle, g["json"])
    assert len(got) == len(g["expect"])
    for a, e in zip(got, g["expect"]):
        assert a["op"] == e["op"] and len(a["rules"]) == len(e["rules"])
        if e["op"] == 0:
            assert (a["ttl_type"], a["ttl_value"]) == (e["ttl_type"], e["ttl_value"])
        for ra, re_ in zip(a["rules"], e["rules"]):
            for k, v in re_.items():
                assert ra[k] == v
    assert _describe_ops(oracle, "") == []
    # creator tables
    L = oracle.lib()
    for row in T["update_ttl_creator"]["rows"]:
        t, v = C.c_int32(), C.c_uint32()
        ok = L.orc_update_ttl_create(b(row["params"]), len(row["params"]), C.byref(t), C.byref(v))
        assert bool(ok) == row["ok"]
        if row["ok"]:
            assert (t.value, v.value) == (row["type"], row["value"])
    # the product's parser produces the same table (binary form)
    ops = pgs.parse_ops(g["json"])
    assert int.from_bytes(ops[:4].tobytes(), "little") == 2
    assert int.from_bytes(pgs.parse_ops("")[:4].tobytes(), "little") == 0
    assert int.from_bytes(pgs.parse_ops('{"ops":[{"type":"COT_DELETE","params":"","rules":[]}]}')[:4].tobytes(), "little") == 0


def test_value_schema_roundtrip(oracle):
    L = oracle.lib()
    L.orc_generate_value.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.orc_extract_expire_ts.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32]
    L.orc_extract_timetag.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32]
    for version, ets, tag, data in T["value_schema_roundtrip"]["rows"]:
        buf = (C.c_uint8 * 64)()
        n = L.orc_generate_value(version, ets, tag, b(data), len(data), buf, 64)
        raw = bytes(buf[:n])
        hdr = L.orc_user_data_offset(version)
        assert hdr == (12 if version == 1 else 4) and n == hdr + len(data)
        assert L.orc_extract_expire_ts(version, raw, n) == ets
        assert raw[:4] == ets.to_bytes(4, "big")
        if version == 1:
            assert L.orc_extract_timetag(1, raw, n) == tag and raw[4:12] == tag.to_bytes(8, "big")
        assert raw[hdr:] == b(data)
        u = T["value_schema_roundtrip"]["update_expire_ts"]
        m = (C.c_uint8 * n).from_buffer_copy(raw)
        L.orc_update_expire_ts(version, m, n, u["to"])
        assert L.orc_extract_expire_ts(version, bytes(m), n) == u["to"] and bytes(m)[4:] == raw[4:]
    # timetag layout: ts_us << 8 | cluster_id << 1 | deleted  (pegasus_value_schema.h:44-47)
    assert L.orc_generate_timetag(0x1234, 5, 1) == (0x1234 << 8) | (5 << 1) | 1


def test_key_schema_and_transform(oracle, pgs):
    L, P = oracle.lib(), pgs.lib()
    rng = np.random.default_rng(9)
    cases = [(b"h1", b"s1"), (b"", b"sk"), (b"hk", b""), (b"\xff\xff", b"\xff"), (b"a" * 300, b"b" * 5), (b"\xff" * 3, b"")]
    cases += [(bytes(rng.integers(0, 256, rng.integers(0, 20), dtype=np.uint8)), bytes(rng.integers(250, 256, rng.integers(0, 6), dtype=np.uint8)))
              for _ in range(200)]
    for hk, sk in cases:
        want = raw_key(hk, sk)
        for lib_, name in ((L, "orc_generate_key"), (P, "pgs_generate_key")):
            buf = (C.c_uint8 * 600)()
            n = getattr(lib_, name)(hk, len(hk), sk, len(sk), buf, 600)
            assert bytes(buf[:n]) == want
        for with_sk in (0, 1):
            src = bytearray(raw_key(hk, sk if with_sk else b""))
            while src[-1] == 0xFF:
                src.pop()
            src[-1] += 1
            for lib_, name in ((L, "orc_generate_next_blob"), (P, "pgs_generate_next_blob")):
                buf = (C.c_uint8 * 600)()
                n = getattr(lib_, name)(hk, len(hk), sk, len(sk), with_sk, buf, 600)
                got = bytes(buf[:n])
                assert got == bytes(src)
                assert got > raw_key(hk, sk if with_sk else b"")  # strictly after every key with that prefix
        assert L.orc_key_hash(want, len(want)) == P.pgs_key_hash(want, len(want))
        assert L.orc_hashkey_transform(want, len(want)) == 2 + len(hk)
    assert L.orc_hashkey_transform(b"x", 1) == -1
    for h1, s1, op, h2, s2 in T["hashkey_transform"]["ordering"]:
        a, c = raw_key(b(h1), b(s1)), raw_key(b(h2), b(s2))
        assert (a < c) if op == "<" else (a > c)


# ---- rrdb semantics on the oracle: the reference's function-test tables ------------------------------------
@pytest.fixture()
def orc_db():
    be = Backend("oracle")
    yield be
    be.close()


@pytest.mark.parametrize("direction", ["forward", "reverse"])
def test_multi_get_basic_tables_oracle(orc_db, direction):
    g = MG[direction]
    hk = b"basic_test_multi_get"
    assert orc_db.multi_put(hk, {b(k): b(v) for k, v in g["fixture"]}) == 0
    assert orc_db.sortkey_count(hk)["count"] == 13
    for case in g["cases"]:
        o = case["options"]
        r = orc_db.multi_get(hk, b(case["start"]), b(case["stop"]), o["start_inclusive"], o["stop_inclusive"],
                             max_kv_count=case["max_count"], max_kv_size=1000000, reverse=o["reverse"],
                             filter_type=o["sort_key_filter_type"], filter_pattern=b(o["sort_key_filter_pattern"]))
        assert r["error"] == case["error"], case["title"]
        assert sorted([k.decode("latin-1"), v.decode("latin-1")] for k, v, _ in r["kvs"]) == case["expect"], case["title"]
        assert [k for k, _, _ in r["kvs"]] == sorted(k for k, _, _ in r["kvs"])  # ascending by sort key
    # the "set a expired value" case (test_basic.cpp:568-578)
    if direction == "forward":
        orc_db.put(hk, b"", b"expire_value", expire_ts=NOW - 1, now=NOW - 5)
        r = orc_db.multi_get(hk, max_kv_count=2, now=NOW)
        assert r["error"] == 7 and [k for k, _, _ in r["kvs"]] == [b"1", b"1-abcdefg"]
    st, cnt = orc_db.multi_remove(hk, [b(k) for k, _ in g["fixture"]])
    assert (st, cnt) == (0, 13)
    assert orc_db.sortkey_count(hk)["count"] == 0


def _prepare_range_read(be, total, expired):
    hk = b"range_read_hashkey"
    if expired:
        be.multi_put(hk, {b"1-%d" % i: b"value" for i in range(expired)}, expire_ts=NOW - 10, now=NOW - 20)
    if total > expired:
        be.multi_put(hk, {b"2-%d" % i: b"value" for i in range(expired, total)})
    return hk


def test_range_read_tables_oracle():
    for exp, total, max_count, want_err, want_n in T["range_read_multiget"]["rows"]:
        be = Backend("oracle")
        hk = _prepare_range_read(be, total, exp)
        r = be.multi_get(hk, max_kv_count=max_count, max_kv_size=1000000, now=NOW)
        assert (r["error"], len(r["kvs"])) == (want_err, want_n), (exp, total, max_count)
        be.close()
    for exp, total, want_err, want_n in T["range_read_sortkey_count"]["rows"]:
        be = Backend("oracle")
        hk = _prepare_range_read(be, total, exp)
        r = be.sortkey_count(hk, now=NOW)
        assert (r["error"], r["count"]) == (want_err, want_n)
        be.close()
    for exp, total, batch, want_n in T["range_read_scan"]["rows"]:
        be = Backend("oracle")
        hk = _prepare_range_read(be, total, exp)
        kvs, batches = be.scan_all(hk, batch_size=batch, now=NOW)
        # the client keeps calling next() until total-expired records came back (test_range_read.cpp:75-103)
        assert len(kvs) == total - exp
        assert all(x["error"] == 0 for x in batches)
        be.close()
