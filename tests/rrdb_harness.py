"""One harness, two backends with the same C structs: the product's rrdb surface (pgs_rrdb_*, CUDA engine)
and the oracle's (orc_rrdb_*, CPU model).  Tests issue identical requests to both and compare responses."""
from __future__ import annotations

import ctypes as C

import incubator_pegasus_b200 as pgs
from incubator_pegasus_b200 import (Blob, FullKey, GetScannerRequest, MultiGetRequest, Response, ServerOptions)

vp = C.c_void_p


def blob(b: bytes, keep: list) -> Blob:
    buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if b else b"\0")
    keep.append(buf)
    return Blob(C.cast(buf, C.POINTER(C.c_uint8)), len(b))


def raw_key(hk: bytes, sk: bytes) -> bytes:
    return len(hk).to_bytes(2, "big") + hk + sk


def next_blob(b: bytes) -> bytes:
    b = bytearray(b)
    while b[-1] == 0xFF:
        b.pop()
    b[-1] += 1
    return bytes(b)


class WriteRequest(C.Structure):
    _fields_ = [("op", C.c_uint32), ("raw_key", Blob), ("value", Blob), ("expire_ts_seconds", C.c_uint32)]


class Mutate(C.Structure):
    _fields_ = [("operation", C.c_uint32), ("sort_key", Blob), ("value", Blob), ("set_expire_ts_seconds", C.c_int32)]


class CheckAndMutateRequest(C.Structure):
    _fields_ = [("hash_key", Blob), ("check_sort_key", Blob), ("check_type", C.c_int32), ("check_operand", Blob),
                ("mutate_list", C.POINTER(Mutate)), ("n_mutate", C.c_uint32), ("return_check_value", C.c_uint8)]


class CheckAndSetRequest(C.Structure):
    _fields_ = [("hash_key", Blob), ("check_sort_key", Blob), ("check_type", C.c_int32), ("check_operand", Blob),
                ("set_diff_sort_key", C.c_uint8), ("set_sort_key", Blob), ("set_value", Blob), ("set_expire_ts_seconds", C.c_int32),
                ("return_check_value", C.c_uint8)]


class CasResult(C.Structure):
    _fields_ = [("error", C.c_int32), ("check_value_returned", C.c_uint8), ("check_value_exist", C.c_uint8), ("reserved", C.c_uint8 * 2),
                ("check_value_len", C.c_uint32)]


class Backend:
    def __init__(self, kind: str, engine=None, app_id=1, pidx=0, opts: dict | None = None, envs: dict | None = None):
        self.kind = kind
        if kind == "gpu":
            self.L = pgs.lib()
            self.p = "pgs_"
        else:
            import oracle_py
            self.L = oracle_py.lib()
            self.p = "orc_"
        L, p = self.L, self.p
        so = ServerOptions()
        so.prefix_filter = 1
        for k, v in (opts or {}).items():
            setattr(so, k, v)
        env_blob, n_env = self._envs(envs or {})
        self.h = vp()
        if kind == "gpu":
            st = L.pgs_rrdb_start(engine.h, app_id, pidx, C.byref(so), env_blob, n_env, C.byref(self.h))
            assert st == 0, st
        else:
            L.orc_rrdb_start.restype = vp
            L.orc_rrdb_start.argtypes = [C.c_int32, C.c_int32, C.POINTER(ServerOptions), C.c_char_p, C.c_uint32]
            self.h = vp(L.orc_rrdb_start(app_id, pidx, C.byref(so), env_blob, n_env))
        f = lambda name: getattr(L, p + name)
        self.f = f
        f("response_new").restype = vp
        f("response_new").argtypes = []
        f("response_free").argtypes = [vp]
        f("response_view").restype = C.POINTER(Response)
        f("response_view").argtypes = [vp]
        sigs = {
            "rrdb_get": [vp, Blob, C.c_uint32, vp], "rrdb_ttl": [vp, Blob, C.c_uint32, vp],
            "rrdb_multi_get": [vp, C.POINTER(MultiGetRequest), C.c_uint32, vp],
            "rrdb_batch_get": [vp, C.POINTER(FullKey), C.c_uint32, C.c_uint32, vp],
            "rrdb_sortkey_count": [vp, Blob, C.c_uint32, vp],
            "rrdb_get_scanner": [vp, C.POINTER(GetScannerRequest), C.c_uint32, vp],
            "rrdb_scan": [vp, C.c_int64, C.c_uint32, vp], "rrdb_clear_scanner": [vp, C.c_int64],
            "rrdb_put": [vp, Blob, Blob, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32],
            "rrdb_remove": [vp, Blob, C.c_int64, C.c_uint32],
            "rrdb_on_batched_writes": [vp, C.POINTER(WriteRequest), C.c_uint32, C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(C.c_int32)],
            "rrdb_check_and_set": [vp, C.POINTER(CheckAndSetRequest), C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(CasResult), vp, C.c_uint32],
            "rrdb_check_and_mutate": [vp, C.POINTER(CheckAndMutateRequest), C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(CasResult), vp, C.c_uint32],
            "rrdb_incr": [vp, Blob, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int64)],
            "rrdb_multi_put": [vp, Blob, C.POINTER(Blob), C.POINTER(Blob), C.c_uint32, C.c_uint32, C.c_int64,
                               C.c_uint64, C.c_uint32],
            "rrdb_multi_remove": [vp, Blob, C.POINTER(Blob), C.c_uint32, C.c_int64, C.POINTER(C.c_int64), C.c_uint32],
            "rrdb_flush": [vp, C.c_uint32], "rrdb_manual_compact": [vp, C.c_uint32, vp],
            "rrdb_update_app_envs": [vp, C.c_char_p, C.c_uint32, C.c_uint32],
            "rrdb_set_partition_version": [vp, C.c_int32], "rrdb_stop": [vp],
            "rrdb_last_flushed_decree": [vp], "rrdb_last_committed_decree": [vp], "rrdb_gc": [vp, C.c_uint32],
            "rrdb_sync_checkpoint": [vp, C.c_char_p, C.c_uint32, C.POINTER(C.c_int64)], "rrdb_last_durable_decree": [vp],
            "rrdb_apply_checkpoint": [vp, C.c_char_p],
        }
        if kind != "gpu":  # checkpoints exist on the product side only
            for n in ("rrdb_sync_checkpoint", "rrdb_last_durable_decree", "rrdb_apply_checkpoint"):
                sigs.pop(n)
        for n, a in sigs.items():
            f(n).argtypes = a
        if kind == "gpu":
            f("rrdb_last_durable_decree").restype = C.c_int64
        f("rrdb_last_flushed_decree").restype = C.c_int64
        f("rrdb_last_committed_decree").restype = C.c_int64
        f("rrdb_gc").restype = C.c_uint32
        f("rrdb_clear_scanner").restype = None
        f("rrdb_set_partition_version").restype = None
        f("rrdb_stop").restype = None
        self.resp = vp(f("response_new")())
        self.decree = 0

    def reader(self) -> "Backend":
        """a second handle on the same server with its own response object, for a concurrent reader thread"""
        import copy
        r = copy.copy(self)
        r.resp = C.c_void_p(self.f("response_new")())
        r.is_reader = True
        return r

    @staticmethod
    def _envs(envs: dict):
        b = b"".join(k.encode() + b"\0" + v.encode() + b"\0" for k, v in envs.items())
        return (b if b else None), len(envs)

    def close(self):
        if self.h:
            self.f("response_free")(self.resp)
            if not getattr(self, "is_reader", False):
                self.f("rrdb_stop")(self.h)
            self.h = None

    # ---- writes ----
    def put(self, hk, sk, value, expire_ts=0, now=0, ts_us=1):
        keep = []
        self.decree += 1
        return self.f("rrdb_put")(self.h, blob(raw_key(hk, sk), keep), blob(value, keep), expire_ts, self.decree, ts_us, now)

    def incr(self, hk, sk, increment, expire_ts_seconds=0, now=0, ts_us=1):
        """-> (return code, response error, response new_value)"""
        keep = []
        self.decree += 1
        e, v = C.c_int32(-1), C.c_int64(0)
        rc = self.f("rrdb_incr")(self.h, blob(raw_key(hk, sk), keep), increment, expire_ts_seconds, self.decree, ts_us, now, C.byref(e), C.byref(v))
        return rc, e.value, v.value

    def batched_writes(self, ops, now=0, ts_us=1):
        """ops: list of ("put", hk, sk, value, expire_ts) / ("remove", hk, sk); one decree -> (apply status, [response errors])"""
        keep = []
        self.decree += 1
        arr = (WriteRequest * max(1, len(ops)))()
        for i, o in enumerate(ops):
            arr[i].op = {"put": 0, "remove": 1}.get(o[0], 7)
            arr[i].raw_key = blob(raw_key(o[1], o[2]), keep)
            arr[i].value = blob(o[3] if len(o) > 3 else b"", keep)
            arr[i].expire_ts_seconds = o[4] if len(o) > 4 else 0
        errs = (C.c_int32 * max(1, len(ops)))(*([-1] * max(1, len(ops))))
        rc = self.f("rrdb_on_batched_writes")(self.h, arr, len(ops), self.decree, ts_us, now, errs)
        return rc, list(errs[:len(ops)])

    def _cas_out(self, rc, res, buf):
        return {"rc": rc, "error": res.error, "returned": bool(res.check_value_returned), "exist": bool(res.check_value_exist),
                "check_value": bytes(buf[:res.check_value_len]) if res.check_value_exist else None}

    def check_and_set(self, hk, check_sk, check_type, operand, set_sk, set_value, return_check_value=True, ttl_ts=0, now=0, ts_us=1):
        keep = []
        self.decree += 1
        q = CheckAndSetRequest()
        q.hash_key, q.check_sort_key, q.check_type, q.check_operand = blob(hk, keep), blob(check_sk, keep), check_type, blob(operand, keep)
        q.set_diff_sort_key, q.set_sort_key, q.set_value = int(set_sk != check_sk), blob(set_sk, keep), blob(set_value, keep)
        q.set_expire_ts_seconds, q.return_check_value = ttl_ts, int(return_check_value)
        res, buf = CasResult(), (C.c_uint8 * 4096)()
        rc = self.f("rrdb_check_and_set")(self.h, C.byref(q), self.decree, ts_us, now, C.byref(res), buf, 4096)
        return self._cas_out(rc, res, buf)

    def check_and_mutate(self, hk, check_sk, check_type, operand, mutations, return_check_value=True, now=0, ts_us=1):
        """mutations: list of ("put", sort_key, value, expire_ts) / ("del", sort_key)"""
        keep = []
        self.decree += 1
        arr = (Mutate * max(1, len(mutations)))()
        for i, m in enumerate(mutations):
            arr[i].operation = {"put": 0, "del": 1}.get(m[0], m[0] if isinstance(m[0], int) else 9)
            arr[i].sort_key = blob(m[1], keep)
            arr[i].value = blob(m[2] if len(m) > 2 else b"", keep)
            arr[i].set_expire_ts_seconds = m[3] if len(m) > 3 else 0
        q = CheckAndMutateRequest()
        q.hash_key, q.check_sort_key, q.check_type, q.check_operand = blob(hk, keep), blob(check_sk, keep), check_type, blob(operand, keep)
        q.mutate_list, q.n_mutate, q.return_check_value = arr, len(mutations), int(return_check_value)
        res, buf = CasResult(), (C.c_uint8 * 4096)()
        rc = self.f("rrdb_check_and_mutate")(self.h, C.byref(q), self.decree, ts_us, now, C.byref(res), buf, 4096)
        return self._cas_out(rc, res, buf)

    def remove(self, hk, sk, now=0):
        keep = []
        self.decree += 1
        return self.f("rrdb_remove")(self.h, blob(raw_key(hk, sk), keep), self.decree, now)

    def multi_put(self, hk, kvs: dict, expire_ts=0, now=0, ts_us=1):
        keep = []
        self.decree += 1
        items = list(kvs.items())
        sks = (Blob * max(1, len(items)))(*[blob(k, keep) for k, _ in items])
        vals = (Blob * max(1, len(items)))(*[blob(v, keep) for _, v in items])
        return self.f("rrdb_multi_put")(self.h, blob(hk, keep), sks, vals, len(items), expire_ts, self.decree, ts_us, now)

    def multi_remove(self, hk, sort_keys, now=0):
        keep = []
        self.decree += 1
        sks = (Blob * max(1, len(sort_keys)))(*[blob(k, keep) for k in sort_keys])
        cnt = C.c_int64()
        st = self.f("rrdb_multi_remove")(self.h, blob(hk, keep), sks, len(sort_keys), self.decree, C.byref(cnt), now)
        return st, cnt.value

    def flush(self, now=0):
        return self.f("rrdb_flush")(self.h, now)

    def manual_compact(self, now=0):
        return self.f("rrdb_manual_compact")(self.h, now, None)

    def update_envs(self, envs: dict, now=0):
        b, n = self._envs(envs)
        return self.f("rrdb_update_app_envs")(self.h, b, n, now)

    def set_partition_version(self, pv):
        self.f("rrdb_set_partition_version")(self.h, pv)

    # ---- reads ----
    def _view(self, with_hk=False):
        v = self.f("response_view")(self.resp).contents
        arena = bytes(C.cast(v.arena, C.POINTER(C.c_uint8 * v.arena_len)).contents) if v.arena_len else b""
        kvs = []
        for i in range(v.n_kvs):
            kv = v.kvs[i]
            item = (arena[kv.key_off:kv.key_off + kv.key_len], arena[kv.value_off:kv.value_off + kv.value_len], kv.expire_ts)
            if with_hk:
                item = item + (v.hk_len[i],)
            kvs.append(item)
        ctx = v.context_id
        return {"error": v.error, "kvs": kvs, "count": v.count, "ttl": v.ttl_seconds, "kv_count": v.kv_count,
                "context_id": ctx, "ctx_class": "valid" if ctx >= 0 and False else None, "expire_count": v.expire_count,
                "filter_count": v.filter_count, "iteration_count": v.iteration_count, "app_id": v.app_id,
                "partition_index": v.partition_index}

    def get(self, hk, sk, now=0):
        keep = []
        self.f("rrdb_get")(self.h, blob(raw_key(hk, sk), keep), now, self.resp)
        return self._view()

    def ttl(self, hk, sk, now=0):
        keep = []
        self.f("rrdb_ttl")(self.h, blob(raw_key(hk, sk), keep), now, self.resp)
        return self._view()

    def multi_get(self, hk, start=b"", stop=b"", start_inclusive=True, stop_inclusive=False, sort_keys=None, max_kv_count=0,
                  max_kv_size=0, no_value=False, reverse=False, filter_type=0, filter_pattern=b"", now=0):
        keep = []
        q = MultiGetRequest()
        q.hash_key = blob(hk, keep)
        sks = sort_keys or []
        arr = (Blob * max(1, len(sks)))(*[blob(k, keep) for k in sks])
        keep.append(arr)
        q.sort_keys = arr
        q.n_sort_keys = len(sks)
        q.max_kv_count, q.max_kv_size = max_kv_count, max_kv_size
        q.no_value, q.start_inclusive, q.stop_inclusive, q.reverse = int(no_value), int(start_inclusive), int(stop_inclusive), int(reverse)
        q.start_sortkey, q.stop_sortkey = blob(start, keep), blob(stop, keep)
        q.sort_key_filter_type, q.sort_key_filter_pattern = filter_type, blob(filter_pattern, keep)
        self.f("rrdb_multi_get")(self.h, C.byref(q), now, self.resp)
        return self._view()

    def batch_get(self, keys, now=0):
        keep = []
        arr = (FullKey * max(1, len(keys)))(*[FullKey(blob(h, keep), blob(s, keep)) for h, s in keys])
        self.f("rrdb_batch_get")(self.h, arr, len(keys), now, self.resp)
        return self._view(with_hk=True)

    def sortkey_count(self, hk, now=0):
        keep = []
        self.f("rrdb_sortkey_count")(self.h, blob(hk, keep), now, self.resp)
        return self._view()

    def get_scanner(self, start_key, stop_key, start_inclusive=True, stop_inclusive=False, batch_size=0, no_value=False,
                    hash_filter=(0, b""), sort_filter=(0, b""), validate_partition_hash=True, return_expire_ts=False,
                    full_scan=False, only_return_count=False, now=0):
        keep = []
        q = GetScannerRequest()
        q.start_key, q.stop_key = blob(start_key, keep), blob(stop_key, keep)
        q.start_inclusive, q.stop_inclusive, q.no_value = int(start_inclusive), int(stop_inclusive), int(no_value)
        q.validate_partition_hash, q.return_expire_ts = int(validate_partition_hash), int(return_expire_ts)
        q.full_scan, q.only_return_count = int(full_scan), int(only_return_count)
        q.batch_size = batch_size
        q.hash_key_filter_type, q.hash_key_filter_pattern = hash_filter[0], blob(hash_filter[1], keep)
        q.sort_key_filter_type, q.sort_key_filter_pattern = sort_filter[0], blob(sort_filter[1], keep)
        self.f("rrdb_get_scanner")(self.h, C.byref(q), now, self.resp)
        return self._view()

    def scan(self, context_id, now=0):
        self.f("rrdb_scan")(self.h, context_id, now, self.resp)
        return self._view()

    def clear_scanner(self, context_id):
        self.f("rrdb_clear_scanner")(self.h, context_id)

    # client-side conveniences (pegasus_client_impl.cpp:1135-1192 builds the same keys)
    def scan_all(self, hk, start_sk=b"", stop_sk=b"", batch_size=0, now=0, **kw):
        start = raw_key(hk, start_sk)
        stop = next_blob(raw_key(hk, b"")) if not stop_sk else raw_key(hk, stop_sk)
        r = self.get_scanner(start, stop, batch_size=batch_size, now=now, **kw)
        out, batches = list(r["kvs"]), [r]
        guard = 0
        while r["error"] == 0 and r["context_id"] >= 0 and guard < 100000:
            r = self.scan(r["context_id"], now=now)
            out += r["kvs"]
            batches.append(r)
            guard += 1
        return out, batches


def same_response(a: dict, b: dict, ignore_ctx_value=True):
    """compare two handler responses; context ids are compared by class (>=0 parked, -1 completed, 0 unset)"""
    ka = dict(a)
    kb = dict(b)
    if ignore_ctx_value:
        for d in (ka, kb):
            c = d["context_id"]
            d["context_id"] = "parked" if c > 0 else c
    return ka == kb, (ka, kb)
