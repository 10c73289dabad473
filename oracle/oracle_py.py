"""ctypes binding of the CPU oracle (oracle/libpegasus_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpegasus_oracle.so")
REF_CRC_PATH = os.path.join(_HERE, "_ref", "libref_crc.so")

vp = C.c_void_p
u8p = C.POINTER(C.c_uint8)


class OrcFilterParams(C.Structure):
    _fields_ = [("enabled", C.c_uint8), ("validate_hash", C.c_uint8), ("data_version", C.c_uint32),
                ("default_ttl", C.c_uint32), ("pidx", C.c_int32), ("partition_version", C.c_int32), ("ops", vp)]


class OrcCompactStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("in_records", "out_records", "in_bytes", "out_bytes", "dropped_shadowed",
                                          "dropped_tombstone", "dropped_expired", "dropped_user", "dropped_stale",
                                          "ttl_rewritten")]


class DecodeSizes(C.Structure):
    _fields_ = [("n_records", C.c_uint64), ("key_bytes", C.c_uint64), ("value_bytes", C.c_uint64)]


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.orc_crc64.restype = C.c_uint64
    L.orc_crc64.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.orc_key_hash.restype = C.c_uint64
    L.orc_generate_timetag.restype = C.c_uint64
    L.orc_generate_timetag.argtypes = [C.c_uint64, C.c_uint8, C.c_int32]
    L.orc_extract_timetag.restype = C.c_uint64
    L.orc_extract_expire_ts.restype = C.c_uint32
    for n in ("orc_ops_create", "orc_ops_build", "orc_run_from_records", "orc_run_from_blocks", "orc_compact",
              "orc_blockrun_build", "orc_blockrun_from_blocks", "orc_blockrun_decode", "orc_compact_blocks",
              "orc_rrdb_start", "orc_response_new", "orc_rrdb_dump"):
        getattr(L, n).restype = vp
    L.orc_blockrun_bytes.restype = C.c_uint64
    L.orc_blockruns_get_many.restype = C.c_uint64
    L.orc_blockruns_prefix_scan_many.restype = C.c_uint64
    L.orc_rrdb_last_flushed_decree.restype = C.c_int64
    L.orc_ops_create.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
    L.orc_ops_free.argtypes = [vp]
    L.orc_ops_count.argtypes = [vp]
    L.orc_run_from_records.argtypes = [C.c_uint64, vp, vp, vp, vp, vp, vp]
    L.orc_run_free.argtypes = [vp]
    L.orc_run_sizes.argtypes = [vp, C.POINTER(DecodeSizes)]
    L.orc_run_export.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.orc_run_from_blocks.argtypes = [vp, vp, vp, C.c_uint32, C.POINTER(C.c_int32)]
    L.orc_compact.argtypes = [C.POINTER(vp), C.c_uint32, C.c_int32, C.POINTER(OrcFilterParams), C.c_uint32,
                              C.POINTER(OrcCompactStats)]
    L.orc_blockrun_build.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.orc_blockrun_from_blocks.argtypes = [vp, C.c_uint64, vp, vp, C.c_uint32]
    L.orc_blockrun_free.argtypes = [vp]
    L.orc_blockrun_bytes.argtypes = [vp]
    L.orc_blockrun_blocks.argtypes = [vp]
    L.orc_blockrun_decode.argtypes = [vp]
    L.orc_compact_blocks.argtypes = [C.POINTER(vp), C.c_uint32, C.c_int32, C.POINTER(OrcFilterParams), C.c_uint32,
                                     C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(OrcCompactStats),
                                     C.POINTER(C.c_double)]
    L.orc_blockruns_get_many.argtypes = [C.POINTER(vp), C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.orc_blockruns_prefix_scan_many.argtypes = [C.POINTER(vp), C.c_uint32, vp, vp, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    _lib = L
    return L


def ref_crc():
    """The reference's own crc.cpp compiled into oracle/_ref (None when it was never built)."""
    if not os.path.exists(REF_CRC_PATH):
        return None
    L = C.CDLL(REF_CRC_PATH)
    L.ref_crc64.restype = C.c_uint64
    L.ref_crc64.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64]
    L.ref_crc32.restype = C.c_uint32
    L.ref_crc32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
    return L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(vp)


class Ops:
    def __init__(self, json_text: str, data_version: int = 1):
        raw = json_text.encode()
        self.h = lib().orc_ops_create(raw, len(raw), data_version)

    def __len__(self):
        return int(lib().orc_ops_count(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ops_free(self.h)
            self.h = None


def filter_params(enabled=True, default_ttl=0, validate_hash=False, pidx=0, partition_version=-1, ops: Ops | None = None,
                  data_version=1) -> OrcFilterParams:
    fp = OrcFilterParams()
    fp.enabled = 1 if enabled else 0
    fp.validate_hash = 1 if validate_hash else 0
    fp.data_version = data_version
    fp.default_ttl = default_ttl
    fp.pidx = pidx
    fp.partition_version = partition_version
    fp.ops = ops.h if ops is not None else None
    return fp


class Run:
    """orc_run handle <-> incubator_pegasus_b200.Records"""

    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def from_records(r) -> "Run":
        return Run(lib().orc_run_from_records(r.n, _ptr(r.keys), _ptr(r.key_off), _ptr(r.vals), _ptr(r.val_off),
                                              _ptr(r.seq), _ptr(r.type)))

    @staticmethod
    def from_blocks(run) -> "Run":
        st = C.c_int32()
        h = lib().orc_run_from_blocks(_ptr(run.data), _ptr(run.blk_off), _ptr(run.blk_size), run.n_blocks, C.byref(st))
        if st.value != 0:
            raise RuntimeError(f"oracle block decode: status {st.value}")
        return Run(h)

    def records(self):
        from incubator_pegasus_b200 import Records
        sz = DecodeSizes()
        lib().orc_run_sizes(self.h, C.byref(sz))
        n = sz.n_records
        r = Records(np.zeros(sz.key_bytes, np.uint8), np.zeros(n + 1, np.uint64), np.zeros(sz.value_bytes, np.uint8),
                    np.zeros(n + 1, np.uint64), np.zeros(n, np.uint64), np.zeros(n, np.uint8))
        lib().orc_run_export(self.h, _ptr(r.keys), _ptr(r.key_off), _ptr(r.vals), _ptr(r.val_off), _ptr(r.seq),
                             _ptr(r.type))
        return r

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_run_free(self.h)
            self.h = None


def compact(runs, bottommost: bool, fp: OrcFilterParams, now: int):
    arr = (vp * len(runs))(*[r.h for r in runs])
    st = OrcCompactStats()
    out = Run(lib().orc_compact(arr, len(runs), 1 if bottommost else 0, C.byref(fp), now, C.byref(st)))
    return out, st


class BlockRunCPU:
    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def from_run(run: Run, block_size=4096, restart_interval=16) -> "BlockRunCPU":
        return BlockRunCPU(lib().orc_blockrun_build(run.h, block_size, restart_interval))

    @staticmethod
    def from_blocks(br) -> "BlockRunCPU":
        h = lib().orc_blockrun_from_blocks(_ptr(br.data), br.data.shape[0], _ptr(br.blk_off), _ptr(br.blk_size),
                                           br.n_blocks)
        if not h:
            raise RuntimeError("oracle: corrupt blocks")
        return BlockRunCPU(h)

    def decode(self) -> Run:
        return Run(lib().orc_blockrun_decode(self.h))

    @property
    def nbytes(self):
        return int(lib().orc_blockrun_bytes(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_blockrun_free(self.h)
            self.h = None


def compact_blocks(runs, bottommost: bool, fp: OrcFilterParams, now: int, threads: int = 1):
    arr = (vp * len(runs))(*[r.h for r in runs])
    st = OrcCompactStats()
    secs = C.c_double()
    out = BlockRunCPU(lib().orc_compact_blocks(arr, len(runs), 1 if bottommost else 0, C.byref(fp), now, threads, 4096,
                                               16, C.byref(st), C.byref(secs)))
    return out, st, secs.value


def get_many(runs, keys: np.ndarray, key_off: np.ndarray, now: int, threads: int = 1):
    arr = (vp * len(runs))(*[r.h for r in runs])
    vb = C.c_uint64()
    secs = C.c_double()
    found = lib().orc_blockruns_get_many(arr, len(runs), _ptr(keys), _ptr(key_off), key_off.shape[0] - 1, now, threads,
                                         C.byref(vb), C.byref(secs))
    return int(found), int(vb.value), secs.value


def prefix_scan_many(runs, hashkeys: np.ndarray, hk_off: np.ndarray, now: int, threads: int = 1):
    arr = (vp * len(runs))(*[r.h for r in runs])
    nb = C.c_uint64()
    secs = C.c_double()
    cnt = lib().orc_blockruns_prefix_scan_many(arr, len(runs), _ptr(hashkeys), _ptr(hk_off), hk_off.shape[0] - 1, now,
                                               threads, C.byref(nb), C.byref(secs))
    return int(cnt), int(nb.value), secs.value
