"""oracle/sst_py.py — TEST INFRASTRUCTURE ONLY: an independent pure-Python restatement of RocksDB's BlockBasedTable file
layout (format_version 2, no compression), used to check the product's SST egress / ingest
(incubator_pegasus_b200/host/sst_format.cpp).  Small cases only.

Follows RocksDB 8.5.3's public format description (table/block_based/block_based_table_builder.cc,
table/format.cc, util/crc32c.h, util/hash.cc, util/bloom_impl.h, none of which are in the reference tree).  Pinned pieces:
crc32c by the RFC 3720 vectors, the filter hash by the known answers of RocksDB's own util/hash_test.cc (HashTest.Values), the
LZ4 codec by the system liblz4 (tests/test_sst_format.py).  PARITY UNPINNED for the container itself (block handles, metaindex,
properties, footer): no RocksDB build is available in this image, so these files have not been read by RocksDB.

  data block   : entries  varint32 shared | varint32 non_shared | varint32 value_len | key delta | value,
                 restart array (fixed32 each) | fixed32 count
  block trailer: 1 byte compression type (0) | fixed32 masked crc32c(block | type)
  index block  : restart interval 1; key = last internal key of the data block, value = varint64 offset | varint64 size
  filter block : legacy full Bloom filter: lines of 64 bytes (odd count), 6 probes, | 1 byte probes | fixed32 lines
  metaindex    : "fullfilter.rocksdb.BuiltinBloomFilter" -> handle, "rocksdb.properties" -> handle
  footer (53 B): checksum type 1 | metaindex handle | index handle | zero padding to 41 | fixed32 version 2 | magic
"""
from __future__ import annotations

import struct

MAGIC = 0x88E241B785F4CFF7
FILTER_NAME = b"fullfilter.rocksdb.BuiltinBloomFilter"
PROPS_NAME = b"rocksdb.properties"

_T = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (0x82F63B78 if _c & 1 else 0)
    _T.append(_c)


def crc32c(data: bytes, crc: int = 0) -> int:
    crc ^= 0xFFFFFFFF
    for b in data:
        crc = (crc >> 8) ^ _T[(crc ^ b) & 0xFF]
    return crc ^ 0xFFFFFFFF


def mask(crc: int) -> int:
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v: int) -> bytes:
    out = bytearray()
    while v >= 128:
        out.append((v & 127) | 128)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(b: bytes, p: int):
    r = s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 127) << s
        if not x & 128:
            return r, p
        s += 7


def lz4_block_decode(b: bytes, raw_size: int) -> bytes:
    """LZ4 block format (lz4_Block_format.md): token | literal length bytes | literals | offset LE16 | match length bytes"""
    out, p = bytearray(), 0
    while p < len(b):
        tok = b[p]; p += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                x = b[p]; p += 1; lit += x
                if x != 255:
                    break
        out += b[p:p + lit]; p += lit
        if p >= len(b):
            break
        off = b[p] | (b[p + 1] << 8); p += 2
        ml = tok & 15
        if ml == 15:
            while True:
                x = b[p]; p += 1; ml += x
                if x != 255:
                    break
        ml += 4
        assert 0 < off <= len(out)
        for _ in range(ml):
            out.append(out[-off])
    assert len(out) == raw_size
    return bytes(out)


def build_block(entries, restart_interval: int) -> bytes:
    buf, restarts, last, counter = bytearray(), [0], b"", 0
    for k, v in entries:
        shared = 0
        if counter < restart_interval:
            m = min(len(last), len(k))
            while shared < m and last[shared] == k[shared]:
                shared += 1
        else:
            restarts.append(len(buf))
            counter = 0
        buf += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        last = k
        counter += 1
    for r in restarts:
        buf += struct.pack("<I", r)
    buf += struct.pack("<I", len(restarts))
    return bytes(buf)


def parse_block(b: bytes):
    n = struct.unpack_from("<I", b, len(b) - 4)[0]
    lim = len(b) - 4 * (n + 1)
    out, key, p = [], b"", 0
    while p < lim:
        sh, p = get_varint(b, p)
        ns, p = get_varint(b, p)
        vl, p = get_varint(b, p)
        key = key[:sh] + b[p:p + ns]
        out.append((key, b[p + ns:p + ns + vl]))
        p += ns + vl
    return out


def bloom_hash(key: bytes) -> int:
    m, h = 0xC6A4A793, (0xBC9F1D34 ^ (len(key) * 0xC6A4A793)) & 0xFFFFFFFF
    p = 0
    while len(key) - p >= 4:
        h = (h + struct.unpack_from("<I", key, p)[0]) & 0xFFFFFFFF
        h = (h * m) & 0xFFFFFFFF
        h ^= h >> 16
        p += 4
    rest = key[p:]
    sx = lambda b: b - 256 if b >= 128 else b  # the tail bytes are sign-extended
    if len(rest) == 3:
        h = (h + (sx(rest[2]) << 16)) & 0xFFFFFFFF
    if len(rest) >= 2:
        h = (h + (sx(rest[1]) << 8)) & 0xFFFFFFFF
    if len(rest) >= 1:
        h = (h + sx(rest[0])) & 0xFFFFFFFF
        h = (h * m) & 0xFFFFFFFF
        h ^= h >> 24
    return h


def bloom_build(hashes, bits_per_key: int = 10) -> bytes:
    lines = (len(hashes) * bits_per_key + 511) // 512
    if not hashes:
        lines = 0
    elif lines % 2 == 0:
        lines += 1
    data = bytearray(lines * 64)
    for h in hashes:
        base = (h % lines) * 64
        delta = ((h >> 17) | (h << 15)) & 0xFFFFFFFF
        for _ in range(6):
            bit = h & 511
            data[base + bit // 8] |= 1 << (bit % 8)
            h = (h + delta) & 0xFFFFFFFF
    return bytes(data) + bytes([6]) + struct.pack("<I", lines)


def bloom_may_match(filt: bytes, key: bytes) -> bool:
    probes, lines = filt[-5], struct.unpack_from("<I", filt, len(filt) - 4)[0]
    if lines == 0:
        return True
    h = bloom_hash(key)
    base = (h % lines) * 64
    delta = ((h >> 17) | (h << 15)) & 0xFFFFFFFF
    for _ in range(probes):
        bit = h & 511
        if not filt[base + bit // 8] & (1 << (bit % 8)):
            return False
        h = (h + delta) & 0xFFFFFFFF
    return True


def hashkey_prefix(ukey: bytes) -> bytes:
    if len(ukey) < 2:
        return b""
    p = 2 + ((ukey[0] << 8) | ukey[1])
    return ukey[:p] if p <= len(ukey) else b""


def internal_key(ukey: bytes, seq: int, typ: int) -> bytes:
    return ukey + struct.pack("<Q", (seq << 8) | typ)


def write_sst(blocks, restart_interval: int = 16) -> bytes:
    """blocks: list of lists of (user_key, seq, type, value) in internal-key order"""
    f = bytearray()

    def put(block: bytes):
        h = (len(f), len(block))
        f.extend(block)
        f.extend(b"\x00" + struct.pack("<I", mask(crc32c(block + b"\x00"))))
        return h

    index, hashes, prev_prefix = [], [], None
    for recs in blocks:
        ents = [(internal_key(k, s, t), v) for k, s, t, v in recs]
        for k, _s, _t, _v in recs:
            h = bloom_hash(k)
            if not hashes or hashes[-1] != h:
                hashes.append(h)
            pf = hashkey_prefix(k)
            if pf and pf != prev_prefix:
                prev_prefix = pf
                hashes.append(bloom_hash(pf))
        off, size = put(build_block(ents, restart_interval))
        index.append((ents[-1][0], varint(off) + varint(size)))
    fh = put(bloom_build(hashes))
    props = sorted({b"rocksdb.num.entries": varint(sum(len(b) for b in blocks)), b"rocksdb.format.version": varint(2),
                    b"rocksdb.comparator": b"leveldb.BytewiseComparator"}.items())
    ph = put(build_block(props, 1))
    mh = put(build_block([(FILTER_NAME, varint(fh[0]) + varint(fh[1])), (PROPS_NAME, varint(ph[0]) + varint(ph[1]))], 1))
    ih = put(build_block(index, 1))
    foot = b"\x01" + varint(mh[0]) + varint(mh[1]) + varint(ih[0]) + varint(ih[1])
    foot += b"\x00" * (41 - len(foot)) + struct.pack("<I", 2) + struct.pack("<Q", MAGIC)
    return bytes(f) + foot


def read_sst(sst: bytes):
    """-> dict(records=[(user_key, seq, type, value)], blocks=[[...]], filter=bytes, props={...}, index_keys=[...]); raises on a bad checksum"""
    assert len(sst) >= 53
    stats = {"lz4_blocks": 0}
    foot = sst[-53:]
    assert struct.unpack_from("<Q", foot, 45)[0] == MAGIC and struct.unpack_from("<I", foot, 41)[0] == 2 and foot[0] == 1
    p = 1
    mo, p = get_varint(foot, p); ms, p = get_varint(foot, p); io, p = get_varint(foot, p); isz, p = get_varint(foot, p)

    def block(off, size):
        b, typ = sst[off:off + size], sst[off + size]
        assert typ in (0, 4), "unsupported block type"
        assert struct.unpack_from("<I", sst, off + size + 1)[0] == mask(crc32c(b + bytes([typ]))), "bad block checksum"
        if typ == 4:  # kLZ4Compression, compress_format_version 2: varint32 raw size | LZ4 block
            raw, p = get_varint(b, 0)
            b = lz4_block_decode(b[p:], raw)
            stats["lz4_blocks"] += 1
        return b

    def handle(v):
        o, q = get_varint(v, 0)
        s, _ = get_varint(v, q)
        return o, s

    meta = dict(parse_block(block(mo, ms)))
    filt = block(*handle(meta[FILTER_NAME])) if FILTER_NAME in meta else b""
    props = dict(parse_block(block(*handle(meta[PROPS_NAME])))) if PROPS_NAME in meta else {}
    idx = parse_block(block(io, isz))
    blocks, records = [], []
    for _k, hv in idx:
        recs = []
        for ik, v in parse_block(block(*handle(hv))):
            t = struct.unpack("<Q", ik[-8:])[0]
            recs.append((ik[:-8], t >> 8, t & 0xFF, v))
        blocks.append(recs)
        records += recs
    return {"records": records, "blocks": blocks, "filter": filt, "props": props, "index_keys": [k for k, _ in idx], **stats}
