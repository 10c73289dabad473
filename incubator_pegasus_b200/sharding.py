"""Partition -> GPU placement and the partitioned synthetic table of BASELINE.json configs[2] (multi_get + prefix scan over 256
partitions, YCSB-C zipfian).

Hash partitions are independent (`pidx = pegasus_key_hash(key) % partition_count`, src/base/pegasus_key_schema.h:150-165,
src/client/partition_resolver.cpp:48-51), so a box shards them over its GPUs with no collective on the data path (SURVEY.md
§8e): the process of rank r owns the partitions with pidx % world == r (the same rule pgs_router_device_for applies inside
one process)."""
from __future__ import annotations

import numpy as np

from . import lib, synth, TYPE_DELETION, TYPE_VALUE


def partition_of(hash_key: bytes, partition_count: int) -> int:
    """client-side routing: crc64 of the hash key modulo the partition count."""
    return int(lib().pgs_partition_index(hash_key, len(hash_key), b"", 0, partition_count))


def partitions_of_rank(partition_count: int, rank: int, world: int) -> list[int]:
    return [p for p in range(partition_count) if p % world == rank]


def owner_rank(pidx: int, world: int) -> int:
    return pidx % world


def whole_job_rate(units_per_rank: float, seconds: float, world: int, reduce_max):
    """weak-scaling aggregate: every rank processed `units_per_rank`; the job took the slowest rank's time."""
    return world * units_per_rank / reduce_max(seconds)


class Table:
    """A hash-partitioned synthetic table: n_hash hash keys x sortkeys_per_hash sort keys (config #2 record shape), every
    hash key routed by crc64 like a client would.  A partition's data is three runs of different age and size -- the bottom
    level holds every record, L1 newer versions of `l1_frac` of them, L0 newer versions (1 % tombstones) of `l0_frac` --
    so reads have to merge overlapping runs, as after a few flushes and one compaction."""

    def __init__(self, partition_count: int = 256, n_hash: int = 65536, sortkeys_per_hash: int = 64, hk_len: int = 16, sk_len: int = 32,
                 user_len: int = 256, l1_frac: float = 0.3, l0_frac: float = 0.1, now: int = synth.NOW, seed: int = 4242):
        self.P, self.n_hash, self.spk = partition_count, n_hash, sortkeys_per_hash
        self.hk_len, self.sk_len, self.user_len = hk_len, sk_len, user_len
        self.l1_frac, self.l0_frac, self.now, self.seed = l1_frac, l0_frac, now, seed
        ids = np.arange(n_hash, dtype=np.uint64)
        self.hashkeys = synth.make_keys(ids, np.zeros(n_hash, np.uint64), hk_len, 0, seed)[:, 2:2 + hk_len]
        L = lib()
        flat = np.ascontiguousarray(self.hashkeys)
        self.pidx = np.fromiter((L.pgs_partition_index(flat[i].tobytes(), hk_len, b"", 0, partition_count) for i in range(n_hash)),
                                dtype=np.int32, count=n_hash)

    def hash_ids_of(self, pidx: int) -> np.ndarray:
        return np.nonzero(self.pidx == pidx)[0].astype(np.uint64)

    def partition_runs(self, pidx: int):
        """[(level, Records)] oldest (bottom) first; empty list when no hash key lands on the partition"""
        h = self.hash_ids_of(pidx)
        if h.size == 0:
            return []
        rng = np.random.default_rng(self.seed * 1000003 + pidx)
        base_h = np.repeat(h, self.spk)
        base_s = np.tile(np.arange(self.spk, dtype=np.uint64), h.size)
        n = base_h.size
        out = []
        seq0 = 1
        for level, frac, tomb in ((2, 1.0, 0.0), (1, self.l1_frac, 0.0), (0, self.l0_frac, 0.01)):
            m = n if frac >= 1.0 else max(1, int(n * frac))
            pick = np.arange(n) if m == n else np.sort(rng.choice(n, m, replace=False))
            keys = synth.make_keys(base_h[pick], base_s[pick], self.hk_len, self.sk_len, self.seed)
            vals = synth.make_values(rng, m, self.user_len, self.now, ts_us=1_700_000_000_000_000 + (2 - level))
            typ = np.full(m, TYPE_VALUE, np.uint8)
            if tomb:
                typ[rng.choice(m, max(1, int(m * tomb)), replace=False)] = TYPE_DELETION
            seq = np.uint64(seq0) + rng.permutation(m).astype(np.uint64)
            seq0 += m
            order = synth._sort_fixed(keys)
            out.append((level, synth.fixed_records(keys[order], vals[order], seq[order], typ[order])))
        return out

    def requests(self, n_get: int, n_scan: int, theta: float = 0.99, seed: int = 7):
        """YCSB-C: zipfian(theta) hash keys over the whole table (scrambled), a uniformly random sort key per get.
        Returns (get_hash_ids, get_sort_ids, scan_hash_ids); every rank draws the same requests and serves its own."""
        rng = np.random.default_rng(seed)
        w = 1.0 / np.power(np.arange(1, self.n_hash + 1, dtype=np.float64), theta)
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        perm = rng.permutation(self.n_hash)
        gh = perm[np.minimum(np.searchsorted(cdf, rng.random(n_get)), self.n_hash - 1)].astype(np.uint64)
        gs = rng.integers(0, self.spk, n_get).astype(np.uint64)
        sh = perm[np.minimum(np.searchsorted(cdf, rng.random(n_scan)), self.n_hash - 1)].astype(np.uint64)
        return gh, gs, sh
