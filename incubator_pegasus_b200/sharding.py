"""Partition -> GPU placement.  Hash partitions are independent (`partition = crc64(hashkey) % partition_count`,
src/base/pegasus_key_schema.h:150-165, src/client/partition_resolver.cpp:48-51), so a box shards them over its
GPUs with no collective on the data path: rank r owns the partitions with pidx % world == r."""
from __future__ import annotations

import ctypes as C

from . import lib


def partition_of(hash_key: bytes, partition_count: int) -> int:
    """client-side routing: crc64 of the hash key modulo the partition count."""
    return int(lib().pgs_crc64(hash_key, len(hash_key), 0) % partition_count)


def partitions_of_rank(partition_count: int, rank: int, world: int) -> list[int]:
    return [p for p in range(partition_count) if p % world == rank]


def owner_rank(pidx: int, world: int) -> int:
    return pidx % world


def whole_job_rate(units_per_rank: float, seconds: float, world: int, reduce_max):
    """weak-scaling aggregate: every rank processed `units_per_rank`; the job took the slowest rank's time."""
    return world * units_per_rank / reduce_max(seconds)
