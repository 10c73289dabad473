"""Checkpoints (first slice; sync_checkpoint / storage_apply_checkpoint, pegasus_server_impl.cpp:1951-2336): a replica
writes its resident runs as BlockBasedTable files + a MANIFEST under checkpoint.<decree>; another replica (or the same one,
later) takes that state over.  Reads after the hand-over are compared with an oracle replica that replayed exactly the writes
the checkpoint covers."""
import os
import random
import sys

import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend, raw_key, next_blob, same_response

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sst_py  # noqa: E402

NOW = synth.NOW
pytestmark = pytest.mark.gpu


def history(backends, seed, rounds):
    rnd = random.Random(seed)
    for r in range(rounds):
        for hk_i in range(6):
            hk = b"hk%02d" % hk_i
            kvs = {b"s%03d" % rnd.randrange(120): bytes(rnd.getrandbits(8) for _ in range(rnd.choice([6, 80, 300]))) for _ in range(25)}
            ets = rnd.choice([0, 0, NOW + 1000, NOW - 3])
            for be in backends:
                be.multi_put(hk, kvs, expire_ts=ets, now=NOW)
        gone = [b"s%03d" % rnd.randrange(120) for _ in range(5)]
        for be in backends:
            be.multi_remove(b"hk%02d" % (r % 6), gone, now=NOW)
            be.flush(NOW)


def same_reads(a, b):
    for hk_i in range(7):
        hk = b"hk%02d" % hk_i
        for sk in (b"s000", b"s017", b"s060", b"s119", b"nope"):
            ok, d = same_response(a.get(hk, sk, now=NOW), b.get(hk, sk, now=NOW))
            assert ok, (hk, sk, d)
        for kw in (dict(), dict(reverse=True, max_kv_count=7), dict(start=b"s020", stop=b"s090", stop_inclusive=True)):
            ok, d = same_response(a.multi_get(hk, now=NOW, **kw), b.multi_get(hk, now=NOW, **kw))
            assert ok, (hk, kw)
        assert a.scan_all(hk, batch_size=17, now=NOW)[0] == b.scan_all(hk, batch_size=17, now=NOW)[0]


def test_checkpoint_and_apply(pgs, engine, tmp_path):
    opts = {"l0_compaction_trigger": 3, "memtable_bytes": 64 << 10}
    g, o = Backend("gpu", engine, pidx=21, opts=opts), Backend("oracle", pidx=21, opts=opts)
    g2 = Backend("gpu", engine, pidx=21, opts=opts)  # a second replica object of the same gpid (the learner)
    try:
        history([g, o], seed=1, rounds=5)
        g.manual_compact(NOW); o.manual_compact(NOW)       # a bottom level (LZ4 in the checkpoint) ...
        history([g, o], seed=2, rounds=2)                  # ... and newer L0 / L1 runs on top
        g.put(b"hk00", b"in-memtable", b"v", now=NOW); o.put(b"hk00", b"in-memtable", b"v", now=NOW)
        import ctypes as C
        d = C.c_int64()
        base = str(tmp_path / "ckpt").encode()
        assert g.f("rrdb_sync_checkpoint")(g.h, base, NOW, C.byref(d)) == 0
        assert d.value == g.decree == g.f("rrdb_last_durable_decree")(g.h) == g.f("rrdb_last_flushed_decree")(g.h)
        cdir = os.path.join(base.decode(), "checkpoint.%d" % d.value)
        files = sorted(os.listdir(cdir))
        assert files[-1] == "MANIFEST" and len(files) >= 3 and all(f.endswith(".sst") for f in files[:-1])
        manifest = open(os.path.join(cdir, "MANIFEST")).read().split("\n")
        assert manifest[0] == "pegasus_b200_checkpoint 1" and "last_flushed_decree %d" % d.value in manifest
        small = min((os.path.getsize(os.path.join(cdir, f)), f) for f in files[:-1])[1]
        parsed = sst_py.read_sst(open(os.path.join(cdir, small), "rb").read())  # every file is a plain SST image
        assert parsed["records"]
        assert g.f("rrdb_sync_checkpoint")(g.h, base, NOW, C.byref(d)) == 0    # same decree again: nothing to do
        # the replica moves on; the checkpoint does not
        history([g], seed=3, rounds=2)
        g.remove(b"hk00", b"in-memtable", now=NOW)
        # a fresh replica takes the checkpoint over
        assert g2.f("rrdb_apply_checkpoint")(g2.h, cdir.encode()) == 0
        g2.decree = d.value
        assert g2.f("rrdb_last_committed_decree")(g2.h) == g2.f("rrdb_last_durable_decree")(g2.h) == d.value
        same_reads(g2, o)
        # it keeps working as a replica: the same writes on both sides, still the same answers
        history([g2, o], seed=4, rounds=1)
        same_reads(g2, o)
        # the original replica rolls back to its own checkpoint
        o2 = Backend("oracle", pidx=21, opts=opts)
        try:
            history([o2], seed=1, rounds=5); o2.manual_compact(NOW); history([o2], seed=2, rounds=2)
            o2.put(b"hk00", b"in-memtable", b"v", now=NOW)
            assert g.f("rrdb_apply_checkpoint")(g.h, cdir.encode()) == 0
            same_reads(g, o2)
        finally:
            o2.close()
        # damaged / missing checkpoints
        assert g2.f("rrdb_apply_checkpoint")(g2.h, str(tmp_path / "nowhere").encode()) == pgs.NOT_FOUND
        bad = tmp_path / "bad"
        bad.mkdir()
        (bad / "MANIFEST").write_text("something else 1\n")
        assert g2.f("rrdb_apply_checkpoint")(g2.h, str(bad).encode()) == pgs.CORRUPTION
        same_reads(g2, o)  # a refused checkpoint left the replica as it was
        import shutil
        torn = tmp_path / "torn"
        shutil.copytree(cdir, torn)
        victim = torn / small
        raw = bytearray(victim.read_bytes())
        raw[1] ^= 0x40                                  # one flipped bit inside the first data block: its checksum no longer matches
        victim.write_bytes(bytes(raw))
        assert g2.f("rrdb_apply_checkpoint")(g2.h, str(torn).encode()) == pgs.CORRUPTION
        same_reads(g2, o)  # every image is decoded before the old state is given up
        (torn / small).write_bytes(bytes(raw[:-7]))     # wrong size
        assert g2.f("rrdb_apply_checkpoint")(g2.h, str(torn).encode()) == pgs.CORRUPTION
        same_reads(g2, o)
    finally:
        for b in (g, o, g2):
            b.close()
