"""BASELINE.json configs[4] at small scale: YCSB-A, 50/50 put + get with zipfian keys through the rrdb surface.  The memtable is
tiny, so the run keeps flushing and compacting L0 under the reads; gets are answered from the memtable in place over the HBM
runs (rocksdb_wrapper.cpp:78-127).  On CPU the oracle is checked against a dictionary; under -m gpu the engine is checked
against the oracle response by response."""
import random

import numpy as np
import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend, same_response

NOW = synth.NOW


def workload(n_keys=400, n_ops=3000, seed=17):
    rng = np.random.default_rng(seed)
    w = 1.0 / np.power(np.arange(1, n_keys + 1), 0.99)
    cdf = np.cumsum(w) / w.sum()
    ids = rng.permutation(n_keys)[np.minimum(np.searchsorted(cdf, rng.random(n_ops)), n_keys - 1)]
    rnd = random.Random(seed)
    ops = []
    for i in range(n_ops):
        hk = b"user%06d" % int(ids[i])
        if rnd.random() < 0.5:
            ops.append(("put", hk, bytes(rnd.getrandbits(8) for _ in range(rnd.choice([10, 100, 400]))), rnd.choice([0, 0, 0, NOW + 50, NOW - 1])))
        elif rnd.random() < 0.05:
            ops.append(("remove", hk, None, 0))
        else:
            ops.append(("get", hk, None, 0))
    return ops


OPTS = {"memtable_bytes": 8 << 10, "l0_compaction_trigger": 3}


def test_oracle_matches_a_dictionary():
    o = Backend("oracle", opts=OPTS)
    model = {}
    try:
        for kind, hk, val, ets in workload():
            if kind == "put":
                o.put(hk, b"f0", val, expire_ts=ets, now=NOW)
                model[hk] = (val, ets)
            elif kind == "remove":
                o.remove(hk, b"f0", now=NOW)
                model.pop(hk, None)
            else:
                r = o.get(hk, b"f0", now=NOW)
                want = model.get(hk)
                alive = want is not None and (want[1] == 0 or want[1] > NOW)
                assert (r["error"] == 0) == alive, (hk, r["error"], want and want[1])
                if alive:
                    assert r["kvs"][0][1] == want[0]
    finally:
        o.close()


@pytest.mark.gpu
def test_engine_matches_the_oracle(engine):
    g, o = Backend("gpu", engine, pidx=4, opts=OPTS), Backend("oracle", pidx=4, opts=OPTS)
    try:
        n_get = 0
        for kind, hk, val, ets in workload(n_ops=4000):
            if kind == "put":
                assert g.put(hk, b"f0", val, expire_ts=ets, now=NOW) == o.put(hk, b"f0", val, expire_ts=ets, now=NOW) == 0
            elif kind == "remove":
                assert g.remove(hk, b"f0", now=NOW) == o.remove(hk, b"f0", now=NOW) == 0
            else:
                ok, d = same_response(g.get(hk, b"f0", now=NOW), o.get(hk, b"f0", now=NOW))
                assert ok, (hk, d)
                n_get += 1
        assert n_get > 1000
        assert g.f("rrdb_last_flushed_decree")(g.h) == o.f("rrdb_last_flushed_decree")(o.h) > 0  # both flushed at the same points
    finally:
        g.close(); o.close()
