"""GPU: the threading contract of the reference's storage app (src/server/config.ini:140-150 runs THREAD_POOL_LOCAL_APP
with many worker threads over the replicas of one process): eight host threads hammer one engine at the same time --
gets, multi_gets, batch_gets and scanners on replicas with different numbers of runs (so different kernel shapes and
shared-memory sizes are in flight together), while one thread writes, flushes and manually compacts another replica.
Every response is compared with the CPU oracle's answer to the same request."""
import random
import threading

import pytest

from incubator_pegasus_b200 import synth
from rrdb_harness import Backend, raw_key, next_blob, same_response

NOW = synth.NOW
pytestmark = pytest.mark.gpu


def fill(backends, n_runs, seed, hashkeys=12, per_run=150):
    rnd = random.Random(seed)
    for round_ in range(n_runs):
        for hk_i in range(hashkeys):
            hk = b"hk%03d" % hk_i
            kvs = {b"s%04d" % rnd.randrange(400): bytes(rnd.getrandbits(8) for _ in range(rnd.choice([8, 90, 600])))
                   for _ in range(per_run // hashkeys)}
            ets = rnd.choice([0, 0, NOW + 500, NOW - 5])
            for be in backends:
                be.multi_put(hk, kvs, expire_ts=ets, now=NOW)
        rm = [b"s%04d" % rnd.randrange(400) for _ in range(6)]
        for be in backends:
            be.multi_remove(b"hk%03d" % (round_ % hashkeys), rm, now=NOW)
            be.flush(NOW)


def requests(seed, n=60):
    rnd = random.Random(seed)
    out = []
    for _ in range(n):
        hk = b"hk%03d" % rnd.randrange(14)  # two of them do not exist
        kind = rnd.choice(["get", "multi_get", "multi_get_rev", "batch_get", "count", "scan", "ttl"])
        if kind in ("get", "ttl"):
            out.append((kind, hk, b"s%04d" % rnd.randrange(400)))
        elif kind == "batch_get":
            out.append((kind, [(b"hk%03d" % rnd.randrange(14), b"s%04d" % rnd.randrange(400)) for _ in range(20)], None))
        else:
            out.append((kind, hk, rnd.choice([0, 7, 50])))
    return out


def answer(be, req):
    kind, a, b = req
    if kind == "get":
        return be.get(a, b, now=NOW)
    if kind == "ttl":
        return be.ttl(a, b, now=NOW)
    if kind == "multi_get":
        return be.multi_get(a, max_kv_count=b, now=NOW)
    if kind == "multi_get_rev":
        return be.multi_get(a, max_kv_count=b, reverse=True, now=NOW)
    if kind == "batch_get":
        return be.batch_get(a, now=NOW)
    if kind == "count":
        return be.sortkey_count(a, now=NOW)
    kvs, batches = be.scan_all(a, batch_size=b or 33, now=NOW)
    r = dict(batches[-1])
    r["kvs"] = kvs
    return r


def test_eight_threads_on_one_engine(pgs, oracle, engine):
    opts = {"l0_compaction_trigger": 100}
    shapes = [1, 3, 6]
    gpu = [Backend("gpu", engine, pidx=i, opts=opts) for i in range(len(shapes))]
    orc = [Backend("oracle", pidx=i, opts=opts) for i in range(len(shapes))]
    wg, wo = Backend("gpu", engine, pidx=7, opts=opts), Backend("oracle", pidx=7, opts=opts)
    readers = []
    try:
        for i, n_runs in enumerate(shapes):
            fill([gpu[i], orc[i]], n_runs, seed=100 + i)
        fill([wg, wo], 2, seed=77)
        plans = []
        for t in range(7):
            reqs = requests(seed=1000 + t)
            which = [(t + j) % len(shapes) for j in range(len(reqs))]
            want = [answer(orc[w], r) for w, r in zip(which, reqs)]
            mine = [g.reader() for g in gpu]
            readers += mine
            plans.append((reqs, which, want, mine))
        failures = []
        start = threading.Barrier(8)

        def read_worker(t):
            reqs, which, want, mine = plans[t]
            start.wait()
            for rep in range(3):
                for w, r, exp in zip(which, reqs, want):
                    got = answer(mine[w], r)
                    ok, d = same_response(got, exp)
                    if not ok:
                        failures.append((t, r, d))
                        return

        def write_worker():
            start.wait()
            for rep in range(3):
                fill([wg], 2, seed=500 + rep)
                wg.manual_compact(NOW)

        threads = [threading.Thread(target=read_worker, args=(t,)) for t in range(7)] + [threading.Thread(target=write_worker)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=600)
        assert not any(th.is_alive() for th in threads)
        assert not failures, failures[:2]
        # the writer's replica: the oracle twin replays the same history, then both answer the same reads
        for rep in range(3):
            fill([wo], 2, seed=500 + rep)
            wo.manual_compact(NOW)
        for r in requests(seed=9, n=40):
            ok, d = same_response(answer(wg, r), answer(wo, r))
            assert ok, (r, d)
    finally:
        for b in readers + gpu + orc + [wg, wo]:
            b.close()


def test_router_places_replicas(pgs):
    with pgs.Router() as router:
        n = router.device_count
        assert n >= 1
        for pidx in range(4 * n):
            assert router.device_for(1, pidx) == pidx % n
        assert router.device_for(1, -1) == -1
        eng = router.engine_for(2, 5)
        part = eng.partition(app_id=2, pidx=5)
        recs = synth.compaction_runs(k=1, n_per_run=2000)[0]
        rid = part.upload_records(recs)
        assert part.run_info(rid).n_records == 2000
        part.close()
