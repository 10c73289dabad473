#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 LSM compaction / read engine (BASELINE.json metric:
compaction merged-GB/s + scan keys/s, next to the CPU path).

One "step" = one L0->L1 compaction of one hash partition (BASELINE.json configs[1]: 4 sorted runs x
2.5 M records, 16 B hashkey / 32 B sortkey / 256 B value, synthetic, fixed seed and `now`), with the
KeyWithTTLCompactionFilter fused.  Each rank (GPU) owns its own partition(s): weak scaling, no
collective on the data path (hash partitions are independent).

  value  = whole-job merged GB/s, sum(user key + value bytes of all input records) / time, inputs already
           resident in HBM, timed with CUDA events on the engine's stream, max over ranks.
  e2e    = the same metric through the C ABI starting from HOST buffers: upload of the 4 runs
           (pinned host memory -> HBM, device index build) + compaction + result struct back.
  roofline = k_merge (the dominant kernel): algorithmic bytes (B_in + B_out, key+value only) / its
           CUDA-event duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline = the oracle's block-level CPU compaction (heap merging iterator -> filter -> block builder,
           all host threads) on a bounded sample of the same workload.  `--impl reference` times the
           same CPU path as a full arm.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

NOW = 300_000_000
RUNS = 4
HK, SK, VAL = 16, 32, 256


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json, torch copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def gen_runs(records_per_run: int, seed: int):
    from incubator_pegasus_b200 import synth
    return synth.compaction_runs(k=RUNS, n_per_run=records_per_run, hk_len=HK, sk_len=SK, user_len=VAL, now=NOW,
                                 seed=seed)


def cpu_block_runs(runs):
    """block-encode the sample for the CPU path (setup, not timed)."""
    import oracle_py as orc
    return [orc.BlockRunCPU.from_run(orc.Run.from_records(r)) for r in runs]


def cpu_compaction(bruns, threads: int):
    """oracle block-level compaction on host cores; returns (merged GB/s, seconds, in_bytes)."""
    import oracle_py as orc
    fp = orc.filter_params(enabled=True)
    _out, st, secs = orc.compact_blocks(bruns, True, fp, NOW, threads)
    return st.in_bytes / secs / 1e9, secs, int(st.in_bytes)


def zipf_ids(rng, n_items: int, n: int, theta: float = 0.99):
    """YCSB zipfian(theta) over n_items, scrambled."""
    w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), theta)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    ranks = np.searchsorted(cdf, rng.random(n))
    perm = rng.permutation(n_items)
    return perm[np.minimum(ranks, n_items - 1)]


def read_workload(records_per_run: int, n_get: int, n_scan: int, seed: int):
    """keys of the read legs: zipfian hash keys of the synthetic data set; gets pick a random sort key."""
    from incubator_pegasus_b200 import synth
    rng = np.random.default_rng(seed + 77)
    per_run_hash = (int(records_per_run * 0.9) + 63) // 64  # own hash keys of runs 1.. (synth.compaction_runs)
    n_hash = (records_per_run + 63) // 64 + (RUNS - 1) * per_run_hash
    gh = zipf_ids(rng, n_hash, n_get).astype(np.uint64)
    gs = rng.integers(0, 64, n_get).astype(np.uint64)
    get_keys = synth.make_keys(gh, gs, HK, SK, seed)
    sh = zipf_ids(rng, n_hash, n_scan).astype(np.uint64)
    scan_keys = synth.make_keys(sh, np.zeros(n_scan, np.uint64), HK, SK, seed)[:, 2:2 + HK]
    return get_keys, scan_keys


def reference_arm(args, rank: int, world: int):
    """--impl reference: the CPU path (oracle port; RocksDB itself is not in the reference tree nor this image)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = min(args.records_per_run, args.cpu_sample_records or max(250_000, 20_000 * threads))
    bruns = cpu_block_runs(gen_runs(sample, 1000))
    vals = []
    for _ in range(args.warmup + args.steps):
        gbs, secs, in_bytes = cpu_compaction(bruns, threads)
        vals.append((gbs, secs))
    timed = vals[args.warmup:]
    ms = 1e3 * sum(s for _, s in timed) / len(timed)
    v = in_bytes / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": "compaction_merged_GBps", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"L0->L1 compaction sample, {RUNS} runs x {sample} records, {HK}B hashkey/{SK}B sortkey/{VAL}B value, TTL filter on"},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{RUNS} x {sample} records ({in_bytes / 1e9:.2f} GB merged) per step"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--records-per-run", type=int, default=2_500_000)
    ap.add_argument("--cpu-sample-records", type=int, default=0, help="records per run of the CPU sample; 0 = scale with cores")
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--no-tma", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-reads", action="store_true")
    ap.add_argument("--n-get", type=int, default=262144)
    ap.add_argument("--n-scan", type=int, default=16384)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import incubator_pegasus_b200 as pgs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: this rank's partition ---------------------------------------------------------
    t0 = time.time()
    runs = gen_runs(args.records_per_run, 1000 + rank)
    host_runs = [pgs.build_run(r) for r in runs]  # flush side: records -> data blocks (host)
    in_bytes = sum(int(r.keys.shape[0] + r.vals.shape[0]) for r in runs)
    n_records = sum(r.n for r in runs)
    gen_s = time.time() - t0
    # pinned host copies of the block bytes for the end-to-end leg
    pinned, pinned_tensors = [], []
    for hr in host_runs:
        t = torch.empty(hr.data.shape[0], dtype=torch.uint8).pin_memory()
        t.numpy()[:] = hr.data
        pinned.append(pgs.BlockRun(t.numpy(), hr.blk_off, hr.blk_size))
        pinned_tensors.append(t)
    h2d_bytes = sum(int(p.data.shape[0]) for p in pinned)

    eng = pgs.Engine(device=local_rank, ctas_per_sm=args.ctas_per_sm, flags=1 if args.no_tma else 0)
    part = eng.partition(app_id=1, pidx=rank)
    ids = [part.upload(p) for p in pinned]
    stream = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local_rank))
    KEEP = 1 | 2  # PGS_COMPACT_KEEP_INPUTS | PGS_COMPACT_DISCARD_OUTPUT: repeat the same job

    def step():
        return part.compact(ids, out_level=1, bottommost=1, now=NOW, enabled=True, flags=KEEP)

    for _ in range(args.warmup):
        res = step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    merge_ms, plan_ms = [], []
    w0 = time.perf_counter()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            res = step()
            merge_ms.append(res.merge_kernel_ms)
            plan_ms.append(res.device_ms - res.merge_kernel_ms)
        ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    sampler.stop_flag.set()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launches - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    ms_per_step = max_ms / args.steps
    value = world * in_bytes / (ms_per_step / 1e3) / 1e9

    # ---- end to end through the C ABI from host buffers ------------------------------------------
    e2e = None
    if not args.skip_e2e:
        part2 = eng.partition(app_id=1, pidx=rank + 1000)

        split = {"upload": 0.0, "compact": 0.0, "drop": 0.0}

        def e2e_step():
            t0 = time.perf_counter()
            rid = [part2.upload(p) for p in pinned]            # H2D of the runs + device index build
            t1 = time.perf_counter()
            r = part2.compact(rid, out_level=1, bottommost=1, now=NOW, enabled=True)  # result struct comes back
            t2 = time.perf_counter()
            if r.new_run_id:
                part2.drop(r.new_run_id)
            t3 = time.perf_counter()
            split["upload"] += (t1 - t0) * 1e3; split["compact"] += (t2 - t1) * 1e3; split["drop"] += (t3 - t2) * 1e3
            return r

        e2e_step()
        barrier()
        split.update(upload=0.0, compact=0.0, drop=0.0)
        e0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e_ms = (time.perf_counter() - e0) * 1e3 / n_e2e
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # what the host link of this box delivers for a plain pinned -> device copy (context for the number above)
        probe = torch.empty(min(1 << 30, int(pinned_tensors[0].numel())), dtype=torch.uint8, device="cuda")
        src = pinned_tensors[0][: probe.numel()]
        probe.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        probe.copy_(src, non_blocking=True)
        p1.record()
        torch.cuda.synchronize()
        h2d_probe = probe.numel() / (p0.elapsed_time(p1) / 1e3) / 1e9
        del probe
        e2e = {"value": world * in_bytes / (float(te.item()) / 1e3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 128, "ms_per_step": float(te.item()),
               "timed": "host wall clock around upload(4 runs)+compact, barrier+synchronize both sides",
               "h2d_link_probe_GBps": round(h2d_probe, 1),
               "host_ms_per_step": {k: round(v / n_e2e, 2) for k, v in split.items()}}
        part2.close()

    # ---- read path on the same partition (4 overlapping runs resident): YCSB-C shaped, zipfian hash keys ----------
    reads = None
    if not args.skip_reads:
        pins = []

        def pinned(n, dt):  # host buffers of the read legs live in pinned memory, like a server's I/O buffers
            t = torch.empty(int(n) * np.dtype(dt).itemsize, dtype=torch.uint8).pin_memory()
            pins.append(t)
            return t.numpy().view(dt)

        gk, sk = read_workload(args.records_per_run, args.n_get, args.n_scan, 1000 + rank)
        gkeys = pinned(gk.size, np.uint8)
        gkeys[:] = gk.reshape(-1)
        goff = pinned(args.n_get + 1, np.uint32)
        goff[:] = np.arange(args.n_get + 1, dtype=np.uint32) * np.uint32(gk.shape[1])
        hashkeys = [bytes(r) for r in sk]
        garena_cap = args.n_get * (VAL + 8)
        garena_buf = pinned(garena_cap, np.uint8)
        gres_buf = (pgs.GetResult * args.n_get)()
        reps = max(3, args.steps)
        # gets
        part.get_batch(gkeys, goff, NOW, arena_cap=garena_cap, arena=garena_buf, results=gres_buf)
        g_ms, g_wall, found, probes = [], [], 0, 0
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            st, gres, garena, gused = part.get_batch(gkeys, goff, NOW, arena_cap=garena_cap, arena=garena_buf, results=gres_buf)
            g_wall.append((time.perf_counter() - t0) * 1e3)
            g_ms.append(eng.last_kernel_ms)
            probes = eng.last_blocks_probed
        found = sum(1 for i in range(0, args.n_get, max(1, args.n_get // 4096)) if gres[i].status == 0)
        found_frac = found / len(range(0, args.n_get, max(1, args.n_get // 4096)))
        # prefix scans = multi_get(hash_key, all sort keys)
        sb = part.prefix_scan_batch(hashkeys, max_records=80, arena_stride=24576, alloc=pinned)  # request structs marshalled once
        assert sb.run(NOW) == 0
        s_ms, s_wall = [], []
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            st = sb.run(NOW)  # host request structs in, packed records out (host buffers)
            s_wall.append((time.perf_counter() - t0) * 1e3)
            s_ms.append(eng.last_kernel_ms)
            assert st == 0, st
        sres, abase, kbase = sb.results, sb.abase, sb.kbase
        returned = int(kbase[-1])
        iterated = int(sum(sres[i].iter_count for i in range(args.n_scan)))
        scan_bytes = int(abase[-1])
        gm, sm = sum(g_ms) / len(g_ms), sum(s_ms) / len(s_ms)
        nb_log = 18
        get_algo = probes * (4096 + nb_log * 58) + args.n_get * (2 + HK + SK) + int(gused)
        scan_algo = returned * 2 * (2 + HK + SK + 12 + VAL) + (iterated - returned) * (2 + HK + SK + 12 + VAL)
        vals = torch.tensor([args.n_get / (gm / 1e3), args.n_get / (min(g_wall) / 1e3), returned / (sm / 1e3), returned / (min(s_wall) / 1e3)],
                            dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.SUM)  # partitions are independent: whole-job keys/s = sum over ranks
        reads = {
            "get": {"metric": "get_keys_per_s", "value": float(vals[0]), "e2e": float(vals[1]), "unit": "keys/s", "batch": args.n_get,
                    "kernel_ms": gm, "found_frac_sampled": found_frac, "blocks_probed_per_key": probes / args.n_get,
                    "roofline": {"bound": "hbm", "kernel": "k_get", "achieved": get_algo / (gm / 1e3) / 1e9, "peak": load_peaks()[0], "unit": "GB/s",
                                 "frac": get_algo / (gm / 1e3) / 1e9 / load_peaks()[0], "algorithmic_bytes_per_launch": get_algo}},
            "scan": {"metric": "scan_keys_per_s", "value": float(vals[2]), "e2e": float(vals[3]), "unit": "keys/s", "requests": args.n_scan,
                     "returned_per_launch": returned, "iterated_per_launch": iterated, "kernel_ms": sm, "d2h_bytes": scan_bytes,
                     "roofline": {"bound": "hbm", "kernel": "k_scan", "achieved": scan_algo / (sm / 1e3) / 1e9, "peak": load_peaks()[0], "unit": "GB/s",
                                  "frac": scan_algo / (sm / 1e3) / 1e9 / load_peaks()[0], "algorithmic_bytes_per_launch": scan_algo}},
            "workload": "YCSB-C shaped: zipfian(0.99) hash keys over the 4 resident overlapping runs; get(hk,sk) and multi_get(hk, all sort keys)",
        }
        if rank == 0 and world == 1 and not args.skip_cpu:
            import oracle_py as orc
            threads = os.cpu_count() or 1
            bruns = [orc.BlockRunCPU.from_blocks(hr) for hr in reversed(host_runs)]  # newest first
            ng, tot_s, tot_n = args.n_get, 0.0, 0
            while tot_s < 2.0 and tot_n < 200 * ng:  # repeat the batch until the sample is a couple of seconds of wall time
                f, vb, secs = orc.get_many(bruns, gkeys, goff, NOW, threads)
                tot_s += secs
                tot_n += ng
            reads["get"]["cpu_baseline"] = {"value": tot_n / tot_s, "unit": "keys/s", "cores": threads, "kind": "port",
                                            "sample": f"{tot_n} gets ({ng}-key batch repeated) over the same 4 block runs, {tot_s:.2f} s"}
            nsc = args.n_scan
            hk_flat = np.ascontiguousarray(sk[:nsc].reshape(-1))
            hk_off = (np.arange(nsc + 1, dtype=np.uint32) * np.uint32(HK))
            tot_s, tot_n, tot_q = 0.0, 0, 0
            while tot_s < 2.0 and tot_q < 200 * nsc:
                cnt, nb_, secs = orc.prefix_scan_many(bruns, hk_flat, hk_off, NOW, threads)
                tot_s += secs
                tot_n += cnt
                tot_q += nsc
            reads["scan"]["cpu_baseline"] = {"value": tot_n / tot_s, "unit": "keys/s", "cores": threads, "kind": "port",
                                             "sample": f"{tot_q} prefix scans ({tot_n} records) over the same 4 block runs, {tot_s:.2f} s"}
            del bruns

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = load_peaks()
    k_ms = sum(merge_ms) / len(merge_ms)
    algo_bytes = int(res.in_bytes + res.out_bytes)
    achieved = algo_bytes / (k_ms / 1e3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "k_merge_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "k_merge", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms, "plan_kernel_ms": sum(plan_ms) / len(plan_ms)}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        sample = min(args.records_per_run, args.cpu_sample_records or max(250_000, 20_000 * threads))
        sruns = runs if sample == args.records_per_run else gen_runs(sample, 1000)
        gbs, secs, sb = cpu_compaction(cpu_block_runs(sruns), threads)
        cpu = {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"{RUNS} runs x {sample} records ({sb / 1e9:.2f} GB merged), oracle block-level compaction, {secs:.2f} s"}

    if rank == 0:
        line = {
            "metric": "compaction_merged_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"single-partition L0->L1 compaction per GPU: {RUNS} SSTs x {args.records_per_run} keys, "
                                   f"{HK}B hashkey/{SK}B sortkey/{VAL}B value (BASELINE.json configs[1])",
                       "records_per_step_per_gpu": n_records, "merged_bytes_per_step_per_gpu": in_bytes,
                       "survivors": int(res.out_records), "tiles": int(res.n_tiles), "filter": "KeyWithTTLCompactionFilter on",
                       "l2": "inputs (2.9 GB of blocks) larger than the 126 MB L2", "ctas_per_sm": args.ctas_per_sm or 1,
                       "tma": not args.no_tma, "input_gen_s": round(gen_s, 1)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "reads": reads, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "wall_ms_per_step": wall_ms / args.steps,
        }
        print(json.dumps(line), flush=True)
    part.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
