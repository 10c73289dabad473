// compact.cu — level compaction on the GPU: k-way merge of HBM-resident sorted runs with
// KeyWithTTLCompactionFilter fused into the merge.
//
// Replaces (reference file:line):
//   DB::CompactRange / background compaction job ....... src/server/pegasus_server_impl.cpp:3373-3394
//   RocksDB MergingIterator + CompactionIterator + BlockBasedTableBuilder (v8.5.3, not in tree;
//   semantics restated in SURVEY.md Appendix A)
//   KeyWithTTLCompactionFilter::Filter .................. src/server/key_ttl_compaction_filter.h:55-121
//   compaction_operation / compaction_filter_rule ....... src/server/compaction_operation.cpp:33-113,
//                                                         src/server/compaction_filter_rule.cpp:31-90
//
// Shape of the computation (B200-first, no tensor cores: this is byte/integer work bound by HBM):
//   k_plan   one thread per input block ranks the block's last user key against every run's block
//            index (binary search) => cumulative shared-memory weight of everything <= that key.
//            Keys where the weight crosses a multiple of the tile budget become tile boundaries:
//            tile q = user keys in (U_q, U_q+1], a contiguous block range per run.
//   k_merge  persistent CTAs take tiles in ticket order.  Per tile:
//              TMA (cp.async.bulk) stages each run's block slice into shared memory,
//              one warp per block decodes restart-interval prefix compression into an arena of
//              full user keys, every record binary-searches the other runs' records for its merge
//              rank (and finds out whether a newer version shadows it), the compaction filter runs
//              per surviving record, survivors are re-encoded into 4 KB-target data blocks
//              (restart interval 16) whose byte offsets come from block-wide scans, the tile's
//              output position comes from a decoupled look-back over tile aggregates, and warps
//              copy entries shared -> global with destination-aligned 16-byte stores.
//            Output blocks stay contiguous and in key order, so the new run needs no second pass.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "device_util.cuh"
#include "engine.h"

namespace pgs {

constexpr uint32_t kMaxTileBlocks = 256;
constexpr uint32_t kMaxOutBlocks = 128;
constexpr uint32_t kRecExtra = 48; // per-record shared-memory bytes besides the key slot

enum : uint8_t { F_VALID = 1, F_SHADOW = 2, F_KEEP = 4, F_TOMB = 8, F_NEWTS = 16 };

// One look-back slot per tile = two 16-byte halves, each written with ONE vector store and carrying its own state
// word, so a reader gets data and validity in a single round trip (no fence, no second load); halves whose states
// differ were caught in the middle of an update and are read again.
struct TileAgg {
    unsigned long long bytes;
    uint32_t blocks, state0;
    uint32_t recs, keyb, state1, pad;
};
static_assert(sizeof(TileAgg) == 32, "TileAgg");

struct MergeStats {
    unsigned long long in_records, in_bytes, out_records, out_bytes;
    unsigned long long dropped_shadowed, dropped_tombstone, dropped_expired, dropped_user, dropped_stale, ttl_rewritten;
    unsigned long long out_tomb, out_raw_key, out_raw_val, min_seq, max_seq;
    uint32_t max_ukey, max_vlen, max_blk_size, max_blk_rec;
    uint32_t error, error_tile;
};

struct MergeParams {
    RunDev runs[kMaxRuns];
    uint32_t k;
    // plan
    uint32_t *split_pos; // [(Q+1)*k]
    uint32_t *split_ref; // [Q+1]  run<<28 | block
    uint32_t Q;
    unsigned long long tile_weight; // T
    uint32_t rec_cost;
    uint32_t total_blocks;
    // tile pipeline
    uint32_t *ticket;
    TileAgg *agg;
    uint32_t KS, pool_bytes, warp_scratch, use_tma, early_tma;
    // filter + policy
    uint32_t now, enabled, validate_hash, data_version, default_ttl;
    int32_t pidx, partition_version;
    const uint8_t *ops;
    uint32_t n_ops;
    uint32_t bottommost, block_size, restart_interval;
    const unsigned long long *crc_table;
    // output run
    uint8_t *out_data;
    unsigned long long out_cap;
    unsigned long long *out_blk_off;
    uint32_t *out_blk_size, *out_blk_rec, *out_ikey_off, *out_rec_off;
    uint8_t *out_ikeys;
    uint32_t out_blk_cap, out_ikey_cap;
    MergeStats *stats;
    unsigned long long *phase_cycles; // [16] or null
};

// ------------------------------------------------------------------------------------------------
// k_plan
// ------------------------------------------------------------------------------------------------
PGS_DEV unsigned long long run_weight(const RunDev &r, uint32_t pos, uint32_t rec_cost)
{
    return r.blk_off[pos] + (unsigned long long)r.blk_rec[pos] * rec_cost;
}

__global__ void __launch_bounds__(256) k_plan(const __grid_constant__ MergeParams P)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.total_blocks) return;
    uint32_t i = 0, b = g;
    while (b >= P.runs[i].nb) { b -= P.runs[i].nb; i++; }
    const RunDev &ri = P.runs[i];
    const uint8_t *U = ri.ikeys + ri.ikey_off[b];
    uint32_t ulen = ri.ikey_off[b + 1] - ri.ikey_off[b];
    uint32_t pos[kMaxRuns];
    unsigned long long Wb = 0, W = 0;
    for (uint32_t j = 0; j < P.k; j++) {
        const RunDev &rj = P.runs[j];
        uint32_t lo = 0, hi = rj.nb; // upper bound: #blocks with last key <= U
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            const uint8_t *kp = rj.ikeys + rj.ikey_off[mid];
            uint32_t kl = rj.ikey_off[mid + 1] - rj.ikey_off[mid];
            if (cmp_bytes4(kp, kl, U, ulen) <= 0) lo = mid + 1; else hi = mid;
        }
        uint32_t ub = lo, lb = lo;
        while (lb > 0) {
            const uint8_t *kp = rj.ikeys + rj.ikey_off[lb - 1];
            uint32_t kl = rj.ikey_off[lb] - rj.ikey_off[lb - 1];
            if (cmp_bytes4(kp, kl, U, ulen) != 0) break;
            lb--;
        }
        pos[j] = ub;
        W += run_weight(rj, ub, P.rec_cost);
        Wb += run_weight(rj, lb, P.rec_cost);
    }
    unsigned long long q_lo = Wb / P.tile_weight + 1, q_hi = W / P.tile_weight;
    if (q_hi > P.Q - 1) q_hi = P.Q - 1;
    for (unsigned long long q = q_lo; q <= q_hi; q++) {
        for (uint32_t j = 0; j < P.k; j++) P.split_pos[q * P.k + j] = pos[j];
        P.split_ref[q] = (i << 28) | b;
    }
}

// ------------------------------------------------------------------------------------------------
// compaction filter on the device
// ------------------------------------------------------------------------------------------------
PGS_DEV bool dev_pattern_match(const uint8_t *v, uint32_t vl, uint32_t match_type, const uint8_t *pat, uint32_t pl)
{
    // string_pattern_match: compaction_filter_rule.cpp:31-54 (empty pattern never matches)
    if (pl == 0 || vl < pl) return false;
    if (match_type == MATCH_PREFIX) {
        for (uint32_t i = 0; i < pl; i++) if (v[i] != pat[i]) return false;
        return true;
    }
    if (match_type == MATCH_POSTFIX) {
        const uint8_t *s = v + vl - pl;
        for (uint32_t i = 0; i < pl; i++) if (s[i] != pat[i]) return false;
        return true;
    }
    if (match_type == MATCH_ANYWHERE) {
        for (uint32_t s = 0; s + pl <= vl; s++) {
            uint32_t i = 0;
            while (i < pl && v[s + i] == pat[i]) i++;
            if (i == pl) return true;
        }
        return false;
    }
    return false;
}

PGS_DEV uint32_t ld_u32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
PGS_DEV uint32_t ld_u16(const uint8_t *p) { return p[0] | (p[1] << 8); }

// user_specified_operation_filter: key_ttl_compaction_filter.h:94-108 over the binary ops table.
// Every op sees the value as of entry (entry_ts); returns true when a delete op fired.
PGS_DEV bool dev_user_ops(const MergeParams &P, const uint8_t *hk, uint32_t hkl, const uint8_t *sk, uint32_t skl,
                          uint32_t entry_ts, uint32_t &new_ts, bool &changed)
{
    const uint8_t *p = P.ops + 4;
    for (uint32_t o = 0; o < P.n_ops; o++) {
        uint32_t op_type = p[0], ttl_type = p[1], n_rules = ld_u16(p + 2), ttl_value = ld_u32(p + 4);
        p += 8;
        bool all = n_rules > 0; // all_rules_match: empty rule set => false (compaction_operation.cpp:37-39)
        for (uint32_t r = 0; r < n_rules; r++) {
            uint32_t rt = p[0], mt = p[1], pl = ld_u16(p + 2), start_ttl = ld_u32(p + 4), stop_ttl = ld_u32(p + 8);
            const uint8_t *pat = p + 12;
            p += 12 + ((pl + 3) & ~3u);
            if (!all) continue;
            bool m;
            if (rt == RULE_HASHKEY) m = dev_pattern_match(hk, hkl, mt, pat, pl);
            else if (rt == RULE_SORTKEY) m = dev_pattern_match(sk, skl, mt, pat, pl);
            else { // ttl_range_rule::match, compaction_filter_rule.cpp:76-90 (u32 arithmetic)
                if (entry_ts == 0 && start_ttl == 0 && stop_ttl == 0) m = true;
                else m = (uint32_t)(start_ttl + P.now) <= entry_ts && (uint32_t)(stop_ttl + P.now) >= entry_ts;
            }
            all = m;
        }
        if (!all) continue;
        if (op_type == OP_DELETE) return true; // delete_key::filter
        // update_ttl::filter, compaction_operation.cpp:77-113
        uint32_t ts;
        if (ttl_type == TTL_FROM_NOW) ts = P.now + ttl_value;
        else if (ttl_type == TTL_FROM_CURRENT) { if (entry_ts == 0) continue; ts = ttl_value + entry_ts; }
        else if (ttl_type == TTL_TIMESTAMP) ts = ttl_value - kEpochBegin;
        else continue;
        new_ts = ts;
        changed = true;
    }
    return false;
}

PGS_DEV unsigned long long dev_crc64(const unsigned long long *tab, const uint8_t *p, uint32_t n)
{
    unsigned long long c = ~0ull; // init 0 -> ~init
    for (uint32_t i = 0; i < n; i++) c = tab[(uint8_t)(c ^ p[i])] ^ (c >> 8);
    return ~c;
}

// KeyWithTTLCompactionFilter::Filter (key_ttl_compaction_filter.h:55-92).
// returns 0 keep, 1 expired, 2 user op, 3 stale split data
PGS_DEV uint32_t dev_filter(const MergeParams &P, const unsigned long long *crc_tab, const uint8_t *ukey, uint32_t klen,
                            const uint8_t *val, uint32_t vlen, uint32_t &new_ts, bool &changed)
{
    changed = false;
    if (!P.enabled || klen < 2 || vlen < 4) return 0;
    uint32_t expire_ts = be32(val);
    if (P.default_ttl != 0 && expire_ts == 0) {
        expire_ts = P.now + P.default_ttl;
        new_ts = expire_ts;
        changed = true;
    }
    uint32_t hkl = be16(ukey);
    if (hkl > klen - 2) hkl = klen - 2; // malformed key: never read outside it
    const uint8_t *hk = ukey + 2, *sk = ukey + 2 + hkl;
    uint32_t skl = klen - 2 - hkl;
    if (P.n_ops) {
        if (dev_user_ops(P, hk, hkl, sk, skl, expire_ts, new_ts, changed)) return 2;
    }
    if (ts_expired(P.now, expire_ts)) return 1;
    if (P.validate_hash && P.partition_version >= 0 && P.pidx <= P.partition_version) {
        // check_pegasus_key_hash: pegasus_key_schema.h:150-183
        unsigned long long h = hkl > 0 ? dev_crc64(crc_tab, hk, hkl) : dev_crc64(crc_tab, sk, skl);
        if ((long long)(h & (unsigned long long)(long long)P.partition_version) != (long long)P.pidx) return 3;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// k_merge
// ------------------------------------------------------------------------------------------------
struct TileShared {
    unsigned long long mbar;
    unsigned long long base_bytes;
    unsigned long long min_seq, max_seq;
    uint32_t base_blocks, base_recs, base_keyb;
    uint32_t tile, error;
    uint32_t in_bytes, n_rec, n_blk_in, n_valid, n_surv, n_ob;
    uint32_t has_lo, has_hi, ulo_len, uhi_len;
    uint32_t tile_bytes, tile_keyb;
    uint32_t max_ukey, max_vlen, max_blk_size, max_blk_rec, max_ch;
    uint32_t lo[kMaxRuns], nblk[kMaxRuns], nrec[kMaxRuns], in_off[kMaxRuns], rec_base[kMaxRuns], blk_base[kMaxRuns];
    uint32_t vlo[kMaxRuns], vhi[kMaxRuns], nabove[kMaxRuns];
    uint32_t nx_tile, nx_err; // next tile's ticket + slice metadata, fetched while this tile is being written
    uint32_t nx_lo[kMaxRuns], nx_nblk[kMaxRuns], nx_nrec[kMaxRuns], nx_bytes[kMaxRuns], nx_grec0[kMaxRuns];
    unsigned long long nx_boff[kMaxRuns]; // byte offset of the slice inside its run
    uint32_t nx_krun[2], nx_koff[2], nx_klen[2]; // boundary keys of the next tile: run, offset and length inside its ikeys
    uint32_t k_run[2], k_off[2];
    uint32_t grec0[kMaxRuns]; // index of the slice's first record inside its run
    uint32_t scan[33];
    unsigned long long scan64[33];
    unsigned long long lb_bytes[4];
    uint32_t lb_blocks[4], lb_recs[4], lb_keyb[4], lb_inc[4];
    long long prev_tile; // this CTA's previous tile and its inclusive prefix
    unsigned long long prev_bytes;
    uint32_t prev_blocks, prev_recs, prev_keyb;
    uint32_t stat[16];
    uint32_t tb_off[kMaxTileBlocks], tb_size[kMaxTileBlocks], tb_rec[kMaxTileBlocks], tb_nrec[kMaxTileBlocks];
    uint32_t cut[kMaxOutBlocks + 1], ob_off[kMaxOutBlocks + 1], ob_size[kMaxOutBlocks], ob_keyoff[kMaxOutBlocks + 1];
    unsigned long long crc[256];
};

enum { ST_IN_REC = 0, ST_IN_BYTES, ST_OUT_REC, ST_OUT_BYTES, ST_SHADOW, ST_TOMB, ST_EXPIRED, ST_USER, ST_STALE, ST_TTL,
       ST_OUT_TOMB, ST_OUT_KEY, ST_OUT_VAL };

struct RecArrays {
    uint8_t *in;
    uint8_t *arena;
    unsigned long long *trailer;
    uint32_t *voff, *vlen, *koff, *R, *E;
    uint16_t *klen, *rank, *order, *surv, *shr, *blkid, *pos;
    uint8_t *flags;
    uint32_t total, arrays_end;
};
constexpr uint32_t kInPad = 64; // readable bytes after the staged blocks (unaligned word loads run a little past a value)
// Record arrays grow from the start of the pool, the staged input blocks sit at its END: the next tile's blocks can
// then be requested from the TMA unit while this tile's arrays are still live (see "early load" in k_merge).
PGS_DEV uint32_t in_start(uint32_t pool_bytes, uint32_t in_bytes) { return pool_bytes - kInPad - ((in_bytes + 15) & ~15u); }
PGS_DEV RecArrays carve(uint8_t *pool, uint32_t pool_bytes, uint32_t in_bytes, uint32_t n, uint32_t KS, uint32_t k)
{
    RecArrays a;
    uint32_t n8 = (n + 8) & ~7u; // >= n+1, multiple of 8
    uint32_t off = 0;
    a.arena = pool + off; off += n8 * KS;
    a.trailer = (unsigned long long *)(pool + off); off += n8 * 8;
    a.voff = (uint32_t *)(pool + off); off += n8 * 4;
    a.vlen = (uint32_t *)(pool + off); off += n8 * 4;
    a.koff = (uint32_t *)(pool + off); off += n8 * 4;
    a.R = (uint32_t *)(pool + off); off += n8 * 4;
    a.E = (uint32_t *)(pool + off); off += n8 * 4;
    a.klen = (uint16_t *)(pool + off); off += n8 * 2;
    a.rank = (uint16_t *)(pool + off); off += n8 * 2;
    a.order = (uint16_t *)(pool + off); off += n8 * 2;
    a.surv = (uint16_t *)(pool + off); off += n8 * 2;
    a.shr = (uint16_t *)(pool + off); off += n8 * 2;
    a.blkid = (uint16_t *)(pool + off); off += n8 * 2;
    a.pos = (uint16_t *)(pool + off); off += n8 * 2 * (k - 1); // merge positions inside the later runs
    a.flags = pool + off; off += n8;
    off = (off + 15) & ~15u;
    a.arrays_end = off;
    a.total = off + ((in_bytes + 15) & ~15u) + kInPad; // <= pool_bytes when the tile fits
    a.in = pool + (a.total <= pool_bytes ? in_start(pool_bytes, in_bytes) : 0);
    return a;
}

// exclusive scan of f(i), i in [0,n), into out[0..n] (out[n] = total); f is evaluated once per element.
// scratch = 33 uint32 of shared memory; two barriers per call.
template <class F>
PGS_DEV uint32_t chunked_scan(uint32_t n, uint32_t *out, uint32_t *scratch, F f)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t ipt = (n + blockDim.x - 1) / blockDim.x;
    uint32_t begin = min(threadIdx.x * ipt, n), end = min(begin + ipt, n);
    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) { uint32_t v = f(i); out[i] = v; local += v; }
    const uint32_t inc = warp_incl_scan(local, lane);
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    // every warp scans the warp totals itself: one barrier less than a designated scanning warp
    const uint32_t w = lane < nw ? scratch[lane] : 0, ws = warp_incl_scan(w, lane);
    const uint32_t total = __shfl_sync(kFull, ws, 31);
    uint32_t pre = __shfl_sync(kFull, ws - w, warp) + inc - local;
    for (uint32_t i = begin; i < end; i++) { uint32_t v = out[i]; out[i] = pre; pre += v; }
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads(); // results visible; scratch reusable
    return total;
}

// two exclusive scans in one pass: f(i) = (a, b) packed as a << 32 | b; neither running sum may pass 2^32.
// scratch64 = 33 x 8 bytes of shared memory.
template <class F>
PGS_DEV unsigned long long chunked_scan2(uint32_t n, uint32_t *outA, uint32_t *outB, unsigned long long *scratch64, F f)
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t ipt = (n + blockDim.x - 1) / blockDim.x;
    uint32_t begin = min(threadIdx.x * ipt, n), end = min(begin + ipt, n);
    unsigned long long local = 0;
    for (uint32_t i = begin; i < end; i++) { unsigned long long v = f(i); outA[i] = (uint32_t)(v >> 32); outB[i] = (uint32_t)v; local += v; }
    unsigned long long inc = local;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) scratch64[warp] = inc;
    __syncthreads();
    // every warp scans the warp totals itself: one barrier less than a designated scanning warp
    unsigned long long w = lane < nw ? scratch64[lane] : 0, ws = w;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(kFull, ws, d);
        if (lane >= d) ws += o;
    }
    const unsigned long long total = __shfl_sync(kFull, ws, 31);
    unsigned long long pre = __shfl_sync(kFull, ws - w, warp) + inc - local;
    for (uint32_t i = begin; i < end; i++) {
        unsigned long long v = ((unsigned long long)outA[i] << 32) | outB[i];
        outA[i] = (uint32_t)(pre >> 32); outB[i] = (uint32_t)pre;
        pre += v;
    }
    __syncthreads(); // results visible; scratch64 reusable
    return total;
}

PGS_DEV uint4 ld_v4_volatile(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
PGS_DEV void st_v4_volatile(void *p, uint4 v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// state 1 = the tile's own aggregate, 2 = inclusive prefix (overwrites the aggregate in place)
PGS_DEV void publish(TileAgg *slot, unsigned long long bytes, uint32_t blocks, uint32_t recs, uint32_t keyb, uint32_t state)
{
    st_v4_volatile(slot, make_uint4((uint32_t)bytes, (uint32_t)(bytes >> 32), blocks, state));
    st_v4_volatile((uint8_t *)slot + 16, make_uint4(recs, keyb, state, 0u));
}

// the varint32 encoding of v as little-endian bytes in a register (at most 5), *len = its length
PGS_DEV unsigned long long varint_pack(uint32_t v, uint32_t &len)
{
    unsigned long long o = 0;
    uint32_t n = 0;
    while (v >= 128) { o |= (unsigned long long)((v & 127u) | 128u) << (8 * n); v >>= 7; n++; }
    o |= (unsigned long long)v << (8 * n);
    len = n + 1;
    return o;
}

// one warp takes the next ticket and loads that tile's per-run slice metadata into S.nx_*
PGS_DEV void fetch_next_tile(const MergeParams &P, TileShared &S, uint32_t lane)
{
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(P.ticket, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (lane == 0) { S.nx_tile = t; S.nx_err = 0; }
    __syncwarp();
    if (t < P.Q && lane < P.k) {
        const RunDev &r = P.runs[lane];
        const bool first = t == 0, last = t == P.Q - 1;
        uint32_t lo = first ? 0 : P.split_pos[t * P.k + lane];
        uint32_t hi = last ? r.nb : P.split_pos[(t + 1) * P.k + lane];
        if (lo == 0xFFFFFFFFu || hi == 0xFFFFFFFFu || lo > r.nb || hi > r.nb || lo > hi) { atomicMax(&S.nx_err, (uint32_t)PGS_ABORTED); lo = hi = 0; }
        uint32_t hi_ex = last ? r.nb : min(hi + 1, r.nb);
        S.nx_lo[lane] = lo;
        S.nx_nblk[lane] = hi_ex - lo;
        const unsigned long long bo = r.blk_off[lo];
        S.nx_boff[lane] = bo;
        S.nx_bytes[lane] = (uint32_t)(r.blk_off[hi_ex] - bo);
        uint32_t g0 = r.blk_rec[lo];
        S.nx_nrec[lane] = r.blk_rec[hi_ex] - g0;
        S.nx_grec0[lane] = g0;
    }
    if (t < P.Q && (lane == 16 || lane == 17)) { // where the tile's boundary user keys live: (U_lo, U_hi]
        const uint32_t which = lane - 16;
        const bool none = which == 0 ? t == 0 : t == P.Q - 1;
        uint32_t run = 0, off = 0, len = 0;
        if (!none) {
            const uint32_t ref = P.split_ref[t + which];
            run = ref >> 28;
            const uint32_t b = ref & 0x0FFFFFFFu;
            if (ref == 0xFFFFFFFFu || run >= P.k || b >= P.runs[run].nb) { atomicMax(&S.nx_err, (uint32_t)PGS_ABORTED); run = 0; }
            else {
                off = P.runs[run].ikey_off[b];
                len = P.runs[run].ikey_off[b + 1] - off;
                if (len > P.KS) { atomicMax(&S.nx_err, (uint32_t)PGS_ABORTED); len = 0; }
            }
        }
        S.nx_krun[which] = run; S.nx_koff[which] = off; S.nx_klen[which] = len;
    }
}

template <uint32_t NT>
__global__ void __launch_bounds__(NT) k_merge(const __grid_constant__ MergeParams P)
{
    constexpr uint32_t NW = NT / 32;
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ TileShared S;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t KS = P.KS, RI = P.restart_interval;
    uint8_t *ulo = dyn, *uhi = dyn + KS + 8;
    uint8_t *pool = dyn + 2 * (KS + 8);

    if (tid == 0) {
        S.prev_tile = -1; S.prev_bytes = 0; S.prev_blocks = 0; S.prev_recs = 0; S.prev_keyb = 0;
        mbar_init((uint64_t *)&S.mbar, 1);
        mbar_fence_init();
    }
    if (P.validate_hash)
        for (uint32_t i = tid; i < 256; i += NT) S.crc[i] = P.crc_table[i];
    if (warp == 0) fetch_next_tile(P, S, lane);
    __syncthreads();
    uint32_t phase = 0;
    long long pt_last = P.phase_cycles ? clock64() : 0; // phase timing (diagnostics): thread 0 stamps every phase boundary
#define PT(i) do { if (P.phase_cycles && tid == 0) { long long t_ = clock64(); atomicAdd(&P.phase_cycles[i], (unsigned long long)(t_ - pt_last)); pt_last = t_; } } while (0)
    bool early = false; // thread 0: this tile's block loads were already issued during the previous tile's write phase

    for (;;) {
        // ---- tile setup (slice metadata and boundary-key references were prefetched into S.nx_*) -------------
        if (tid == 0) {
            const uint32_t t = S.nx_tile;
            S.tile = t;
            S.min_seq = ~0ull; S.max_seq = 0; S.max_ukey = 0; S.max_vlen = 0; S.max_blk_size = 0; S.max_blk_rec = 0; S.max_ch = 0;
            
            uint32_t err = S.nx_err;
            if (t < P.Q) {
                uint32_t bytes = 0, recs = 0, blks = 0;
                for (uint32_t j = 0; j < P.k; j++) {
                    S.lo[j] = S.nx_lo[j]; S.nblk[j] = S.nx_nblk[j]; S.nrec[j] = S.nx_nrec[j]; S.grec0[j] = S.nx_grec0[j];
                    S.in_off[j] = bytes; S.rec_base[j] = recs; S.blk_base[j] = blks;
                    bytes += S.nx_bytes[j]; recs += S.nx_nrec[j]; blks += S.nx_nblk[j];
                }
                S.in_bytes = bytes; S.n_rec = recs; S.n_blk_in = blks;
                S.k_run[0] = S.nx_krun[0]; S.k_run[1] = S.nx_krun[1];
                S.k_off[0] = S.nx_koff[0]; S.k_off[1] = S.nx_koff[1];
                S.ulo_len = S.nx_klen[0]; S.uhi_len = S.nx_klen[1];
                S.has_lo = t != 0;
                S.has_hi = t != P.Q - 1;
                const RecArrays a0 = carve(pool, P.pool_bytes, bytes, recs, KS, P.k);
                if (a0.total > P.pool_bytes || blks > kMaxTileBlocks || recs > 65000) err = max(err, (uint32_t)PGS_ABORTED);
                if (!err && P.use_tma && !early) {
                    // generic-proxy writes of the previous tile precede async-proxy writes to the same bytes
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (bytes) {
                        mbar_expect_tx((uint64_t *)&S.mbar, bytes);
                        for (uint32_t j = 0; j < P.k; j++) {
                            const uint32_t bj = S.nx_bytes[j];
                            if (bj) tma_load_1d(a0.in + S.in_off[j], P.runs[j].data + S.nx_boff[j], bj, (uint64_t *)&S.mbar);
                        }
                    }
                }
                early = false;
            }
            S.error = err;
        }
        if (tid >= 32 && tid < 48) S.stat[tid - 32] = 0;
        if (tid >= 64 && tid < 64 + kMaxRuns) { S.vlo[tid - 64] = 0; S.nabove[tid - 64] = 0; }
        __syncthreads();
        const uint32_t tile = S.tile;
        if (tile >= P.Q) break;
        const bool first = tile == 0, last = tile == P.Q - 1;
        const RecArrays A = carve(pool, P.pool_bytes, S.in_bytes, S.n_rec, KS, P.k);
        bool tile_ok = S.error == 0;
        if (tile_ok) {
            if (!P.use_tma) {
                for (uint32_t j = 0; j < P.k; j++) {
                    uint32_t bytes = (j + 1 < P.k ? S.in_off[j + 1] : S.in_bytes) - S.in_off[j];
                    const uint4 *src = (const uint4 *)(P.runs[j].data + P.runs[j].blk_off[S.lo[j]]);
                    uint4 *dst = (uint4 *)(A.in + S.in_off[j]);
                    for (uint32_t i = tid; i < bytes / 16; i += NT) dst[i] = src[i];
                }
            }
            // the three index reads below are independent global-memory round trips: different warps take them so
            // that they overlap instead of queueing behind each other
            if (warp < 2) { // boundary keys (zero padded slots): warp 0 -> U_lo, warp 1 -> U_hi
                if (!(warp == 0 ? first : last)) {
                    const uint8_t *src = P.runs[S.k_run[warp]].ikeys + S.k_off[warp];
                    const uint32_t len = warp == 0 ? S.ulo_len : S.uhi_len;
                    uint8_t *dst = warp == 0 ? ulo : uhi;
                    for (uint32_t i = lane; i < KS + 8; i += 64) {
                        const uint32_t i2 = i + 32;
                        const uint8_t v0 = i < len ? src[i] : 0, v1 = i2 < len ? src[i2] : 0;
                        dst[i] = v0;
                        if (i2 < KS + 8) dst[i2] = v1;
                    }
                }
            } else if (warp < 4) { // block table
                for (uint32_t t = tid - 64; t < S.n_blk_in; t += 64) {
                    uint32_t j = 0;
                    while (j + 1 < P.k && t >= S.blk_base[j + 1]) j++;
                    const RunDev &r = P.runs[j];
                    uint32_t gb = S.lo[j] + (t - S.blk_base[j]);
                    S.tb_off[t] = S.in_off[j] + (uint32_t)(r.blk_off[gb] - r.blk_off[S.lo[j]]);
                    S.tb_size[t] = r.blk_size[gb];
                    S.tb_rec[t] = S.rec_base[j] + (r.blk_rec[gb] - r.blk_rec[S.lo[j]]);
                    S.tb_nrec[t] = r.blk_rec[gb + 1] - r.blk_rec[gb];
                }
            } else { // every record's entry offset inside its block (device-built index): the header walk needs no chain
                for (uint32_t r = tid - 128; r < S.n_rec; r += NT - 128) {
                    uint32_t j = 0;
                    while (j + 1 < P.k && r >= S.rec_base[j + 1]) j++;
                    A.koff[r] = P.runs[j].rec_off[S.grec0[j] + (r - S.rec_base[j])];
                }
            }
            if (P.use_tma && S.in_bytes) {
                mbar_wait((uint64_t *)&S.mbar, phase);
                phase ^= 1;
            }
        }
        __syncthreads();
        PT(0);

        // ---- decode step 1: one THREAD per record parses its entry header (offsets come from the run's rec_off index) ----
        if (tile_ok) {
            const uint32_t nblk = S.n_blk_in;
            for (uint32_t r = tid; r < S.n_rec; r += NT) {
                uint32_t lo = 0, hi = nblk; // block of record r: last t with tb_rec[t] <= r
                while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (S.tb_rec[mid] <= r) lo = mid; else hi = mid; }
                const uint32_t t = lo;
                const uint8_t *base = A.in + S.tb_off[t];
                const uint32_t size = S.tb_size[t], i = r - S.tb_rec[t], cnt = S.tb_nrec[t];
                uint32_t err = 0, nr = 0;
                if (size < 8) err = PGS_CORRUPTION;
                if (!err) {
                    nr = ld_u32(base + size - 4);
                    if (nr == 0 || (unsigned long long)nr * 4 + 4 > size) err = PGS_CORRUPTION;
                }
                const uint32_t limit = err ? 0 : size - 4 - 4 * nr;
                const uint32_t p = A.koff[r];
                if (!err && (p >= limit || (i == 0 && p != 0))) err = PGS_CORRUPTION;
                if (!err) {
                    uint32_t shared, non_shared, vlen, h, c;
                    h = c = parse_header8(lds_u64_at(A.in, S.tb_off[t] + p), shared, non_shared, vlen); // header bytes from registers
                    if (!c) { // uncommon shape: byte-wise decoder
                        h = 0;
                        c = get_varint32(base + p, limit - p, shared);
                        h += c;
                        if (c) { c = get_varint32(base + p + h, limit - p - h, non_shared); h += c; }
                        if (c) { c = get_varint32(base + p + h, limit - p - h, vlen); h += c; }
                    }
                    const uint32_t klen = shared + non_shared;
                    const unsigned long long end = (unsigned long long)p + h + non_shared + vlen;
                    if (!c || klen < 8 || klen - 8 > KS || end > limit || (i == 0 && shared != 0) || (i + 1 == cnt && end != limit)) {
                        err = PGS_CORRUPTION;
                    } else {
                        {
                            A.rank[r] = (uint16_t)shared;      // scratch until the rank phase
                            A.order[r] = (uint16_t)non_shared; // scratch until the scatter phase
                        }
                        A.koff[r] = S.tb_off[t] + p + h;
                        A.klen[r] = (uint16_t)(klen - 8);
                        A.voff[r] = S.tb_off[t] + p + h + non_shared;
                        A.vlen[r] = vlen;
                        if (non_shared >= 8) { // the (seq<<8|type) trailer sits wholly in this entry's delta
                            A.trailer[r] = lds_u64_at(A.in, S.tb_off[t] + p + h + non_shared - 8);
                            A.flags[r] = 0;
                        } else {               // part of it is shared with the previous key: rebuilt in step 2
                            A.trailer[r] = 0;
                            A.flags[r] = 1;
                            
                        }
                    }
                }
                if (err) atomicMax(&S.error, err);
            }
        }
        __syncthreads();
        PT(1);
        tile_ok = S.error == 0;

        // ---- decode step 2: HALF a warp per block rebuilds the keys, four key bytes per lane -----------
        // lane L (0..15) of a half-warp owns internal-key positions 4L..4L+3 (+64 per pass) as one 32-bit word: an
        // entry overwrites the bytes of the word that its delta covers (one unaligned load + byte mask) and inherits
        // the rest from the entry before it.
        
        if (tile_ok) {
            const uint32_t hl = lane & 15, sub = lane >> 4;
            const uint32_t hmask = sub ? 0xffff0000u : 0x0000ffffu;
            for (uint32_t t = 2 * warp + sub; t < S.n_blk_in; t += 2 * NW) {
                const uint32_t rec0 = S.tb_rec[t], nrec = S.tb_nrec[t];
                uint32_t maxk = 0;
                for (uint32_t i = hl; i < nrec; i += 16) maxk = max(maxk, (uint32_t)A.klen[rec0 + i] + 8);
                maxk = __reduce_max_sync(hmask, maxk);
                for (uint32_t pass = 0; pass * 64 < maxk; pass++) {
                    const uint32_t p0 = pass * 64 + 4 * hl;
                    uint32_t cur = 0, prev_klen = 0; // the four running bytes, little endian
                    uint32_t sh = 0, ns = 0, ulen = 0, ko = 0, fl = 0;
                    {
                        if (nrec) { sh = A.rank[rec0]; ns = A.order[rec0]; ulen = A.klen[rec0]; ko = A.koff[rec0]; fl = A.flags[rec0]; }
                    }
                    for (uint32_t i = 0; i < nrec; i++) {
                        const uint32_t r = rec0 + i;
                        const uint32_t c_sh = sh, c_ns = ns, c_ulen = ulen, c_ko = ko, c_fl = fl;
                        {
                            if (i + 1 < nrec) { sh = A.rank[r + 1]; ns = A.order[r + 1]; ulen = A.klen[r + 1]; ko = A.koff[r + 1]; fl = A.flags[r + 1]; } // next entry's metadata in flight
                        }
                        if (c_sh > prev_klen) { if (hl == 0) atomicMax(&S.error, (uint32_t)PGS_CORRUPTION); break; } // a prefix longer than the previous key
                        prev_klen = c_ulen + 8;
                        const uint32_t a = max(c_sh, p0), b = min(c_sh + c_ns, p0 + 4);
                        if (a < b) {
                            const uint32_t so = c_ko + (a - c_sh); // delta bytes for positions a..a+3 (offset inside IN, which is 16-aligned)
                            const uint32_t *w = (const uint32_t *)A.in + (so >> 2);
                            const uint32_t x = __funnelshift_r(w[0], w[1], (so & 3) * 8);
                            const uint32_t s0 = 8 * (a - p0), s1 = 8 * (p0 + 4 - b);
                            const uint32_t msk = (0xffffffffu << s0) & (0xffffffffu >> s1);
                            cur = (cur & ~msk) | ((x << s0) & msk);
                        }
                        const uint32_t pad = (c_ulen + 7) & ~7u; // slots are zero padded to 8 bytes
                        if (p0 < pad) {
                            const uint32_t keep = c_ulen > p0 ? c_ulen - p0 : 0; // bytes of this word that belong to the user key
                            *(uint32_t *)(A.arena + (size_t)r * KS + p0) = keep >= 4 ? cur : (cur & ((1u << (8 * keep)) - 1u));
                        }
                        // rare: the 8 trailer bytes after the user key straddle the shared prefix
                        if (c_fl && pass * 64 < c_ulen + 8 && pass * 64 + 64 > c_ulen) {
                            unsigned long long c = 0;
                            if (p0 >= c_ulen) { if (p0 < c_ulen + 8) c = (unsigned long long)cur << (8 * (p0 - c_ulen)); }
                            else if (c_ulen - p0 < 4) c = cur >> (8 * (c_ulen - p0));
                            const uint32_t lo = __reduce_or_sync(hmask, (uint32_t)c), hi = __reduce_or_sync(hmask, (uint32_t)(c >> 32));
                            if (hl == 0) A.trailer[r] |= ((unsigned long long)hi << 32) | lo;
                        }
                    }
                }
            }
        }
        __syncthreads();
        PT(2);

        // ---- valid range of every run's slice: user keys in (U_lo, U_hi]; records at or below U_lo form a
        //      prefix of a slice, records above U_hi a suffix, so counting them gives the window -----------
        if (tile_ok) {
            for (uint32_t r = tid; r < S.n_rec; r += NT) {
                uint32_t j = 0;
                while (j + 1 < P.k && r >= S.rec_base[j + 1]) j++;
                const uint8_t *key = A.arena + (size_t)r * KS;
                uint32_t kl = A.klen[r];
                if (S.has_lo && cmp_slots(key, kl, ulo, S.ulo_len) <= 0) atomicAdd(&S.vlo[j], 1u);
                else if (S.has_hi && cmp_slots(key, kl, uhi, S.uhi_len) > 0) atomicAdd(&S.nabove[j], 1u);
            }
        }
        __syncthreads();
        if (tile_ok && tid < P.k) {
            uint32_t vhi = S.nrec[tid] - S.nabove[tid];
            if (vhi < S.vlo[tid]) vhi = S.vlo[tid];
            S.vhi[tid] = vhi;
        }
        __syncthreads();
        PT(3);
        if (tile_ok && tid == 0) {
            uint32_t nv = 0;
            for (uint32_t j = 0; j < P.k; j++) nv += S.vhi[j] - S.vlo[j];
            S.n_valid = nv;
        }

        // ---- merge rank + shadow detection -----------------------------------------------------------------
        // (1) one thread per record: validity, position inside its own run, predecessor of the same run;
        if (tile_ok) {
            for (uint32_t r = tid; r < S.n_rec; r += NT) {
                uint32_t j = 0;
                while (j + 1 < P.k && r >= S.rec_base[j + 1]) j++;
                uint32_t idx = r - S.rec_base[j];
                if (idx < S.vlo[j] || idx >= S.vhi[j]) { A.flags[r] = 0; continue; }
                const uint32_t kl = A.klen[r];
                bool shadow = idx > 0 && A.klen[r - 1] == kl && cmp_slots(A.arena + (size_t)(r - 1) * KS, kl, A.arena + (size_t)r * KS, kl) == 0;
                A.R[r] = idx - S.vlo[j];
                A.E[r] = shadow ? 1u : 0u;
                A.flags[r] = F_VALID;
            }
        }
        __syncthreads();
        // (2) one thread per (record of run j, LATER run o): LCP-aware binary search for the number of o's records that
        //     sort before it (kept in pos[]); (3) one thread per (record of run o, EARLIER run j): the number of j's
        //     records before it is an upper bound over j's monotone pos[] column -- integer compares only.  Every pair
        //     of runs pays key compares in one direction; ranks accumulate with shared-memory atomics.
        PT(12);
        if (tile_ok && P.k > 1) {
            const uint32_t k = P.k, km1 = k - 1;
            uint32_t ntask = 0;
            for (uint32_t j = 0; j < k; j++) ntask += S.nrec[j] * (km1 - j);
            for (uint32_t id = tid; id < ntask; id += NT) {
                uint32_t j = 0, local = id;
                while (local >= S.nrec[j] * (km1 - j)) { local -= S.nrec[j] * (km1 - j); j++; }
                const uint32_t nt = km1 - j, ri = local / nt;
                const uint32_t o = j + 1 + (local - ri * nt), r = S.rec_base[j] + ri;
                if (!(A.flags[r] & F_VALID)) continue;
                const uint8_t *key = A.arena + (size_t)r * KS;
                const uint32_t kl = A.klen[r];
                const unsigned long long tr = A.trailer[r];
                uint32_t base = S.rec_base[o], lo = S.vlo[o], hi = S.vhi[o];
                uint32_t lcp_lo = 0, lcp_hi = 0; // words shared with the keys just outside [lo, hi)
                while (lo < hi) { // first position whose internal key is not before ours
                    uint32_t mid = (lo + hi) >> 1, q = base + mid, d;
                    int c = cmp_slots_from(A.arena + (size_t)q * KS, A.klen[q], key, kl, min(lcp_lo, lcp_hi), &d);
                    bool before;
                    if (c != 0) before = c < 0;
                    else {
                        unsigned long long tq = A.trailer[q];
                        before = tq > tr || (tq == tr && o < j);
                    }
                    if (before) { lo = mid + 1; lcp_lo = d; } else { hi = mid; lcp_hi = d; }
                }
                const uint32_t cnt = lo - S.vlo[o];
                A.pos[(size_t)r * km1 + (o - 1)] = (uint16_t)cnt;
                if (cnt) {
                    atomicAdd(&A.R[r], cnt);
                    uint32_t q = base + lo - 1;
                    if (A.klen[q] == kl && cmp_slots(A.arena + (size_t)q * KS, kl, key, kl) == 0) atomicOr(&A.E[r], 1u);
                }
            }
            __syncthreads();
            PT(13);
            ntask = 0;
            for (uint32_t o = 1; o < k; o++) ntask += S.nrec[o] * o;
            for (uint32_t id = tid; id < ntask; id += NT) {
                uint32_t o = 1, local = id;
                while (local >= S.nrec[o] * o) { local -= S.nrec[o] * o; o++; }
                const uint32_t qi = local / o, j = local - qi * o, q = S.rec_base[o] + qi;
                if (!(A.flags[q] & F_VALID)) continue;
                const uint32_t me = qi - S.vlo[o]; // j's record r sorts before q  <=>  pos[r -> o] <= me
                const uint32_t base = S.rec_base[j];
                uint32_t lo = S.vlo[j], hi = S.vhi[j];
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (A.pos[(size_t)(base + mid) * km1 + (o - 1)] <= me) lo = mid + 1; else hi = mid;
                }
                const uint32_t cnt = lo - S.vlo[j];
                if (cnt) {
                    atomicAdd(&A.R[q], cnt);
                    const uint32_t r = base + lo - 1, kl = A.klen[q];
                    if (A.klen[r] == kl && cmp_slots(A.arena + (size_t)r * KS, kl, A.arena + (size_t)q * KS, kl) == 0) atomicOr(&A.E[q], 1u);
                }
            }
        }
        __syncthreads();
        PT(4);

        // ---- compaction filter + tombstone policy; scatter into merged order -----------------------------
        if (tile_ok) {
            uint32_t s_in = 0, s_inb = 0, s_sh = 0, s_tomb = 0, s_exp = 0, s_user = 0, s_stale = 0, s_ttl = 0;
            for (uint32_t r = tid; r < S.n_rec; r += NT) {
                uint8_t f = A.flags[r];
                if (!(f & F_VALID)) continue;
                uint32_t kl = A.klen[r], vl = A.vlen[r];
                s_in++;
                s_inb += kl + vl;
                A.order[A.R[r]] = (uint16_t)r;
                if (A.E[r]) { s_sh++; continue; }
                uint8_t type = (uint8_t)A.trailer[r];
                if (type == PGS_TYPE_VALUE) {
                    uint32_t nts = 0;
                    bool changed;
                    uint8_t *val = A.in + A.voff[r];
                    uint32_t why = dev_filter(P, S.crc, A.arena + (size_t)r * KS, kl, val, vl, nts, changed);
                    if (why) {
                        if (why == 1) s_exp++; else if (why == 2) s_user++; else s_stale++;
                        // Decision::kRemove turns the entry into a deletion; it disappears only at the bottommost level
                        if (!P.bottommost) { f |= F_KEEP | F_TOMB; A.vlen[r] = 0; }
                    } else {
                        f |= F_KEEP;
                        if (changed) {
                            s_ttl++;
                            val[0] = (uint8_t)(nts >> 24); val[1] = (uint8_t)(nts >> 16); val[2] = (uint8_t)(nts >> 8); val[3] = (uint8_t)nts;
                        }
                    }
                } else if (type == PGS_TYPE_DELETION) {
                    if (P.bottommost) s_tomb++; else f |= F_KEEP | F_TOMB;
                } else {
                    f |= F_KEEP;
                }
                A.flags[r] = f;
            }
            // one shared-memory atomic per warp and counter (a tile's counts fit 32 bits)
            s_in = __reduce_add_sync(kFull, s_in); s_inb = __reduce_add_sync(kFull, s_inb);
            s_sh = __reduce_add_sync(kFull, s_sh); s_tomb = __reduce_add_sync(kFull, s_tomb);
            s_exp = __reduce_add_sync(kFull, s_exp); s_user = __reduce_add_sync(kFull, s_user);
            s_stale = __reduce_add_sync(kFull, s_stale); s_ttl = __reduce_add_sync(kFull, s_ttl);
            if (lane == 0) {
                if (s_in) atomicAdd(&S.stat[ST_IN_REC], s_in);
                if (s_inb) atomicAdd(&S.stat[ST_IN_BYTES], s_inb);
                if (s_sh) atomicAdd(&S.stat[ST_SHADOW], s_sh);
                if (s_tomb) atomicAdd(&S.stat[ST_TOMB], s_tomb);
                if (s_exp) atomicAdd(&S.stat[ST_EXPIRED], s_exp);
                if (s_user) atomicAdd(&S.stat[ST_USER], s_user);
                if (s_stale) atomicAdd(&S.stat[ST_STALE], s_stale);
                if (s_ttl) atomicAdd(&S.stat[ST_TTL], s_ttl);
            }
        }
        __syncthreads();
        PT(5);

        // ---- survivors in merged order, output block layout ---------------------------------------------------
        uint32_t m = 0;
        if (tile_ok) {
            const uint32_t nv = S.n_valid;
            // one pass scans both the survivor count and the survivors' raw sizes over the merged order.  Raw sizes
            // decide the block cuts: an entry starts a new block when its raw offset enters the next block_size
            // window (blocks hold ~block_size raw bytes; every entry's block is known in parallel)
            m = (uint32_t)(chunked_scan2(nv, A.E, A.koff, S.scan64, [&](uint32_t p) -> unsigned long long {
                const uint32_t r = A.order[p];
                return (A.flags[r] & F_KEEP) ? ((1ull << 32) | (A.klen[r] + 8u + A.vlen[r] + 3u)) : 0ull;
            }) >> 32);
            for (uint32_t p = tid; p < nv; p += NT) {
                uint32_t r = A.order[p];
                if (A.flags[r] & F_KEEP) { const uint32_t q = A.E[p]; A.surv[q] = (uint16_t)r; A.R[q] = A.koff[p]; }
            }
            __syncthreads();
            const uint32_t BS = P.block_size;
            uint32_t nob = chunked_scan(m, A.E, S.scan, [&](uint32_t p) -> uint32_t {
                return (p == 0 || A.R[p] / BS != A.R[p - 1] / BS) ? 1u : 0u;
            });
            if (nob > kMaxOutBlocks) { if (tid == 0) atomicMax(&S.error, (uint32_t)PGS_ABORTED); nob = 0; }
            for (uint32_t p = tid; p < m; p += NT) {
                bool firstp = p == 0 || A.R[p] / BS != A.R[p - 1] / BS;
                uint32_t b = A.E[p] + (firstp ? 1u : 0u) - 1u; // E = exclusive count of block starts before p
                A.blkid[p] = (uint16_t)b;
                if (firstp && b < kMaxOutBlocks) S.cut[b] = p;
            }
            if (tid == 0) { S.cut[nob] = m; S.n_ob = nob; S.n_surv = m; }
            __syncthreads();
            tile_ok = S.error == 0;
        }
        if (tile_ok) {
            const uint32_t nob = S.n_ob;
            // prefix compression against the previous survivor + encoded sizes
            chunked_scan(m, A.E, S.scan, [&](uint32_t p) -> uint32_t {
                uint32_t b = A.blkid[p];
                uint32_t r = A.surv[p], kl = A.klen[r], shared = 0;
                if ((p - S.cut[b]) % RI != 0) {
                    uint32_t q = A.surv[p - 1];
                    shared = lcp_slots(A.arena + (size_t)q * KS, A.klen[q], A.arena + (size_t)r * KS, kl);
                }
                A.shr[p] = (uint16_t)shared;
                uint32_t ns = kl + 8 - shared, vl = A.vlen[r];
                uint32_t hs = varint_len(shared) + varint_len(ns) + varint_len(vl) + ns;
                A.rank[p] = (uint16_t)hs; // entry head bytes (varints + key delta + trailer)
                return hs + vl;
            });
            if (warp == 0) { // block sizes, 16-byte aligned offsets and index-key offsets: one warp, shuffle scans
                uint32_t off = 0, koff = 0;
                for (uint32_t b0 = 0; b0 < nob; b0 += 32) {
                    uint32_t b = b0 + lane, sz = 0, al = 0, kl = 0;
                    if (b < nob) {
                        uint32_t cnt = S.cut[b + 1] - S.cut[b];
                        uint32_t nrest = (cnt + RI - 1) / RI;
                        sz = A.E[S.cut[b + 1]] - A.E[S.cut[b]] + 4 * (nrest + 1);
                        al = (sz + kBlockAlign - 1) & ~(kBlockAlign - 1);
                        kl = A.klen[A.surv[S.cut[b + 1] - 1]];
                        S.ob_size[b] = sz;
                    }
                    uint32_t ia = warp_incl_scan(al, lane), ik = warp_incl_scan(kl, lane);
                    if (b < nob) { S.ob_off[b] = off + ia - al; S.ob_keyoff[b] = koff + ik - kl; }
                    off += __shfl_sync(kFull, ia, 31);
                    koff += __shfl_sync(kFull, ik, 31);
                }
                if (lane == 0) { S.ob_off[nob] = off; S.ob_keyoff[nob] = koff; S.tile_bytes = off; S.tile_keyb = koff; }
            }
        } else if (tid == 0) {
            S.n_ob = 0; S.n_surv = 0; S.tile_bytes = 0; S.tile_keyb = 0;
        }
        
        __syncthreads();
        PT(6);

        // ---- decoupled look-back: where does this tile's output start? ----------------------------------------
        // Warps 0..3 each resolve a window of 64 predecessors at the same time (one global round trip covers 256
        // tiles, more than are ever in flight with one CTA per SM); warp 0 adds the windows up to the nearest
        // inclusive prefix.
        constexpr uint32_t LBW = 4;
        if (warp < LBW) {
            const unsigned long long my_bytes = S.tile_bytes;
            const uint32_t my_blocks = S.n_ob, my_recs = S.n_surv, my_keyb = S.tile_keyb;
            if (tid == 0) publish(&P.agg[tile], my_bytes, my_blocks, my_recs, my_keyb, 1u);
            unsigned long long ex_bytes = 0; // thread 0 only
            uint32_t ex_blocks = 0, ex_recs = 0, ex_keyb = 0;
            int64_t look = (int64_t)tile - 1 - 64 * (int64_t)warp;
            for (;;) { // lane L of window w reads tiles look-L and look-32-L; a tile before tile 0 is an inclusive prefix of zero
                uint32_t have_inc[2] = {0, 0};
                unsigned long long b[2] = {0, 0};
                uint32_t bl[2] = {0, 0}, rc[2] = {0, 0}, kb[2] = {0, 0};
                const int64_t idx0 = look - lane, idx1 = look - 32 - lane;
                // Tiles at or before this CTA's previous tile need no load: its inclusive prefix is still in shared memory
                // (every CTA takes tickets in increasing order), so a look-back reads only the tiles in between.
                uint32_t done = 0; // bit h: slot h is resolved
                const int64_t prev = S.prev_tile; // -1 before the CTA's first tile: "tile -1" has an inclusive prefix of zero
#pragma unroll
                for (uint32_t h = 0; h < 2; h++) {
                    const int64_t idx = h ? idx1 : idx0;
                    if (idx <= prev) {
                        done |= 1u << h; have_inc[h] = 1;
                        if (idx == prev) { b[h] = S.prev_bytes; bl[h] = S.prev_blocks; rc[h] = S.prev_recs; kb[h] = S.prev_keyb; }
                    }
                }
                for (;;) { // both halves of each predecessor's slot in one round trip
#pragma unroll
                    for (uint32_t h = 0; h < 2; h++) {
                        if (done >> h & 1) continue;
                        const int64_t idx = h ? idx1 : idx0;
                        const uint4 a0 = ld_v4_volatile(&P.agg[idx]), a1 = ld_v4_volatile((const uint8_t *)&P.agg[idx] + 16);
                        if (a0.w && a0.w == a1.z) {
                            have_inc[h] = a0.w == 2u;
                            b[h] = ((unsigned long long)a0.y << 32) | a0.x; bl[h] = a0.z; rc[h] = a1.x; kb[h] = a1.y; done |= 1u << h;
                        }
                    }
                    // finished when every position before the window's nearest inclusive prefix is resolved (positions:
                    // the 32 lanes of half 0, then the 32 lanes of half 1)
                    const uint32_t f0 = __ballot_sync(kFull, have_inc[0]), n0 = __ballot_sync(kFull, !(done & 1u));
                    if (f0) { if (!(n0 & ((1u << (__ffs(f0) - 1)) - 1u))) break; continue; }
                    if (n0) continue;
                    const uint32_t f1 = __ballot_sync(kFull, have_inc[1]), n1 = __ballot_sync(kFull, !(done & 2u));
                    if (f1) { if (!(n1 & ((1u << (__ffs(f1) - 1)) - 1u))) break; continue; }
                    if (!n1) break;
                }
                const uint32_t inc0 = __ballot_sync(kFull, have_inc[0]), inc1 = __ballot_sync(kFull, have_inc[1]);
                // nearest predecessor with an inclusive prefix: position = lane (first half) or 32 + lane (second half)
                const uint32_t stop = inc0 ? (uint32_t)__ffs(inc0) - 1 : (inc1 ? 32u + (uint32_t)__ffs(inc1) - 1 : 63u);
                if (lane > stop) { b[0] = 0; bl[0] = 0; rc[0] = 0; kb[0] = 0; }
                if (32 + lane > stop) { b[1] = 0; bl[1] = 0; rc[1] = 0; kb[1] = 0; }
                unsigned long long sb = b[0] + b[1];
                uint32_t sbl = bl[0] + bl[1], src = rc[0] + rc[1], skb = kb[0] + kb[1];
                for (uint32_t d = 16; d; d >>= 1) {
                    sb += __shfl_xor_sync(kFull, sb, d);
                    sbl += __shfl_xor_sync(kFull, sbl, d);
                    src += __shfl_xor_sync(kFull, src, d);
                    skb += __shfl_xor_sync(kFull, skb, d);
                }
                if (lane == 0) { S.lb_bytes[warp] = sb; S.lb_blocks[warp] = sbl; S.lb_recs[warp] = src; S.lb_keyb[warp] = skb; S.lb_inc[warp] = inc0 | inc1; }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                uint32_t found = 0;
#pragma unroll
                for (uint32_t w = 0; w < LBW; w++) found |= S.lb_inc[w];
                if (tid == 0) {
                    for (uint32_t w = 0; w < LBW; w++) {
                        ex_bytes += S.lb_bytes[w]; ex_blocks += S.lb_blocks[w]; ex_recs += S.lb_recs[w]; ex_keyb += S.lb_keyb[w];
                        if (S.lb_inc[w]) break;
                    }
                }
                if (found) break;
                asm volatile("bar.sync 1, 128;" ::: "memory"); // the windows' sums are consumed before the next round overwrites them
                look -= 64 * LBW;
            }
            if (tid == 0) {
                publish(&P.agg[tile], ex_bytes + my_bytes, ex_blocks + my_blocks, ex_recs + my_recs, ex_keyb + my_keyb, 2u);
                S.prev_tile = (long long)tile;
                S.prev_bytes = ex_bytes + my_bytes; S.prev_blocks = ex_blocks + my_blocks; S.prev_recs = ex_recs + my_recs; S.prev_keyb = ex_keyb + my_keyb;
                S.base_bytes = ex_bytes;
                S.base_blocks = ex_blocks;
                S.base_recs = ex_recs;
                S.base_keyb = ex_keyb;
                if (ex_bytes + my_bytes > P.out_cap || ex_blocks + my_blocks > P.out_blk_cap || ex_keyb + my_keyb > P.out_ikey_cap)
                    atomicMax(&S.error, (uint32_t)PGS_ABORTED);
            }
            PT(8);
        } else {
          // ... while the other warps prepare the writes: nothing here needs the tile's output base
          if (warp == NW - 1) fetch_next_tile(P, S, lane); // next ticket + slice metadata: global round trips under the look-back
          if (tile_ok && m > 0) {
            uint32_t s_outb = 0, s_otomb = 0, s_okey = 0, s_oval = 0, mx_k = 0, mx_v = 0;
            unsigned long long mn_seq = ~0ull, mx_seq = 0;
            // (a) entry start offsets inside the tile's output + per-survivor stats, one thread per survivor
            uint32_t max_chunks = 0;
            for (uint32_t p = tid - 32 * LBW; p < m; p += NT - 32 * LBW) {
                const uint32_t r = A.surv[p], b = A.blkid[p];
                const uint32_t kl = A.klen[r], vl = A.vlen[r];
                const uint32_t eoff = S.ob_off[b] + (A.E[p] - A.E[S.cut[b]]);
                A.R[p] = eoff;
                const uint8_t f = A.flags[r];
                const unsigned long long tr = A.trailer[r];
                const uint8_t type = (f & F_TOMB) ? (uint8_t)PGS_TYPE_DELETION : (uint8_t)tr;
                const unsigned long long seq = (P.bottommost && type == PGS_TYPE_VALUE) ? 0ull : (tr >> 8);
                A.trailer[r] = (seq << 8) | type; // the trailer as written
                s_outb += kl + vl;
                s_otomb += type == PGS_TYPE_DELETION;
                s_okey += kl;
                s_oval += vl;
                mx_k = max(mx_k, kl);
                mx_v = max(mx_v, vl);
                mn_seq = seq < mn_seq ? seq : mn_seq;
                mx_seq = seq > mx_seq ? seq : mx_seq;
                max_chunks = max(max_chunks, ((vl >> 4) + 3) >> 1); // pairs of 16-byte chunks
                
            }
            s_outb = __reduce_add_sync(kFull, s_outb); s_otomb = __reduce_add_sync(kFull, s_otomb);
            s_okey = __reduce_add_sync(kFull, s_okey); s_oval = __reduce_add_sync(kFull, s_oval);
            mx_k = __reduce_max_sync(kFull, mx_k); mx_v = __reduce_max_sync(kFull, mx_v);
            max_chunks = __reduce_max_sync(kFull, max_chunks);
            
            for (uint32_t d = 16; d; d >>= 1) {
                unsigned long long o1 = __shfl_xor_sync(kFull, mn_seq, d), o2 = __shfl_xor_sync(kFull, mx_seq, d);
                mn_seq = o1 < mn_seq ? o1 : mn_seq;
                mx_seq = o2 > mx_seq ? o2 : mx_seq;
            }
            if (lane == 0) {
                atomicMax(&S.max_ch, max_chunks);
                
                atomicAdd(&S.stat[ST_OUT_BYTES], s_outb);
                atomicAdd(&S.stat[ST_OUT_TOMB], s_otomb);
                atomicAdd(&S.stat[ST_OUT_KEY], s_okey);
                atomicAdd(&S.stat[ST_OUT_VAL], s_oval);
                atomicMax(&S.max_ukey, mx_k);
                atomicMax(&S.max_vlen, mx_v);
                atomicMin(&S.min_seq, mn_seq);
                atomicMax(&S.max_seq, mx_seq);
            }
          }
        }
        __syncthreads();
        PT(7);
        tile_ok = S.error == 0;

        // ---- write the tile's blocks ------------------------------------------------------------------------------
        if (tile_ok && m > 0) {
            const uint32_t nob = S.n_ob;
            uint8_t *out = P.out_data + S.base_bytes;
            const uint32_t CH = S.max_ch;
            // (b) values first: one thread per PAIR of 16-byte destination-aligned chunks (CH pairs per survivor),
            //     source words re-aligned with funnel shifts.  The first and last chunk of a value are written as
            //     FULL 16-byte stores whenever the bytes that do not belong to the value fall inside this entry's
            //     own head or the next entry's head of the same block: those heads are written after the barrier
            //     below and overwrite the spill.  Where a spill could touch foreign bytes, a tail is stored as
            //     8/4/2/1-byte pieces (the chunk start is 16-aligned) and a head byte by byte.
            const uint32_t ch_magic = (uint32_t)((0x100000000ull + CH - 1) / CH); // id / CH by multiply-high (exact for id*CH < 2^32)
            
            for (uint32_t id = tid; id < m * CH; id += NT) {
                const uint32_t p = CH == 1 ? id : __umulhi(id, ch_magic), c0 = (id - p * CH) << 1;
                const uint32_t r = A.surv[p];
                const uint32_t vl = A.vlen[r];
                if (vl == 0) continue;
                const uint32_t hs = A.rank[p];
                uint8_t *dv = out + A.R[p] + hs;
                const uint32_t lead = (uint32_t)((uintptr_t)dv & 15);
                const uint32_t nch = (lead + vl + 15) >> 4;
                if (c0 >= nch) continue;
                const int32_t so = (int32_t)(A.voff[r] + (c0 << 4)) - (int32_t)lead; // may start a few bytes before the value (even before IN): still inside the pool
                const uint8_t *sp = A.in + so;
                const uint32_t sh = (uint32_t)(so & 3) * 8;
                const uint32_t *w = (const uint32_t *)A.in + (so >> 2); // IN is 16-aligned; the offset arithmetic keeps the loads in the shared window
                uint32_t wv[9];
#pragma unroll
                for (uint32_t x = 0; x < 9; x++) wv[x] = w[x];
#pragma unroll
                for (uint32_t half = 0; half < 2; half++) {
                    const uint32_t c = c0 + half;
                    if (c >= nch) break;
                    uint4 o4;
                    o4.x = __funnelshift_r(wv[4 * half + 0], wv[4 * half + 1], sh); o4.y = __funnelshift_r(wv[4 * half + 1], wv[4 * half + 2], sh);
                    o4.z = __funnelshift_r(wv[4 * half + 2], wv[4 * half + 3], sh); o4.w = __funnelshift_r(wv[4 * half + 3], wv[4 * half + 4], sh);
                    uint8_t *addr = dv - lead + (c << 4);
                    const uint32_t lo = c == 0 ? lead : 0;
                    const uint32_t rem = lead + vl - (c << 4);
                    const uint32_t hi = rem < 16 ? rem : 16;
                    bool full = lo <= hs; // else the spill before the value would reach the previous entry
                    if (hi < 16) {
                        const bool block_last = p + 1 == S.cut[A.blkid[p] + 1];
                        if (block_last || 16 - hi > A.rank[p + 1]) full = false;
                    }
                    if (full) {
                        *reinterpret_cast<uint4 *>(addr) = o4;
                    } else if (lo == 0) { // exact tail [0, hi), hi < 16
                        uint32_t at = 0;
                        if (hi & 8) { *reinterpret_cast<uint2 *>(addr) = make_uint2(o4.x, o4.y); at = 8; }
                        const uint32_t q0 = (hi & 8) ? o4.z : o4.x, q1 = (hi & 8) ? o4.w : o4.y;
                        uint32_t q = q0;
                        if (hi & 4) { *reinterpret_cast<uint32_t *>(addr + at) = q0; at += 4; q = q1; }
                        if (hi & 2) { *reinterpret_cast<uint16_t *>(addr + at) = (uint16_t)q; at += 2; q >>= 16; }
                        if (hi & 1) addr[at] = (uint8_t)q;
                    } else {
                        const uint8_t *sb = sp + (half << 4);
                        for (uint32_t x = lo; x < hi; x++) addr[x] = sb[x];
                    }
                }
            }
            __syncthreads();
            PT(9);
            // early load: the staged blocks are dead now.  When the next tile's blocks (end of the pool) do not reach
            // into this tile's record arrays (start of the pool, still needed below), their TMA copies start here
            // and run under the head writes, the block trailers and the next tile's setup.
            if (tid == 0 && P.use_tma && P.early_tma && S.nx_tile < P.Q && !S.nx_err) {
                uint32_t nb = 0, nr = 0, nbl = 0;
                for (uint32_t j = 0; j < P.k; j++) { nb += S.nx_bytes[j]; nr += S.nx_nrec[j]; nbl += S.nx_nblk[j]; }
                const RecArrays nx = carve(pool, P.pool_bytes, nb, nr, KS, P.k);
                if (nb && nx.total <= P.pool_bytes && nbl <= kMaxTileBlocks && nr <= 65000 && in_start(P.pool_bytes, nb) >= A.arrays_end) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx((uint64_t *)&S.mbar, nb);
                    uint8_t *dst = nx.in;
                    for (uint32_t j = 0; j < P.k; j++) {
                        const uint32_t bytes = S.nx_bytes[j];
                        if (bytes) tma_load_1d(dst, P.runs[j].data + S.nx_boff[j], bytes, (uint64_t *)&S.mbar);
                        dst += bytes;
                    }
                    early = true;
                }
            }
            // (c) entry heads = 3 varints | key delta | trailer.  A quarter warp per survivor; every lane assembles one
            //     destination-aligned 32-bit word of the head from its three sources (packed varints in a register, key
            //     bytes in the arena slot, trailer) with shifts and byte masks, and stores it whole; only a first or
            //     last word that the head covers partly goes out byte by byte.
            for (uint32_t p = 4 * warp + (lane >> 3); p < m; p += 4 * NW) {
                const uint32_t ql = lane & 7;
                const uint32_t r = A.surv[p];
                const uint32_t kl = A.klen[r], vl = A.vlen[r], shared = A.shr[p], hs = A.rank[p], doff = A.R[p];
                const unsigned long long otr = A.trailer[r];
                const uint32_t kd = kl - shared;
                uint32_t l1, l2, l3;
                unsigned long long hv = varint_pack(shared, l1);
                const unsigned long long hv2 = varint_pack(kd + 8, l2), hv3 = varint_pack(vl, l3);
                uint32_t h = l1 + l2 + l3;
                bool packed = false;
                
                if (ql == 0) P.out_rec_off[S.base_recs + p] = doff - S.ob_off[A.blkid[p]]; // entry offset inside its block
                uint8_t *dst = out + doff;
                if (h > 8) { // lengths this large do not fit the packed register: one lane writes the head serially
                    if (ql == 0) {
                        uint8_t *d = dst;
                        d += put_varint32(d, shared); d += put_varint32(d, kd + 8); d += put_varint32(d, vl);
                        const uint8_t *ks = A.arena + (size_t)r * KS + shared;
                        for (uint32_t i = 0; i < kd; i++) d[i] = ks[i];
                        d += kd;
                        for (uint32_t x = 0; x < 8; x++) d[x] = (uint8_t)(otr >> (8 * x));
                    }
                    continue;
                }
                if (!packed) hv |= (hv2 << (8 * l1)) | (hv3 << (8 * (l1 + l2)));
                const uint32_t lead4 = doff & 3; // the tile's output base is 16-byte aligned
                const uint32_t nwords = (lead4 + hs + 3) >> 2;
                const int32_t kbase = (int32_t)(r * KS + shared) - (int32_t)h; // arena byte offset of head byte 0, were the key delta to start at byte h
                for (uint32_t w = ql; w < nwords; w += 8) {
                    const int32_t i0 = (int32_t)(4 * w) - (int32_t)lead4; // head byte held by the low byte of this word
                    // bytes below index n inside this word: a byte mask
                    auto lt = [&](uint32_t n) -> uint32_t {
                        const int32_t d = (int32_t)n - i0;
                        return d >= 4 ? 0xffffffffu : (d <= 0 ? 0u : ((1u << (8 * d)) - 1u));
                    };
                    const uint32_t m_h = lt(h), m_k = lt(h + kd), m_0 = lt(0), m_e = lt(hs);
                    uint32_t word = 0;
                    if (m_h) word |= (i0 >= 0 ? (uint32_t)(hv >> (8 * i0)) : (uint32_t)(hv << (8 * -i0))) & m_h;
                    if (m_k & ~m_h) {
                        const int32_t ko = kbase + i0;
                        const uint32_t *kw = (const uint32_t *)A.arena + (ko >> 2);
                        word |= __funnelshift_r(kw[0], kw[1], (uint32_t)(ko & 3) * 8) & m_k & ~m_h;
                    }
                    if (~m_k) {
                        const int32_t j0 = i0 - (int32_t)(h + kd);
                        const uint32_t tw = j0 >= 0 ? (j0 < 8 ? (uint32_t)(otr >> (8 * j0)) : 0u) : (j0 > -4 ? (uint32_t)(otr << (8 * -j0)) : 0u);
                        word |= tw & ~m_k;
                    }
                    const uint32_t vm = m_e & ~m_0; // bytes of this word that belong to the head
                    uint8_t *wp = dst + i0;
                    if (vm == 0xffffffffu) *reinterpret_cast<uint32_t *>(wp) = word;
                    else {
#pragma unroll
                        for (uint32_t x = 0; x < 4; x++) if ((vm >> (8 * x)) & 1u) wp[x] = (uint8_t)(word >> (8 * x));
                    }
                }
            }
            // restart arrays, padding, index entries: one warp per output block
            for (uint32_t b = warp; b < nob; b += NW) {
                const uint32_t c0 = S.cut[b], c1 = S.cut[b + 1], cnt = c1 - c0;
                const uint32_t nrest = (cnt + RI - 1) / RI;
                const uint32_t ent = A.E[c1] - A.E[c0];
                uint8_t *bp = out + S.ob_off[b];
                uint8_t *rp = bp + ent;
                for (uint32_t i = lane; i <= nrest; i += 32) {
                    uint32_t v = i < nrest ? A.E[c0 + i * RI] - A.E[c0] : nrest;
                    rp[4 * i] = (uint8_t)v; rp[4 * i + 1] = (uint8_t)(v >> 8); rp[4 * i + 2] = (uint8_t)(v >> 16); rp[4 * i + 3] = (uint8_t)(v >> 24);
                }
                for (uint32_t x = S.ob_size[b] + lane; x < S.ob_off[b + 1] - S.ob_off[b]; x += 32) bp[x] = 0;
                const uint32_t lr = A.surv[c1 - 1], lk = A.klen[lr];
                uint8_t *kd = P.out_ikeys + S.base_keyb + S.ob_keyoff[b];
                const uint8_t *ksrc = A.arena + (size_t)lr * KS;
                for (uint32_t x = lane; x < lk; x += 32) kd[x] = ksrc[x];
                if (lane == 0) {
                    const uint32_t g = S.base_blocks + b;
                    P.out_blk_off[g] = S.base_bytes + S.ob_off[b];
                    P.out_blk_size[g] = S.ob_size[b];
                    P.out_blk_rec[g] = S.base_recs + c0;
                    P.out_ikey_off[g] = S.base_keyb + S.ob_keyoff[b];
                    atomicMax(&S.max_blk_size, S.ob_size[b]);
                    atomicMax(&S.max_blk_rec, cnt);
                }
            }
            if (tid == 0) S.stat[ST_OUT_REC] = m;
        }
        if (last && tid == 0) { // sentinels of the new run's index
            uint32_t g = S.base_blocks + S.n_ob;
            if (g <= P.out_blk_cap) {
                P.out_blk_off[g] = S.base_bytes + S.tile_bytes;
                P.out_blk_rec[g] = S.base_recs + S.n_surv;
                P.out_ikey_off[g] = S.base_keyb + S.tile_keyb;
            }
        }
        __syncthreads();
        PT(10);
        if (tid < 16 && S.stat[tid]) {
            unsigned long long *g = &P.stats->in_records;
            static_assert(ST_OUT_VAL == 12, "stat layout");
            // MergeStats starts with in_records,in_bytes,out_records,out_bytes,dropped_shadowed,dropped_tombstone,
            // dropped_expired,dropped_user,dropped_stale,ttl_rewritten,out_tomb,out_raw_key,out_raw_val
            atomicAdd(g + tid, (unsigned long long)S.stat[tid]);
        }
        if (tid == 0) {
            if (S.error) {
                atomicMax(&P.stats->error, S.error);
                atomicMin(&P.stats->error_tile, tile);
            } else if (S.n_surv) {
                atomicMax(&P.stats->max_ukey, S.max_ukey);
                atomicMax(&P.stats->max_vlen, S.max_vlen);
                atomicMax(&P.stats->max_blk_size, S.max_blk_size);
                atomicMax(&P.stats->max_blk_rec, S.max_blk_rec);
                atomicMin(&P.stats->min_seq, S.min_seq);
                atomicMax(&P.stats->max_seq, S.max_seq);
            }
        }
        __syncthreads();
        PT(11);
    }
}

static uint64_t *g_crc_dev[16] = {nullptr};

} // namespace pgs

using namespace pgs;

namespace pgs {
const uint64_t *crc64_table();
}

extern "C" int32_t pgs_compact(pgs_partition *ph, const uint64_t *run_ids, uint32_t k, int32_t out_level,
                               int32_t bottommost, const pgs_filter_params *fp, uint32_t now,
                               pgs_compact_result *out)
{
    return pgs_compact_ex(ph, run_ids, k, out_level, bottommost, fp, now, 0, out);
}

extern "C" int32_t pgs_compact_ex(pgs_partition *ph, const uint64_t *run_ids, uint32_t k, int32_t out_level,
                                  int32_t bottommost, const pgs_filter_params *fp, uint32_t now, uint32_t flags,
                                  pgs_compact_result *out)
{
    if (!ph || !run_ids || k == 0 || out_level < 0) return PGS_INVALID_ARGUMENT;
    Partition &part = ph->p;
    Engine *e = part.eng;
    pgs_compact_result res{};
    std::vector<std::shared_ptr<Run>> in;
    {
        std::lock_guard<std::mutex> g(part.mu);
        for (uint32_t i = 0; i < k; i++) {
            auto r = part.find(run_ids[i]);
            if (!r) { set_error("compact: unknown run %llu", (unsigned long long)run_ids[i]); return PGS_NOT_FOUND; }
            for (auto &x : in) if (x == r) return PGS_INVALID_ARGUMENT;
            in.push_back(r);
        }
        if (bottommost < 0) { // true iff every run outside the input set is newer than every input
            size_t first_in = part.runs.size();
            for (size_t i = 0; i < part.runs.size(); i++)
                if (std::find(in.begin(), in.end(), part.runs[i]) != in.end()) { first_in = i; break; }
            bottommost = 1;
            for (size_t i = first_in; i < part.runs.size(); i++)
                if (std::find(in.begin(), in.end(), part.runs[i]) == in.end()) bottommost = 0;
        }
    }
    if (k > kMaxRuns) { set_error("compact: %u runs > %u per merge", k, kMaxRuns); return PGS_NOT_SUPPORTED; }
    PGS_CUDA(cudaSetDevice(e->device));
  for (int rigorous = 0; rigorous < 2; rigorous++) {
    cudaStream_t st = e->stream;

    MergeParams P{};
    P.k = k;
    uint32_t max_ukey = 0, max_blk = 0, max_blk_rec = 0;
    uint64_t total_blocks = 0, n_rec = 0, raw_key = 0, raw_val = 0, in_block_bytes = 0;
    for (uint32_t i = 0; i < k; i++) {
        P.runs[i] = in[i]->dev();
        const pgs_run_info &fi = in[i]->info;
        max_ukey = std::max(max_ukey, fi.max_ukey_len);
        max_blk = std::max(max_blk, fi.max_block_size);
        max_blk_rec = std::max(max_blk_rec, fi.max_block_records);
        total_blocks += fi.n_blocks;
        n_rec += fi.n_records;
        raw_key += fi.raw_key_bytes;
        raw_val += fi.raw_value_bytes;
        in_block_bytes += fi.data_bytes;
        if (fi.n_blocks >= (1u << 28)) return PGS_NOT_SUPPORTED;
    }
    if (max_ukey > kMaxUkeyLen) { set_error("compact: user key of %u bytes > %u", max_ukey, kMaxUkeyLen); return PGS_NOT_SUPPORTED; }
    const uint32_t KS = std::max(8u, (max_ukey + 7) & ~7u);
    P.KS = KS;
    P.rec_cost = KS + kRecExtra + 2 * (k - 1);
    P.warp_scratch = (KS + 48 + 15) & ~15u;
    P.total_blocks = (uint32_t)total_blocks;
    P.use_tma = (e->cfg.flags & PGS_ENGINE_NO_TMA) ? 0 : 1;
    {
        const char *ev = getenv("PGS_EARLY_TMA"); // diagnostics: 0 turns the early block load off
        P.early_tma = (ev && ev[0] == '0') ? 0 : 1;
    }
    P.block_size = e->cfg.block_size;
    P.restart_interval = e->cfg.restart_interval;
    P.bottommost = bottommost ? 1 : 0;
    P.now = now;
    P.data_version = part.data_version;
    std::vector<uint8_t> ops_host;
    if (fp) {
        P.enabled = fp->enabled;
        P.validate_hash = fp->validate_hash;
        P.default_ttl = fp->default_ttl;
        P.pidx = fp->pidx;
        P.partition_version = fp->partition_version;
        if (fp->ops && fp->ops_len >= 4) {
            memcpy(&P.n_ops, fp->ops, 4);
            ops_host.assign(fp->ops, fp->ops + fp->ops_len);
        }
    }

    // ctas_per_sm: 2 -> two 512-thread CTAs per SM (default); 1 -> one 1024-thread CTA per SM with tiles twice as large
    const bool big = e->cfg.ctas_per_sm == 1;
    auto kern = big ? k_merge<1024> : k_merge<512>;
    const uint32_t nthreads = big ? 1024 : 512;
    cudaFuncAttributes attr;
    PGS_CUDA(cudaFuncGetAttributes(&attr, kern));
    uint32_t ctas = e->cfg.ctas_per_sm;
    const uint32_t fixed_dyn = 2 * (KS + 8);
    const uint64_t maxw = (((uint64_t)max_blk + 15) & ~15ull) + 16 + (uint64_t)max_blk_rec * P.rec_cost;
    uint32_t dyn = 0;
    uint64_t T = 0;
    for (;; ctas--) {
        if (ctas == 0) { set_error("compact: blocks/records too large for shared memory (k=%u, max block %u B)", k, max_blk); return PGS_NOT_SUPPORTED; }
        uint64_t per_cta = (228ull * 1024) / ctas - 1024; // SM shared memory split, 1 KB reserved per CTA
        per_cta = std::min<uint64_t>(per_cta, (uint64_t)e->max_smem_optin);
        if (per_cta < attr.sharedSizeBytes + fixed_dyn + 1024) continue;
        dyn = (uint32_t)((per_cta - attr.sharedSizeBytes) & ~127ull);
        uint64_t pool = dyn - fixed_dyn;
        // a tile holds < T of whole blocks, the group of blocks that end on the boundary key
        // (one per run at worst) and one partial block per run.  Groups larger than two blocks are
        // rare, so the first attempt budgets k+2 blocks of slack; the kernel verifies every tile and
        // the host retries with the rigorous 2k bound if one did not fit.
        uint64_t overhead = (rigorous ? 2ull * k : (uint64_t)k + 2) * maxw + 1024;
        if (pool > overhead + maxw) { T = pool - overhead; P.pool_bytes = (uint32_t)pool; break; }
    }
    uint64_t W_total = 0;
    for (uint32_t i = 0; i < k; i++) W_total += in[i]->info.data_bytes + in[i]->info.n_records * P.rec_cost;
    uint64_t Q = std::max<uint64_t>(1, (W_total + T - 1) / T);
    if (Q > 0x7FFFFFF0ull) return PGS_NOT_SUPPORTED;
    P.Q = (uint32_t)Q;
    P.tile_weight = T;

    // output capacity bounds
    const uint64_t RIv = P.restart_interval;
    uint64_t raw_total = raw_key + raw_val + 11 * n_rec;
    uint64_t blk_cap = raw_total / P.block_size + Q + 2;
    uint64_t out_cap = raw_key + raw_val + 23 * n_rec + 19 * blk_cap + 4 * (n_rec / RIv + blk_cap) + 256;
    out_cap = (out_cap + 255) & ~255ull;
    uint64_t ikey_cap = std::min<uint64_t>(raw_key, blk_cap * (uint64_t)std::max(1u, max_ukey)) + 16;
    if (blk_cap > 0xFFFFFFF0ull || ikey_cap > 0xFFFFFFF0ull) return PGS_NOT_SUPPORTED;

    auto outr = std::make_shared<Run>();
    outr->level = out_level;
    outr->data_cap = out_cap + 256;
    uint32_t *d_split_pos = nullptr, *d_split_ref = nullptr, *d_ticket = nullptr;
    TileAgg *d_agg = nullptr;
    MergeStats *d_stats = nullptr;
    uint8_t *d_ops = nullptr;
    auto cleanup = [&]() {
        cudaFreeAsync(d_split_pos, st); cudaFreeAsync(d_split_ref, st); cudaFreeAsync(d_ticket, st);
        cudaFreeAsync(d_agg, st); cudaFreeAsync(d_stats, st); cudaFreeAsync(d_ops, st);
        d_split_pos = d_split_ref = d_ticket = nullptr; d_agg = nullptr; d_stats = nullptr; d_ops = nullptr;
    };
    outr->pool_stream = st;
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr); } } while (0)
    outr->eng = e;
    { uint64_t cap = 0; outr->d_data = outr->data_cap >= (64ull << 20) ? e->take_data(outr->data_cap, &cap) : nullptr; if (outr->d_data) outr->data_cap = cap; }
    if (!outr->d_data) CK(cudaMallocAsync(&outr->d_data, outr->data_cap, st));
    CK(cudaMallocAsync(&outr->d_blk_off, sizeof(uint64_t) * (blk_cap + 1), st));
    CK(cudaMallocAsync(&outr->d_blk_size, sizeof(uint32_t) * (blk_cap + 1), st));
    CK(cudaMallocAsync(&outr->d_blk_rec, sizeof(uint32_t) * (blk_cap + 1), st));
    CK(cudaMallocAsync(&outr->d_ikey_off, sizeof(uint32_t) * (blk_cap + 1), st));
    CK(cudaMallocAsync(&outr->d_ikeys, ikey_cap, st));
    CK(cudaMallocAsync(&outr->d_rec_off, sizeof(uint32_t) * (n_rec + 1), st));
    CK(cudaMallocAsync(&d_split_pos, sizeof(uint32_t) * (Q + 1) * k, st));
    CK(cudaMallocAsync(&d_split_ref, sizeof(uint32_t) * (Q + 1), st));
    CK(cudaMallocAsync(&d_ticket, 256, st));
    CK(cudaMallocAsync(&d_agg, sizeof(TileAgg) * Q, st));
    CK(cudaMallocAsync(&d_stats, sizeof(MergeStats), st));
    CK(cudaMemsetAsync(d_split_pos, 0xFF, sizeof(uint32_t) * (Q + 1) * k, st));
    CK(cudaMemsetAsync(d_split_ref, 0xFF, sizeof(uint32_t) * (Q + 1), st));
    CK(cudaMemsetAsync(d_ticket, 0, 256, st));
    CK(cudaMemsetAsync(d_agg, 0, sizeof(TileAgg) * Q, st));
    MergeStats hs{};
    hs.min_seq = ~0ull;
    hs.error_tile = 0xFFFFFFFFu;
    CK(cudaMemcpyAsync(d_stats, &hs, sizeof hs, cudaMemcpyHostToDevice, st));
    if (!ops_host.empty()) {
        CK(cudaMallocAsync(&d_ops, ops_host.size(), st));
        CK(cudaMemcpyAsync(d_ops, ops_host.data(), ops_host.size(), cudaMemcpyHostToDevice, st));
        P.ops = d_ops;
    }
    if (P.validate_hash) {
        int dev = e->device;
        if (!g_crc_dev[dev & 15]) {
            uint64_t *t = nullptr;
            CK(cudaMalloc(&t, 256 * 8));
            CK(cudaMemcpyAsync(t, crc64_table(), 256 * 8, cudaMemcpyHostToDevice, st));
            g_crc_dev[dev & 15] = t;
        }
        P.crc_table = (const unsigned long long *)g_crc_dev[dev & 15];
    }
    P.split_pos = d_split_pos;
    P.split_ref = d_split_ref;
    P.ticket = d_ticket;
    const char *pt_env = getenv("PGS_PHASE_TIMING"); // diagnostics: per-phase cycle totals of k_merge on stderr
    const bool phase_timing = pt_env && pt_env[0] == '1';
    P.phase_cycles = phase_timing ? (unsigned long long *)(d_ticket + 16) : nullptr;
    P.agg = d_agg;
    P.out_data = outr->d_data;
    P.out_cap = out_cap;
    P.out_blk_off = (unsigned long long *)outr->d_blk_off;
    P.out_blk_size = outr->d_blk_size;
    P.out_blk_rec = outr->d_blk_rec;
    P.out_ikey_off = outr->d_ikey_off;
    P.out_ikeys = outr->d_ikeys;
    P.out_rec_off = outr->d_rec_off;
    P.out_blk_cap = (uint32_t)blk_cap;
    P.out_ikey_cap = (uint32_t)ikey_cap;
    P.stats = d_stats;

    cudaEvent_t ev0, ev1, ev2;
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaEventCreate(&ev2));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CK(cudaEventRecord(ev0, st));
    k_plan<<<(uint32_t)((total_blocks + 255) / 256), 256, 0, st>>>(P);
    CK(cudaEventRecord(ev1, st));
    uint32_t grid = (uint32_t)std::min<uint64_t>(Q, (uint64_t)ctas * e->sm_count);
    kern<<<grid, nthreads, dyn, st>>>(P);
    CK(cudaEventRecord(ev2, st));
    e->launches += 2;
    CK(cudaMemcpyAsync(&hs, d_stats, sizeof hs, cudaMemcpyDeviceToHost, st));
    TileAgg fin{};
    CK(cudaMemcpyAsync(&fin, d_agg + (Q - 1), sizeof fin, cudaMemcpyDeviceToHost, st));
    unsigned long long h_phase[16] = {0};
    if (phase_timing) CK(cudaMemcpyAsync(h_phase, d_ticket + 16, sizeof h_phase, cudaMemcpyDeviceToHost, st));
    cudaError_t se = cudaStreamSynchronize(st);
    if (se != cudaSuccess) { cleanup(); return cuda_fail(se, "compaction kernels"); }
    if (phase_timing) {
        static const char *names[14] = {"setup", "decode1", "decode2", "window", "rank3", "filter", "layout", "a-wait", "lookback", "write_b", "write_c", "flush", "rank1", "rank2"};
        unsigned long long tot = 0;
        for (int i = 0; i < 14; i++) tot += h_phase[i];
        fprintf(stderr, "[k_merge phases] tiles=%u", P.Q);
        for (int i = 0; i < 14; i++) fprintf(stderr, " %s=%.1f%%", names[i], tot ? 100.0 * (double)h_phase[i] / (double)tot : 0.0);
        fprintf(stderr, " cycles/tile=%.0f\n", P.Q ? (double)tot / P.Q : 0.0);
    }
    float ms_total = 0, ms_merge = 0;
    cudaEventElapsedTime(&ms_total, ev0, ev2);
    cudaEventElapsedTime(&ms_merge, ev1, ev2);
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    cleanup();
#undef CK
    if (hs.error) {
        set_error("compaction kernel failed with status %u at tile %u of %u (T=%llu, dyn smem %u)", hs.error, hs.error_tile,
                  P.Q, (unsigned long long)T, dyn);
        if (hs.error == PGS_ABORTED && !rigorous) continue; // a tile did not fit: retry with the rigorous bound
        return (int32_t)hs.error;
    }
    res.in_records = hs.in_records; res.out_records = hs.out_records;
    res.in_bytes = hs.in_bytes; res.out_bytes = hs.out_bytes;
    res.in_block_bytes = in_block_bytes; res.out_block_bytes = fin.bytes;
    res.dropped_shadowed = hs.dropped_shadowed; res.dropped_tombstone = hs.dropped_tombstone;
    res.dropped_expired = hs.dropped_expired; res.dropped_user = hs.dropped_user; res.dropped_stale = hs.dropped_stale;
    res.ttl_rewritten = hs.ttl_rewritten;
    res.n_tiles = P.Q; res.n_launches = 2;
    res.device_ms = ms_total; res.merge_kernel_ms = ms_merge;

    outr->info.level = out_level;
    outr->info.n_blocks = fin.blocks;
    outr->info.n_records = fin.recs;
    outr->info.n_tombstones = hs.out_tomb;
    outr->info.data_bytes = fin.bytes;
    outr->info.raw_key_bytes = hs.out_raw_key;
    outr->info.raw_value_bytes = hs.out_raw_val;
    outr->info.max_ukey_len = hs.max_ukey;
    outr->info.max_value_len = hs.max_vlen;
    outr->info.max_block_size = hs.max_blk_size;
    outr->info.max_block_records = hs.max_blk_rec;
    outr->info.smallest_seq = hs.min_seq;
    outr->info.largest_seq = hs.max_seq;
    {
        std::lock_guard<std::mutex> g(part.mu);
        if (!(flags & PGS_COMPACT_KEEP_INPUTS))
            for (auto &r : in) part.runs.erase(std::find(part.runs.begin(), part.runs.end(), r));
        if (fin.blocks > 0 && !(flags & PGS_COMPACT_DISCARD_OUTPUT)) {
            outr->id = e->next_run_id++;
            outr->info.run_id = outr->id;
            part.insert(outr);
            res.new_run_id = outr->id;
        }
    }
    if (out) *out = res;
    return PGS_OK;
  }
    return PGS_ABORTED;
}
