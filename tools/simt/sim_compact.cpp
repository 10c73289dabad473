// sim_compact.cpp — runs the compaction kernels (incubator_pegasus_b200/csrc/compact_kernels.cuh, the same source nvcc compiles
// for sm_100a) inside the host SIMT interpreter of simt.h.  Test / development tool: lets the CPU test-suite execute the kernels'
// logic against the oracle without a GPU.  NOT part of the product library and never a fallback for it.
#include "simt.h"

#define PGS_SIM 1
#include "../../incubator_pegasus_b200/csrc/compact_kernels.cuh"
#include "../../incubator_pegasus_b200/csrc/read_kernels.cuh"

#include <string>
#include <vector>

using namespace pgs;

namespace {

struct HostRun {
    std::vector<uint8_t> data;
    std::vector<uint64_t> blk_off;
    std::vector<uint32_t> blk_size, blk_rec, ikey_off, rec_off;
    std::vector<uint8_t> ikeys;
    std::vector<uint32_t> bloom;
    pgs_run_info info{};
    RunDev dev() const
    {
        return RunDev{data.data(), blk_off.data(), blk_size.data(), blk_rec.data(), ikey_off.data(), ikeys.data(), rec_off.data(),
                      bloom.data(), (uint32_t)(bloom.size() / 16), (uint32_t)blk_size.size(), info.max_ukey_len, 0};
    }
};

uint32_t varint(const uint8_t *p, uint32_t &v)
{
    v = 0;
    for (uint32_t i = 0; i < 5; i++) {
        v |= (uint32_t)(p[i] & 127) << (7 * i);
        if (!(p[i] & 128)) return i + 1;
    }
    return 0;
}

// what k_index_walk builds on the device (engine.cu): record counts, last user keys, entry offsets, run info
bool build_index(HostRun &r)
{
    const uint32_t nb = (uint32_t)r.blk_size.size();
    r.blk_rec.assign(nb + 1, 0);
    r.ikey_off.assign(nb + 1, 0);
    r.info.n_blocks = nb;
    r.info.smallest_seq = ~0ull;
    std::string key;
    std::vector<std::string> all_keys;
    for (uint32_t b = 0; b < nb; b++) {
        const uint8_t *base = r.data.data() + r.blk_off[b];
        const uint32_t size = r.blk_size[b];
        if (size < 8) return false;
        uint32_t nr;
        memcpy(&nr, base + size - 4, 4);
        const uint32_t limit = size - 4 - 4 * nr;
        uint32_t p = 0, n = 0;
        key.clear();
        while (p < limit) {
            uint32_t sh, ns, vl, h = 0, c;
            c = varint(base + p, sh); h += c;
            c = varint(base + p + h, ns); h += c;
            c = varint(base + p + h, vl); h += c;
            if (!c || sh > key.size()) return false;
            key.resize(sh);
            key.append((const char *)base + p + h, ns);
            r.rec_off.push_back(p);
            all_keys.emplace_back(key.data(), key.size() - 8);
            unsigned long long tr;
            memcpy(&tr, key.data() + key.size() - 8, 8);
            r.info.n_tombstones += (uint8_t)tr == PGS_TYPE_DELETION;
            r.info.smallest_seq = std::min<uint64_t>(r.info.smallest_seq, tr >> 8);
            r.info.largest_seq = std::max<uint64_t>(r.info.largest_seq, tr >> 8);
            r.info.raw_key_bytes += key.size() - 8;
            r.info.raw_value_bytes += vl;
            r.info.max_ukey_len = std::max<uint32_t>(r.info.max_ukey_len, (uint32_t)key.size() - 8);
            r.info.max_value_len = std::max(r.info.max_value_len, vl);
            p += h + ns + vl;
            n++;
        }
        r.blk_rec[b + 1] = r.blk_rec[b] + n;
        r.ikeys.insert(r.ikeys.end(), key.begin(), key.end() - 8);
        r.ikey_off[b + 1] = (uint32_t)r.ikeys.size();
        r.info.max_block_size = std::max(r.info.max_block_size, size);
        r.info.max_block_records = std::max(r.info.max_block_records, n);
        r.info.n_records += n;
    }
    r.ikeys.resize(r.ikeys.size() + 16); // the product's slack (engine.cu)
    r.rec_off.push_back(0);
    // the Bloom filter k_index_walk builds at upload: whole keys + hash-key prefixes
    std::vector<std::string> pres;
    for (auto &k : all_keys) {
        const uint32_t pl = hashkey_prefix_len((const uint8_t *)k.data(), (uint32_t)k.size());
        if (pl && (pres.empty() || pres.back() != k.substr(0, pl))) pres.push_back(k.substr(0, pl));
    }
    const uint32_t lines = bloom_lines_for(all_keys.size() + pres.size());
    r.bloom.assign((size_t)lines * 16, 0);
    for (auto *v : {&all_keys, &pres})
        for (auto &k : *v) {
            const unsigned long long h = bloom_hash_bytes((const uint8_t *)k.data(), (uint32_t)k.size());
            for (uint32_t s = 0; s < 6; s++) bloom_add_bit(r.bloom.data(), lines, h, s);
        }
    return true;
}

uint64_t crc_tab[256];
void make_crc()
{
    // reflected CRC-64 with the polynomial of src/utils/crc.cpp:289-295 (generated, as host_util.cpp does)
    const uint64_t poly = 0x9a6c9329ac4bc9b5ull;
    for (uint32_t i = 0; i < 256; i++) {
        uint64_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
        crc_tab[i] = c;
    }
}

struct Result {
    std::vector<uint8_t> data;
    std::vector<uint64_t> blk_off;
    std::vector<uint32_t> blk_size, blk_rec, ikey_off, rec_off;
    std::vector<uint8_t> ikeys;
    std::vector<uint32_t> bloom;
    uint32_t bloom_lines = 0;
    MergeStats st{};
    uint32_t Q = 0;
} g_res;

} // namespace

extern "C" {

// returns a status code; the merged run stays in a static buffer until the next call (sim_result_* read it)
int32_t sim_compact(uint32_t k, const uint8_t **data, const uint64_t *data_bytes, const uint64_t **blk_off, const uint32_t **blk_size,
                    const uint32_t *n_blocks, uint32_t block_size, uint32_t restart_interval, int32_t bottommost,
                    const pgs_filter_params *fp, uint32_t now, uint32_t group_lanes, uint64_t seg_weight, const uint64_t *crc_table)
{
    std::vector<HostRun> runs(k);
    MergeParams P{};
    P.k = k;
    CompactTotals T{};
    for (uint32_t i = 0; i < k; i++) {
        HostRun &r = runs[i];
        r.data.assign(data[i], data[i] + data_bytes[i]);
        r.data.resize(r.data.size() + 256, 0); // the product's slack after a run's blocks (engine.cu, compact.cu)
        r.blk_off.assign(blk_off[i], blk_off[i] + n_blocks[i]);
        r.blk_size.assign(blk_size[i], blk_size[i] + n_blocks[i]);
        uint64_t end = n_blocks[i] ? r.blk_off.back() + r.blk_size.back() : 0;
        end = (end + 15) & ~15ull;
        r.blk_off.push_back(end);
        r.info.data_bytes = end;
        if (!build_index(r)) return PGS_CORRUPTION;
        P.runs[i] = r.dev();
        T.max_ukey = std::max(T.max_ukey, r.info.max_ukey_len);
        T.max_blk = std::max(T.max_blk, r.info.max_block_size);
        T.max_blk_rec = std::max(T.max_blk_rec, r.info.max_block_records);
        T.total_blocks += r.info.n_blocks;
        T.n_rec += r.info.n_records;
        T.raw_key += r.info.raw_key_bytes;
        T.raw_val += r.info.raw_value_bytes;
        T.in_block_bytes += r.info.data_bytes;
    }
    P.block_size = block_size;
    P.restart_interval = restart_interval;
    P.bottommost = bottommost ? 1 : 0;
    P.now = now;
    P.data_version = 1;
    std::vector<uint8_t> ops;
    if (fp) {
        P.enabled = fp->enabled; P.validate_hash = fp->validate_hash; P.default_ttl = fp->default_ttl;
        P.pidx = fp->pidx; P.partition_version = fp->partition_version;
        if (fp->ops && fp->ops_len >= 4) { memcpy(&P.n_ops, fp->ops, 4); ops.assign(fp->ops, fp->ops + fp->ops_len); ops.resize(ops.size() + 16); P.ops = ops.data(); }
    }
    CompactGeometry geo{};
    if (group_lanes && group_lanes != 1 && group_lanes != 2 && group_lanes != 4 && group_lanes != 8 && group_lanes != 16) return PGS_INVALID_ARGUMENT;
    if (!compact_geometry(P, T, 227 * 1024, geo, group_lanes)) return PGS_NOT_SUPPORTED;
    if (seg_weight) { // smaller segments: more boundaries per record in a small test
        P.tile_weight = seg_weight;
        const uint64_t W = T.in_block_bytes + T.n_rec * P.rec_cost;
        P.Q = (uint32_t)std::max<uint64_t>(1, (W + seg_weight - 1) / seg_weight);
        MergeParams P2 = P;
        // redo the bounds that depend on Q
        const uint64_t Q = P.Q;
        const uint64_t raw_total = T.raw_key + T.raw_val + 23 * T.n_rec;
        geo.blk_cap = 2 * (raw_total / P.block_size) + Q + 2;
        geo.out_cap = (raw_total + 19 * geo.blk_cap + 4 * (T.n_rec / restart_interval + geo.blk_cap) + 256 + 255) & ~255ull;
        geo.ikey_cap = std::min<uint64_t>(T.raw_key, geo.blk_cap * (uint64_t)std::max(1u, T.max_ukey)) + 16;
        P.out_cap = geo.out_cap; P.out_blk_cap = (uint32_t)geo.blk_cap; P.out_ikey_cap = (uint32_t)geo.ikey_cap;
        const uint64_t Nb = T.n_rec + Q * k * (uint64_t)T.max_blk_rec, Bb = T.in_block_bytes + Q * k * ((uint64_t)T.max_blk + 16);
        const uint64_t per_head = 15 + P.KS + 8 + 4;
        P.desc_cap = Nb + 1;
        P.head_cap = Nb * per_head + (2 * (Bb + Nb * per_head) / P.block_size + 2 * Q + 2) * (uint64_t)(P.KS + 8) + 64 * Q + 64;
        (void)P2;
    }
    const uint64_t Q = P.Q;
    std::vector<uint32_t> split_pos((Q + 1) * k, 0xFFFFFFFFu), split_ref(Q + 1, 0xFFFFFFFFu), ticket(64, 0);
    std::vector<SegLayout> seg(Q);
    std::vector<SegAgg> agg(Q);
    std::vector<SegBase> base(Q);
    std::vector<Desc> desc(P.desc_cap);
    std::vector<uint8_t> heads(P.head_cap + 64, 0xEE);
    MergeStats st{};
    st.error_seg = 0xFFFFFFFFu;
    Result &R = g_res;
    R = Result{};
    R.data.assign(geo.out_cap + 256, 0xDD);
    R.blk_off.assign(geo.blk_cap + 1, 0);
    R.blk_size.assign(geo.blk_cap + 1, 0);
    R.blk_rec.assign(geo.blk_cap + 1, 0);
    R.ikey_off.assign(geo.blk_cap + 1, 0);
    R.ikeys.assign(geo.ikey_cap, 0);
    R.rec_off.assign(T.n_rec + 1, 0);
    R.bloom_lines = bloom_lines_for(2 * T.n_rec);
    R.bloom.assign((size_t)R.bloom_lines * 16, 0);
    P.out_bloom = R.bloom.data();
    P.out_bloom_lines = R.bloom_lines;
    P.split_pos = split_pos.data(); P.split_ref = split_ref.data(); P.ticket = ticket.data();
    P.seg = seg.data(); P.agg = agg.data(); P.base = base.data(); P.desc = desc.data(); P.heads = heads.data();
    P.out_data = R.data.data(); P.out_blk_off = (unsigned long long *)R.blk_off.data(); P.out_blk_size = R.blk_size.data();
    P.out_blk_rec = R.blk_rec.data(); P.out_ikey_off = R.ikey_off.data(); P.out_ikeys = R.ikeys.data(); P.out_rec_off = R.rec_off.data();
    P.stats = &st;
    if (P.validate_hash) {
        if (crc_table) memcpy(crc_tab, crc_table, sizeof crc_tab); else make_crc();
        P.crc_table = (const unsigned long long *)crc_tab;
    }
    PGS_LAUNCH(k_plan, (T.total_blocks + 255) / 256, 256, 0, 0, P);
    PGS_LAUNCH(k_seg_bounds, (P.Q + 255) / 256, 256, 0, 0, P);
    PGS_LAUNCH(k_seg_layout, 1, 1024, 0, 0, P);
    if (!st.error) {
        if (geo.G == 1) PGS_LAUNCH(k_walk<1>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else if (geo.G == 2) PGS_LAUNCH(k_walk<2>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else if (geo.G == 4) PGS_LAUNCH(k_walk<4>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else if (geo.G == 8) PGS_LAUNCH(k_walk<8>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else PGS_LAUNCH(k_walk<16>, 2, kWalkThreads, geo.walk_dyn, 0, P);
    }
    if (getenv("PGS_SIM_DUMP")) {
        for (uint32_t q = 0; q < P.Q; q++) {
            const Desc *d = desc.data() + seg[q].desc_off;
            const uint8_t *h = heads.data() + seg[q].head_off;
            uint32_t hp = 0;
            for (uint32_t e = 0; e < agg[q].n_entries; e++) {
                uint32_t fl = d[e].loc >> 60, hl = (d[e].loc >> 44) & 0xffff;
                fprintf(stderr, "seg %u e %u fl %u hl %u vlen %u aux %u", q, e, fl, hl, d[e].vlen, d[e].aux);
                if (fl & 1) { fprintf(stderr, " prevkey ..%.*s", 6, h + hp + (d[e].aux > 6 ? d[e].aux - 6 : 0)); hp += d[e].aux; }
                if (fl & 2) hp += 4;
                fprintf(stderr, " head:");
                for (uint32_t i = 0; i < hl && i < 70; i++) fprintf(stderr, "%02x", h[hp + i]);
                fprintf(stderr, "\n");
                hp += hl;
            }
        }
    }
    if (!st.error) PGS_LAUNCH(k_seg_scan, 1, 1024, 0, 0, P);
    if (!st.error) PGS_LAUNCH(k_emit, 2, geo.emit_warps * 32, geo.emit_dyn, 0, P);
    R.st = st;
    R.Q = P.Q;
    if (st.error) { fprintf(stderr, "sim_compact: status %u at segment %u of %u\n", st.error, st.error_seg, P.Q); return (int32_t)st.error; }
    return PGS_OK;
}

void sim_result_sizes(uint64_t *data_bytes, uint32_t *n_blocks, uint32_t *n_segments)
{
    *data_bytes = g_res.st.tot_bytes;
    *n_blocks = (uint32_t)g_res.st.tot_blocks;
    *n_segments = g_res.Q;
}
void sim_result_copy(uint8_t *data, uint64_t *blk_off, uint32_t *blk_size, uint32_t *blk_rec, uint32_t *ikey_off, uint8_t *ikeys, uint32_t *rec_off)
{
    const uint32_t nb = (uint32_t)g_res.st.tot_blocks;
    memcpy(data, g_res.data.data(), g_res.st.tot_bytes);
    memcpy(blk_off, g_res.blk_off.data(), 8 * (size_t)(nb + 1));
    memcpy(blk_size, g_res.blk_size.data(), 4 * (size_t)nb);
    if (blk_rec) memcpy(blk_rec, g_res.blk_rec.data(), 4 * (size_t)(nb + 1));
    if (ikey_off) memcpy(ikey_off, g_res.ikey_off.data(), 4 * (size_t)(nb + 1));
    if (ikeys) memcpy(ikeys, g_res.ikeys.data(), g_res.st.tot_keyb);
    if (rec_off) memcpy(rec_off, g_res.rec_off.data(), 4 * (size_t)g_res.st.tot_recs);
}
// the counters of pgs_compact_result, in its order: in_records, out_records, in_bytes, out_bytes, dropped_shadowed,
// dropped_tombstone, dropped_expired, dropped_user, dropped_stale, ttl_rewritten, + run info: tombstones, raw key, raw value,
// max_ukey, max_vlen, max_blk_size, max_blk_rec, smallest_seq, largest_seq, index key bytes
void sim_result_stats(uint64_t *o)
{
    const MergeStats &s = g_res.st;
    const uint64_t v[20] = {s.cnt[EV_IN], s.cnt[EV_OUT], s.bytes[SB_IN], s.bytes[SB_OUT], s.cnt[EV_SHADOW], s.cnt[EV_TOMB], s.cnt[EV_EXPIRED],
                            s.cnt[EV_USER], s.cnt[EV_STALE], s.cnt[EV_TTL], s.cnt[EV_OUT_TOMB], s.bytes[SB_OUT_KEY], s.bytes[SB_OUT_VAL], s.mx[SM_UKEY],
                            s.mx[SM_VLEN], s.mx[SM_BLK_SIZE], s.mx[SM_BLK_REC], s.tot_recs ? (~s.mx[SM_MIN_SEQ_INV] & ((1ull << 56) - 1)) : ~0ull,
                            s.mx[SM_MAX_SEQ], s.tot_keyb};
    memcpy(o, v, sizeof v);
}

// 1 when the merged run's Bloom filter admits the byte string (a user key or a hash-key prefix)
int32_t sim_result_bloom_check(const uint8_t *key, uint32_t len)
{
    return bloom_may_contain(g_res.bloom.data(), g_res.bloom_lines, bloom_hash_bytes(key, len)) ? 1 : 0;
}

static bool load_runs(uint32_t k, const uint8_t **data, const uint64_t *data_bytes, const uint64_t **blk_off, const uint32_t **blk_size,
                      const uint32_t *n_blocks, std::vector<HostRun> &runs, ReadRuns &rr, uint32_t &max_ukey)
{
    runs.resize(k);
    rr.n = k;
    max_ukey = 0;
    for (uint32_t i = 0; i < k; i++) {
        HostRun &r = runs[i];
        r.data.assign(data[i], data[i] + data_bytes[i]);
        r.data.resize(r.data.size() + 256, 0); // the product's slack after a run's blocks (engine.cu, compact.cu)
        r.blk_off.assign(blk_off[i], blk_off[i] + n_blocks[i]);
        r.blk_size.assign(blk_size[i], blk_size[i] + n_blocks[i]);
        uint64_t end = n_blocks[i] ? r.blk_off.back() + r.blk_size.back() : 0;
        r.blk_off.push_back((end + 15) & ~15ull);
        if (!build_index(r)) return false;
        rr.runs[i] = r.dev();
        max_ukey = std::max(max_ukey, r.info.max_ukey_len);
    }
    return true;
}

// k_get over k runs (newest first); results / arena as pgs_get_batch.  stats[0] = arena bytes, [1] = blocks probed, [2] = runs skipped
int32_t sim_get(uint32_t k, const uint8_t **data, const uint64_t *data_bytes, const uint64_t **blk_off, const uint32_t **blk_size,
                const uint32_t *n_blocks, const uint8_t *keys, const uint32_t *key_off, uint32_t n, uint32_t now, uint8_t *arena,
                uint64_t arena_cap, pgs_get_result *results, uint64_t *stats, uint32_t use_bloom)
{
    std::vector<HostRun> runs;
    GetParams P{};
    uint32_t mk = 0;
    if (!load_runs(k, data, data_bytes, blk_off, blk_size, n_blocks, runs, P.rr, mk)) return PGS_CORRUPTION;
    if (!(use_bloom & 1)) for (uint32_t i = 0; i < k; i++) { P.rr.runs[i].bloom = nullptr; P.rr.runs[i].bloom_lines = 0; }
    std::vector<uint8_t> kcopy(keys, keys + key_off[n]);
    kcopy.resize(kcopy.size() + 16); // as lookup.cu allocates the key buffer
    unsigned long long cur[4] = {0, 0, 0, 0};
    uint32_t err[4] = {0, 0, 0, 0};
    P.keys = kcopy.data(); P.key_off = key_off; P.n = n; P.now = now; P.data_version = 1;
    P.results = results; P.arena = arena; P.arena_cap = arena_cap; P.arena_cursor = cur; P.error = err; P.ticket = err + 1;
    P.KS = std::max(8u, (mk + 3) & ~3u);
    P.KSW = (P.KS + 8) / 4 + 1;
    P.group_smem = (uint32_t)((sizeof(CurState) + 2 * P.KSW * 4 + 15) & ~(size_t)15);
    const uint32_t dyn = kMaxReadRuns * (uint32_t)sizeof(RunDev) + (kReadThreads / 8) * P.group_smem;
    // use_bloom & 2: the multi-partition shape of pgs_get_batch_multi -- slot 0 owns the runs [0, k/2), slot 1 the rest, slot 2
    // is an empty partition; key i belongs to slot i % 3
    std::vector<RunDev> packed(P.rr.runs, P.rr.runs + k);
    std::vector<uint32_t> begin = {0, k / 2, k, k}, kp(n);
    if (use_bloom & 2) {
        for (uint32_t i = 0; i < n; i++) kp[i] = i % 3;
        P.multi_runs = packed.data(); P.multi_begin = begin.data(); P.key_part = kp.data();
    }
    if (use_bloom & 2) PGS_LAUNCH((k_get<8, true>), 2, kReadThreads, dyn, 0, P);
    else PGS_LAUNCH((k_get<8, false>), 2, kReadThreads, dyn, 0, P);
    stats[0] = cur[0]; stats[1] = cur[1]; stats[2] = cur[2];
    return err[0] ? (int32_t)err[0] : PGS_OK;
}

// k_scan_fwd over k runs; outputs as the device side of scan_many (request i uses arena + i*arena_stride, kvs + i*kv_stride)
int32_t sim_scan(uint32_t k, const uint8_t **data, const uint64_t *data_bytes, const uint64_t **blk_off, const uint32_t **blk_size,
                 const uint32_t *n_blocks, const pgs_scan_request *reqs, uint32_t n, uint32_t now, uint64_t arena_stride, uint32_t kv_stride,
                 uint8_t *arena, pgs_kv *kvs, uint8_t *resume, uint32_t resume_stride, pgs_scan_result *results, uint32_t lanes)
{
    std::vector<HostRun> runs;
    ScanParams P{};
    uint32_t mk = 0;
    if (!load_runs(k, data, data_bytes, blk_off, blk_size, n_blocks, runs, P.rr, mk)) return PGS_CORRUPTION;
    std::vector<ScanReqDev> dev(n);
    std::string blob;
    bool need_crc = false;
    for (uint32_t i = 0; i < n; i++) {
        const pgs_scan_request &q = reqs[i];
        ScanReqDev &d = dev[i];
        memset(&d, 0, sizeof d);
        auto put = [&](const pgs_blob &b, uint32_t &off, uint32_t &len) { off = (uint32_t)blob.size(); len = b.len; if (b.len) blob.append((const char *)b.data, b.len); };
        put(q.start, d.start_off, d.start_len); put(q.stop, d.stop_off, d.stop_len);
        put(q.hash_filter, d.hf_off, d.hf_len); put(q.sort_filter, d.sf_off, d.sf_len);
        d.start_inclusive = q.start_inclusive; d.stop_inclusive = q.stop_inclusive; d.reverse = q.reverse;
        d.no_value = q.no_value; d.key_mode = q.key_mode; d.return_expire_ts = q.return_expire_ts;
        d.count_only = q.count_only; d.validate_hash = q.validate_hash; d.prefix_same_as_start = q.prefix_same_as_start;
        d.has_upper = q.reserved[0];
        d.hash_filter_type = q.hash_filter_type; d.sort_filter_type = q.sort_filter_type;
        d.max_count = q.max_count; d.max_iter_count = q.max_iter_count; d.max_iter_size = q.max_iter_size;
        d.pidx = q.pidx; d.partition_version = q.partition_version;
        need_crc |= q.validate_hash != 0;
        if (q.reverse) return PGS_NOT_SUPPORTED;
    }
    blob.append(16, '\0'); // as lookup.cu
    uint32_t err[16] = {0};
    P.reqs = dev.data(); P.blob = (const uint8_t *)blob.data(); P.n = n; P.now = now; P.data_version = 1;
    P.results = results; P.kvs = kvs; P.kv_stride = kv_stride; P.arena = arena; P.arena_stride = arena_stride;
    P.resume = resume; P.resume_stride = resume_stride; P.error = err; P.ticket = err + 8;
    if (need_crc) { make_crc(); P.crc_table = (const unsigned long long *)crc_tab; }
    // lanes & 0x100: the multi-partition shape of pgs_range_scan_many_multi -- slot 0 owns the runs [0, k/2), slot 1 the rest,
    // slot 2 is an empty partition; request i belongs to slot i % 3
    const bool multi = (lanes & 0x100) != 0;
    lanes &= 0xFF;
    std::vector<RunDev> packed(P.rr.runs, P.rr.runs + k);
    std::vector<uint32_t> begin = {0, k / 2, k, k}, rp(n);
    if (multi) {
        for (uint32_t i = 0; i < n; i++) rp[i] = i % 3;
        P.multi_runs = packed.data(); P.multi_begin = begin.data(); P.req_part = rp.data();
        P.rr.n = k - k / 2;
    }
    const uint32_t G = lanes ? lanes : (k <= 8 ? 8 : k <= 16 ? 16 : 32);
    if (G < k) return PGS_INVALID_ARGUMENT;
    P.KS = std::max(8u, (mk + 3) & ~3u);
    P.KSW = (P.KS + 8) / 4 + 1;
    P.group_smem = (uint32_t)((k * (sizeof(CurState) + P.KSW * 4) + 3 * P.KSW * 4 + 15) & ~(size_t)15);
    const uint32_t dyn = 2048 + kMaxReadRuns * (uint32_t)sizeof(RunDev) + (kReadThreads / G) * P.group_smem;
    if (multi) {
        if (G == 8) PGS_LAUNCH((k_scan_fwd<8, true>), 2, kReadThreads, dyn, 0, P);
        else if (G == 16) PGS_LAUNCH((k_scan_fwd<16, true>), 2, kReadThreads, dyn, 0, P);
        else PGS_LAUNCH((k_scan_fwd<32, true>), 2, kReadThreads, dyn, 0, P);
    } else if (G == 8) PGS_LAUNCH((k_scan_fwd<8, false>), 2, kReadThreads, dyn, 0, P);
    else if (G == 16) PGS_LAUNCH((k_scan_fwd<16, false>), 2, kReadThreads, dyn, 0, P);
    else PGS_LAUNCH((k_scan_fwd<32, false>), 2, kReadThreads, dyn, 0, P);
    return err[0] ? (int32_t)err[0] : PGS_OK;
}

} // extern "C"
