// sim_compact.cpp — runs the compaction kernels (incubator_pegasus_b200/csrc/compact_kernels.cuh, the same source nvcc compiles
// for sm_100a) inside the host SIMT interpreter of simt.h.  Test / development tool: lets the CPU test-suite execute the kernels'
// logic against the oracle without a GPU.  NOT part of the product library and never a fallback for it.
#include "simt.h"

#define PGS_SIM 1
#include "../../incubator_pegasus_b200/csrc/compact_kernels.cuh"

#include <string>
#include <vector>

using namespace pgs;

namespace {

struct HostRun {
    std::vector<uint8_t> data;
    std::vector<uint64_t> blk_off;
    std::vector<uint32_t> blk_size, blk_rec, ikey_off, rec_off;
    std::vector<uint8_t> ikeys;
    pgs_run_info info{};
    RunDev dev() const
    {
        return RunDev{data.data(), blk_off.data(), blk_size.data(), blk_rec.data(), ikey_off.data(), ikeys.data(), rec_off.data(),
                      (uint32_t)blk_size.size(), info.max_ukey_len};
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
    r.ikeys.resize(r.ikeys.size() + 64);
    r.rec_off.push_back(0);
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
        r.data.resize(r.data.size() + 512, 0);
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
    if (!compact_geometry(P, T, 227 * 1024, geo)) return PGS_NOT_SUPPORTED;
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
    if (group_lanes) {
        if (group_lanes < k || (group_lanes != 8 && group_lanes != 16 && group_lanes != 32)) return PGS_INVALID_ARGUMENT;
        geo.G = group_lanes;
        geo.walk_dyn = 2048 + kMaxRuns * (uint32_t)sizeof(RunDev) + (kWalkThreads / geo.G) * P.group_smem;
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
        if (geo.G == 8) PGS_LAUNCH(k_walk<8>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else if (geo.G == 16) PGS_LAUNCH(k_walk<16>, 2, kWalkThreads, geo.walk_dyn, 0, P);
        else PGS_LAUNCH(k_walk<32>, 2, kWalkThreads, geo.walk_dyn, 0, P);
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
    const uint64_t v[20] = {s.in_records, s.out_records, s.in_bytes, s.out_bytes, s.dropped_shadowed, s.dropped_tombstone, s.dropped_expired,
                            s.dropped_user, s.dropped_stale, s.ttl_rewritten, s.out_tomb, s.out_raw_key, s.out_raw_val, s.max_ukey, s.max_vlen,
                            s.max_blk_size, s.max_blk_rec, s.tot_recs ? ~s.min_seq_inv : ~0ull, s.max_seq, s.tot_keyb};
    memcpy(o, v, sizeof v);
}

} // extern "C"
