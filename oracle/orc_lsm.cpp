// orc_lsm.cpp — CPU ORACLE (test infrastructure only): the LSM engine semantics the reference
// gets from RocksDB v8.5.3 (not in the reference tree; restated from its published block format
// and compaction behaviour):
//   * data block codec (varint shared/non_shared/value_len entries, restart array, flush policy)
//   * semantic compaction over flat records
//   * the CPU baseline: block-level compaction (heap MergingIterator -> compaction loop ->
//     BlockBuilder) with sub-compactions on threads, point gets and prefix scans over block runs.
#include "orc_internal.h"

#include <algorithm>
#include <chrono>
#include <queue>
#include <thread>

namespace orc {

// ---- varint / fixed (RocksDB util/coding.h) ----
static inline void put_varint32(std::string &dst, uint32_t v)
{
    while (v >= 128) { dst.push_back((char)(v | 128)); v >>= 7; }
    dst.push_back((char)v);
}
static inline const uint8_t *get_varint32(const uint8_t *p, const uint8_t *limit, uint32_t *v)
{
    uint32_t r = 0;
    for (uint32_t shift = 0; shift <= 28 && p < limit; shift += 7) {
        uint32_t b = *p++;
        if (b & 128) r |= (b & 127) << shift;
        else { r |= b << shift; *v = r; return p; }
    }
    return nullptr;
}
static inline void put_fixed32(std::string &dst, uint32_t v) { dst.append((const char *)&v, 4); }
static inline uint32_t get_fixed32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t get_fixed64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline int varint_len(uint64_t v) { int l = 1; while (v >= 128) { v >>= 7; l++; } return l; }

static inline std::string internal_key(sv ukey, uint64_t seq, uint8_t type)
{
    std::string k(ukey);
    uint64_t t = (seq << 8) | type;
    k.append((const char *)&t, 8);
    return k;
}

// ---- BlockBuilder (RocksDB table/block_based/block_builder.cc semantics) ----
struct BlockBuilder {
    uint32_t interval;
    std::string buf;
    std::vector<uint32_t> restarts{0};
    uint32_t counter = 0;
    std::string last_key;
    uint32_t n = 0;
    explicit BlockBuilder(uint32_t ri) : interval(ri) {}
    bool empty() const { return buf.empty(); }
    size_t size_estimate() const { return buf.size() + restarts.size() * 4 + 4; }
    size_t estimate_after(size_t klen, size_t vlen) const
    {
        size_t e = size_estimate() + klen + vlen;
        if (counter >= interval) e += 4;
        e += 4; // varint for shared prefix length (RocksDB uses sizeof(int32_t))
        e += varint_len(klen) + varint_len(vlen);
        return e;
    }
    void add(sv key, sv value)
    {
        uint32_t shared = 0;
        if (counter >= interval) {
            restarts.push_back((uint32_t)buf.size());
            counter = 0;
        } else {
            size_t m = std::min(last_key.size(), key.size());
            while (shared < m && last_key[shared] == key[shared]) shared++;
        }
        uint32_t non_shared = (uint32_t)key.size() - shared;
        put_varint32(buf, shared);
        put_varint32(buf, non_shared);
        put_varint32(buf, (uint32_t)value.size());
        buf.append(key.data() + shared, non_shared);
        buf.append(value.data(), value.size());
        last_key.assign(key.data(), key.size());
        counter++;
        n++;
    }
    std::string finish()
    {
        for (uint32_t r : restarts) put_fixed32(buf, r);
        put_fixed32(buf, (uint32_t)restarts.size());
        return std::move(buf);
    }
};
// FlushBlockBySizePolicy::Update (flush_block_policy.cc): block_size_deviation 10
static bool should_flush(const BlockBuilder &b, size_t klen, size_t vlen, uint32_t block_size)
{
    if (b.empty()) return false;
    size_t cur = b.size_estimate();
    if (cur >= block_size) return true;
    size_t limit = ((size_t)block_size * 90 + 99) / 100;
    return b.estimate_after(klen, vlen) > block_size && cur > limit;
}

struct BlockRun {
    std::string data;               // blocks back to back, each start 16-aligned
    std::vector<uint64_t> off;
    std::vector<uint32_t> size;
    std::vector<std::string> last_ukey; // index: last user key of each block
    uint64_t n_records = 0;
};

struct RunBuilder {
    BlockRun out;
    BlockBuilder bb;
    uint32_t block_size;
    std::string last_internal;
    RunBuilder(uint32_t bs, uint32_t ri) : bb(ri), block_size(bs) {}
    void flush()
    {
        if (bb.empty()) return;
        uint32_t ri = bb.interval;
        std::string last = bb.last_key;
        std::string blk = bb.finish();
        while (out.data.size() % 16) out.data.push_back(0);
        out.off.push_back(out.data.size());
        out.size.push_back((uint32_t)blk.size());
        out.data += blk;
        out.last_ukey.emplace_back(last.data(), last.size() - 8);
        bb = BlockBuilder(ri);
    }
    void add(sv ukey, uint64_t seq, uint8_t type, sv value)
    {
        std::string ik = internal_key(ukey, seq, type);
        if (should_flush(bb, ik.size(), value.size(), block_size)) flush();
        bb.add(ik, value);
        out.n_records++;
    }
    BlockRun finish() { flush(); return std::move(out); }
};

// ---- block iterator: sequential decode + seek via restart array ----
struct BlockIter {
    const uint8_t *base = nullptr;
    uint32_t restart_off = 0, n_restarts = 0;
    const uint8_t *p = nullptr, *limit = nullptr;
    std::string key; // current internal key
    sv value;
    bool valid = false, corrupt = false;
    bool init(const uint8_t *b, uint32_t size)
    {
        base = b;
        valid = false;
        if (size < 4) return !(corrupt = true);
        n_restarts = get_fixed32(b + size - 4);
        if (n_restarts == 0 || (uint64_t)n_restarts * 4 + 4 > size) return !(corrupt = true);
        restart_off = size - 4 - n_restarts * 4;
        limit = b + restart_off;
        return true;
    }
    uint32_t restart_point(uint32_t i) const { return get_fixed32(base + restart_off + 4 * i); }
    void seek_restart(uint32_t i) { key.clear(); p = base + restart_point(i); next_from(p); }
    void seek_first() { seek_restart(0); }
    void next() { next_from(p); }
    void next_from(const uint8_t *q)
    {
        if (q >= limit) { valid = false; return; }
        uint32_t shared, non_shared, vlen;
        q = get_varint32(q, limit, &shared);
        if (q) q = get_varint32(q, limit, &non_shared);
        if (q) q = get_varint32(q, limit, &vlen);
        if (!q || shared > key.size() || (uint64_t)(limit - q) < (uint64_t)non_shared + vlen) {
            corrupt = true;
            valid = false;
            return;
        }
        key.resize(shared);
        key.append((const char *)q, non_shared);
        value = sv((const char *)q + non_shared, vlen);
        p = q + non_shared + vlen;
        valid = key.size() >= 8;
        if (!valid) corrupt = true;
    }
    sv ukey() const { return sv(key.data(), key.size() - 8); }
    uint64_t trailer() const { return get_fixed64((const uint8_t *)key.data() + key.size() - 8); }
    // position at first entry with internal key >= (uk, trailer t)   [t = max for "newest"]
    void seek(sv uk, uint64_t t)
    {
        // binary search over restart points (keys there are stored whole)
        uint32_t lo = 0, hi = n_restarts - 1;
        while (lo < hi) {
            uint32_t mid = (lo + hi + 1) / 2;
            const uint8_t *q = base + restart_point(mid);
            uint32_t shared, non_shared, vlen;
            q = get_varint32(q, limit, &shared);
            if (q) q = get_varint32(q, limit, &non_shared);
            if (q) q = get_varint32(q, limit, &vlen);
            if (!q || shared != 0 || non_shared < 8) { corrupt = true; valid = false; return; }
            sv k((const char *)q, non_shared - 8);
            uint64_t kt = get_fixed64(q + non_shared - 8);
            int c = k.compare(uk);
            if (c == 0) c = kt > t ? -1 : (kt < t ? 1 : 0);
            if (c < 0) lo = mid; else hi = mid - 1;
        }
        seek_restart(lo);
        while (valid) {
            int c = ukey().compare(uk);
            if (c == 0) { uint64_t kt = trailer(); c = kt > t ? -1 : (kt < t ? 1 : 0); }
            if (c >= 0) break;
            next();
        }
    }
};

// iterator over one block run (index bsearch + block iterators)
struct RunIter {
    const BlockRun *r;
    uint32_t blk = 0;
    BlockIter it;
    bool valid = false, corrupt = false;
    explicit RunIter(const BlockRun *run) : r(run) {}
    void load(uint32_t b)
    {
        blk = b;
        if (b >= r->off.size()) { valid = false; return; }
        if (!it.init((const uint8_t *)r->data.data() + r->off[b], r->size[b])) { corrupt = true; valid = false; }
    }
    void seek_first()
    {
        load(0);
        if (blk < r->off.size() && !corrupt) { it.seek_first(); settle(); }
    }
    void settle()
    {
        while (!it.valid && !it.corrupt) {
            load(blk + 1);
            if (blk >= r->off.size() || corrupt) { valid = false; return; }
            it.seek_first();
        }
        corrupt |= it.corrupt;
        valid = it.valid;
    }
    void seek(sv uk, uint64_t t)
    {
        // first block whose last user key >= uk
        auto pos = std::lower_bound(r->last_ukey.begin(), r->last_ukey.end(), uk,
                                    [](const std::string &a, sv b) { return sv(a).compare(b) < 0; });
        load((uint32_t)(pos - r->last_ukey.begin()));
        if (blk >= r->off.size() || corrupt) { valid = false; return; }
        it.seek(uk, t);
        settle();
    }
    void next() { it.next(); settle(); }
    sv ukey() const { return it.ukey(); }
    uint64_t trailer() const { return it.trailer(); }
    sv value() const { return it.value; }
};

// min-heap merging iterator (RocksDB MergingIterator): internal key order
struct MergeIter {
    std::vector<RunIter> its;
    struct Cmp {
        std::vector<RunIter> *v;
        bool operator()(uint32_t a, uint32_t b) const
        {
            const RunIter &x = (*v)[a], &y = (*v)[b];
            int c = x.ukey().compare(y.ukey());
            if (c) return c > 0;
            if (x.trailer() != y.trailer()) return x.trailer() < y.trailer();
            return a > b;
        }
    };
    std::vector<uint32_t> heap;
    explicit MergeIter(const std::vector<const BlockRun *> &runs)
    {
        for (auto *r : runs) its.emplace_back(r);
    }
    void rebuild()
    {
        heap.clear();
        for (uint32_t i = 0; i < its.size(); i++)
            if (its[i].valid) heap.push_back(i);
        std::make_heap(heap.begin(), heap.end(), Cmp{&its});
    }
    void seek_first() { for (auto &i : its) i.seek_first(); rebuild(); }
    void seek(sv uk, uint64_t t) { for (auto &i : its) i.seek(uk, t); rebuild(); }
    bool valid() const { return !heap.empty(); }
    RunIter &top() { return its[heap.front()]; }
    void next()
    {
        std::pop_heap(heap.begin(), heap.end(), Cmp{&its});
        uint32_t i = heap.back();
        its[i].next();
        if (its[i].valid) std::push_heap(heap.begin(), heap.end(), Cmp{&its});
        else heap.pop_back();
    }
    bool corrupt() const { for (auto &i : its) if (i.corrupt) return true; return false; }
};

// -------------------------------------------------------------------------------------------
// semantic compaction over flat runs.  RocksDB CompactionIterator (no snapshots): per user key
// only the newest entry survives; kTypeValue entries go through CompactionFilter::Filter, a
// "remove" decision turns the entry into a deletion; deletions are dropped only when the output
// is the bottommost level, otherwise kept; at the bottommost level seqnos are zeroed.
// -------------------------------------------------------------------------------------------
template <class Emit>
static void compaction_step(sv ukey, uint64_t trailer, sv value, bool bottommost,
                            const FilterParams &fp, uint32_t now, orc_compact_stats *st, Emit &&emit)
{
    uint64_t seq = trailer >> 8;
    uint8_t type = (uint8_t)trailer;
    std::string nv;
    bool changed = false;
    if (type == PGS_TYPE_VALUE) {
        DropReason d = compaction_filter(fp, ukey, value, now, &nv, &changed);
        if (d != kKeep) {
            if (d == kDropExpired) st->dropped_expired++;
            else if (d == kDropUser) st->dropped_user++;
            else st->dropped_stale++;
            type = PGS_TYPE_DELETION;
            value = sv();
            changed = false;
            if (bottommost) return;
            emit(ukey, seq, type, value);
            return;
        }
        if (changed) { st->ttl_rewritten++; value = nv; }
        emit(ukey, bottommost ? 0 : seq, type, value);
        return;
    }
    if (type == PGS_TYPE_DELETION) {
        if (bottommost) { st->dropped_tombstone++; return; }
        emit(ukey, seq, type, value);
        return;
    }
    emit(ukey, seq, type, value);
}

Run compact(const std::vector<const Run *> &runs, bool bottommost, const FilterParams &fp,
            uint32_t now, orc_compact_stats *st)
{
    orc_compact_stats local{};
    if (!st) st = &local;
    std::vector<const Rec *> all;
    for (auto *r : runs)
        for (auto &rec : r->recs) {
            all.push_back(&rec);
            st->in_records++;
            st->in_bytes += rec.ukey.size() + rec.value.size();
        }
    std::stable_sort(all.begin(), all.end(), [](const Rec *a, const Rec *b) {
        return cmp_internal(a->ukey, a->seq, a->type, b->ukey, b->seq, b->type) < 0;
    });
    Run out;
    const Rec *prev = nullptr;
    for (const Rec *r : all) {
        if (prev && prev->ukey == r->ukey) { st->dropped_shadowed++; continue; }
        prev = r;
        compaction_step(r->ukey, (r->seq << 8) | r->type, r->value, bottommost, fp, now, st,
                        [&](sv k, uint64_t seq, uint8_t type, sv v) {
                            out.recs.push_back(Rec{std::string(k), seq, type, std::string(v)});
                            st->out_records++;
                            st->out_bytes += k.size() + v.size();
                        });
    }
    return out;
}

static Run decode_blockrun(const BlockRun &b, bool *corrupt)
{
    Run out;
    RunIter it(&b);
    for (it.seek_first(); it.valid; it.next())
        out.recs.push_back(Rec{std::string(it.ukey()), it.trailer() >> 8, (uint8_t)it.trailer(), std::string(it.value())});
    if (corrupt) *corrupt = it.corrupt;
    return out;
}

static void add_stats(orc_compact_stats &a, const orc_compact_stats &b)
{
    a.in_records += b.in_records; a.out_records += b.out_records;
    a.in_bytes += b.in_bytes; a.out_bytes += b.out_bytes;
    a.dropped_shadowed += b.dropped_shadowed; a.dropped_tombstone += b.dropped_tombstone;
    a.dropped_expired += b.dropped_expired; a.dropped_user += b.dropped_user;
    a.dropped_stale += b.dropped_stale; a.ttl_rewritten += b.ttl_rewritten;
}

// one sub-compaction over user keys in [lo, hi)  (hi empty & !has_hi = +inf)
static BlockRun subcompact(const std::vector<const BlockRun *> &runs, sv lo, bool has_lo, sv hi,
                           bool has_hi, bool bottommost, const FilterParams &fp, uint32_t now,
                           uint32_t block_size, uint32_t ri, orc_compact_stats *st)
{
    MergeIter mi(runs);
    if (has_lo) mi.seek(lo, UINT64_MAX); else mi.seek_first();
    RunBuilder rb(block_size, ri);
    std::string cur;
    bool have_cur = false;
    while (mi.valid()) {
        RunIter &t = mi.top();
        sv uk = t.ukey();
        if (has_hi && uk.compare(hi) >= 0) break;
        st->in_records++;
        st->in_bytes += uk.size() + t.value().size();
        if (have_cur && sv(cur) == uk) {
            st->dropped_shadowed++;
        } else {
            cur.assign(uk.data(), uk.size());
            have_cur = true;
            compaction_step(uk, t.trailer(), t.value(), bottommost, fp, now, st,
                            [&](sv k, uint64_t seq, uint8_t type, sv v) {
                                rb.add(k, seq, type, v);
                                st->out_records++;
                                st->out_bytes += k.size() + v.size();
                            });
        }
        mi.next();
    }
    return rb.finish();
}

} // namespace orc

using namespace orc;
struct orc_blockrun { BlockRun br; };

static inline sv mk(const uint8_t *p, uint64_t n) { return sv((const char *)p, n); }

extern "C" {

orc_run *orc_run_from_records(uint64_t n, const uint8_t *keys, const uint64_t *key_off,
                              const uint8_t *vals, const uint64_t *val_off, const uint64_t *seq,
                              const uint8_t *type)
{
    auto *r = new orc_run;
    r->run.recs.reserve(n);
    for (uint64_t i = 0; i < n; i++)
        r->run.recs.push_back(Rec{std::string((const char *)keys + key_off[i], key_off[i + 1] - key_off[i]), seq[i],
                                  type[i],
                                  std::string((const char *)vals + val_off[i], val_off[i + 1] - val_off[i])});
    return r;
}
void orc_run_free(orc_run *r) { delete r; }
void orc_run_sizes(const orc_run *r, pgs_decode_sizes *out)
{
    out->n_records = r->run.recs.size();
    out->key_bytes = out->value_bytes = 0;
    for (auto &rec : r->run.recs) { out->key_bytes += rec.ukey.size(); out->value_bytes += rec.value.size(); }
}
void orc_run_export(const orc_run *r, uint8_t *keys, uint64_t *key_off, uint8_t *vals,
                    uint64_t *val_off, uint64_t *seq, uint8_t *type)
{
    uint64_t ko = 0, vo = 0, i = 0;
    for (auto &rec : r->run.recs) {
        key_off[i] = ko; val_off[i] = vo;
        memcpy(keys + ko, rec.ukey.data(), rec.ukey.size());
        memcpy(vals + vo, rec.value.data(), rec.value.size());
        ko += rec.ukey.size(); vo += rec.value.size();
        seq[i] = rec.seq; type[i] = rec.type;
        i++;
    }
    key_off[i] = ko; val_off[i] = vo;
}

static BlockRun blockrun_from_raw(const uint8_t *data, uint64_t data_bytes, const uint64_t *blk_off,
                                  const uint32_t *blk_size, uint32_t n_blocks, bool *corrupt)
{
    BlockRun br;
    br.data.assign((const char *)data, data_bytes);
    br.off.assign(blk_off, blk_off + n_blocks);
    br.size.assign(blk_size, blk_size + n_blocks);
    *corrupt = false;
    for (uint32_t b = 0; b < n_blocks; b++) {
        BlockIter it;
        std::string last;
        uint64_t n = 0;
        if (blk_off[b] + blk_size[b] > data_bytes || !it.init(data + blk_off[b], blk_size[b])) { *corrupt = true; break; }
        for (it.seek_first(); it.valid; it.next()) { last.assign(it.ukey().data(), it.ukey().size()); n++; }
        if (it.corrupt || n == 0) { *corrupt = true; break; }
        br.last_ukey.push_back(last);
        br.n_records += n;
    }
    return br;
}

orc_run *orc_run_from_blocks(const uint8_t *data, const uint64_t *blk_off, const uint32_t *blk_size,
                             uint32_t n_blocks, int32_t *status)
{
    uint64_t bytes = n_blocks ? blk_off[n_blocks - 1] + blk_size[n_blocks - 1] : 0;
    bool corrupt = false;
    BlockRun br = blockrun_from_raw(data, bytes, blk_off, blk_size, n_blocks, &corrupt);
    auto *r = new orc_run;
    if (!corrupt) r->run = decode_blockrun(br, &corrupt);
    if (status) *status = corrupt ? PGS_CORRUPTION : PGS_OK;
    return r;
}

static FilterParams to_fp(const orc_filter_params *p)
{
    FilterParams fp;
    if (!p) return fp;
    fp.enabled = p->enabled; fp.validate_hash = p->validate_hash;
    fp.data_version = p->data_version; fp.default_ttl = p->default_ttl;
    fp.pidx = p->pidx; fp.partition_version = p->partition_version;
    fp.ops = p->ops ? &p->ops->ops : nullptr;
    return fp;
}

orc_run *orc_compact(const orc_run *const *runs, uint32_t k, int32_t bottommost,
                     const orc_filter_params *p, uint32_t now, orc_compact_stats *st)
{
    std::vector<const Run *> rs;
    for (uint32_t i = 0; i < k; i++) rs.push_back(&runs[i]->run);
    orc_compact_stats local{};
    auto *out = new orc_run;
    out->run = compact(rs, bottommost != 0, to_fp(p), now, &local);
    if (st) *st = local;
    return out;
}

orc_blockrun *orc_blockrun_build(const orc_run *r, uint32_t block_size, uint32_t restart_interval)
{
    RunBuilder rb(block_size ? block_size : 4096, restart_interval ? restart_interval : 16);
    for (auto &rec : r->run.recs) rb.add(rec.ukey, rec.seq, rec.type, rec.value);
    auto *b = new orc_blockrun;
    b->br = rb.finish();
    return b;
}
orc_blockrun *orc_blockrun_from_blocks(const uint8_t *data, uint64_t data_bytes,
                                       const uint64_t *blk_off, const uint32_t *blk_size,
                                       uint32_t n_blocks)
{
    bool corrupt = false;
    auto *b = new orc_blockrun;
    b->br = blockrun_from_raw(data, data_bytes, blk_off, blk_size, n_blocks, &corrupt);
    if (corrupt) { delete b; return nullptr; }
    return b;
}
void orc_blockrun_free(orc_blockrun *b) { delete b; }
uint64_t orc_blockrun_bytes(const orc_blockrun *b) { return b->br.data.size(); }
uint32_t orc_blockrun_blocks(const orc_blockrun *b) { return (uint32_t)b->br.off.size(); }
orc_run *orc_blockrun_decode(const orc_blockrun *b)
{
    auto *r = new orc_run;
    r->run = decode_blockrun(b->br, nullptr);
    return r;
}

orc_blockrun *orc_compact_blocks(const orc_blockrun *const *runs, uint32_t k, int32_t bottommost,
                                 const orc_filter_params *p, uint32_t now, uint32_t threads,
                                 uint32_t block_size, uint32_t restart_interval,
                                 orc_compact_stats *st, double *seconds)
{
    if (!block_size) block_size = 4096;
    if (!restart_interval) restart_interval = 16;
    if (!threads) threads = 1;
    std::vector<const BlockRun *> rs;
    for (uint32_t i = 0; i < k; i++) rs.push_back(&runs[i]->br);
    FilterParams fp = to_fp(p);
    // sub-compaction boundaries: evenly spaced index keys of the union of all runs' indexes
    std::vector<std::string> cand;
    for (auto *r : rs)
        for (auto &lk : r->last_ukey) cand.push_back(lk);
    std::sort(cand.begin(), cand.end());
    std::vector<std::string> bounds;
    for (uint32_t t = 1; t < threads && !cand.empty(); t++) {
        const std::string &c = cand[(size_t)cand.size() * t / threads];
        if (bounds.empty() || bounds.back() < c) bounds.push_back(c);
    }
    uint32_t parts = (uint32_t)bounds.size() + 1;
    std::vector<BlockRun> outs(parts);
    std::vector<orc_compact_stats> sts(parts);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < parts; i++) {
        th.emplace_back([&, i] {
            memset(&sts[i], 0, sizeof(sts[i]));
            sv lo = i > 0 ? sv(bounds[i - 1]) : sv();
            sv hi = i < parts - 1 ? sv(bounds[i]) : sv();
            outs[i] = subcompact(rs, lo, i > 0, hi, i < parts - 1, bottommost != 0, fp, now, block_size,
                                 restart_interval, &sts[i]);
        });
    }
    for (auto &t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    auto *out = new orc_blockrun;
    orc_compact_stats total{};
    for (uint32_t i = 0; i < parts; i++) {
        add_stats(total, sts[i]);
        BlockRun &o = outs[i];
        while (out->br.data.size() % 16) out->br.data.push_back(0);
        uint64_t base = out->br.data.size();
        out->br.data += o.data;
        for (size_t b = 0; b < o.off.size(); b++) {
            out->br.off.push_back(base + o.off[b]);
            out->br.size.push_back(o.size[b]);
            out->br.last_ukey.push_back(std::move(o.last_ukey[b]));
        }
        out->br.n_records += o.n_records;
    }
    if (st) *st = total;
    return out;
}

// DBImpl::Get over runs newest->oldest: index bsearch -> block seek -> newest visible version
static bool runs_get(const std::vector<const BlockRun *> &rs, sv key, std::string *val)
{
    for (auto *r : rs) {
        RunIter it(r);
        it.seek(key, UINT64_MAX);
        if (it.valid && it.ukey() == key) {
            if ((uint8_t)it.trailer() != PGS_TYPE_VALUE) return false;
            val->assign(it.value().data(), it.value().size());
            return true;
        }
    }
    return false;
}

uint64_t orc_blockruns_get_many(const orc_blockrun *const *runs, uint32_t k, const uint8_t *keys,
                                const uint32_t *key_off, uint32_t n, uint32_t now, uint32_t threads,
                                uint64_t *value_bytes, double *seconds)
{
    std::vector<const BlockRun *> rs;
    for (uint32_t i = 0; i < k; i++) rs.push_back(&runs[i]->br);
    if (!threads) threads = 1;
    std::vector<uint64_t> found(threads, 0), bytes(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            std::string v;
            for (uint64_t i = (uint64_t)n * t / threads; i < (uint64_t)n * (t + 1) / threads; i++) {
                sv key = mk(keys + key_off[i], key_off[i + 1] - key_off[i]);
                // on_get: Get -> TTL hide -> strip header (pegasus_server_impl.cpp:441-490)
                if (runs_get(rs, key, &v) && !ts_expired(now, extract_expire_ts(1, v))) {
                    found[t]++;
                    bytes[t] += v.size() - user_data_offset(1);
                }
            }
        });
    for (auto &t : th) t.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t f = 0, b = 0;
    for (uint32_t t = 0; t < threads; t++) { f += found[t]; b += bytes[t]; }
    if (value_bytes) *value_bytes = b;
    return f;
}

uint64_t orc_blockruns_prefix_scan_many(const orc_blockrun *const *runs, uint32_t k,
                                        const uint8_t *hashkeys, const uint32_t *hk_off, uint32_t n,
                                        uint32_t now, uint32_t threads, uint64_t *bytes_out,
                                        double *seconds)
{
    std::vector<const BlockRun *> rs;
    for (uint32_t i = 0; i < k; i++) rs.push_back(&runs[i]->br);
    if (!threads) threads = 1;
    std::vector<uint64_t> cnt(threads, 0), bytes(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            std::string out; // the response the handler would fill (key + user data copied)
            for (uint64_t i = (uint64_t)n * t / threads; i < (uint64_t)n * (t + 1) / threads; i++) {
                sv hk = mk(hashkeys + hk_off[i], hk_off[i + 1] - hk_off[i]);
                std::string start = generate_key(hk, sv()), stop = next_blob(hk);
                MergeIter mi(rs);
                mi.seek(start, UINT64_MAX);
                std::string cur;
                bool have = false;
                out.clear();
                while (mi.valid()) {
                    RunIter &tp = mi.top();
                    sv uk = tp.ukey();
                    if (uk.compare(stop) >= 0) break;
                    if (!(have && sv(cur) == uk)) {
                        cur.assign(uk.data(), uk.size());
                        have = true;
                        if ((uint8_t)tp.trailer() == PGS_TYPE_VALUE && !ts_expired(now, extract_expire_ts(1, tp.value()))) {
                            cnt[t]++;
                            out.append(uk.data(), uk.size());
                            out.append(tp.value().data() + 12, tp.value().size() - 12);
                        }
                    }
                    mi.next();
                }
                bytes[t] += out.size();
            }
        });
    for (auto &t : th) t.join();
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    uint64_t c = 0, b = 0;
    for (uint32_t t = 0; t < threads; t++) { c += cnt[t]; b += bytes[t]; }
    if (bytes_out) *bytes_out = b;
    return c;
}

} // extern "C"
