// orc_rrdb.cpp — CPU ORACLE (test infrastructure only): the rrdb operator surface of one
// replica (pegasus_server_impl) on a semantic LSM model: memtable + sorted runs of
// (user key, seq, type, value), newest version wins, tombstones hide.  Handlers restate
// src/server/pegasus_server_impl.cpp:418-1549 and the helpers at :2350-2504; the write side
// restates src/server/rocksdb_wrapper.cpp:129-219 / pegasus_write_service_impl.h:90-169.
// Wall-clock limits of range_read_limiter (range_read_limiter.h:64-84) are not modelled: `now`
// is an argument and no test runs 30 s.
#include "orc_internal.h"

#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <ctime>
#include <map>
#include <random>
#include <sstream>
#include <unordered_map>

namespace orc {

struct View { // what a RocksDB iterator sees: newest non-deleted version of every user key
    std::vector<std::pair<std::string, std::string>> kv;
};

struct Iter { // rocksdb::Iterator over a pinned View
    std::shared_ptr<View> v;
    size_t pos = 0;
    bool ok = false;
    bool prefix_mode = false;
    std::string prefix;
    bool has_upper = false;
    std::string upper;
    bool Valid() const { return ok; }
    sv key() const { return v->kv[pos].first; }
    sv value() const { return v->kv[pos].second; }
    void check()
    {
        ok = pos < v->kv.size();
        if (ok && prefix_mode && hashkey_prefix(key()) != sv(prefix)) ok = false;
        if (ok && has_upper && key().compare(upper) >= 0) ok = false;
    }
    void Seek(sv target, bool prefix_same_as_start)
    {
        pos = std::lower_bound(v->kv.begin(), v->kv.end(), target,
                               [](const std::pair<std::string, std::string> &a, sv b) { return sv(a.first).compare(b) < 0; }) -
              v->kv.begin();
        prefix_mode = prefix_same_as_start && target.size() >= 2; // InDomain
        if (prefix_mode) prefix = std::string(hashkey_prefix(target));
        check();
    }
    void SeekForPrev(sv target) // total order
    {
        size_t ub = std::upper_bound(v->kv.begin(), v->kv.end(), target,
                                     [](sv b, const std::pair<std::string, std::string> &a) { return b.compare(a.first) < 0; }) -
                    v->kv.begin();
        prefix_mode = false;
        if (ub == 0) { ok = false; return; }
        pos = ub - 1;
        ok = true;
    }
    void Next() { pos++; check(); }
    void Prev()
    {
        if (pos == 0) { ok = false; return; }
        pos--;
        ok = true;
    }
};

struct ScanContext { // pegasus_scan_context.h:33-95
    Iter it;
    std::string stop;
    bool stop_inclusive;
    int hash_key_filter_type, sort_key_filter_type;
    std::string hash_key_filter_pattern, sort_key_filter_pattern;
    int32_t batch_size;
    bool no_value, validate_partition_hash, return_expire_ts, only_return_count;
    uint32_t parked_at = 0; // epoch seconds; a parked context lives 5 minutes (pegasus_server_impl.cpp:1377-1385)
};

struct Resp {
    pgs_response view{};
    std::vector<pgs_kv> kvs;
    std::vector<uint32_t> hk_len;
    std::string arena;
    void reset(int32_t app_id, int32_t pidx)
    {
        view = pgs_response{};
        view.app_id = app_id;
        view.partition_index = pidx;
        view.kv_count = -1;
        view.context_id = 0;
        kvs.clear();
        hk_len.clear();
        arena.clear();
    }
    void add(sv key, sv value, uint32_t expire_ts)
    {
        pgs_kv kv;
        kv.key_off = (uint32_t)arena.size(); kv.key_len = (uint32_t)key.size();
        arena.append(key.data(), key.size());
        kv.value_off = (uint32_t)arena.size(); kv.value_len = (uint32_t)value.size();
        arena.append(value.data(), value.size());
        kv.expire_ts = expire_ts;
        kvs.push_back(kv);
    }
    void seal()
    {
        view.n_kvs = (uint32_t)kvs.size();
        view.kvs = kvs.data();
        view.hk_len = hk_len.empty() ? nullptr : hk_len.data();
        view.arena = (const uint8_t *)arena.data();
        view.arena_len = arena.size();
    }
};

enum RangeState { kNormal, kExpired, kFiltered, kHashInvalid };

struct Limiter { // range_read_limiter.h:37-103 without the clock
    uint32_t max_count; uint64_t max_size;
    uint32_t count = 0; uint64_t size = 0;
    Limiter(uint32_t c, uint64_t s) : max_count(c), max_size(s) {}
    bool valid() const { return count < max_count && !(max_size > 0 && size >= max_size); }
};

struct Server {
    int32_t app_id, pidx;
    pgs_server_options opt;
    uint32_t data_version = 1; // pegasus_server_impl_test.cpp:356-360
    uint32_t default_ttl = 0;
    bool validate_partition_hash = false; // server flag, env replica.split.validate_partition_hash
    int32_t partition_version = -1;
    std::vector<Op> ops;
    uint64_t last_seq = 0;
    int64_t last_committed_decree = 0, last_flushed_decree = 0;
    uint64_t manual_compact_last_finish_ms = 0;
    std::map<std::string, Rec> mem;
    uint64_t mem_bytes = 0;
    struct LRun { int level; Run run; };
    std::vector<LRun> runs; // read order: L0 newest first, then L1, L2, ...
    std::shared_ptr<View> view;
    int64_t ctx_counter;
    std::unordered_map<int64_t, std::unique_ptr<ScanContext>> ctx;

    Server()
    {
        std::mt19937_64 rng(12345);
        ctx_counter = (int64_t)(rng() % (1ull << 31)) << 32; // pegasus_scan_context.h:113-114 (kept non-negative: the reference's 2^32 range can overflow into the reserved negative ids)
    }
    FilterParams fparams() const
    {
        FilterParams fp;
        fp.enabled = true;
        fp.validate_hash = validate_partition_hash;
        fp.data_version = data_version;
        fp.default_ttl = default_ttl;
        fp.pidx = pidx;
        fp.partition_version = partition_version;
        fp.ops = &ops;
        return fp;
    }
    void write(Rec r, uint32_t now)
    {
        view.reset();
        mem_bytes += r.ukey.size() + r.value.size() + 16;
        std::string k = r.ukey;
        mem[k] = std::move(r);
        uint64_t cap = opt.memtable_bytes ? opt.memtable_bytes : 64ull << 20;
        if (mem_bytes >= cap) { flush_mem(); maybe_compact(now); }
    }
    void flush_mem()
    {
        mem_bytes = 0;
        if (mem.empty()) return;
        last_flushed_decree = last_committed_decree;
        LRun lr{0, {}};
        for (auto &kv : mem) lr.run.recs.push_back(kv.second);
        mem.clear();
        runs.insert(runs.begin(), std::move(lr));
        view.reset();
    }
    void compact_runs(size_t first, size_t last /*exclusive*/, int out_level, uint32_t now, orc_compact_stats *st)
    {
        std::vector<const Run *> in;
        for (size_t i = first; i < last; i++) in.push_back(&runs[i].run);
        bool bottommost = last == runs.size();
        Run out = compact(in, bottommost, fparams(), now, st);
        runs.erase(runs.begin() + first, runs.begin() + last);
        size_t pos = 0;
        while (pos < runs.size() && runs[pos].level < out_level) pos++;
        if (!out.recs.empty()) runs.insert(runs.begin() + pos, LRun{out_level, std::move(out)}); // an empty output installs no run
        view.reset();
    }
    void compact_l0(uint32_t now)
    {
        size_t last = 0;
        while (last < runs.size() && runs[last].level <= 1) last++;
        if (last < 2) return;
        compact_runs(0, last, 1, now, nullptr);
    }
    void maybe_compact(uint32_t now)
    {
        uint32_t trigger = opt.l0_compaction_trigger ? opt.l0_compaction_trigger : 4;
        size_t l0 = 0;
        while (l0 < runs.size() && runs[l0].level == 0) l0++;
        if (l0 >= trigger) compact_l0(now);
    }
    // the product flushes the memtable before a RANGE read (and folds L0 once more than 12 runs pile up); point reads
    // see the memtable in place.  Both steps are invisible in RocksDB terms but the fold runs the compaction filter at `now`.
    void prepare_read(uint32_t now, bool range)
    {
        if (!range) return;
        if (mem.empty() && runs.size() <= 12) return;
        flush_mem();
        if (runs.size() > 12) compact_l0(now);
    }
    uint32_t gc_contexts(uint32_t now)
    {
        uint32_t n = 0;
        for (auto it = ctx.begin(); it != ctx.end();) {
            if (now >= it->second->parked_at && now - it->second->parked_at >= 300) { it = ctx.erase(it); n++; }
            else ++it;
        }
        return n;
    }
    std::shared_ptr<View> get_view()
    {
        if (view) return view;
        std::vector<const Rec *> all;
        for (auto &kv : mem) all.push_back(&kv.second); // the memtable holds the newest versions
        for (auto &lr : runs)
            for (auto &r : lr.run.recs) all.push_back(&r);
        std::stable_sort(all.begin(), all.end(), [](const Rec *a, const Rec *b) {
            return cmp_internal(a->ukey, a->seq, a->type, b->ukey, b->seq, b->type) < 0;
        });
        auto v = std::make_shared<View>();
        const Rec *prev = nullptr;
        for (const Rec *r : all) {
            if (prev && prev->ukey == r->ukey) continue;
            prev = r;
            if (r->type == PGS_TYPE_VALUE) v->kv.emplace_back(r->ukey, r->value);
        }
        view = v;
        return v;
    }
    bool db_get(sv key, std::string *value)
    {
        auto v = get_view();
        auto it = std::lower_bound(v->kv.begin(), v->kv.end(), key,
                                   [](const std::pair<std::string, std::string> &a, sv b) { return sv(a.first).compare(b) < 0; });
        if (it == v->kv.end() || sv(it->first) != key) return false;
        *value = it->second;
        return true;
    }
    sv user_data(sv raw) const { return raw.substr(user_data_offset(data_version)); }

    // validate_key_value_for_scan: pegasus_server_impl.cpp:2382-2432
    RangeState validate_for_scan(sv key, sv value, int hft, sv hpat, int sft, sv spat, uint32_t now,
                                 bool request_validate_hash) const
    {
        if (ts_expired(now, extract_expire_ts(data_version, value))) return kExpired;
        if (request_validate_hash && validate_partition_hash) {
            if (partition_version < 0 || pidx > partition_version || !check_key_hash(key, pidx, partition_version))
                return kHashInvalid;
        }
        if (hft != PGS_FT_NO_FILTER || sft != PGS_FT_NO_FILTER) {
            sv hk, sk;
            restore_key(key, hk, sk);
            if (hft != PGS_FT_NO_FILTER && !validate_filter(hft, hpat, hk)) return kFiltered;
            if (sft != PGS_FT_NO_FILTER && !validate_filter(sft, spat, sk)) return kFiltered;
        }
        return kNormal;
    }
    // append_key_value: :2434-2460
    void append_kv(Resp &r, sv key, sv value, bool no_value, bool request_expire_ts) const
    {
        uint32_t ets = request_expire_ts ? extract_expire_ts(data_version, value) : 0;
        r.add(key, no_value ? sv() : user_data(value), ets);
    }
    // append_key_value_for_multi_get: :2462-2504
    RangeState append_for_multi_get(Resp &r, sv key, sv value, int sft, sv spat, uint32_t now, bool no_value) const
    {
        if (ts_expired(now, extract_expire_ts(data_version, value))) return kExpired;
        sv hk, sk;
        restore_key(key, hk, sk);
        if (sft != PGS_FT_NO_FILTER && !validate_filter(sft, spat, sk)) return kFiltered;
        r.add(sk, no_value ? sv() : user_data(value), 0);
        return kNormal;
    }
};

static inline sv bsv(const pgs_blob &b) { return sv((const char *)b.data, b.len); }
static inline bool filter_type_supported(int t) { return t >= PGS_FT_NO_FILTER && t <= PGS_FT_MATCH_POSTFIX; }

// on_get: pegasus_server_impl.cpp:418-494
static int32_t on_get(Server &s, sv key, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    s.prepare_read(now, false);
    std::string value;
    int32_t st = s.db_get(key, &value) ? PGS_OK : PGS_NOT_FOUND;
    if (st == PGS_OK && ts_expired(now, extract_expire_ts(s.data_version, value))) {
        st = PGS_NOT_FOUND;
        r.view.expire_count = 1;
    }
    r.view.error = st;
    if (st == PGS_OK) r.add(sv(), s.user_data(value), 0);
    r.seal();
    return st;
}

// on_ttl: :1088-1149
static int32_t on_ttl(Server &s, sv key, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    s.prepare_read(now, false);
    std::string value;
    int32_t st = s.db_get(key, &value) ? PGS_OK : PGS_NOT_FOUND;
    uint32_t expire_ts = 0;
    if (st == PGS_OK) {
        expire_ts = extract_expire_ts(s.data_version, value);
        if (ts_expired(now, expire_ts)) { st = PGS_NOT_FOUND; r.view.expire_count = 1; }
    }
    r.view.error = st;
    if (st == PGS_OK) r.view.ttl_seconds = expire_ts > 0 ? (int32_t)(expire_ts - now) : -1;
    r.seal();
    return st;
}

// on_multi_get: :496-904
static int32_t on_multi_get(Server &s, const pgs_multi_get_request &q, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    if (!filter_type_supported(q.sort_key_filter_type)) {
        r.view.error = PGS_INVALID_ARGUMENT;
        r.seal();
        return r.view.error;
    }
    s.prepare_read(now, q.n_sort_keys == 0);
    uint32_t cfg_count = s.opt.rocksdb_multi_get_max_iteration_count ? s.opt.rocksdb_multi_get_max_iteration_count : 3000;
    uint64_t cfg_size = s.opt.rocksdb_multi_get_max_iteration_size ? s.opt.rocksdb_multi_get_max_iteration_size : 30ull << 20;
    uint32_t max_kv_count = cfg_count, max_iteration_count = cfg_count;
    if (q.max_kv_count > 0 && (uint32_t)q.max_kv_count < max_kv_count) max_kv_count = q.max_kv_count;
    int32_t max_kv_size = q.max_kv_size > 0 ? q.max_kv_size : INT_MAX;
    int32_t max_iteration_size_config = cfg_size > 0 ? (int32_t)std::min<uint64_t>(cfg_size, INT_MAX) : INT_MAX;
    int32_t max_iteration_size = std::min(max_kv_size, max_iteration_size_config);
    int32_t count = 0;
    int64_t size = 0;
    sv hash_key = bsv(q.hash_key);

    if (q.n_sort_keys == 0) {
        std::string start = generate_key(hash_key, bsv(q.start_sortkey));
        bool start_inclusive = q.start_inclusive;
        std::string stop;
        bool stop_inclusive;
        if (q.stop_sortkey.len == 0) { stop = next_blob(hash_key); stop_inclusive = false; }
        else { stop = generate_key(hash_key, bsv(q.stop_sortkey)); stop_inclusive = q.stop_inclusive; }
        if (q.sort_key_filter_type == PGS_FT_MATCH_PREFIX && q.sort_key_filter_pattern.len > 0) {
            std::string ps = generate_key(hash_key, bsv(q.sort_key_filter_pattern));
            std::string pe = next_blob(hash_key, bsv(q.sort_key_filter_pattern));
            if (sv(ps).compare(start) > 0) { start = ps; start_inclusive = true; }
            if (sv(pe).compare(stop) <= 0) { stop = pe; stop_inclusive = false; }
        }
        int c = sv(start).compare(stop);
        if (c > 0 || (c == 0 && (!start_inclusive || !stop_inclusive))) {
            r.view.error = PGS_OK;
            r.seal();
            return PGS_OK;
        }
        Iter it;
        it.v = s.get_view();
        bool complete = false;
        Limiter lim(max_iteration_count, (uint64_t)max_iteration_size);
        bool prefix = s.opt.prefix_filter;
        if (!q.reverse) {
            it.Seek(start, prefix);
            bool first_exclusive = !start_inclusive;
            while ((uint32_t)count < max_kv_count && lim.valid() && it.Valid()) {
                int c2 = it.key().compare(stop);
                if (c2 > 0 || (c2 == 0 && !stop_inclusive)) { complete = true; break; }
                if (first_exclusive) {
                    first_exclusive = false;
                    if (it.key().compare(start) == 0) { it.Next(); continue; }
                }
                lim.count++;
                size_t before = r.kvs.size();
                RangeState st = s.append_for_multi_get(r, it.key(), it.value(), q.sort_key_filter_type,
                                                       bsv(q.sort_key_filter_pattern), now, q.no_value);
                if (st == kNormal) {
                    count++;
                    uint64_t kv_size = r.kvs[before].key_len + r.kvs[before].value_len;
                    size += kv_size;
                    lim.size += kv_size;
                } else if (st == kExpired) r.view.expire_count++;
                else if (st == kFiltered) r.view.filter_count++;
                if (c2 == 0) { complete = true; break; }
                it.Next();
            }
        } else {
            it.SeekForPrev(stop);
            bool first_exclusive = !stop_inclusive;
            Resp rev;
            rev.reset(0, 0);
            while ((uint32_t)count < max_kv_count && lim.valid() && it.Valid()) {
                int c2 = it.key().compare(start);
                if (c2 < 0 || (c2 == 0 && !start_inclusive)) { complete = true; break; }
                if (first_exclusive) {
                    first_exclusive = false;
                    if (it.key().compare(stop) == 0) { it.Prev(); continue; }
                }
                lim.count++;
                size_t before = rev.kvs.size();
                RangeState st = s.append_for_multi_get(rev, it.key(), it.value(), q.sort_key_filter_type,
                                                       bsv(q.sort_key_filter_pattern), now, q.no_value);
                if (st == kNormal) {
                    count++;
                    uint64_t kv_size = rev.kvs[before].key_len + rev.kvs[before].value_len;
                    size += kv_size;
                    lim.size += kv_size;
                } else if (st == kExpired) r.view.expire_count++;
                else if (st == kFiltered) r.view.filter_count++;
                if (c2 == 0) { complete = true; break; }
                it.Prev();
            }
            for (size_t i = rev.kvs.size(); i-- > 0;) {
                const pgs_kv &kv = rev.kvs[i];
                r.add(sv(rev.arena).substr(kv.key_off, kv.key_len), sv(rev.arena).substr(kv.value_off, kv.value_len), 0);
            }
        }
        r.view.iteration_count = lim.count;
        r.view.error = PGS_OK;
        if (it.Valid() && !complete) r.view.error = PGS_INCOMPLETE;
    } else {
        bool exceed_limit = false;
        for (uint32_t i = 0; i < q.n_sort_keys; i++) {
            std::string key = generate_key(hash_key, bsv(q.sort_keys[i]));
            std::string value;
            if (!s.db_get(key, &value)) continue;
            if (ts_expired(now, extract_expire_ts(s.data_version, value))) { r.view.expire_count++; continue; }
            if (count >= (int32_t)max_kv_count || size >= max_kv_size) { exceed_limit = true; break; }
            sv ud = q.no_value ? sv() : s.user_data(value);
            r.add(bsv(q.sort_keys[i]), ud, 0);
            count++;
            size += q.sort_keys[i].len + ud.size();
        }
        r.view.error = exceed_limit ? PGS_INCOMPLETE : PGS_OK;
    }
    r.seal();
    return r.view.error;
}

// on_batch_get: :906-1016
static int32_t on_batch_get(Server &s, const pgs_full_key *keys, uint32_t n, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    if (n == 0) { r.view.error = PGS_INVALID_ARGUMENT; r.seal(); return r.view.error; }
    s.prepare_read(now, false);
    for (uint32_t i = 0; i < n; i++) {
        std::string key = generate_key(bsv(keys[i].hash_key), bsv(keys[i].sort_key));
        std::string value;
        if (!s.db_get(key, &value)) continue;
        if (ts_expired(now, extract_expire_ts(s.data_version, value))) { r.view.expire_count++; continue; }
        std::string hs(bsv(keys[i].hash_key));
        hs += bsv(keys[i].sort_key);
        r.add(hs, s.user_data(value), 0);
        r.hk_len.push_back(keys[i].hash_key.len);
    }
    r.view.error = PGS_OK;
    r.seal();
    return PGS_OK;
}

// on_sortkey_count: :1018-1086
static int32_t on_sortkey_count(Server &s, sv hash_key, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    s.prepare_read(now, true);
    std::string start = generate_key(hash_key, sv()), stop = next_blob(hash_key);
    Iter it;
    it.v = s.get_view();
    it.has_upper = true;
    it.upper = stop;
    it.Seek(start, s.opt.prefix_filter);
    int64_t cnt = 0;
    while (it.Valid()) {
        r.view.iteration_count++;
        if (ts_expired(now, extract_expire_ts(s.data_version, it.value()))) r.view.expire_count++;
        else cnt++;
        it.Next();
    }
    r.view.count = cnt;
    r.view.error = PGS_OK;
    r.seal();
    return PGS_OK;
}

// the shared batch loop of on_get_scanner (:1266-1320) and on_scan (:1444-1490)
static void scan_loop(Server &s, Iter &it, sv start, sv stop, bool stop_inclusive, bool &first_exclusive,
                      uint32_t batch_count, uint32_t limiter_max, int hft, sv hpat, int sft, sv spat,
                      bool no_value, bool validate_hash, bool return_expire_ts, bool only_return_count,
                      uint32_t now, Resp &r, bool &complete, int32_t &count)
{
    Limiter lim(limiter_max, 0);
    while ((uint32_t)count < batch_count && lim.valid() && it.Valid()) {
        int c = it.key().compare(stop);
        if (c > 0 || (c == 0 && !stop_inclusive)) { complete = true; break; }
        if (first_exclusive) {
            first_exclusive = false;
            if (it.key().compare(start) == 0) { it.Next(); continue; }
        }
        lim.count++;
        RangeState st = s.validate_for_scan(it.key(), it.value(), hft, hpat, sft, spat, now, validate_hash);
        if (st == kNormal) {
            count++;
            if (!only_return_count) s.append_kv(r, it.key(), it.value(), no_value, return_expire_ts);
        } else if (st == kExpired) r.view.expire_count++;
        else if (st == kFiltered) r.view.filter_count++;
        if (c == 0) { complete = true; break; }
        it.Next();
    }
    r.view.iteration_count = lim.count;
}

// on_get_scanner: :1151-1397
static int32_t on_get_scanner(Server &s, const pgs_get_scanner_request &q, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    if (!filter_type_supported(q.hash_key_filter_type) || !filter_type_supported(q.sort_key_filter_type)) {
        r.view.error = PGS_INVALID_ARGUMENT;
        r.seal();
        return r.view.error;
    }
    s.gc_contexts(now);
    s.prepare_read(now, true);
    bool prefix_same_as_start = s.opt.prefix_filter;
    if (s.opt.prefix_filter) {
        sv hk, sk;
        sv sk_in = bsv(q.start_key);
        if (sk_in.size() >= 2) restore_key(sk_in, hk, sk);
        if (hk.empty() || q.full_scan) prefix_same_as_start = false; // total_order_seek
    }
    bool start_inclusive = q.start_inclusive, stop_inclusive = q.stop_inclusive;
    std::string start(bsv(q.start_key)), stop(bsv(q.stop_key));
    if (q.hash_key_filter_type == PGS_FT_MATCH_PREFIX && q.hash_key_filter_pattern.len > 0) {
        std::string ps = generate_key(bsv(q.hash_key_filter_pattern), sv());
        if (sv(ps).compare(start) > 0) { start = ps; start_inclusive = true; }
    }
    int c = sv(start).compare(stop);
    if (c > 0 || (c == 0 && (!start_inclusive || !stop_inclusive))) {
        r.view.error = PGS_OK;
        r.seal();
        return PGS_OK;
    }
    Iter it;
    it.v = s.get_view();
    it.Seek(start, prefix_same_as_start);
    bool complete = false, first_exclusive = !start_inclusive;
    int32_t count = 0;
    uint32_t cfg = s.opt.rocksdb_max_iteration_count ? s.opt.rocksdb_max_iteration_count : 1000;
    uint32_t batch_count = cfg;
    if (q.batch_size > 0 && (uint32_t)q.batch_size < batch_count) batch_count = q.batch_size;
    scan_loop(s, it, start, stop, stop_inclusive, first_exclusive, batch_count, cfg, q.hash_key_filter_type,
              bsv(q.hash_key_filter_pattern), q.sort_key_filter_type, bsv(q.sort_key_filter_pattern), q.no_value,
              q.validate_partition_hash, q.return_expire_ts, q.only_return_count, now, r, complete, count);
    if (q.only_return_count) r.view.kv_count = count;
    r.view.error = PGS_OK;
    if (it.Valid() && !complete) {
        auto ctx = std::make_unique<ScanContext>();
        ctx->it = it;
        ctx->stop = stop;
        ctx->stop_inclusive = q.stop_inclusive; // NB: the request's flag, as the reference (:1368)
        ctx->hash_key_filter_type = q.hash_key_filter_type;
        ctx->hash_key_filter_pattern = std::string(bsv(q.hash_key_filter_pattern));
        ctx->sort_key_filter_type = q.sort_key_filter_type;
        ctx->sort_key_filter_pattern = std::string(bsv(q.sort_key_filter_pattern));
        ctx->batch_size = (int32_t)batch_count;
        ctx->no_value = q.no_value;
        ctx->validate_partition_hash = q.validate_partition_hash;
        ctx->return_expire_ts = q.return_expire_ts;
        ctx->only_return_count = q.only_return_count;
        ctx->parked_at = now;
        int64_t handle = s.ctx_counter++;
        s.ctx[handle] = std::move(ctx);
        r.view.context_id = handle;
    } else {
        r.view.context_id = -1; // SCAN_CONTEXT_ID_COMPLETED
    }
    r.seal();
    return r.view.error;
}

// on_scan: :1399-1547
static int32_t on_scan(Server &s, int64_t context_id, uint32_t now, Resp &r)
{
    r.reset(s.app_id, s.pidx);
    s.gc_contexts(now);
    auto f = s.ctx.find(context_id);
    if (f == s.ctx.end()) { r.view.error = PGS_NOT_FOUND; r.seal(); return r.view.error; }
    std::unique_ptr<ScanContext> ctx = std::move(f->second);
    s.ctx.erase(f);
    bool complete = false, first_exclusive = false;
    int32_t count = 0;
    uint32_t cfg = s.opt.rocksdb_max_iteration_count ? s.opt.rocksdb_max_iteration_count : 1000;
    uint32_t batch_count = cfg;
    if (ctx->batch_size > 0 && (uint32_t)ctx->batch_size < batch_count) batch_count = ctx->batch_size;
    scan_loop(s, ctx->it, sv(), ctx->stop, ctx->stop_inclusive, first_exclusive, batch_count, batch_count,
              ctx->hash_key_filter_type, ctx->hash_key_filter_pattern, ctx->sort_key_filter_type,
              ctx->sort_key_filter_pattern, ctx->no_value, ctx->validate_partition_hash, ctx->return_expire_ts,
              ctx->only_return_count, now, r, complete, count);
    if (ctx->only_return_count) r.view.kv_count = count;
    r.view.error = PGS_OK;
    if (ctx->it.Valid() && !complete) {
        ctx->parked_at = now;
        int64_t handle = s.ctx_counter++;
        s.ctx[handle] = std::move(ctx);
        r.view.context_id = handle;
    } else {
        r.view.context_id = -1;
    }
    r.seal();
    return r.view.error;
}

static void parse_envs(const char *envs, uint32_t n, std::vector<std::pair<std::string, std::string>> &out)
{
    const char *p = envs;
    for (uint32_t i = 0; i < n; i++) {
        std::string k(p);
        p += k.size() + 1;
        std::string v(p);
        p += v.size() + 1;
        out.emplace_back(std::move(k), std::move(v));
    }
}

} // namespace orc

using namespace orc;
struct orc_server { Server s; };
static inline Resp &R(pgs_response_buf *r) { return *reinterpret_cast<Resp *>(r); }
static inline sv bsv2(pgs_blob b) { return sv((const char *)b.data, b.len); }

extern "C" {

pgs_response_buf *orc_response_new(void) { return reinterpret_cast<pgs_response_buf *>(new Resp); }
void orc_response_free(pgs_response_buf *r) { delete reinterpret_cast<Resp *>(r); }
const pgs_response *orc_response_view(pgs_response_buf *r) { return &R(r).view; }

int32_t orc_rrdb_update_app_envs(orc_server *h, const char *envs, uint32_t n_envs, uint32_t now)
{
    Server &s = h->s;
    std::vector<std::pair<std::string, std::string>> kv;
    parse_envs(envs, n_envs, kv);
    // update_app_envs hands over the table's whole env map (pegasus_server_impl.cpp:2728-2741): an absent key means "deleted"
    {
        std::map<std::string, std::string> em(kv.begin(), kv.end());
        auto fd = em.find("default_ttl"); // update_default_ttl :2814-2826: buf2int32 and >= 0, otherwise the old value stays
        if (fd != em.end()) {
            char *endp = nullptr;
            errno = 0;
            const long long v = strtoll(fd->second.c_str(), &endp, 10);
            if (!fd->second.empty() && !*endp && errno == 0 && v >= 0 && v <= INT32_MAX) s.default_ttl = (uint32_t)v;
        }
        auto fv = em.find("replica.split.validate_partition_hash"); // :2966-2983: absent -> false, unparsable -> unchanged (buf2bool)
        if (fv == em.end()) s.validate_partition_hash = false;
        else {
            std::string v = fv->second;
            for (auto &c : v) c = (char)tolower((unsigned char)c);
            if (v == "true") s.validate_partition_hash = true;
            else if (v == "false") s.validate_partition_hash = false;
        }
        auto fo = em.find("user_specified_compaction"); // :2985-3001: absent -> cleared
        if (fo == em.end()) s.ops.clear();
        else s.ops = fo->second.empty() ? std::vector<Op>() : ops_from_json(fo->second, s.data_version);
    }
    // pegasus_manual_compact_service.cpp:83-121 (order of the checks), :122-166 (disabled, running-count limit), :168-184 (`once`),
    // :186-219 (`periodic`: a "H:M" of today's local day between the last finish and now), :231-272 (options of the rule that fired)
    std::map<std::string, std::string> m(kv.begin(), kv.end());
    auto env = [&](const std::string &k) -> const std::string * { auto f = m.find(k); return f == m.end() ? nullptr : &f->second; };
    auto as_int = [](const std::string *v, long long &out) { // dsn::buf2int*: one whole strtoll(base 0) number
        if (!v || v->empty()) return false;
        char *e = nullptr;
        errno = 0;
        out = strtoll(v->c_str(), &e, 0);
        return *e == 0 && errno == 0;
    };
    if (const std::string *d = env("manual_compact.disabled"); d && *d == "true") return PGS_OK;
    long long num = 0;
    if (as_int(env("manual_compact.max_concurrent_running_count"), num) && num >= INT32_MIN && num <= INT32_MAX && num <= 0) return PGS_OK;
    const uint64_t last_ms = s.manual_compact_last_finish_ms, now_ms = ((uint64_t)now + kEpochBegin) * 1000;
    const char *rule = nullptr;
    if (as_int(env("manual_compact.once.trigger_time"), num) && num > 0 && (uint64_t)num > last_ms / 1000) rule = "manual_compact.once.";
    if (!rule && now != 0) {
        if (const std::string *times = env("manual_compact.periodic.trigger_time")) {
            time_t tt = (time_t)(now_ms / 1000);
            struct tm day;
            localtime_r(&tt, &day);
            day.tm_hour = 0; day.tm_min = 0; day.tm_sec = 0;
            const long long midnight = (long long)mktime(&day);
            std::stringstream ss(*times);
            std::string item;
            while (std::getline(ss, item, ',')) {
                int hh = 0, mm = 0;
                if (sscanf(item.c_str(), "%d:%d", &hh, &mm) != 2 || hh < 0 || hh > 23 || mm < 0 || mm > 59) continue;
                const uint64_t at_ms = (uint64_t)(midnight + hh * 3600 + mm * 60) * 1000;
                if (last_ms < at_ms && at_ms < now_ms) { rule = "manual_compact.periodic."; break; }
            }
        }
    }
    if (!rule) return PGS_OK;
    int target_level = -1;
    if (as_int(env(std::string(rule) + "target_level"), num) && (num == -1 || (num >= 1 && num <= 6))) target_level = (int)num;
    const std::string *bl = env(std::string(rule) + "bottommost_level_compaction");
    const bool force = bl && *bl == "force";
    s.flush_mem();
    if (!s.runs.empty() && (force || !(s.runs.size() == 1 && s.runs[0].level >= 1))) {
        int level = 1;
        for (auto &lr : s.runs) level = std::max(level, lr.level);
        if (target_level >= 1) level = target_level;
        s.compact_runs(0, s.runs.size(), level, now, nullptr);
    }
    s.manual_compact_last_finish_ms = ((uint64_t)now + kEpochBegin) * 1000;
    return PGS_OK;
}

orc_server *orc_rrdb_start(int32_t app_id, int32_t pidx, const pgs_server_options *opt,
                           const char *envs, uint32_t n_envs)
{
    auto *h = new orc_server;
    h->s.app_id = app_id;
    h->s.pidx = pidx;
    h->s.opt = opt ? *opt : pgs_server_options{};
    if (!opt) h->s.opt.prefix_filter = 1;
    if (!h->s.opt.cluster_id) h->s.opt.cluster_id = 1;
    if (envs && n_envs) orc_rrdb_update_app_envs(h, envs, n_envs, 0);
    return h;
}
void orc_rrdb_stop(orc_server *s) { delete s; }
void orc_rrdb_set_partition_version(orc_server *s, int32_t pv) { s->s.partition_version = pv; }

int32_t orc_rrdb_get(orc_server *s, pgs_blob key, uint32_t now, pgs_response_buf *r) { return on_get(s->s, bsv2(key), now, R(r)); }
int32_t orc_rrdb_ttl(orc_server *s, pgs_blob key, uint32_t now, pgs_response_buf *r) { return on_ttl(s->s, bsv2(key), now, R(r)); }
int32_t orc_rrdb_multi_get(orc_server *s, const pgs_multi_get_request *q, uint32_t now, pgs_response_buf *r)
{
    return on_multi_get(s->s, *q, now, R(r));
}
int32_t orc_rrdb_batch_get(orc_server *s, const pgs_full_key *keys, uint32_t n, uint32_t now, pgs_response_buf *r)
{
    return on_batch_get(s->s, keys, n, now, R(r));
}
int32_t orc_rrdb_sortkey_count(orc_server *s, pgs_blob hk, uint32_t now, pgs_response_buf *r)
{
    return on_sortkey_count(s->s, bsv2(hk), now, R(r));
}
int32_t orc_rrdb_get_scanner(orc_server *s, const pgs_get_scanner_request *q, uint32_t now, pgs_response_buf *r)
{
    return on_get_scanner(s->s, *q, now, R(r));
}
int32_t orc_rrdb_scan(orc_server *s, int64_t context_id, uint32_t now, pgs_response_buf *r)
{
    return on_scan(s->s, context_id, now, R(r));
}
void orc_rrdb_clear_scanner(orc_server *s, int64_t context_id) { s->s.ctx.erase(context_id); }
uint32_t orc_rrdb_gc(orc_server *s, uint32_t now) { return s->s.gc_contexts(now); }

// write_batch_put_ctx: rocksdb_wrapper.cpp:129-183 (local write, no timetag verification)
static void put_one(Server &s, sv raw_key, sv user_value, uint32_t expire_ts, uint64_t timestamp_us, uint32_t now)
{
    uint64_t timetag = timestamp_us << 8u | (uint64_t)(s.opt.cluster_id << 1u);
    if (s.default_ttl != 0 && expire_ts == 0) expire_ts = now + s.default_ttl; // db_expire_ts :280-288
    Rec r;
    r.ukey = std::string(raw_key);
    r.seq = ++s.last_seq;
    r.type = PGS_TYPE_VALUE;
    r.value = generate_value(s.data_version, expire_ts, timetag, user_value);
    s.write(std::move(r), now);
}
static void del_one(Server &s, sv raw_key, uint32_t now)
{
    Rec r;
    r.ukey = std::string(raw_key);
    r.seq = ++s.last_seq;
    r.type = PGS_TYPE_DELETION;
    s.write(std::move(r), now);
}

int32_t orc_rrdb_put(orc_server *h, pgs_blob key, pgs_blob value, uint32_t expire_ts, int64_t decree,
                     uint64_t timestamp_us, uint32_t now)
{
    h->s.last_committed_decree = decree;
    put_one(h->s, bsv2(key), bsv2(value), expire_ts, timestamp_us, now);
    return PGS_OK;
}
// on_batched_write_requests: src/server/pegasus_server_write.cpp:92-222 (puts and removes of one decree; count 0 = empty write)
int32_t orc_rrdb_on_batched_writes(orc_server *h, const pgs_write_request *reqs, uint32_t count, int64_t decree, uint64_t timestamp_us,
                                   uint32_t now, int32_t *resp_errors)
{
    for (uint32_t i = 0; i < count; i++)
        if (reqs[i].op > 1) return PGS_INVALID_ARGUMENT;
    Server &s = h->s;
    s.last_committed_decree = decree;
    if (count == 0) { put_one(s, sv(), sv(), 0, timestamp_us, now); return PGS_OK; }
    for (uint32_t i = 0; i < count; i++) {
        if (reqs[i].op == 0) put_one(s, bsv2(reqs[i].raw_key), bsv2(reqs[i].value), reqs[i].expire_ts_seconds, timestamp_us, now);
        else del_one(s, bsv2(reqs[i].raw_key), now);
        if (resp_errors) resp_errors[i] = PGS_OK;
    }
    return PGS_OK;
}
// incr: pegasus_write_service_impl.h:264-342 (buf2int64: src/utils/string_conv.h:35-62; safe_add: src/utils/safe_arithmetic.h:44, int64 overflow check)
int32_t orc_rrdb_incr(orc_server *h, pgs_blob key, int64_t increment, int32_t expire_ts_seconds, int64_t decree, uint64_t timestamp_us,
                      uint32_t now, int32_t *resp_error, int64_t *new_value)
{
    Server &s = h->s;
    s.last_committed_decree = decree;
    *new_value = 0;
    s.prepare_read(now, false);
    std::string value;
    const bool found = s.db_get(bsv2(key), &value);
    const uint32_t old_ets = found ? extract_expire_ts(s.data_version, value) : 0;
    int64_t nv = increment;
    uint32_t new_ets = expire_ts_seconds > 0 ? (uint32_t)expire_ts_seconds : 0u;
    if (found && !ts_expired(now, old_ets)) {
        const std::string old(s.user_data(value));
        if (!old.empty()) {
            errno = 0;
            char *p = nullptr;
            const long long base = std::strtoll(old.c_str(), &p, 0);
            if ((size_t)(p - old.c_str()) != old.size() || errno != 0) {
                *resp_error = PGS_INVALID_ARGUMENT;
                put_one(s, sv(), sv(), 0, timestamp_us, now); // empty_put: the decree still advances
                return PGS_OK;
            }
            long long sum;
            if (__builtin_add_overflow(base, (long long)increment, &sum)) {
                *resp_error = PGS_INVALID_ARGUMENT;
                *new_value = base;
                put_one(s, sv(), sv(), 0, timestamp_us, now);
                return PGS_OK;
            }
            nv = sum;
        }
        new_ets = expire_ts_seconds == 0 ? old_ets : expire_ts_seconds < 0 ? 0u : (uint32_t)expire_ts_seconds;
    }
    put_one(s, bsv2(key), std::to_string(nv), new_ets, timestamp_us, now);
    *resp_error = PGS_OK;
    *new_value = nv;
    return PGS_OK;
}
static bool orc_buf2int64(sv buf, int64_t &out) // dsn::buf2int64, src/utils/string_conv.h:35-62
{
    if (buf.empty()) return false;
    const std::string str(buf);
    errno = 0;
    char *p = nullptr;
    const long long v = std::strtoll(str.c_str(), &p, 0);
    if ((size_t)(p - str.c_str()) != str.size() || errno != 0) return false;
    out = v;
    return true;
}

// validate_check (pegasus_write_service_impl.h:1144-1270)
static bool cas_validate(int32_t type, sv operand, bool exist, sv value, bool &invalid)
{
    invalid = false;
    switch (type) {
    case 0: return true;                              // CT_NO_CHECK
    case 1: return !exist;                            // CT_VALUE_NOT_EXIST
    case 2: return !exist || value.empty();           // CT_VALUE_NOT_EXIST_OR_EMPTY
    case 3: return exist;                             // CT_VALUE_EXIST
    case 4: return exist && !value.empty();           // CT_VALUE_NOT_EMPTY
    case 5: case 6: case 7:                           // CT_VALUE_MATCH_ANYWHERE / PREFIX / POSTFIX
        if (!exist) return false;
        if (operand.empty()) return true;
        if (value.size() < operand.size()) return false;
        if (type == 5) return value.find(operand) != sv::npos;
        if (type == 6) return value.substr(0, operand.size()) == operand;
        return value.substr(value.size() - operand.size()) == operand;
    case 8: case 9: case 10: case 11: case 12: {      // CT_VALUE_BYTES_LESS .. GREATER
        if (!exist) return false;
        const int c = value.compare(operand);
        if (c < 0) return type <= 9;
        if (c > 0) return type >= 11;
        return type >= 9 && type <= 11;
    }
    case 13: case 14: case 15: case 16: case 17: {    // CT_VALUE_INT_LESS .. GREATER
        if (!exist) return false;
        int64_t a = 0, b = 0;
        if (!orc_buf2int64(value, a) || !orc_buf2int64(operand, b)) { invalid = true; return false; }
        if (a < b) return type <= 14;
        if (a > b) return type >= 16;
        return type >= 14 && type <= 16;
    }
    }
    return false;
}

// check_and_mutate: pegasus_write_service_impl.h:710-840 (check_and_set :436-530 is the one-put case)
int32_t orc_rrdb_check_and_mutate(orc_server *h, const pgs_check_and_mutate_request *q, int64_t decree, uint64_t timestamp_us, uint32_t now,
                                  pgs_cas_result *res, uint8_t *cv_out, uint32_t cv_cap)
{
    Server &s = h->s;
    s.last_committed_decree = decree;
    *res = pgs_cas_result{};
    bool bad = q->n_mutate == 0 || q->check_type < 0 || q->check_type > 17;
    for (uint32_t i = 0; i < q->n_mutate && !bad; i++) bad = q->mutate_list[i].operation > 1;
    if (bad) {
        res->error = PGS_INVALID_ARGUMENT;
        put_one(s, sv(), sv(), 0, timestamp_us, now);
        return PGS_OK;
    }
    s.prepare_read(now, false);
    std::string raw;
    const std::string ck = generate_key(bsv2(q->hash_key), bsv2(q->check_sort_key));
    bool exist = s.db_get(ck, &raw);
    if (exist && ts_expired(now, extract_expire_ts(s.data_version, raw))) exist = false;
    const std::string value = exist ? std::string(s.user_data(raw)) : std::string();
    if (q->return_check_value) {
        res->check_value_returned = 1;
        if (exist) {
            res->check_value_exist = 1;
            res->check_value_len = (uint32_t)value.size();
            if (cv_out && cv_cap) memcpy(cv_out, value.data(), std::min<size_t>(cv_cap, value.size()));
        }
    }
    bool invalid = false;
    const bool passed = cas_validate(q->check_type, bsv2(q->check_operand), exist, value, invalid);
    if (passed) {
        for (uint32_t i = 0; i < q->n_mutate; i++) {
            const pgs_mutate &m = q->mutate_list[i];
            const std::string key = generate_key(bsv2(q->hash_key), bsv2(m.sort_key));
            if (m.operation == 0) put_one(s, key, bsv2(m.value), (uint32_t)m.set_expire_ts_seconds, timestamp_us, now);
            else del_one(s, key, now);
        }
        res->error = PGS_OK;
    } else {
        put_one(s, sv(), sv(), 0, timestamp_us, now);
        res->error = invalid ? PGS_INVALID_ARGUMENT : PGS_TRY_AGAIN;
    }
    return PGS_OK;
}
int32_t orc_rrdb_check_and_set(orc_server *h, const pgs_check_and_set_request *q, int64_t decree, uint64_t timestamp_us, uint32_t now,
                               pgs_cas_result *res, uint8_t *cv_out, uint32_t cv_cap)
{
    pgs_mutate m{0, q->set_diff_sort_key ? q->set_sort_key : q->check_sort_key, q->set_value, q->set_expire_ts_seconds};
    pgs_check_and_mutate_request r{q->hash_key, q->check_sort_key, q->check_type, q->check_operand, &m, 1, q->return_check_value};
    return orc_rrdb_check_and_mutate(h, &r, decree, timestamp_us, now, res, cv_out, cv_cap);
}
int32_t orc_rrdb_remove(orc_server *h, pgs_blob key, int64_t decree, uint32_t now)
{
    h->s.last_committed_decree = decree;
    del_one(h->s, bsv2(key), now);
    return PGS_OK;
}
int32_t orc_rrdb_multi_put(orc_server *h, pgs_blob hash_key, const pgs_blob *sort_keys, const pgs_blob *values,
                           uint32_t n, uint32_t expire_ts, int64_t decree, uint64_t timestamp_us, uint32_t now)
{
    Server &s = h->s;
    s.last_committed_decree = decree;
    if (n == 0) { // pegasus_write_service_impl.h:112-119: empty_put + kInvalidArgument
        put_one(s, sv(), sv(), 0, timestamp_us, now);
        return PGS_INVALID_ARGUMENT;
    }
    for (uint32_t i = 0; i < n; i++)
        put_one(s, generate_key(bsv2(hash_key), bsv2(sort_keys[i])), bsv2(values[i]), expire_ts, timestamp_us, now);
    return PGS_OK;
}
int32_t orc_rrdb_multi_remove(orc_server *h, pgs_blob hash_key, const pgs_blob *sort_keys, uint32_t n,
                              int64_t decree, int64_t *count, uint32_t now)
{
    Server &s = h->s;
    s.last_committed_decree = decree;
    if (count) *count = 0;
    if (n == 0) {
        put_one(s, sv(), sv(), 0, 0, now);
        return PGS_INVALID_ARGUMENT;
    }
    for (uint32_t i = 0; i < n; i++) del_one(s, generate_key(bsv2(hash_key), bsv2(sort_keys[i])), now);
    if (count) *count = n;
    return PGS_OK;
}
int32_t orc_rrdb_flush(orc_server *h, uint32_t now)
{
    h->s.flush_mem();
    h->s.maybe_compact(now);
    return PGS_OK;
}
int32_t orc_rrdb_manual_compact(orc_server *h, uint32_t now, orc_compact_stats *st)
{
    Server &s = h->s;
    s.flush_mem();
    orc_compact_stats local{};
    if (!s.runs.empty()) {
        int level = 1;
        for (auto &lr : s.runs) level = std::max(level, lr.level);
        s.compact_runs(0, s.runs.size(), level, now, &local);
    }
    s.manual_compact_last_finish_ms = ((uint64_t)now + kEpochBegin) * 1000;
    if (st) *st = local;
    return PGS_OK;
}
int64_t orc_rrdb_last_flushed_decree(orc_server *h) { return h->s.last_flushed_decree; }
int64_t orc_rrdb_last_committed_decree(orc_server *h) { return h->s.last_committed_decree; }
uint32_t orc_rrdb_run_count(orc_server *h) { return (uint32_t)h->s.runs.size(); }
orc_run *orc_rrdb_dump(orc_server *h)
{
    Server &s = h->s;
    s.flush_mem();
    std::vector<const Run *> in;
    for (auto &lr : s.runs) in.push_back(&lr.run);
    FilterParams nofilter;
    auto *out = new orc_run;
    out->run = compact(in, false, nofilter, 0, nullptr);
    return out;
}

} // extern "C"
