// orc_internal.h — shared declarations of the CPU oracle (TEST INFRASTRUCTURE ONLY, see
// pegasus_oracle.h).  Plain C++17, no dependencies.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#include "pegasus_oracle.h"

namespace orc {

using sv = std::string_view;

// ---- endian (src/utils/endians.h:68-156: big-endian fixed ints) ----
inline uint16_t be16(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }
inline uint32_t be32(const uint8_t *p)
{
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
inline uint64_t be64(const uint8_t *p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
inline void put_be16(uint8_t *p, uint16_t v) { p[0] = v >> 8; p[1] = (uint8_t)v; }
inline void put_be32(uint8_t *p, uint32_t v)
{
    p[0] = v >> 24; p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
inline void put_be64(uint8_t *p, uint64_t v) { put_be32(p, (uint32_t)(v >> 32)); put_be32(p + 4, (uint32_t)v); }

constexpr uint32_t kEpochBegin = 1451606400u; // src/base/pegasus_utils.h:39

uint64_t crc64(const void *p, size_t n, uint64_t init);

// ---- key schema (src/base/pegasus_key_schema.h) ----
std::string generate_key(sv hk, sv sk);
std::string next_blob(sv hk);
std::string next_blob(sv hk, sv sk);
void restore_key(sv key, sv &hk, sv &sk);
uint64_t key_hash(sv key);
bool check_key_hash(sv key, int32_t pidx, int32_t pv);
inline sv hashkey_prefix(sv key) // HashkeyTransform::Transform (hashkey_transform.h:40-54)
{
    if (key.size() < 2) return key;
    return key.substr(0, 2 + be16((const uint8_t *)key.data()));
}

// ---- value schema (src/base/pegasus_value_schema.h) ----
inline uint32_t extract_expire_ts(uint32_t /*version*/, sv v) { return be32((const uint8_t *)v.data()); }
inline size_t user_data_offset(uint32_t version) { return version == 1 ? 12 : 4; }
inline bool ts_expired(uint32_t now, uint32_t ts) { return ts > 0 && ts <= now; }
std::string generate_value(uint32_t version, uint32_t expire_ts, uint64_t timetag, sv data);

// ---- tiny JSON (subset the envs use) ----
struct JVal {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    bool is_int = false, neg = false;
    uint64_t u = 0; // magnitude when is_int
    std::string s;
    std::vector<JVal> a;
    std::vector<std::pair<std::string, JVal>> o;
    const JVal *get(const char *k) const
    {
        for (auto &kv : o)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
bool json_parse(sv text, JVal &out);

// ---- rules / ops (compaction_filter_rule.{h,cpp}, compaction_operation.{h,cpp}) ----
enum { FRT_HASHKEY_PATTERN = 0, FRT_SORTKEY_PATTERN, FRT_TTL_RANGE, FRT_INVALID };
enum { SMT_MATCH_ANYWHERE = 0, SMT_MATCH_PREFIX, SMT_MATCH_POSTFIX, SMT_INVALID };
enum { COT_UPDATE_TTL = 0, COT_DELETE, COT_INVALID };
enum { UTOT_FROM_NOW = 0, UTOT_FROM_CURRENT, UTOT_TIMESTAMP, UTOT_INVALID };

struct Rule {
    int type = FRT_INVALID;
    std::string pattern;
    int match_type = SMT_INVALID;
    uint32_t start_ttl = 0, stop_ttl = 0;
    uint32_t data_version = 1;
    bool match(sv hk, sv sk, sv value, uint32_t now) const;
};
struct Op {
    int type = COT_INVALID;
    int ttl_type = UTOT_INVALID;
    uint32_t ttl_value = 0;
    uint32_t data_version = 1;
    std::vector<Rule> rules;
    bool all_rules_match(sv hk, sv sk, sv value, uint32_t now) const;
    // returns true = delete.  new_value/value_changed as compaction_operation::filter
    bool filter(sv hk, sv sk, sv value, uint32_t now, std::string *new_value, bool *changed) const;
};
bool string_pattern_match(sv value, int match_type, sv pattern);
bool validate_filter(int filter_type, sv pattern, sv value);
bool rule_from_json(int type, sv params, uint32_t data_version, Rule &out);
bool update_ttl_from_json(sv params, Op &out);
std::vector<Op> ops_from_json(sv json, uint32_t data_version);

struct FilterParams {
    bool enabled = false, validate_hash = false;
    uint32_t data_version = 1, default_ttl = 0;
    int32_t pidx = 0, partition_version = -1;
    const std::vector<Op> *ops = nullptr;
};
enum DropReason { kKeep = 0, kDropExpired, kDropUser, kDropStale };
// KeyWithTTLCompactionFilter::Filter; returns reason != kKeep when the record must be removed
DropReason compaction_filter(const FilterParams &fp, sv key, sv value, uint32_t now,
                             std::string *new_value, bool *changed);

// ---- LSM model ----
struct Rec {
    std::string ukey;
    uint64_t seq = 0;
    uint8_t type = 1;
    std::string value;
};
inline int cmp_internal(sv ak, uint64_t aseq, uint8_t atype, sv bk, uint64_t bseq, uint8_t btype)
{
    int c = ak.compare(bk);
    if (c) return c < 0 ? -1 : 1;
    uint64_t at = (aseq << 8) | atype, bt = (bseq << 8) | btype;
    if (at > bt) return -1; // larger seq sorts first
    if (at < bt) return 1;
    return 0;
}
struct Run { std::vector<Rec> recs; };

struct CompactStats : orc_compact_stats {};
Run compact(const std::vector<const Run *> &runs, bool bottommost, const FilterParams &fp,
            uint32_t now, orc_compact_stats *st);

} // namespace orc

struct orc_ops { std::vector<orc::Op> ops; };
struct orc_run { orc::Run run; };
