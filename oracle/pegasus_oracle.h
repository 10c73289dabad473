/*
 * pegasus_oracle.h — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A restatement, on the CPU, of the reference algorithms on the LSM read/compaction hot path of
 * apache/incubator-pegasus.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product (libpegasus_b200.so)
 * never links or calls it.
 *
 * Pinning: the Pegasus-side pieces (key/value schema, compaction rules/ops, range semantics) are
 * pinned against the golden tables of the reference's own tests (tests/golden/ json files, transcribed
 * from the _test.cpp files of src/server/test, src/base/test/value_schema_test.cpp,
 * src/test/function_test/base_api); crc64 is pinned against the reference's own
 * src/utils/crc.cpp compiled into oracle/_ref/ (see oracle/Makefile).  The LSM engine semantics
 * (newest-seqno-wins, tombstones, bottommost drop, data-block encoding) live in RocksDB v8.5.3
 * (thirdparty/CMakeLists.txt:507-521), which is NOT in the reference tree nor in this image:
 * for that part the oracle restates RocksDB's published format/behaviour and is
 * "parity unpinned" against a real RocksDB binary.
 *
 * It shares the plain request/response struct typedefs of include/pegasus_b200.h (types only).
 */
#ifndef PEGASUS_ORACLE_H_
#define PEGASUS_ORACLE_H_

#include "../include/pegasus_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- crc / key schema / value schema ---------------------------------------------------- */
ORC_API uint64_t orc_crc64(const uint8_t *p, uint64_t n, uint64_t init);
ORC_API int32_t orc_generate_key(const uint8_t *hk, uint32_t hk_len, const uint8_t *sk,
                                 uint32_t sk_len, uint8_t *out, uint32_t cap);
ORC_API int32_t orc_generate_next_blob(const uint8_t *hk, uint32_t hk_len, const uint8_t *sk,
                                       uint32_t sk_len, int32_t with_sort_key, uint8_t *out,
                                       uint32_t cap);
ORC_API int32_t orc_restore_key(const uint8_t *key, uint32_t len, uint32_t *hk_len,
                                uint32_t *sk_len);
ORC_API uint64_t orc_key_hash(const uint8_t *key, uint32_t len);
ORC_API int32_t orc_check_key_hash(const uint8_t *key, uint32_t len, int32_t pidx,
                                   int32_t partition_version);
ORC_API int32_t orc_hashkey_transform(const uint8_t *key, uint32_t len); /* prefix length, -1 = not in domain */
ORC_API uint64_t orc_generate_timetag(uint64_t timestamp, uint8_t cluster_id, int32_t deleted);
ORC_API int32_t orc_generate_value(uint32_t version, uint32_t expire_ts, uint64_t timetag,
                                   const uint8_t *data, uint32_t len, uint8_t *out, uint32_t cap);
ORC_API uint32_t orc_extract_expire_ts(uint32_t version, const uint8_t *v, uint32_t len);
ORC_API uint64_t orc_extract_timetag(uint32_t version, const uint8_t *v, uint32_t len);
ORC_API int32_t orc_user_data_offset(uint32_t version);
ORC_API void orc_update_expire_ts(uint32_t version, uint8_t *v, uint32_t len, uint32_t ts);
ORC_API int32_t orc_check_if_ts_expired(uint32_t now, uint32_t expire_ts);

/* ---- compaction rules / operations / filter ---------------------------------------------- */
ORC_API int32_t orc_string_pattern_match(const uint8_t *v, uint32_t vlen, int32_t match_type,
                                         const uint8_t *pat, uint32_t plen);
ORC_API int32_t orc_validate_filter(int32_t filter_type, const uint8_t *pat, uint32_t plen,
                                    const uint8_t *v, uint32_t vlen);
ORC_API int32_t orc_ttl_range_rule_match(uint32_t start_ttl, uint32_t stop_ttl, uint32_t expire_ts,
                                         uint32_t now);
/* rule JSON ("params" of one rule): returns 1 ok / 0 nullptr; outputs decoded fields */
ORC_API int32_t orc_rule_create(int32_t rule_type, const char *params, uint32_t len,
                                char *pattern_out, uint32_t cap, int32_t *match_type,
                                uint32_t *start_ttl, uint32_t *stop_ttl);
ORC_API int32_t orc_update_ttl_create(const char *params, uint32_t len, int32_t *type,
                                      uint32_t *value);

typedef struct orc_ops orc_ops;
ORC_API orc_ops *orc_ops_create(const char *json, uint32_t len, uint32_t data_version);
ORC_API void orc_ops_free(orc_ops *o);
ORC_API uint32_t orc_ops_count(const orc_ops *o);
/* describe op i: op type (0 UPDATE_TTL,1 DELETE), ttl type/value, number of rules */
ORC_API int32_t orc_ops_describe(const orc_ops *o, uint32_t i, int32_t *op_type, int32_t *ttl_type,
                                 uint32_t *ttl_value, uint32_t *n_rules);
ORC_API int32_t orc_ops_describe_rule(const orc_ops *o, uint32_t i, uint32_t r, int32_t *rule_type,
                                      int32_t *match_type, char *pattern, uint32_t cap,
                                      uint32_t *start_ttl, uint32_t *stop_ttl);
/* one op built by hand for table tests: rules given as parallel arrays */
ORC_API orc_ops *orc_ops_build(int32_t op_type, int32_t ttl_type, uint32_t ttl_value,
                               uint32_t n_rules, const int32_t *rule_type,
                               const int32_t *match_type, const char *const *pattern,
                               const uint32_t *start_ttl, const uint32_t *stop_ttl,
                               uint32_t data_version);
ORC_API int32_t orc_op_all_rules_match(const orc_ops *o, uint32_t i, const uint8_t *hk,
                                       uint32_t hk_len, const uint8_t *sk, uint32_t sk_len,
                                       const uint8_t *v, uint32_t vlen, uint32_t now);
/* compaction_operation::filter for op i; new_value must hold vlen bytes */
ORC_API int32_t orc_op_filter(const orc_ops *o, uint32_t i, const uint8_t *hk, uint32_t hk_len,
                              const uint8_t *sk, uint32_t sk_len, const uint8_t *v, uint32_t vlen,
                              uint32_t now, uint8_t *new_value, int32_t *value_changed);

typedef struct {
    uint8_t enabled, validate_hash;
    uint32_t data_version, default_ttl;
    int32_t pidx, partition_version;
    const orc_ops *ops;
} orc_filter_params;
/* KeyWithTTLCompactionFilter::Filter: returns 1 = remove. new_value: vlen bytes. */
ORC_API int32_t orc_filter(const orc_filter_params *fp, const uint8_t *key, uint32_t klen,
                           const uint8_t *v, uint32_t vlen, uint32_t now, uint8_t *new_value,
                           int32_t *value_changed);

/* ---- flat record sets -------------------------------------------------------------------- */
typedef struct orc_run orc_run; /* sorted run of (user key, seq, type, value) */
ORC_API orc_run *orc_run_from_records(uint64_t n, const uint8_t *keys, const uint64_t *key_off,
                                      const uint8_t *vals, const uint64_t *val_off,
                                      const uint64_t *seq, const uint8_t *type);
ORC_API void orc_run_free(orc_run *r);
ORC_API void orc_run_sizes(const orc_run *r, pgs_decode_sizes *out);
ORC_API void orc_run_export(const orc_run *r, uint8_t *keys, uint64_t *key_off, uint8_t *vals,
                            uint64_t *val_off, uint64_t *seq, uint8_t *type);
/* decode RocksDB-format data blocks (independent decoder) */
ORC_API orc_run *orc_run_from_blocks(const uint8_t *data, const uint64_t *blk_off,
                                     const uint32_t *blk_size, uint32_t n_blocks, int32_t *status);

typedef struct {
    uint64_t in_records, out_records, in_bytes, out_bytes;
    uint64_t dropped_shadowed, dropped_tombstone, dropped_expired, dropped_user, dropped_stale;
    uint64_t ttl_rewritten;
} orc_compact_stats;
/* semantic compaction: merge k runs, newest wins, filter, tombstone rules */
ORC_API orc_run *orc_compact(const orc_run *const *runs, uint32_t k, int32_t bottommost,
                             const orc_filter_params *fp, uint32_t now, orc_compact_stats *st);

/* ---- CPU baseline: block-level compaction the way the RocksDB CPU path does it ------------
 * inputs are block-encoded runs; a min-heap MergingIterator over BlockIters feeds the
 * compaction loop which calls the filter and a BlockBuilder; `threads` sub-compactions split
 * the key space. */
typedef struct orc_blockrun orc_blockrun;
ORC_API orc_blockrun *orc_blockrun_build(const orc_run *r, uint32_t block_size,
                                         uint32_t restart_interval);
ORC_API orc_blockrun *orc_blockrun_from_blocks(const uint8_t *data, uint64_t data_bytes,
                                               const uint64_t *blk_off, const uint32_t *blk_size,
                                               uint32_t n_blocks);
ORC_API void orc_blockrun_free(orc_blockrun *b);
ORC_API uint64_t orc_blockrun_bytes(const orc_blockrun *b);
ORC_API uint32_t orc_blockrun_blocks(const orc_blockrun *b);
ORC_API orc_run *orc_blockrun_decode(const orc_blockrun *b);
ORC_API orc_blockrun *orc_compact_blocks(const orc_blockrun *const *runs, uint32_t k,
                                         int32_t bottommost, const orc_filter_params *fp,
                                         uint32_t now, uint32_t threads, uint32_t block_size,
                                         uint32_t restart_interval, orc_compact_stats *st,
                                         double *seconds);
/* CPU baseline for reads: point gets / prefix scans over block runs (index bsearch + block
 * decode), `threads` workers; returns found / returned-record counts */
ORC_API uint64_t orc_blockruns_get_many(const orc_blockrun *const *runs, uint32_t k,
                                        const uint8_t *keys, const uint32_t *key_off, uint32_t n,
                                        uint32_t now, uint32_t threads, uint64_t *value_bytes,
                                        double *seconds);
ORC_API uint64_t orc_blockruns_prefix_scan_many(const orc_blockrun *const *runs, uint32_t k,
                                                const uint8_t *hashkeys, const uint32_t *hk_off,
                                                uint32_t n, uint32_t now, uint32_t threads,
                                                uint64_t *bytes, double *seconds);

/* ---- rrdb surface on the semantic LSM model ------------------------------------------------ */
typedef struct orc_server orc_server;
ORC_API orc_server *orc_rrdb_start(int32_t app_id, int32_t pidx, const pgs_server_options *opt,
                                   const char *envs, uint32_t n_envs);
ORC_API void orc_rrdb_stop(orc_server *s);
ORC_API int32_t orc_rrdb_update_app_envs(orc_server *s, const char *envs, uint32_t n_envs,
                                         uint32_t now);
ORC_API void orc_rrdb_set_partition_version(orc_server *s, int32_t pv);
ORC_API pgs_response_buf *orc_response_new(void);
ORC_API void orc_response_free(pgs_response_buf *r);
ORC_API const pgs_response *orc_response_view(pgs_response_buf *r);
ORC_API int32_t orc_rrdb_get(orc_server *s, pgs_blob key, uint32_t now, pgs_response_buf *r);
ORC_API int32_t orc_rrdb_ttl(orc_server *s, pgs_blob key, uint32_t now, pgs_response_buf *r);
ORC_API int32_t orc_rrdb_multi_get(orc_server *s, const pgs_multi_get_request *q, uint32_t now,
                                   pgs_response_buf *r);
ORC_API int32_t orc_rrdb_batch_get(orc_server *s, const pgs_full_key *keys, uint32_t n,
                                   uint32_t now, pgs_response_buf *r);
ORC_API int32_t orc_rrdb_sortkey_count(orc_server *s, pgs_blob hash_key, uint32_t now,
                                       pgs_response_buf *r);
ORC_API int32_t orc_rrdb_get_scanner(orc_server *s, const pgs_get_scanner_request *q, uint32_t now,
                                     pgs_response_buf *r);
ORC_API int32_t orc_rrdb_scan(orc_server *s, int64_t context_id, uint32_t now,
                              pgs_response_buf *r);
ORC_API void orc_rrdb_clear_scanner(orc_server *s, int64_t context_id);
ORC_API int32_t orc_rrdb_put(orc_server *s, pgs_blob key, pgs_blob value, uint32_t expire_ts,
                             int64_t decree, uint64_t timestamp_us, uint32_t now);
ORC_API int32_t orc_rrdb_on_batched_writes(orc_server *s, const pgs_write_request *reqs, uint32_t count, int64_t decree,
                                           uint64_t timestamp_us, uint32_t now, int32_t *resp_errors);
ORC_API int32_t orc_rrdb_incr(orc_server *s, pgs_blob raw_key, int64_t increment, int32_t expire_ts_seconds, int64_t decree,
                              uint64_t timestamp_us, uint32_t now, int32_t *resp_error, int64_t *new_value);
ORC_API int32_t orc_rrdb_check_and_set(orc_server *s, const pgs_check_and_set_request *req, int64_t decree, uint64_t timestamp_us,
                                       uint32_t now, pgs_cas_result *res, uint8_t *check_value_out, uint32_t check_value_cap);
ORC_API int32_t orc_rrdb_check_and_mutate(orc_server *s, const pgs_check_and_mutate_request *req, int64_t decree, uint64_t timestamp_us,
                                          uint32_t now, pgs_cas_result *res, uint8_t *check_value_out, uint32_t check_value_cap);
ORC_API int32_t orc_rrdb_remove(orc_server *s, pgs_blob key, int64_t decree, uint32_t now);
ORC_API int32_t orc_rrdb_multi_put(orc_server *s, pgs_blob hash_key, const pgs_blob *sort_keys,
                                   const pgs_blob *values, uint32_t n, uint32_t expire_ts,
                                   int64_t decree, uint64_t timestamp_us, uint32_t now);
ORC_API int32_t orc_rrdb_multi_remove(orc_server *s, pgs_blob hash_key, const pgs_blob *sort_keys,
                                      uint32_t n, int64_t decree, int64_t *count, uint32_t now);
ORC_API int32_t orc_rrdb_flush(orc_server *s, uint32_t now);
ORC_API int32_t orc_rrdb_manual_compact(orc_server *s, uint32_t now, orc_compact_stats *st);
ORC_API int64_t orc_rrdb_last_flushed_decree(orc_server *s);
ORC_API int64_t orc_rrdb_last_committed_decree(orc_server *s);
ORC_API uint32_t orc_rrdb_gc(orc_server *s, uint32_t now);
/* test hook: number of runs / export of the whole visible state as one run */
ORC_API uint32_t orc_rrdb_run_count(orc_server *s);
ORC_API orc_run *orc_rrdb_dump(orc_server *s);

#ifdef __cplusplus
}
#endif
#endif
