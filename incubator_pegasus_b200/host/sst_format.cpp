// sst_format.cpp — RocksDB BlockBasedTable images of a sorted run (egress) and back (ingest).
//
// First slice of SURVEY.md §8 f1: format_version 2, no compression, kBinarySearch index with full internal keys,
// legacy (format_version < 5) cache-local full Bloom filter with whole keys and HashkeyTransform prefixes, crc32c block
// trailers.  This is what a Pegasus replica writes for L0/L1 (`parse_compression_types`: none below L2,
// src/server/pegasus_server_impl.cpp:3040-3056; table options src/server/pegasus_server_impl_init.cpp:560-581, 817-843) and
// what IngestExternalFile reads (src/server/rocksdb_wrapper.cpp:248-270).  RocksDB 8.5.3 is not in the reference tree: the
// layout follows its public format description and is NOT yet checked against a RocksDB build.
//
//   [data block i][type=0][masked crc32c]* [filter block][t][crc] [properties block][t][crc] [metaindex block][t][crc]
//   [index block][t][crc] [footer: checksum type(1) | metaindex handle | index handle | padding to 41 | version(4) | magic(8)]
#include "../../include/pegasus_b200.h"
#include "host_internal.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace pgs {

// ---- crc32c (Castagnoli), slicing-by-8 ------------------------------------------------------------------------------
static uint32_t g_c32[8][256];
static bool g_c32_ready = false;
static void crc32c_init()
{
    if (g_c32_ready) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (c & 1 ? 0x82F63B78u : 0);
        g_c32[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++) g_c32[t][i] = (g_c32[t - 1][i] >> 8) ^ g_c32[0][g_c32[t - 1][i] & 0xff];
    g_c32_ready = true;
}
uint32_t crc32c(uint32_t crc, const uint8_t *p, uint64_t n)
{
    crc32c_init();
    crc = ~crc;
    while (n >= 8) {
        uint64_t w;
        memcpy(&w, p, 8);
        w ^= crc;
        crc = g_c32[7][w & 0xff] ^ g_c32[6][(w >> 8) & 0xff] ^ g_c32[5][(w >> 16) & 0xff] ^ g_c32[4][(w >> 24) & 0xff] ^
              g_c32[3][(w >> 32) & 0xff] ^ g_c32[2][(w >> 40) & 0xff] ^ g_c32[1][(w >> 48) & 0xff] ^ g_c32[0][w >> 56];
        p += 8;
        n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ g_c32[0][(crc ^ *p++) & 0xff];
    return ~crc;
}
static uint32_t crc_mask(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; } // util/crc32c.h Mask()

// ---- LZ4 block format (the raw block codec RocksDB's kLZ4Compression wraps; lz4 is not linked: the format is small) ---------
// sequence = token (literal length : match length) | [length bytes] | literals | offset LE16 | [length bytes]; the last sequence
// has only literals; a match is at least 4 bytes, the last 5 bytes of a block are literals and the last match starts at
// least 12 bytes before the end.
static bool lz4_decompress(const uint8_t *ip, size_t n, uint8_t *out, size_t out_n)
{
    const uint8_t *iend = ip + n;
    size_t op = 0;
    while (ip < iend) {
        const uint32_t token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) { uint8_t b; do { if (ip >= iend) return false; b = *ip++; lit += b; } while (b == 255); }
        if ((size_t)(iend - ip) < lit || out_n - op < lit) return false;
        memcpy(out + op, ip, lit);
        op += lit; ip += lit;
        if (ip >= iend) break; // the last sequence
        if (iend - ip < 2) return false;
        const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        if (off == 0 || off > op) return false;
        size_t ml = token & 15;
        if (ml == 15) { uint8_t b; do { if (ip >= iend) return false; b = *ip++; ml += b; } while (b == 255); }
        ml += 4;
        if (out_n - op < ml) return false;
        for (size_t i = 0; i < ml; i++) out[op + i] = out[op + i - off]; // may overlap
        op += ml;
    }
    return op == out_n;
}
static void lz4_put_len(std::string &d, size_t v) { while (v >= 255) { d.push_back((char)255); v -= 255; } d.push_back((char)v); }
static std::string lz4_compress(const uint8_t *in, size_t n)
{
    std::string out;
    std::vector<uint32_t> table(1u << 13, 0xFFFFFFFFu);
    size_t anchor = 0, i = 0;
    auto emit = [&](size_t lit_end, size_t match_len, size_t offset) { // literals [anchor, lit_end), then a match (match_len 0: none)
        const size_t lit = lit_end - anchor, ml = match_len ? match_len - 4 : 0;
        out.push_back((char)(((lit < 15 ? lit : 15) << 4) | (match_len ? (ml < 15 ? ml : 15) : 0)));
        if (lit >= 15) lz4_put_len(out, lit - 15);
        out.append((const char *)in + anchor, lit);
        if (match_len) {
            out.push_back((char)(offset & 255));
            out.push_back((char)(offset >> 8));
            if (ml >= 15) lz4_put_len(out, ml - 15);
        }
    };
    if (n >= 13) {
        const size_t mflimit = n - 12, matchlimit = n - 5;
        while (i < mflimit) {
            uint32_t v;
            memcpy(&v, in + i, 4);
            const uint32_t h = (v * 2654435761u) >> 19;
            const uint32_t cand = table[h];
            table[h] = (uint32_t)i;
            uint32_t cv = 0;
            if (cand != 0xFFFFFFFFu) memcpy(&cv, in + cand, 4);
            if (cand != 0xFFFFFFFFu && cv == v && i - cand <= 65535) {
                size_t ml = 4;
                while (i + ml < matchlimit && in[cand + ml] == in[i + ml]) ml++;
                emit(i, ml, i - cand);
                i += ml;
                anchor = i;
            } else i++;
        }
    }
    emit(n, 0, 0);
    return out;
}

// ---- small encoders --------------------------------------------------------------------------------------------------
static void put_v32(std::string &d, uint32_t v)
{
    while (v >= 128) { d.push_back((char)(v | 128)); v >>= 7; }
    d.push_back((char)v);
}
static void put_v64(std::string &d, uint64_t v)
{
    while (v >= 128) { d.push_back((char)(v | 128)); v >>= 7; }
    d.push_back((char)v);
}
static void put_f32(std::string &d, uint32_t v) { d.append((const char *)&v, 4); }
static const uint8_t *get_v64(const uint8_t *p, const uint8_t *lim, uint64_t *v)
{
    uint64_t r = 0;
    for (uint32_t s = 0; s <= 63 && p < lim; s += 7) {
        const uint64_t b = *p++;
        r |= (b & 127) << s;
        if (!(b & 128)) { *v = r; return p; }
    }
    return nullptr;
}
static const uint8_t *get_v32(const uint8_t *p, const uint8_t *lim, uint32_t *v)
{
    uint64_t x;
    p = get_v64(p, lim, &x);
    if (!p || x > 0xffffffffull) return nullptr;
    *v = (uint32_t)x;
    return p;
}

// a block of (key, value) entries in BlockBuilder layout; keys must be added in order
struct KvBlock {
    std::string buf, last;
    std::vector<uint32_t> restarts{0};
    uint32_t interval, counter = 0;
    explicit KvBlock(uint32_t restart_interval) : interval(restart_interval) {}
    void add(std::string_view k, std::string_view v)
    {
        uint32_t shared = 0;
        if (counter < interval) {
            const size_t m = std::min(last.size(), k.size());
            while (shared < m && last[shared] == k[shared]) shared++;
        } else {
            restarts.push_back((uint32_t)buf.size());
            counter = 0;
        }
        put_v32(buf, shared);
        put_v32(buf, (uint32_t)(k.size() - shared));
        put_v32(buf, (uint32_t)v.size());
        buf.append(k.data() + shared, k.size() - shared);
        buf.append(v.data(), v.size());
        last.assign(k.data(), k.size());
        counter++;
    }
    std::string finish()
    {
        std::string out = buf;
        for (uint32_t r : restarts) put_f32(out, r);
        put_f32(out, (uint32_t)restarts.size());
        return out;
    }
};

// iterate the entries of one block; fn(key, value) gets the full key
template <class F>
static bool walk_block(const uint8_t *b, uint64_t size, F &&fn)
{
    if (size < 4) return false;
    uint32_t nr;
    memcpy(&nr, b + size - 4, 4);
    if ((nr & 0x7fffffffu) == 0 || 4ull * ((nr & 0x7fffffffu) + 1) > size) return false;
    const uint8_t *lim = b + size - 4ull * ((nr & 0x7fffffffu) + 1), *p = b;
    std::string key;
    while (p < lim) {
        uint32_t sh, ns, vl;
        if (!(p = get_v32(p, lim, &sh)) || !(p = get_v32(p, lim, &ns)) || !(p = get_v32(p, lim, &vl))) return false;
        if (sh > key.size() || (uint64_t)(lim - p) < (uint64_t)ns + vl) return false;
        key.resize(sh);
        key.append((const char *)p, ns);
        fn(std::string_view(key), std::string_view((const char *)p + ns, vl));
        p += ns + vl;
    }
    return true;
}

// ---- legacy full Bloom filter (FullFilterBitsBuilder of format_version < 5) ---------------------------------------------
static uint32_t bloom_hash(std::string_view key) // util/hash.cc Hash(data, n, 0xbc9f1d34)
{
    const uint32_t m = 0xc6a4a793u, seed = 0xbc9f1d34u;
    const uint8_t *d = (const uint8_t *)key.data();
    size_t n = key.size();
    uint32_t h = seed ^ (uint32_t)(n * m);
    while (n >= 4) {
        uint32_t w;
        memcpy(&w, d, 4);
        d += 4; n -= 4;
        h += w; h *= m; h ^= h >> 16;
    }
    switch (n) { // the tail bytes are sign-extended (a historical quirk every reader depends on)
    case 3: h += (uint32_t)(int32_t)(int8_t)d[2] << 16; [[fallthrough]];
    case 2: h += (uint32_t)(int32_t)(int8_t)d[1] << 8; [[fallthrough]];
    case 1: h += (uint32_t)(int32_t)(int8_t)d[0]; h *= m; h ^= h >> 24;
    }
    return h;
}
static std::string bloom_build(const std::vector<uint32_t> &hashes, uint32_t bits_per_key)
{
    const uint32_t num_probes = 6; // 10 bits per key
    uint64_t total_bits = (uint64_t)hashes.size() * bits_per_key;
    uint32_t num_lines = (uint32_t)((total_bits + 511) / 512);
    if (hashes.empty()) num_lines = 0;
    else if (num_lines % 2 == 0) num_lines++; // an odd number of 64-byte lines
    std::string out((size_t)num_lines * 64 + 5, '\0');
    for (uint32_t h : hashes) {
        uint8_t *line = (uint8_t *)out.data() + (size_t)(h % num_lines) * 64;
        const uint32_t delta = (h >> 17) | (h << 15);
        for (uint32_t i = 0; i < num_probes; i++) {
            const uint32_t bit = h & 511u;
            line[bit / 8] |= (uint8_t)(1u << (bit % 8));
            h += delta;
        }
    }
    out[(size_t)num_lines * 64] = (char)num_probes;
    memcpy(&out[(size_t)num_lines * 64 + 1], &num_lines, 4);
    return out;
}
static bool bloom_may_match(std::string_view filter, std::string_view key)
{
    if (filter.size() < 5) return true;
    const uint32_t num_probes = (uint8_t)filter[filter.size() - 5];
    uint32_t num_lines;
    memcpy(&num_lines, filter.data() + filter.size() - 4, 4);
    if (num_lines == 0 || (uint64_t)num_lines * 64 + 5 != filter.size()) return true;
    uint32_t h = bloom_hash(key);
    const uint8_t *line = (const uint8_t *)filter.data() + (size_t)(h % num_lines) * 64;
    const uint32_t delta = (h >> 17) | (h << 15);
    for (uint32_t i = 0; i < num_probes; i++) {
        const uint32_t bit = h & 511u;
        if (!(line[bit / 8] & (1u << (bit % 8)))) return false;
        h += delta;
    }
    return true;
}
static size_t hashkey_prefix(std::string_view ukey) // HashkeyTransform (src/server/hashkey_transform.h:40-60); 0 = not in domain
{
    if (ukey.size() < 2) return 0;
    const size_t p = 2 + (((size_t)(uint8_t)ukey[0] << 8) | (uint8_t)ukey[1]);
    return p <= ukey.size() ? p : 0;
}

static const uint64_t kMagic = 0x88e241b785f4cff7ull; // kBlockBasedTableMagicNumber
static const char *kFilterName = "fullfilter.rocksdb.BuiltinBloomFilter";
static const char *kPropsName = "rocksdb.properties";

struct Handle { uint64_t off = 0, size = 0; };
static void put_handle(std::string &d, Handle h) { put_v64(d, h.off); put_v64(d, h.size); }

static const uint8_t kNoCompression = 0, kLZ4Compression = 4;
static Handle append_block(std::string &file, const uint8_t *b, uint64_t n, uint8_t compression = kNoCompression)
{
    std::string packed;
    uint8_t type = kNoCompression;
    if (compression == kLZ4Compression && n < 0xffffffffull) { // compress_format_version 2: varint32 raw size | LZ4 block
        put_v32(packed, (uint32_t)n);
        packed += lz4_compress(b, n);
        if (packed.size() < n - n / 8) type = kLZ4Compression; // kept only when it saves 12.5 % (RocksDB's GoodCompressionRatio)
    }
    const uint8_t *p = type ? (const uint8_t *)packed.data() : b;
    const uint64_t pn = type ? packed.size() : n;
    Handle h{file.size(), pn};
    file.append((const char *)p, pn);
    uint32_t crc = crc32c(0, p, pn);
    crc = crc32c(crc, &type, 1);
    file.push_back((char)type);
    put_f32(file, crc_mask(crc));
    return h;
}

int32_t sst_encode(const uint8_t *data, const uint64_t *blk_off, const uint32_t *blk_size, uint32_t nb, std::string &file,
                   uint8_t compression = kNoCompression)
{
    file.clear();
    KvBlock index(1); // index_block_restart_interval = 1
    std::vector<uint32_t> hashes;
    std::string prev_prefix;
    bool have_prefix = false;
    uint64_t n_entries = 0, raw_key = 0, raw_val = 0, n_del = 0, data_size = 0;
    for (uint32_t b = 0; b < nb; b++) {
        std::string last;
        const bool ok = walk_block(data + blk_off[b], blk_size[b], [&](std::string_view ik, std::string_view v) {
            if (ik.size() < 8) return;
            const std::string_view uk = ik.substr(0, ik.size() - 8);
            const uint32_t h = bloom_hash(uk);
            if (hashes.empty() || h != hashes.back()) hashes.push_back(h); // whole key (consecutive versions of a key add once)
            const size_t pl = hashkey_prefix(uk);
            if (pl && (!have_prefix || prev_prefix != uk.substr(0, pl))) {
                prev_prefix.assign(uk.data(), pl);
                have_prefix = true;
                hashes.push_back(bloom_hash(uk.substr(0, pl)));
            }
            n_entries++;
            raw_key += ik.size();
            raw_val += v.size();
            if ((uint8_t)ik[ik.size() - 8] == PGS_TYPE_DELETION) n_del++;
            last.assign(ik.data(), ik.size());
        });
        if (!ok || last.empty()) return PGS_CORRUPTION;
        const Handle h = append_block(file, data + blk_off[b], blk_size[b], compression);
        data_size = file.size();
        std::string hv;
        put_handle(hv, h);
        index.add(last, hv); // separator = the block's last internal key (no shortening)
    }
    const std::string filter = bloom_build(hashes, 10);
    const Handle fh = append_block(file, (const uint8_t *)filter.data(), filter.size());
    const std::string index_blk = index.finish();
    // properties (sorted by name; numbers are varint64)
    std::map<std::string, std::string> props;
    auto num = [&](const char *k, uint64_t v) { std::string s; put_v64(s, v); props[k] = s; };
    num("rocksdb.data.size", data_size);
    num("rocksdb.deleted.keys", n_del);
    num("rocksdb.filter.size", filter.size());
    num("rocksdb.format.version", 2);
    num("rocksdb.index.size", index_blk.size() + 5);
    num("rocksdb.num.data.blocks", nb);
    num("rocksdb.num.entries", n_entries);
    num("rocksdb.raw.key.size", raw_key);
    num("rocksdb.raw.value.size", raw_val);
    props["rocksdb.comparator"] = "leveldb.BytewiseComparator";
    props["rocksdb.compression"] = compression == kLZ4Compression ? "LZ4" : "NoCompression";
    props["rocksdb.filter.policy"] = "rocksdb.BuiltinBloomFilter";
    props["rocksdb.prefix.extractor.name"] = "HashkeyTransform";
    KvBlock pb(1);
    for (auto &kv : props) pb.add(kv.first, kv.second);
    const std::string props_blk = pb.finish();
    const Handle ph = append_block(file, (const uint8_t *)props_blk.data(), props_blk.size());
    KvBlock meta(1);
    { std::string hv; put_handle(hv, fh); meta.add(kFilterName, hv); }
    { std::string hv; put_handle(hv, ph); meta.add(kPropsName, hv); }
    const std::string meta_blk = meta.finish();
    const Handle mh = append_block(file, (const uint8_t *)meta_blk.data(), meta_blk.size());
    const Handle ih = append_block(file, (const uint8_t *)index_blk.data(), index_blk.size());
    std::string footer;
    footer.push_back((char)1); // kCRC32c
    put_handle(footer, mh);
    put_handle(footer, ih);
    footer.resize(1 + 40, '\0');
    put_f32(footer, 2); // format_version
    footer.append((const char *)&kMagic, 8);
    file += footer;
    return PGS_OK;
}

// `inflated` keeps the bytes of a block that was stored compressed (out points into it then)
static int32_t read_block(const uint8_t *sst, uint64_t size, Handle h, std::string_view &out, std::string &inflated)
{
    if (h.off > size || h.size > size - h.off || size - h.off - h.size < 5) return PGS_CORRUPTION; // block + 5-byte trailer inside the file
    const uint8_t *b = sst + h.off;
    const uint8_t type = b[h.size];
    if (type != kNoCompression && type != kLZ4Compression) return PGS_NOT_SUPPORTED; // snappy / zstd / ...: later slices
    uint32_t stored;
    memcpy(&stored, b + h.size + 1, 4);
    if (crc_mask(crc32c(0, b, h.size + 1)) != stored) return PGS_CORRUPTION;
    if (type == kNoCompression) { out = std::string_view((const char *)b, h.size); return PGS_OK; }
    uint32_t raw = 0;
    const uint8_t *p = get_v32(b, b + h.size, &raw);
    if (!p || raw > (64u << 20)) return PGS_CORRUPTION;
    inflated.assign(raw, '\0');
    if (!lz4_decompress(p, (size_t)(b + h.size - p), (uint8_t *)inflated.data(), raw)) return PGS_CORRUPTION;
    out = inflated;
    return PGS_OK;
}

// SST image -> block run (blocks padded to 16-byte starts, as pgs_run_upload wants them) + the filter block
int32_t sst_decode(const uint8_t *sst, uint64_t size, std::string &data, std::vector<uint64_t> &off, std::vector<uint32_t> &sz, std::string *filter_out)
{
    data.clear(); off.clear(); sz.clear();
    if (size < 53) return PGS_CORRUPTION;
    const uint8_t *f = sst + size - 53;
    uint64_t magic;
    uint32_t version;
    memcpy(&magic, f + 45, 8);
    memcpy(&version, f + 41, 4);
    if (magic != kMagic) return PGS_CORRUPTION;
    if (version != 2 || f[0] != 1) return PGS_NOT_SUPPORTED; // other format versions / checksum types: later slices
    Handle mh, ih;
    const uint8_t *p = f + 1, *lim = f + 41;
    if (!(p = get_v64(p, lim, &mh.off)) || !(p = get_v64(p, lim, &mh.size)) || !(p = get_v64(p, lim, &ih.off)) || !(p = get_v64(p, lim, &ih.size)))
        return PGS_CORRUPTION;
    std::string_view meta, index;
    std::string meta_buf, index_buf, blk_buf;
    int32_t rc = read_block(sst, size, mh, meta, meta_buf);
    if (rc == PGS_OK) rc = read_block(sst, size, ih, index, index_buf);
    if (rc != PGS_OK) return rc;
    if (filter_out) {
        filter_out->clear();
        Handle fh;
        bool found = false, bad = false;
        if (!walk_block((const uint8_t *)meta.data(), meta.size(), [&](std::string_view k, std::string_view v) {
                if (k == kFilterName) {
                    const uint8_t *q = (const uint8_t *)v.data(), *ql = q + v.size();
                    if ((q = get_v64(q, ql, &fh.off)) && get_v64(q, ql, &fh.size)) found = true; else bad = true;
                }
            }) || bad)
            return PGS_CORRUPTION;
        if (found) {
            std::string_view fb;
            if ((rc = read_block(sst, size, fh, fb, blk_buf)) != PGS_OK) return rc;
            filter_out->assign(fb.data(), fb.size());
        }
    }
    std::vector<Handle> blocks;
    bool bad = false;
    if (!walk_block((const uint8_t *)index.data(), index.size(), [&](std::string_view, std::string_view v) {
            Handle h;
            const uint8_t *q = (const uint8_t *)v.data(), *ql = q + v.size();
            if ((q = get_v64(q, ql, &h.off)) && get_v64(q, ql, &h.size)) blocks.push_back(h); else bad = true;
        }) || bad)
        return PGS_CORRUPTION;
    for (const Handle &h : blocks) {
        std::string_view blk;
        if ((rc = read_block(sst, size, h, blk, blk_buf)) != PGS_OK) return rc;
        if (blk.size() > 0xffffffffull) return PGS_NOT_SUPPORTED;
        data.resize((data.size() + 15) & ~(size_t)15, '\0');
        off.push_back(data.size());
        sz.push_back((uint32_t)blk.size());
        data.append(blk.data(), blk.size());
    }
    return PGS_OK;
}

} // namespace pgs

using namespace pgs;

extern "C" {

uint32_t pgs_crc32c(const uint8_t *data, uint64_t len, uint32_t init) { return crc32c(init, data, len); }

int32_t pgs_sst_encode(const uint8_t *data, const uint64_t *blk_off, const uint32_t *blk_size, uint32_t n_blocks, uint8_t *out,
                       uint64_t out_cap, uint64_t *out_size)
{
    if (!out_size || (n_blocks && (!data || !blk_off || !blk_size))) return PGS_INVALID_ARGUMENT;
    std::string file;
    const int32_t rc = sst_encode(data, blk_off, blk_size, n_blocks, file);
    if (rc != PGS_OK) return rc;
    *out_size = file.size();
    if (!out || out_cap < file.size()) return PGS_INCOMPLETE; // *out_size tells how much room the image needs
    memcpy(out, file.data(), file.size());
    return PGS_OK;
}

int32_t pgs_sst_encode_ex(const uint8_t *data, const uint64_t *blk_off, const uint32_t *blk_size, uint32_t n_blocks, uint32_t compression,
                          uint8_t *out, uint64_t out_cap, uint64_t *out_size)
{
    if (!out_size || (n_blocks && (!data || !blk_off || !blk_size))) return PGS_INVALID_ARGUMENT;
    if (compression != kNoCompression && compression != kLZ4Compression) return PGS_NOT_SUPPORTED;
    std::string file;
    const int32_t rc = sst_encode(data, blk_off, blk_size, n_blocks, file, (uint8_t)compression);
    if (rc != PGS_OK) return rc;
    *out_size = file.size();
    if (!out || out_cap < file.size()) return PGS_INCOMPLETE;
    memcpy(out, file.data(), file.size());
    return PGS_OK;
}

int32_t pgs_lz4_block(int32_t decompress, const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_size)
{
    if (!in || !out || !out_size) return PGS_INVALID_ARGUMENT;
    if (decompress) {
        if (!lz4_decompress(in, n, out, out_cap)) return PGS_CORRUPTION; // out_cap = the exact raw size
        *out_size = out_cap;
        return PGS_OK;
    }
    const std::string c = lz4_compress(in, n);
    *out_size = c.size();
    if (c.size() > out_cap) return PGS_INCOMPLETE;
    memcpy(out, c.data(), c.size());
    return PGS_OK;
}

int32_t pgs_sst_decode(const uint8_t *sst, uint64_t size, uint8_t *data, uint64_t data_cap, uint64_t *blk_off, uint32_t *blk_size,
                       uint32_t blk_cap, uint64_t *data_bytes, uint32_t *n_blocks)
{
    if (!sst || !data_bytes || !n_blocks) return PGS_INVALID_ARGUMENT;
    std::string d;
    std::vector<uint64_t> off;
    std::vector<uint32_t> sz;
    const int32_t rc = sst_decode(sst, size, d, off, sz, nullptr);
    if (rc != PGS_OK) return rc;
    *data_bytes = d.size();
    *n_blocks = (uint32_t)off.size();
    if (!data || !blk_off || !blk_size || data_cap < d.size() || blk_cap < off.size()) return PGS_INCOMPLETE;
    memcpy(data, d.data(), d.size());
    memcpy(blk_off, off.data(), off.size() * 8);
    memcpy(blk_size, sz.data(), sz.size() * 4);
    return PGS_OK;
}

int32_t pgs_sst_filter_may_match(const uint8_t *sst, uint64_t size, const uint8_t *key, uint32_t key_len)
{
    std::string d, filter;
    std::vector<uint64_t> off;
    std::vector<uint32_t> sz;
    const int32_t rc = sst_decode(sst, size, d, off, sz, &filter);
    if (rc != PGS_OK) return -rc;
    return bloom_may_match(filter, std::string_view((const char *)key, key_len)) ? 1 : 0;
}

int32_t pgs_sst_export(pgs_partition *p, uint64_t run_id, uint8_t *out, uint64_t out_cap, uint64_t *out_size)
{
    if (!p || !out_size) return PGS_INVALID_ARGUMENT;
    pgs_run_info info;
    int32_t rc = pgs_run_info_get(p, run_id, &info);
    if (rc != PGS_OK) return rc;
    std::vector<uint8_t> data(info.data_bytes + 16);
    std::vector<uint64_t> off(info.n_blocks + 1);
    std::vector<uint32_t> sz(info.n_blocks + 1);
    rc = pgs_run_download(p, run_id, data.data(), data.size(), off.data(), sz.data(), info.n_blocks);
    if (rc != PGS_OK) return rc;
    return pgs_sst_encode(data.data(), off.data(), sz.data(), info.n_blocks, out, out_cap, out_size);
}

int32_t pgs_sst_export_ex(pgs_partition *p, uint64_t run_id, uint32_t compression, uint8_t *out, uint64_t out_cap, uint64_t *out_size)
{
    if (!p || !out_size) return PGS_INVALID_ARGUMENT;
    pgs_run_info info;
    int32_t rc = pgs_run_info_get(p, run_id, &info);
    if (rc != PGS_OK) return rc;
    std::vector<uint8_t> data(info.data_bytes + 16);
    std::vector<uint64_t> off(info.n_blocks + 1);
    std::vector<uint32_t> sz(info.n_blocks + 1);
    rc = pgs_run_download(p, run_id, data.data(), data.size(), off.data(), sz.data(), info.n_blocks);
    if (rc != PGS_OK) return rc;
    return pgs_sst_encode_ex(data.data(), off.data(), sz.data(), info.n_blocks, compression, out, out_cap, out_size);
}

int32_t pgs_sst_ingest(pgs_partition *p, int32_t level, const uint8_t *sst, uint64_t size, uint64_t *run_id_out)
{
    if (!p || !sst) return PGS_INVALID_ARGUMENT;
    std::string d;
    std::vector<uint64_t> off;
    std::vector<uint32_t> sz;
    const int32_t rc = sst_decode(sst, size, d, off, sz, nullptr);
    if (rc != PGS_OK) return rc;
    return pgs_run_upload(p, level, (const uint8_t *)d.data(), d.size(), off.data(), sz.data(), (uint32_t)off.size(), run_id_out);
}

} // extern "C"
