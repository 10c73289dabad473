// lookup.cu — host side of the read path (pgs_get_batch, pgs_range_scan, pgs_range_scan_many) and the reverse-scan kernel.
//
//   k_get / k_scan_fwd (read_kernels.cuh)  point lookups and forward range scans: lane-group iterators straight over HBM.
//   k_scan (this file)  REVERSE range scans (SeekForPrev + Prev loops of on_multi_get reverse mode,
//           src/server/pegasus_server_impl.cpp:689-756): one CTA per request; chunks of blocks of every run are staged with TMA,
//           decoded, merged by rank, newest version / tombstone visibility applied, then the reference's loop (stop key,
//           first-exclusive, range_read_limiter counts and sizes, TTL / sort-key filters) is evaluated with block-wide scans
//           walking backwards chunk by chunk.  (It also handles forward requests; a batch that mixes directions uses it.)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "read_kernels.cuh"

namespace pgs {

PGS_DEV uint32_t ld32le(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

// warp-cooperative 33-ary search over the block index: every round the 32 lanes probe 32 pivots at once, so the
// chain of dependent global loads is ~log33(nb) long instead of log2(nb).  upper=false: first block whose last key
// >= key; upper=true: first block whose last key > key.  All lanes return the same value.
PGS_DEV uint32_t warp_index_bound(const RunDev &r, const uint8_t *key, uint32_t klen, uint32_t lane, bool upper)
{
    uint32_t lo = 0, hi = r.nb;
    while (hi - lo > 32) {
        uint32_t span = hi - lo;
        uint32_t piv = lo + (uint32_t)(((unsigned long long)span * (lane + 1)) / 33);
        uint32_t o = r.ikey_off[piv], l = r.ikey_off[piv + 1] - o;
        int c = cmp_bytes4(r.ikeys + o, l, key, klen);
        bool before = upper ? c <= 0 : c < 0; // pivot block lies strictly before the answer
        uint32_t m = __ballot_sync(kFull, before);
        uint32_t cnt = __popc(m); // monotone: lanes 0..cnt-1 are true
        uint32_t nlo = cnt == 0 ? lo : __shfl_sync(kFull, piv, cnt - 1) + 1;
        uint32_t nhi = cnt == 32 ? hi : __shfl_sync(kFull, piv, cnt & 31);
        lo = nlo;
        hi = nhi;
    }
    bool before = false;
    if (lo + lane < hi) {
        uint32_t o = r.ikey_off[lo + lane], l = r.ikey_off[lo + lane + 1] - o;
        int c = cmp_bytes4(r.ikeys + o, l, key, klen);
        before = upper ? c <= 0 : c < 0;
    }
    return lo + __popc(__ballot_sync(kFull, before));
}

// ------------------------------------------------------------------------------------------------
// k_scan (reverse scans)
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kScanThreads = 256;
constexpr uint32_t kScanWarps = kScanThreads / 32;
constexpr uint32_t kScanMaxBlocks = 128;
constexpr uint32_t kScanRecExtra = 48;

struct ScanShared {
    unsigned long long mbar;
    uint32_t error, done;
    uint32_t in_bytes, n_rec, n_blk, n_valid, n_vis;
    uint32_t lo_len, hi_len, has_lo, has_hi, lo_incl, hi_incl; // chunk validity bounds
    uint32_t cur[kMaxReadRuns], nblk[kMaxReadRuns], nrec[kMaxReadRuns], in_off[kMaxReadRuns], rec_base[kMaxReadRuns], blk_base[kMaxReadRuns];
    uint32_t vlo[kMaxReadRuns], vhi[kMaxReadRuns], more[kMaxReadRuns], want_end[kMaxReadRuns];
    uint32_t tb_off[kScanMaxBlocks], tb_size[kScanMaxBlocks], tb_rec[kScanMaxBlocks], tb_nrec[kScanMaxBlocks];
    uint32_t scan[33];
    // carried loop state
    uint32_t count, iter_count, expire_count, filter_count, n_out;
    unsigned long long size, arena_used;
    uint32_t complete, iter_valid, lookahead, resume_len, first_chunk;
    uint32_t P, F; // per chunk
    uint32_t cand_len[kMaxReadRuns];
    uint32_t grec0[kMaxReadRuns]; // index of the slice's first record inside its run
    unsigned long long crc[256];
};

struct ScanArrays {
    uint8_t *in, *arena;
    unsigned long long *trailer;
    uint32_t *voff, *vlen, *A1, *A2, *A3;
    uint16_t *klen, *rank, *order, *vis;
    uint8_t *flags, *state;
    uint32_t total;
};
PGS_DEV ScanArrays scan_carve(uint8_t *pool, uint32_t in_bytes, uint32_t n, uint32_t KS)
{
    ScanArrays a;
    uint32_t n8 = (n + 8) & ~7u;
    uint32_t off = ((in_bytes + 15) & ~15u) + 16;
    a.in = pool;
    a.arena = pool + off; off += n8 * KS;
    a.trailer = (unsigned long long *)(pool + off); off += n8 * 8;
    a.voff = (uint32_t *)(pool + off); off += n8 * 4;
    a.vlen = (uint32_t *)(pool + off); off += n8 * 4;
    a.A1 = (uint32_t *)(pool + off); off += n8 * 4;
    a.A2 = (uint32_t *)(pool + off); off += n8 * 4;
    a.A3 = (uint32_t *)(pool + off); off += n8 * 4;
    a.klen = (uint16_t *)(pool + off); off += n8 * 2;
    a.rank = (uint16_t *)(pool + off); off += n8 * 2;
    a.order = (uint16_t *)(pool + off); off += n8 * 2;
    a.vis = (uint16_t *)(pool + off); off += n8 * 2;
    a.flags = pool + off; off += n8;
    a.state = pool + off; off += n8;
    a.total = off;
    return a;
}

template <class F>
PGS_DEV uint32_t scan_chunked(uint32_t n, uint32_t *out, uint32_t *scratch, F f)
{
    uint32_t ipt = (n + blockDim.x - 1) / blockDim.x;
    uint32_t begin = min(threadIdx.x * ipt, n), end = min(begin + ipt, n);
    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) local += f(i);
    uint32_t total;
    uint32_t pre = block_excl_scan(local, scratch, &total);
    for (uint32_t i = begin; i < end; i++) { uint32_t v = f(i); out[i] = pre; pre += v; }
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads();
    return total;
}

enum : uint8_t { SF_VALID = 1, SF_SHADOW = 2 };

__global__ void __launch_bounds__(kScanThreads, 4) k_scan(const __grid_constant__ ScanParams P)
{
    extern __shared__ __align__(128) uint8_t dyn[];
    __shared__ ScanShared S;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t KS = P.KS, NR = P.rr.n;
    // bounds of the current chunk, zero padded slots: lo = exclusive/inclusive lower, hi = upper
    const uint32_t slot = (KS + 8 + 15) & ~15u; // keeps `pool` (the TMA destination) 16-byte aligned
    uint8_t *klo = dyn, *khi = dyn + slot, *kpre = dyn + 2 * slot;
    uint8_t *kend = dyn + 3 * slot;  // the range end (first KS+8 bytes: no stored key is longer than KS)
    uint8_t *cand = dyn + 4 * slot;  // one slot per run: its candidate for the chunk's far bound
    uint8_t *pool = dyn + (4 + NR) * slot + kScanWarps * P.warp_scratch;

    if (tid == 0) { mbar_init((uint64_t *)&S.mbar, 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t phase = 0;
    long long pt_last = P.phase_cycles ? clock64() : 0; // phase timing (diagnostics): thread 0 stamps phase boundaries
#define SPT(i) do { if (P.phase_cycles && tid == 0) { long long t_ = clock64(); atomicAdd(&P.phase_cycles[i], (unsigned long long)(t_ - pt_last)); pt_last = t_; } } while (0)

    for (uint32_t rq = blockIdx.x; rq < P.n; rq += gridDim.x) {
        const ScanReqDev &Q = P.reqs[rq];
        const uint8_t *start = P.blob + Q.start_off, *stop = P.blob + Q.stop_off;
        const uint8_t *hf = P.blob + Q.hf_off, *sf = P.blob + Q.sf_off;
        const bool rev = Q.reverse != 0;
        // the range end in iteration direction ("stop" forward, "start" reverse) and its inclusiveness
        const uint8_t *endk = rev ? start : stop;
        const uint32_t endl = rev ? Q.start_len : Q.stop_len;
        const bool end_incl = rev ? Q.start_inclusive : Q.stop_inclusive;
        pgs_kv *kvs = P.kvs + (size_t)rq * P.kv_stride;
        uint8_t *arena = P.arena + (size_t)rq * P.arena_stride;

        // prefix_same_as_start: the iterator only lives inside the seek key's hash-key prefix
        uint32_t pre_len = 0;
        if (Q.prefix_same_as_start && !rev && Q.start_len >= 2) {
            uint32_t hl = be16(start);
            if (2 + hl <= Q.start_len) pre_len = 2 + hl;
        }
        if (tid == 0) {
            S.count = S.iter_count = S.expire_count = S.filter_count = S.n_out = 0;
            S.size = 0; S.arena_used = 0;
            S.complete = 0; S.iter_valid = 0; S.lookahead = 0; S.resume_len = 0; S.done = 0; S.error = 0; S.first_chunk = 1;
        }
        if (P.crc_table && Q.validate_hash)
            for (uint32_t i = tid; i < 256; i += kScanThreads) S.crc[i] = P.crc_table[i];
        // initial cursors (one warp per run, 33-ary index search): forward: first block whose last key >= start;
        // reverse: first block whose last key >= stop.  want_end: first block whose last key >= the range end.
        // (the two searches of a run are independent chains of global round trips: different warps take them)
        for (uint32_t task = warp; task < 2 * NR; task += kScanWarps) {
            const uint32_t j = task >> 1;
            const RunDev &r = P.rr.runs[j];
            if (task & 1) {
                uint32_t we = warp_index_bound(r, endk, endl, lane, false);
                if (lane == 0) S.want_end[j] = we;
            } else {
                const uint8_t *sk = rev ? stop : start;
                uint32_t sl = rev ? Q.stop_len : Q.start_len;
                uint32_t b = warp_index_bound(r, sk, sl, lane, false);
                if (rev && b >= r.nb) b = r.nb ? r.nb - 1 : 0;
                if (lane == 0) S.cur[j] = b;
            }
        }
        // first chunk bound in iteration direction = the seek key
        for (uint32_t i = tid; i < KS + 8; i += kScanThreads) {
            const uint8_t *sk = rev ? stop : start;
            uint32_t sl = rev ? Q.stop_len : Q.start_len;
            uint8_t v = i < sl && i < KS ? sk[i] : 0;
            if (rev) khi[i] = v; else klo[i] = v;
            kpre[i] = i < pre_len ? start[i] : 0;
            kend[i] = i < endl ? endk[i] : 0;
        }
        if (tid == 0) {
            uint32_t sl = rev ? Q.stop_len : Q.start_len;
            if (sl > KS) sl = KS; // longer than any stored key: the truncated prefix compares the same way below
            if (rev) { S.hi_len = sl; S.has_hi = 1; S.hi_incl = Q.stop_inclusive; S.has_lo = 0; S.lo_len = 0; S.lo_incl = 0; }
            else { S.lo_len = sl; S.has_lo = 1; S.lo_incl = Q.start_inclusive; S.has_hi = 0; S.hi_len = 0; S.hi_incl = 1; }
        }
        __syncthreads();
        // keys longer than KS cannot exist in the runs; a seek key longer than KS that shares its first KS
        // bytes with a stored key sorts after it: make the truncated bound exclusive/inclusive accordingly
        if (tid == 0) {
            uint32_t sl = rev ? Q.stop_len : Q.start_len;
            if (sl > KS) { if (rev) S.hi_incl = 1; else S.lo_incl = 0; }
        }
        __syncthreads();
        SPT(0);

        // ================================ chunk loop ==========================================
        for (;;) {
            __syncthreads();
            const bool stop_now = S.done || S.error;
            __syncthreads();
            if (stop_now) break;
            // ---- choose blocks: warp 0, lane j = run j; a couple of independent global loads per run -----------
            if (warp == 0) {
                const uint32_t j = lane;
                bool has = false;
                uint32_t c = 0, m = 0, lo_b = 0, bytes_j = 0, recs_j = 0, more_j = 0;
                const RunDev *rp = nullptr;
                if (j < NR) {
                    rp = &P.rr.runs[j];
                    c = S.cur[j];
                    has = rev ? (rp->nb > 0 && c != 0xFFFFFFFFu) : (c < rp->nb);
                }
                const uint32_t active = __popc(__ballot_sync(kFull, has));
                if (has) {
                    const RunDev &r = *rp;
                    const uint32_t budget = (P.pool_bytes - 64) / active; // per-run share of the pool, at least one block each
                    uint32_t maxm = 1;
                    if (!S.lookahead) { // the wanted range end limits the first fetches
                        uint32_t want_end = S.want_end[j];
                        maxm = rev ? (c >= want_end ? c - want_end + 1 : 1) : (want_end >= c ? want_end - c + 1 : 1);
                        if (!rev && maxm > r.nb - c) maxm = r.nb - c;
                        if (rev && maxm > c + 1) maxm = c + 1;
                    }
                    auto weight = [&](uint32_t mm) -> unsigned long long {
                        uint32_t l = rev ? c + 1 - mm : c, h = l + mm;
                        return (r.blk_off[h] - r.blk_off[l]) + 32 + (unsigned long long)(r.blk_rec[h] - r.blk_rec[l]) * (KS + kScanRecExtra);
                    };
                    // largest m in [1, maxm] whose blocks and records fit the budget (cumulative arrays).  The weights of the
                    // first eight candidates come from loads issued together (one round trip); only a run that may take
                    // more than eight blocks continues with a binary search.
                    constexpr uint32_t kProbe = 8;
                    const uint32_t np = maxm < kProbe ? maxm : kProbe;
                    const uint32_t b0 = rev ? c + 1 : c;
                    unsigned long long o[kProbe + 1];
                    uint32_t rc[kProbe + 1];
#pragma unroll
                    for (uint32_t x = 0; x <= kProbe; x++) {
                        const uint32_t idx = x <= np ? (rev ? b0 - x : b0 + x) : b0;
                        o[x] = r.blk_off[idx];
                        rc[x] = r.blk_rec[idx];
                    }
                    m = 1;
#pragma unroll
                    for (uint32_t x = 2; x <= kProbe; x++) {
                        const unsigned long long wb = rev ? o[0] - o[x] : o[x] - o[0];
                        const uint32_t wr = rev ? rc[0] - rc[x] : rc[x] - rc[0];
                        if (x <= np && wb + 32 + (unsigned long long)wr * (KS + kScanRecExtra) <= budget) m = x; // weights grow with x
                    }
                    if (m == kProbe && maxm > kProbe) {
                        if (weight(maxm) <= budget) m = maxm;
                        else {
                            uint32_t lo = kProbe, hi = maxm;
                            while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (weight(mid) <= budget) lo = mid; else hi = mid; }
                            m = lo;
                        }
                    }
                    lo_b = rev ? c + 1 - m : c;
                    bytes_j = (uint32_t)(r.blk_off[lo_b + m] - r.blk_off[lo_b]);
                    const uint32_t g0 = r.blk_rec[lo_b];
                    recs_j = r.blk_rec[lo_b + m] - g0;
                    S.grec0[j] = g0;
                    more_j = rev ? (lo_b > 0) : (lo_b + m < r.nb);
                }
                const uint32_t ib = warp_incl_scan(bytes_j, lane), ir = warp_incl_scan(recs_j, lane), im = warp_incl_scan(m, lane);
                if (j < NR) {
                    S.nblk[j] = m;
                    S.in_off[j] = ib - bytes_j;
                    S.rec_base[j] = ir - recs_j;
                    S.blk_base[j] = im - m;
                    S.nrec[j] = recs_j;
                    S.more[j] = more_j;
                }
                const uint32_t bytes = __shfl_sync(kFull, ib, 31), recs = __shfl_sync(kFull, ir, 31), blks = __shfl_sync(kFull, im, 31);
                if (lane == 0) {
                    S.in_bytes = bytes; S.n_rec = recs; S.n_blk = blks;
                    ScanArrays a0 = scan_carve(pool, bytes, recs, KS);
                    if (a0.total > P.pool_bytes || blks > kScanMaxBlocks || recs > 65000) S.error = PGS_NOT_SUPPORTED;
                    if (!active) S.done = 1; // every run exhausted: the iterator is invalid
                }
            }
            __syncthreads();
            SPT(1);
            if (S.done || S.error) break;
            const ScanArrays A = scan_carve(pool, S.in_bytes, S.n_rec, KS);
            // ---- stage + block table ----------------------------------------------------------------
            if (P.use_tma) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx((uint64_t *)&S.mbar, S.in_bytes);
                    for (uint32_t j = 0; j < NR; j++) {
                        uint32_t m = S.nblk[j];
                        if (!m) continue;
                        const RunDev &r = P.rr.runs[j];
                        uint32_t lo_b = rev ? S.cur[j] + 1 - m : S.cur[j];
                        tma_load_1d(A.in + S.in_off[j], r.data + r.blk_off[lo_b], (uint32_t)(r.blk_off[lo_b + m] - r.blk_off[lo_b]), (uint64_t *)&S.mbar);
                    }
                }
            } else {
                for (uint32_t j = 0; j < NR; j++) {
                    uint32_t m = S.nblk[j];
                    if (!m) continue;
                    const RunDev &r = P.rr.runs[j];
                    uint32_t lo_b = rev ? S.cur[j] + 1 - m : S.cur[j];
                    uint32_t bytes = (uint32_t)(r.blk_off[lo_b + m] - r.blk_off[lo_b]);
                    const uint4 *src = (const uint4 *)(r.data + r.blk_off[lo_b]);
                    uint4 *dst = (uint4 *)(A.in + S.in_off[j]);
                    for (uint32_t i = tid; i < bytes / 16; i += kScanThreads) dst[i] = src[i];
                }
            }
            for (uint32_t t = tid; t < S.n_blk; t += kScanThreads) {
                uint32_t j = 0;
                while (j + 1 < NR && t >= S.blk_base[j + 1]) j++;
                const RunDev &r = P.rr.runs[j];
                uint32_t m = S.nblk[j];
                uint32_t lo_b = rev ? S.cur[j] + 1 - m : S.cur[j];
                uint32_t gb = lo_b + (t - S.blk_base[j]);
                S.tb_off[t] = S.in_off[j] + (uint32_t)(r.blk_off[gb] - r.blk_off[lo_b]);
                S.tb_size[t] = r.blk_size[gb];
                S.tb_rec[t] = S.rec_base[j] + (r.blk_rec[gb] - r.blk_rec[lo_b]);
                S.tb_nrec[t] = r.blk_rec[gb + 1] - r.blk_rec[gb];
            }
            // chunk's far bound: everything up to the nearest "last loaded block" key of a run that has more blocks.
            // Every candidate key is first staged in shared memory (one warp per run, coalesced), then compared there:
            // a byte-wise compare straight out of global memory would pay one round trip per byte.
            for (uint32_t j = warp; j < NR; j += kScanWarps) {
                uint32_t l = 0xFFFFFFFFu;
                if (S.nblk[j] && S.more[j]) {
                    const RunDev &r = P.rr.runs[j];
                    const uint32_t m = S.nblk[j];
                    // forward: last key of the last loaded block; reverse: last key of the block before the first loaded one
                    const uint32_t bb = rev ? (S.cur[j] + 1 - m) - 1 : S.cur[j] + m - 1;
                    const uint32_t o = r.ikey_off[bb];
                    l = r.ikey_off[bb + 1] - o;
                    for (uint32_t i = lane; i < KS + 8; i += 32) cand[j * slot + i] = i < l ? r.ikeys[o + i] : 0;
                }
                if (lane == 0) S.cand_len[j] = l;
            }
            if (P.use_tma) { mbar_wait((uint64_t *)&S.mbar, phase); phase ^= 1; }
            __syncthreads();
            SPT(2);
            if (tid == 0) {
                int best = -1;
                for (uint32_t j = 0; j < NR; j++) {
                    if (S.cand_len[j] == 0xFFFFFFFFu) continue;
                    if (best < 0) { best = (int)j; continue; }
                    int c = cmp_bytes(cand + j * slot, S.cand_len[j], cand + best * slot, S.cand_len[best]);
                    if (rev ? c > 0 : c < 0) best = (int)j;
                }
                S.P = (uint32_t)best; // reuse as scratch: run of the far bound
            }
            __syncthreads();
            {
                int best = (int)S.P;
                uint8_t *dst = rev ? klo : khi;
                if (best >= 0) {
                    const uint32_t l = S.cand_len[best];
                    for (uint32_t i = tid; i < KS + 8; i += kScanThreads) dst[i] = cand[best * slot + i];
                    if (tid == 0) { if (rev) { S.lo_len = l; S.has_lo = 1; S.lo_incl = 0; } else { S.hi_len = l; S.has_hi = 1; S.hi_incl = 1; } }
                } else if (tid == 0) {
                    if (rev) S.has_lo = 0; else S.has_hi = 0;
                }
            }
            __syncthreads();
            SPT(3);

            // ---- decode step 1: one THREAD per record parses its entry header; the entry's offset inside its block
            //      comes from the run's rec_off index, so no thread walks a block's entry chain -------------------------
            {
                const uint32_t nblk = S.n_blk;
                for (uint32_t r = tid; r < S.n_rec; r += kScanThreads) {
                    uint32_t j = 0;
                    while (j + 1 < NR && r >= S.rec_base[j + 1]) j++;
                    uint32_t lo = 0, hi = nblk; // block of record r: last t with tb_rec[t] <= r
                    while (lo + 1 < hi) { uint32_t mid = (lo + hi) >> 1; if (S.tb_rec[mid] <= r) lo = mid; else hi = mid; }
                    const uint32_t t = lo;
                    const uint8_t *base = A.in + S.tb_off[t];
                    const uint32_t size = S.tb_size[t], i = r - S.tb_rec[t], cnt = S.tb_nrec[t];
                    uint32_t err = 0, nr = 0;
                    if (size < 8) err = PGS_CORRUPTION;
                    if (!err) { nr = ld32le(base + size - 4); if (nr == 0 || (unsigned long long)nr * 4 + 4 > size) err = PGS_CORRUPTION; }
                    const uint32_t limit = err ? 0 : size - 4 - 4 * nr;
                    const uint32_t p = err ? 0 : P.rr.runs[j].rec_off[S.grec0[j] + (r - S.rec_base[j])];
                    if (!err && (p >= limit || (i == 0 && p != 0))) err = PGS_CORRUPTION;
                    if (!err) {
                        uint32_t sh, ns, vl, h, c;
                        h = c = parse_header8(lds_u64_at(A.in, S.tb_off[t] + p), sh, ns, vl); // header bytes from registers
                        if (!c) { // uncommon shape: byte-wise decoder
                            h = 0;
                            c = get_varint32(base + p, limit - p, sh); h += c;
                            if (c) { c = get_varint32(base + p + h, limit - p - h, ns); h += c; }
                            if (c) { c = get_varint32(base + p + h, limit - p - h, vl); h += c; }
                        }
                        const uint32_t kl = sh + ns;
                        const unsigned long long end = (unsigned long long)p + h + ns + vl;
                        if (!c || kl < 8 || kl - 8 > KS || end > limit || (i == 0 && sh != 0) || (i + 1 == cnt && end != limit)) err = PGS_CORRUPTION;
                        else {
                            A.rank[r] = (uint16_t)sh;  // scratch until the rank phase
                            A.order[r] = (uint16_t)ns; // scratch until the scatter phase
                            A.A1[r] = S.tb_off[t] + p + h; // the key delta
                            A.klen[r] = (uint16_t)(kl - 8);
                            A.voff[r] = S.tb_off[t] + p + h + ns;
                            A.vlen[r] = vl;
                            if (ns >= 8) { A.trailer[r] = lds_u64_at(A.in, S.tb_off[t] + p + h + ns - 8); A.flags[r] = 0; }
                            else { A.trailer[r] = 0; A.flags[r] = 1; } // part of the trailer is shared with the previous key: step 2
                        }
                    }
                    if (err) atomicMax(&S.error, err);
                }
            }
            __syncthreads();
            SPT(4);
            if (S.error) break;
            // ---- decode step 2: HALF a warp per block rebuilds the keys, four key bytes per lane ---------
            {
                const uint32_t hl = lane & 15, sub = lane >> 4;
                const uint32_t hmask = sub ? 0xffff0000u : 0x0000ffffu;
                for (uint32_t t = 2 * warp + sub; t < S.n_blk; t += 2 * kScanWarps) {
                    const uint32_t rec0 = S.tb_rec[t], nrec = S.tb_nrec[t];
                    uint32_t maxk = 0;
                    for (uint32_t i = hl; i < nrec; i += 16) maxk = max(maxk, (uint32_t)A.klen[rec0 + i] + 8);
                    maxk = __reduce_max_sync(hmask, maxk);
                    for (uint32_t pass = 0; pass * 64 < maxk; pass++) {
                        const uint32_t p0 = pass * 64 + 4 * hl;
                        uint32_t cur = 0, prev_klen = 0; // the four running bytes, little endian
                        for (uint32_t i = 0; i < nrec; i++) {
                            const uint32_t r = rec0 + i;
                            const uint32_t sh = A.rank[r], ns = A.order[r], ulen = A.klen[r], ko = A.A1[r], fl = A.flags[r];
                            if (sh > prev_klen) { if (hl == 0) atomicMax(&S.error, (uint32_t)PGS_CORRUPTION); break; } // a prefix longer than the previous key
                            prev_klen = ulen + 8;
                            const uint32_t a = max(sh, p0), b = min(sh + ns, p0 + 4);
                            if (a < b) {
                                const uint32_t so = ko + (a - sh); // delta bytes for positions a..a+3
                                const uint32_t *w = (const uint32_t *)A.in + (so >> 2);
                                const uint32_t x = __funnelshift_r(w[0], w[1], (so & 3) * 8);
                                const uint32_t s0 = 8 * (a - p0), s1 = 8 * (p0 + 4 - b);
                                const uint32_t msk = (0xffffffffu << s0) & (0xffffffffu >> s1);
                                cur = (cur & ~msk) | ((x << s0) & msk);
                            }
                            const uint32_t pad = (ulen + 7) & ~7u; // slots are zero padded to 8 bytes
                            if (p0 < pad) {
                                const uint32_t keep = ulen > p0 ? ulen - p0 : 0;
                                *(uint32_t *)(A.arena + (size_t)r * KS + p0) = keep >= 4 ? cur : (cur & ((1u << (8 * keep)) - 1u));
                            }
                            if (fl && pass * 64 < ulen + 8 && pass * 64 + 64 > ulen) { // rare: the trailer straddles the shared prefix
                                unsigned long long c = 0;
                                if (p0 >= ulen) { if (p0 < ulen + 8) c = (unsigned long long)cur << (8 * (p0 - ulen)); }
                                else if (ulen - p0 < 4) c = cur >> (8 * (ulen - p0));
                                const uint32_t lo = __reduce_or_sync(hmask, (uint32_t)c), hi = __reduce_or_sync(hmask, (uint32_t)(c >> 32));
                                if (hl == 0) A.trailer[r] |= ((unsigned long long)hi << 32) | lo;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            SPT(5);
            if (S.error) break;

            // ---- validity window per run: lo (<|<=) key (<=) hi.  Records at or below the lower bound form a prefix of a
            //      run's slice and records above the upper bound a suffix: counting them in parallel gives the window ------
            if (tid < NR) { S.vlo[tid] = 0; S.vhi[tid] = 0; } // vhi counts the records above the bound first
            __syncthreads();
            for (uint32_t r = tid; r < S.n_rec; r += kScanThreads) {
                uint32_t j = 0;
                while (j + 1 < NR && r >= S.rec_base[j + 1]) j++;
                const uint8_t *key = A.arena + (size_t)r * KS;
                const uint32_t kl = A.klen[r];
                bool below = false;
                if (S.has_lo) { int c = cmp_slots(key, kl, klo, S.lo_len); below = S.lo_incl ? c < 0 : c <= 0; }
                if (below) atomicAdd(&S.vlo[j], 1u);
                else if (S.has_hi) { int c = cmp_slots(key, kl, khi, S.hi_len); if (S.hi_incl ? c > 0 : c >= 0) atomicAdd(&S.vhi[j], 1u); }
            }
            __syncthreads();
            if (tid < NR) {
                uint32_t vhi = S.nrec[tid] - S.vhi[tid];
                if (vhi < S.vlo[tid]) vhi = S.vlo[tid];
                S.vhi[tid] = vhi;
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t nv = 0;
                for (uint32_t j = 0; j < NR; j++) nv += S.vhi[j] - S.vlo[j];
                S.n_valid = nv;
            }
            // ---- merge rank + shadowing ----------------------------------------------------------------------------
            // (1) one thread per record: validity, position inside its own run, predecessor of the same run;
            // (2) one thread per (record, other run): LCP-aware binary search for the number of that run's records that sort
            //     before it; ranks accumulate with shared-memory atomics (A1 = rank, A2 = shadowed)
            for (uint32_t r = tid; r < S.n_rec; r += kScanThreads) {
                uint32_t j = 0;
                while (j + 1 < NR && r >= S.rec_base[j + 1]) j++;
                uint32_t idx = r - S.rec_base[j];
                if (idx < S.vlo[j] || idx >= S.vhi[j]) { A.flags[r] = 0; continue; }
                const uint32_t kl = A.klen[r];
                const bool shadow = idx > 0 && A.klen[r - 1] == kl && cmp_slots(A.arena + (size_t)(r - 1) * KS, kl, A.arena + (size_t)r * KS, kl) == 0;
                A.A1[r] = idx - S.vlo[j];
                A.A2[r] = shadow ? 1u : 0u;
                A.flags[r] = SF_VALID;
            }
            __syncthreads();
            if (NR > 1) {
                const uint32_t km1 = NR - 1, ntask = S.n_rec * km1;
                for (uint32_t id = tid; id < ntask; id += kScanThreads) {
                    const uint32_t r = id / km1, oi = id - r * km1;
                    if (!(A.flags[r] & SF_VALID)) continue;
                    uint32_t j = 0;
                    while (j + 1 < NR && r >= S.rec_base[j + 1]) j++;
                    const uint32_t o = oi < j ? oi : oi + 1;
                    if (S.vhi[o] == S.vlo[o]) continue;
                    const uint8_t *key = A.arena + (size_t)r * KS;
                    const uint32_t kl = A.klen[r];
                    const unsigned long long tr = A.trailer[r];
                    uint32_t base = S.rec_base[o], lo = S.vlo[o], hi = S.vhi[o];
                    uint32_t lcp_lo = 0, lcp_hi = 0; // words shared with the keys just outside [lo, hi)
                    while (lo < hi) {
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
                    if (lo > S.vlo[o]) {
                        atomicAdd(&A.A1[r], lo - S.vlo[o]);
                        uint32_t q = base + lo - 1;
                        if (A.klen[q] == kl && cmp_slots(A.arena + (size_t)q * KS, kl, key, kl) == 0) atomicOr(&A.A2[r], 1u);
                    }
                }
            }
            __syncthreads();
            for (uint32_t r = tid; r < S.n_rec; r += kScanThreads)
                if (A.flags[r] & SF_VALID) {
                    A.order[A.A1[r]] = (uint16_t)r;
                    if (A.A2[r]) A.flags[r] = SF_VALID | SF_SHADOW;
                }
            __syncthreads();
            SPT(7);
            // ---- visible records in iteration order ---------------------------------------------------------------------
            const uint32_t nv = S.n_valid;
            auto at = [&](uint32_t p) -> uint32_t { return A.order[rev ? nv - 1 - p : p]; };
            uint32_t nvis = scan_chunked(nv, A.A1, S.scan, [&](uint32_t p) -> uint32_t {
                uint32_t r = at(p);
                return (!(A.flags[r] & SF_SHADOW) && (uint8_t)A.trailer[r] == PGS_TYPE_VALUE) ? 1u : 0u;
            });
            for (uint32_t p = tid; p < nv; p += kScanThreads) {
                uint32_t r = at(p);
                if (!(A.flags[r] & SF_SHADOW) && (uint8_t)A.trailer[r] == PGS_TYPE_VALUE) A.vis[A.A1[p]] = (uint16_t)r;
            }
            __syncthreads();
            // prefix bound: visible records outside the seek prefix end the iterator
            // per visible record: in-prefix, in-range, state, sizes
            //   A2 <- 1 if state==normal (count prefix), A3 <- output bytes if normal (size prefix)
            for (uint32_t v = tid; v < nvis; v += kScanThreads) {
                uint32_t r = A.vis[v];
                const uint8_t *key = A.arena + (size_t)r * KS;
                uint32_t kl = A.klen[r];
                uint8_t st;
                bool in_prefix = true;
                if (pre_len) {
                    in_prefix = kl >= pre_len;
                    for (uint32_t i = 0; in_prefix && i < pre_len; i++) in_prefix = key[i] == kpre[i];
                }
                int c = cmp_bytes(key, kl, kend, endl); // reads at most min(kl, endl) <= KS bytes of the staged range end
                bool in_range = rev ? (c > 0 || (c == 0 && end_incl)) : (c < 0 || (c == 0 && end_incl));
                if (Q.has_upper && !rev) in_prefix = in_prefix && c < 0; // iterate_upper_bound (sortkey_count)
                const uint8_t *val = A.in + A.voff[r];
                uint32_t vl = A.vlen[r];
                uint32_t ets = vl >= 4 ? be32(val) : 0;
                uint32_t hkl = kl >= 2 ? be16(key) : 0;
                if (hkl + 2 > kl) hkl = kl >= 2 ? kl - 2 : 0;
                const uint8_t *hk = key + 2, *sk = key + 2 + hkl;
                uint32_t skl = kl >= 2 ? kl - 2 - hkl : 0;
                if (ts_expired(P.now, ets)) st = RS_EXPIRED;
                else {
                    st = RS_NORMAL;
                    if (Q.validate_hash) { // validate_key_value_for_scan: :2397-2404
                        bool bad = Q.partition_version < 0 || Q.pidx > Q.partition_version;
                        if (!bad && kl >= 2) {
                            unsigned long long hcrc = ~0ull;
                            const uint8_t *hp = hkl ? hk : sk;
                            uint32_t hn = hkl ? hkl : skl;
                            for (uint32_t i = 0; i < hn; i++) hcrc = S.crc[(uint8_t)(hcrc ^ hp[i])] ^ (hcrc >> 8);
                            hcrc = ~hcrc;
                            bad = (long long)(hcrc & (unsigned long long)(long long)Q.partition_version) != (long long)Q.pidx;
                        }
                        if (bad) st = RS_HASH_INVALID;
                    }
                    if (st == RS_NORMAL && Q.hash_filter_type != PGS_FT_NO_FILTER && !dev_validate_filter(Q.hash_filter_type, hf, Q.hf_len, hk, hkl)) st = RS_FILTERED;
                    if (st == RS_NORMAL && Q.sort_filter_type != PGS_FT_NO_FILTER && !dev_validate_filter(Q.sort_filter_type, sf, Q.sf_len, sk, skl)) st = RS_FILTERED;
                }
                uint32_t hdr = user_data_offset(P.data_version);
                uint32_t out_k = Q.key_mode == 1 ? skl : kl;
                uint32_t out_v = Q.no_value ? 0 : (vl >= hdr ? vl - hdr : 0);
                A.state[v] = st | (in_prefix ? 0x10 : 0) | (in_range ? 0x20 : 0);
                A.A2[v] = st == RS_NORMAL ? 1u : 0u;
                A.A3[v] = st == RS_NORMAL ? out_k + out_v : 0u;
            }
            __syncthreads();
            SPT(8);
            // count / size prefixes over the visible list (in place: A2, A3 become exclusive prefixes)
            scan_chunked(nvis, A.A1, S.scan, [&](uint32_t v) -> uint32_t { return A.A2[v]; });
            for (uint32_t v = tid; v <= nvis; v += kScanThreads) A.A2[v] = A.A1[v];
            __syncthreads();
            scan_chunked(nvis, A.A1, S.scan, [&](uint32_t v) -> uint32_t { return A.A3[v]; });
            // A1 = size prefix, A2 = count prefix
            // ---- the reference loop, evaluated for all positions at once ---------------------------------------------
            if (tid == 0) { S.P = nvis; S.F = nvis; S.n_vis = nvis; }
            __syncthreads();
            for (uint32_t v = tid; v < nvis; v += kScanThreads) {
                uint8_t s = A.state[v];
                // F: first position where the iterator is out of its prefix or beyond the range end
                if (!(s & 0x10) || !(s & 0x20)) atomicMin(&S.F, v);
                // P: first position where `count < max_count && limiter.valid()` fails
                bool ok = !S.lookahead && (S.count + A.A2[v] < Q.max_count) && (S.iter_count + v < Q.max_iter_count) &&
                          (Q.max_iter_size == 0 || S.size + A.A1[v] < Q.max_iter_size);
                if (!ok) atomicMin(&S.P, v);
            }
            __syncthreads();
            SPT(9);
            const uint32_t Pp = S.P, Ff = S.F;
            const uint32_t nproc = min(Pp, Ff); // processed positions [0, nproc)
            // ---- emit ------------------------------------------------------------------------------------------------------
            if (!S.lookahead && nproc > 0 && !Q.count_only) {
                uint32_t hdr = user_data_offset(P.data_version);
                for (uint32_t v = warp; v < nproc; v += kScanWarps) {
                    if ((A.state[v] & 0xF) != RS_NORMAL) continue;
                    uint32_t r = A.vis[v], kl = A.klen[r], vl = A.vlen[r];
                    const uint8_t *key = A.arena + (size_t)r * KS;
                    uint32_t koff = 0, klen_out = kl;
                    if (Q.key_mode == 1) { uint32_t hkl = kl >= 2 ? be16(key) : 0; if (hkl + 2 > kl) hkl = kl >= 2 ? kl - 2 : 0; koff = 2 + hkl; klen_out = kl >= 2 ? kl - koff : 0; }
                    uint32_t vlen_out = Q.no_value ? 0 : (vl >= hdr ? vl - hdr : 0);
                    uint32_t slot = S.n_out + (A.A2[v]);
                    unsigned long long aoff = S.arena_used + A.A1[v];
                    if (slot >= P.kv_stride || aoff + klen_out + vlen_out > P.arena_stride) { if (lane == 0) atomicMax(&S.error, (uint32_t)PGS_ABORTED); continue; }
                    warp_copy_bytes(arena + aoff, key + koff, klen_out, lane);
                    if (vlen_out) warp_copy_s2g(arena + aoff + klen_out, A.in + A.voff[r] + hdr, vlen_out, lane);
                    if (lane == 0) {
                        pgs_kv kv;
                        kv.key_off = (uint32_t)aoff; kv.key_len = klen_out;
                        kv.value_off = (uint32_t)aoff + klen_out; kv.value_len = vlen_out;
                        kv.expire_ts = Q.return_expire_ts && vl >= 4 ? be32(A.in + A.voff[r]) : 0;
                        kvs[slot] = kv;
                    }
                }
            }
            __syncthreads();
            SPT(10);
            // ---- advance the loop state -----------------------------------------------------------------------------------------
            if (tid == 0) {
                uint32_t nvis_ = S.n_vis;
                if (!S.lookahead) {
                    uint32_t exp = 0, fil = 0;
                    for (uint32_t v = 0; v < nproc; v++) { uint8_t s = A.state[v] & 0xF; exp += s == RS_EXPIRED; fil += s == RS_FILTERED; }
                    uint32_t normals = A.A2[nproc];
                    S.expire_count += exp; S.filter_count += fil;
                    S.iter_count += nproc;
                    S.count += normals;
                    if (!Q.count_only) { S.n_out += normals; S.arena_used += A.A1[nproc]; }
                    S.size += A.A1[nproc];
                    // a processed record equal to the range end completes the scan (`if (c == 0) complete`)
                    bool hit_end = false;
                    if (nproc > 0 && end_incl) {
                        uint32_t r = A.vis[nproc - 1];
                        hit_end = cmp_bytes(A.arena + (size_t)r * KS, A.klen[r], kend, endl) == 0;
                    }
                    if (hit_end) { S.complete = 1; S.iter_valid = 1; S.done = 1; }
                    else if (Pp <= Ff && Pp < nvis_) { // limits ended the loop while the iterator stands on vis[Pp]
                        uint32_t r = A.vis[Pp];
                        bool valid = (A.state[Pp] & 0x10) != 0;
                        S.iter_valid = valid; S.done = 1;
                        if (valid) { S.resume_len = A.klen[r]; for (uint32_t i = 0; i < A.klen[r] && i < P.resume_stride; i++) P.resume[(size_t)rq * P.resume_stride + i] = A.arena[(size_t)r * KS + i]; }
                    } else if (Ff < nvis_) { // reached a record outside the prefix (iterator invalid) or past the end (complete)
                        uint8_t s = A.state[Ff];
                        if (!(s & 0x10)) { S.iter_valid = 0; S.done = 1; }
                        else { S.complete = 1; S.iter_valid = 1; S.done = 1; }
                    } else {
                        // chunk fully consumed.  Did the limits run out exactly here?
                        bool ok = (S.count < Q.max_count) && (S.iter_count < Q.max_iter_count) && (Q.max_iter_size == 0 || S.size < Q.max_iter_size);
                        if (!ok) S.lookahead = 1; // need to know whether the iterator is still valid
                    }
                } else if (nvis_ > 0) { // look-ahead: the iterator stands on the first visible record
                    uint32_t r = A.vis[0];
                    bool valid = (A.state[0] & 0x10) != 0;
                    S.iter_valid = valid; S.done = 1;
                    if (valid) { S.resume_len = A.klen[r]; for (uint32_t i = 0; i < A.klen[r] && i < P.resume_stride; i++) P.resume[(size_t)rq * P.resume_stride + i] = A.arena[(size_t)r * KS + i]; }
                }
                if (!S.done) { // move every run's cursor past the consumed key range
                    bool any_more = false;
                    for (uint32_t j = 0; j < NR; j++) any_more |= S.more[j] != 0;
                    if (!any_more) { S.done = 1; S.iter_valid = 0; }
                }
                S.first_chunk = 0;
            }
            __syncthreads();
            SPT(11);
            if (!S.done) {
                // next chunk: forward: lower bound = this chunk's far bound (exclusive); cursors = first block whose
                // last key > bound.  reverse: upper bound = far bound (inclusive), cursor = first block with last key >= bound
                for (uint32_t i = tid; i < KS + 8; i += kScanThreads) { if (rev) khi[i] = klo[i]; else klo[i] = khi[i]; }
                if (tid == 0) {
                    if (rev) { S.hi_len = S.lo_len; S.has_hi = 1; S.hi_incl = 1; }
                    else { S.lo_len = S.hi_len; S.has_lo = 1; S.lo_incl = 0; }
                }
                __syncthreads();
                for (uint32_t j = warp; j < NR; j += kScanWarps) {
                    const RunDev &r = P.rr.runs[j];
                    uint32_t b;
                    if (rev) { // blocks after b hold only keys > bound; b itself may hold keys <= bound
                        b = warp_index_bound(r, khi, S.hi_len, lane, false);
                        if (b >= r.nb) b = r.nb ? r.nb - 1 : 0xFFFFFFFFu;
                        if (!r.nb) b = 0xFFFFFFFFu;
                    } else {
                        // forward: the bound is the smallest "last key of the last loaded block" over the runs, so the first
                        // block whose last key is > bound lies at or right behind this chunk's loaded blocks; their last keys
                        // are decoded in the arena -- no index search in global memory
                        const uint32_t m = S.nblk[j];
                        uint32_t cnt = 0;
                        for (uint32_t t0 = 0; t0 < m; t0 += 32) {
                            const uint32_t t = t0 + lane;
                            bool le = false;
                            if (t < m) {
                                const uint32_t tt = S.blk_base[j] + t, nrec = S.tb_nrec[tt];
                                const uint32_t rl = S.tb_rec[tt] + nrec - 1;
                                le = nrec == 0 || cmp_slots(A.arena + (size_t)rl * KS, A.klen[rl], klo, S.lo_len) <= 0;
                            }
                            cnt += __popc(__ballot_sync(kFull, le));
                        }
                        b = S.cur[j] + cnt; // a run without loaded blocks keeps its (exhausted) cursor
                    }
                    if (lane == 0) S.cur[j] = b;
                }
                __syncthreads();
            }
        }
        // ---- result -----------------------------------------------------------------------------------------------------------------
        if (tid == 0) {
            pgs_scan_result res;
            memset(&res, 0, sizeof res);
            res.status = S.error ? (int32_t)S.error : PGS_OK;
            res.n_kvs = S.n_out;
            res.count = S.count;
            res.iter_count = S.iter_count;
            res.expire_count = S.expire_count;
            res.filter_count = S.filter_count;
            res.size = S.size;
            res.complete = (uint8_t)S.complete;
            res.iter_valid = (uint8_t)S.iter_valid;
            res.resume_len = S.iter_valid ? S.resume_len : 0;
            res.arena_used = S.arena_used;
            P.results[rq] = res;
            if (S.error) atomicMax(P.error, S.error);
        }
        __syncthreads();
        SPT(12);
    }
}


// pack the per-request output slices densely so that one D2H copy brings a whole batch back
__global__ void k_pack_offsets(const pgs_scan_result *__restrict__ res, uint32_t n, unsigned long long *__restrict__ abase,
                               uint32_t *__restrict__ kbase)
{
    __shared__ uint32_t scratch[33];
    __shared__ unsigned long long carry_a;
    __shared__ uint32_t carry_k;
    if (threadIdx.x == 0) { carry_a = 0; carry_k = 0; }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t a = i < n ? (uint32_t)((res[i].arena_used + 15) & ~15ull) : 0;
        uint32_t k = i < n ? res[i].n_kvs : 0;
        uint32_t ta, tk;
        uint32_t pa = block_excl_scan(a, scratch, &ta);
        uint32_t pk = block_excl_scan(k, scratch, &tk);
        if (i < n) { abase[i] = carry_a + pa; kbase[i] = carry_k + pk; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_a += ta; carry_k += tk; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { abase[n] = carry_a; kbase[n] = carry_k; }
}
__global__ void k_pack_copy(const pgs_scan_result *__restrict__ res, uint32_t n, const uint8_t *__restrict__ arena,
                            unsigned long long arena_stride, const pgs_kv *__restrict__ kvs, uint32_t kv_stride,
                            const unsigned long long *__restrict__ abase, const uint32_t *__restrict__ kbase,
                            uint8_t *__restrict__ parena, pgs_kv *__restrict__ pkvs)
{
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        uint32_t chunks = (uint32_t)((res[i].arena_used + 15) >> 4);
        const uint4 *src = (const uint4 *)(arena + (size_t)i * arena_stride);
        uint4 *dst = (uint4 *)(parena + abase[i]);
        for (uint32_t c = threadIdx.x; c < chunks; c += blockDim.x) dst[c] = src[c];
        for (uint32_t k = threadIdx.x; k < res[i].n_kvs; k += blockDim.x) pkvs[kbase[i] + k] = kvs[(size_t)i * kv_stride + k];
    }
}

static uint64_t *g_crc_dev_rd[16] = {nullptr};
static std::mutex g_crc_rd_mu;
const uint64_t *crc64_table();
void set_last_read_stats(float ms, uint64_t probed, uint64_t skipped);

static int32_t snapshot_runs(Partition &part, std::vector<std::shared_ptr<Run>> &runs, ReadRuns &rr, uint32_t &KS,
                             const std::vector<std::shared_ptr<Run>> *pinned = nullptr)
{
    if (pinned) {
        runs = *pinned;
    } else {
        std::lock_guard<std::mutex> g(part.mu);
        runs = part.runs;
    }
    if (runs.size() > kMaxReadRuns) {
        set_error("read: %zu runs > %u (compact first)", runs.size(), kMaxReadRuns);
        return PGS_NOT_SUPPORTED;
    }
    rr.n = (uint32_t)runs.size();
    uint32_t mk = 0;
    for (uint32_t i = 0; i < rr.n; i++) { rr.runs[i] = runs[i]->dev(); mk = std::max(mk, runs[i]->info.max_ukey_len); }
    if (mk > kMaxUkeyLen) return PGS_NOT_SUPPORTED;
    KS = std::max(8u, (mk + 7) & ~7u);
    return PGS_OK;
}

// kernels of this file take their dynamic shared-memory size per launch; the opt-in maximum is set once per device here
// (a per-call cudaFuncSetAttribute would race between reader threads)
template <class K>
static cudaError_t allow_max_smem(K kernel, int max_smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - (int)a.sharedSizeBytes);
}
int32_t lookup_init_kernels(int max_smem)
{
    PGS_CUDA(allow_max_smem(k_scan, max_smem));
    PGS_CUDA(cudaFuncSetAttribute(k_scan, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    PGS_CUDA(allow_max_smem(k_get<8, false>, max_smem)); PGS_CUDA(allow_max_smem(k_get<8, true>, max_smem));
    PGS_CUDA(allow_max_smem(k_scan_fwd<8, false>, max_smem)); PGS_CUDA(allow_max_smem(k_scan_fwd<8, true>, max_smem));
    PGS_CUDA(allow_max_smem(k_scan_fwd<16, false>, max_smem)); PGS_CUDA(allow_max_smem(k_scan_fwd<16, true>, max_smem));
    PGS_CUDA(allow_max_smem(k_scan_fwd<32, false>, max_smem)); PGS_CUDA(allow_max_smem(k_scan_fwd<32, true>, max_smem));
    return PGS_OK;
}

// the runs of several partitions, packed for one launch (pgs_range_scan_many_multi)
struct ScanMulti {
    std::vector<RunDev> packed;
    std::vector<uint32_t> begin;
    const uint32_t *req_part;
};

static int32_t scan_launch(Partition &part, std::vector<std::shared_ptr<Run>> &runs, ScanParams &P, const ScanMulti *multi,
                           const pgs_scan_request *reqs, uint32_t n, uint32_t now, unsigned long long arena_stride, uint32_t kv_stride,
                           uint8_t *arena, uint64_t arena_cap, pgs_kv *kvs, uint64_t kv_cap, uint8_t *resume, uint32_t resume_stride,
                           pgs_scan_result *results, uint64_t *arena_base, uint32_t *kv_base);

int32_t scan_many(Partition &part, const pgs_scan_request *reqs, uint32_t n, uint32_t now, unsigned long long arena_stride,
                  uint32_t kv_stride, uint8_t *arena, uint64_t arena_cap, pgs_kv *kvs, uint64_t kv_cap, uint8_t *resume,
                  uint32_t resume_stride, pgs_scan_result *results, uint64_t *arena_base, uint32_t *kv_base,
                  const std::vector<std::shared_ptr<Run>> *pinned)
{
    std::vector<std::shared_ptr<Run>> runs;
    ScanParams P{};
    int32_t rc = snapshot_runs(part, runs, P.rr, P.KS, pinned);
    if (rc != PGS_OK) return rc;
    if (n == 0) return PGS_OK;
    return scan_launch(part, runs, P, nullptr, reqs, n, now, arena_stride, kv_stride, arena, arena_cap, kvs, kv_cap, resume, resume_stride,
                       results, arena_base, kv_base);
}

// forward scans over several partitions of one engine in one launch: request i reads partition slot req_part[i]
static int32_t scan_many_multi(pgs_partition *const *parts, uint32_t n_parts, const pgs_scan_request *reqs, const uint32_t *req_part, uint32_t n,
                               uint32_t now, unsigned long long arena_stride, uint32_t kv_stride, uint8_t *arena, uint64_t arena_cap, pgs_kv *kvs,
                               uint64_t kv_cap, uint8_t *resume, uint32_t resume_stride, pgs_scan_result *results, uint64_t *arena_base,
                               uint32_t *kv_base)
{
    Partition &part = parts[0]->p;
    std::vector<std::shared_ptr<Run>> runs; // every run a request may touch stays alive until the launch is done
    ScanParams P{};
    ScanMulti M;
    M.req_part = req_part;
    M.begin.push_back(0);
    P.KS = 8;
    uint32_t max_nr = 0;
    for (uint32_t p = 0; p < n_parts; p++) {
        Partition &pp = parts[p]->p;
        if (pp.eng != part.eng || pp.data_version != part.data_version) { set_error("range_scan_many_multi: partitions of different engines / data versions"); return PGS_INVALID_ARGUMENT; }
        std::vector<std::shared_ptr<Run>> rs;
        ReadRuns rr;
        uint32_t ks = 0;
        int32_t rc = snapshot_runs(pp, rs, rr, ks);
        if (rc != PGS_OK) return rc;
        for (uint32_t i = 0; i < rr.n; i++) M.packed.push_back(rr.runs[i]);
        M.begin.push_back((uint32_t)M.packed.size());
        runs.insert(runs.end(), rs.begin(), rs.end());
        P.KS = std::max(P.KS, ks);
        max_nr = std::max(max_nr, rr.n);
    }
    for (uint32_t i = 0; i < n; i++) {
        if (req_part[i] >= n_parts) { set_error("range_scan_many_multi: request %u names partition slot %u of %u", i, req_part[i], n_parts); return PGS_INVALID_ARGUMENT; }
        if (reqs[i].reverse) { set_error("range_scan_many_multi: reverse scans go through pgs_range_scan_many"); return PGS_NOT_SUPPORTED; }
    }
    if (n == 0) return PGS_OK;
    P.rr.n = max_nr;
    if (!M.packed.empty())
        for (uint32_t i = 0; i < kMaxReadRuns; i++) P.rr.runs[i] = M.packed[0]; // a valid dummy for idle groups
    return scan_launch(part, runs, P, &M, reqs, n, now, arena_stride, kv_stride, arena, arena_cap, kvs, kv_cap, resume, resume_stride, results,
                       arena_base, kv_base);
}

static int32_t scan_launch(Partition &part, std::vector<std::shared_ptr<Run>> &runs, ScanParams &P, const ScanMulti *multi,
                           const pgs_scan_request *reqs, uint32_t n, uint32_t now, unsigned long long arena_stride, uint32_t kv_stride,
                           uint8_t *arena, uint64_t arena_cap, pgs_kv *kvs, uint64_t kv_cap, uint8_t *resume, uint32_t resume_stride,
                           pgs_scan_result *results, uint64_t *arena_base, uint32_t *kv_base)
{
    Engine *e = part.eng;
    PGS_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->read_stream();
    // flatten requests
    std::vector<ScanReqDev> dev(n);
    std::string blob;
    bool need_crc = false, any_reverse = false;
    for (uint32_t i = 0; i < n; i++) {
        const pgs_scan_request &q = reqs[i];
        ScanReqDev &d = dev[i];
        memset(&d, 0, sizeof d);
        auto put = [&](const pgs_blob &b, uint32_t &off, uint32_t &len) {
            off = (uint32_t)blob.size();
            len = b.len;
            if (b.len) blob.append((const char *)b.data, b.len);
        };
        put(q.start, d.start_off, d.start_len);
        put(q.stop, d.stop_off, d.stop_len);
        put(q.hash_filter, d.hf_off, d.hf_len);
        put(q.sort_filter, d.sf_off, d.sf_len);
        d.start_inclusive = q.start_inclusive; d.stop_inclusive = q.stop_inclusive; d.reverse = q.reverse;
        d.no_value = q.no_value; d.key_mode = q.key_mode; d.return_expire_ts = q.return_expire_ts;
        d.count_only = q.count_only; d.validate_hash = q.validate_hash; d.prefix_same_as_start = q.prefix_same_as_start;
        d.has_upper = q.reserved[0]; // iterate_upper_bound (internal flag used by sortkey_count)
        d.hash_filter_type = q.hash_filter_type; d.sort_filter_type = q.sort_filter_type;
        d.max_count = q.max_count; d.max_iter_count = q.max_iter_count; d.max_iter_size = q.max_iter_size;
        d.pidx = q.pidx; d.partition_version = q.partition_version;
        need_crc |= q.validate_hash != 0;
        any_reverse |= q.reverse != 0;
    }
    blob.append(16, '\0');
    if (resume_stride < P.KS) resume_stride = 0; // caller gave no room: resume keys are not reported
    P.n = n; P.now = now; P.data_version = part.data_version;
    P.use_tma = (e->cfg.flags & PGS_ENGINE_NO_TMA) ? 0 : 1;
    P.kv_stride = kv_stride; P.arena_stride = (arena_stride + 15) & ~15ull; P.resume_stride = resume_stride ? resume_stride : P.KS;
    if (multi ? multi->packed.empty() : P.rr.n == 0) { // empty DB: every iterator is invalid from the start
        memset(results, 0, sizeof(pgs_scan_result) * n);
        if (arena_base) for (uint32_t i = 0; i <= n; i++) arena_base[i] = 0;
        if (kv_base) for (uint32_t i = 0; i <= n; i++) kv_base[i] = 0;
        return PGS_OK;
    }

    ScanReqDev *d_reqs = nullptr; uint8_t *d_blob = nullptr, *d_arena = nullptr, *d_resume = nullptr, *d_parena = nullptr;
    pgs_scan_result *d_res = nullptr; pgs_kv *d_kvs = nullptr, *d_pkvs = nullptr; uint32_t *d_err = nullptr, *d_kbase = nullptr;
    unsigned long long *d_abase = nullptr;
    RunDev *d_multi = nullptr;
    uint32_t *d_begin = nullptr, *d_part = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    auto cleanup = [&]() {
        if (d_multi) cudaFreeAsync(d_multi, st);
        if (d_begin) cudaFreeAsync(d_begin, st);
        if (d_part) cudaFreeAsync(d_part, st);
        cudaFreeAsync(d_reqs, st); cudaFreeAsync(d_blob, st); cudaFreeAsync(d_arena, st); cudaFreeAsync(d_resume, st);
        cudaFreeAsync(d_parena, st); cudaFreeAsync(d_res, st); cudaFreeAsync(d_kvs, st); cudaFreeAsync(d_pkvs, st);
        cudaFreeAsync(d_err, st); cudaFreeAsync(d_kbase, st); cudaFreeAsync(d_abase, st);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
    };
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr); } } while (0)
    CK(cudaEventCreate(&ev_a));
    CK(cudaEventCreate(&ev_b));
    CK(cudaMallocAsync(&d_reqs, sizeof(ScanReqDev) * n, st));
    CK(cudaMallocAsync(&d_blob, blob.size(), st));
    CK(cudaMallocAsync(&d_arena, P.arena_stride * n + 16, st));
    CK(cudaMallocAsync(&d_kvs, sizeof(pgs_kv) * (size_t)kv_stride * n + 16, st));
    CK(cudaMallocAsync(&d_resume, (size_t)P.resume_stride * n + 16, st));
    CK(cudaMallocAsync(&d_res, sizeof(pgs_scan_result) * n, st));
    CK(cudaMallocAsync(&d_err, 256, st));
    CK(cudaMemcpyAsync(d_reqs, dev.data(), sizeof(ScanReqDev) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_err, 0, 256, st));
    if (multi) {
        CK(cudaMallocAsync(&d_multi, sizeof(RunDev) * multi->packed.size(), st));
        CK(cudaMallocAsync(&d_begin, sizeof(uint32_t) * multi->begin.size(), st));
        CK(cudaMallocAsync(&d_part, sizeof(uint32_t) * n, st));
        CK(cudaMemcpyAsync(d_multi, multi->packed.data(), sizeof(RunDev) * multi->packed.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_begin, multi->begin.data(), sizeof(uint32_t) * multi->begin.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_part, multi->req_part, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
        P.multi_runs = d_multi; P.multi_begin = d_begin; P.req_part = d_part;
    }
    if (need_crc) {
        std::lock_guard<std::mutex> g(g_crc_rd_mu);
        int dv = e->device & 15;
        if (!g_crc_dev_rd[dv]) {
            uint64_t *t = nullptr;
            CK(cudaMalloc(&t, 2048));
            CK(cudaMemcpy(t, crc64_table(), 2048, cudaMemcpyHostToDevice));
            g_crc_dev_rd[dv] = t;
        }
        P.crc_table = (const unsigned long long *)g_crc_dev_rd[dv];
    }
    P.reqs = d_reqs; P.blob = d_blob; P.results = d_res; P.kvs = d_kvs; P.arena = d_arena; P.resume = d_resume; P.error = d_err;
    P.ticket = d_err + 8;
    const char *pt_env = getenv("PGS_PHASE_TIMING"); // diagnostics: per-phase cycle totals of the reverse kernel on stderr
    const bool phase_timing = any_reverse && pt_env && pt_env[0] == '1';
    P.phase_cycles = phase_timing ? (unsigned long long *)(d_err + 16) : nullptr;
    if (!any_reverse) {
        // ---- forward scans: lane-group merging iterators (read_kernels.cuh) ------------------------------------------
        const uint32_t NR = P.rr.n, G = NR <= 8 ? 8 : NR <= 16 ? 16 : 32;
        P.KS = (P.KS + 3) & ~3u;
        P.KSW = (P.KS + 8) / 4 + 1;
        P.group_smem = (uint32_t)((NR * (sizeof(CurState) + P.KSW * 4) + 3 * P.KSW * 4 + 15) & ~(size_t)15);
        const uint32_t dyn = 2048 + kMaxReadRuns * (uint32_t)sizeof(RunDev) + (kReadThreads / G) * P.group_smem;
        if (dyn > (uint32_t)e->max_smem_optin) { cleanup(); set_error("scan: %u runs with keys of %u bytes do not fit shared memory", NR, P.KS); return PGS_NOT_SUPPORTED; }
        auto kern = multi ? (G == 8 ? k_scan_fwd<8, true> : G == 16 ? k_scan_fwd<16, true> : k_scan_fwd<32, true>)
                          : (G == 8 ? k_scan_fwd<8, false> : G == 16 ? k_scan_fwd<16, false> : k_scan_fwd<32, false>);
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)kReadThreads, (size_t)dyn));
        const uint32_t per_cta = kReadThreads / G;
        const uint32_t grid = std::min<uint32_t>((n + per_cta - 1) / per_cta, (uint32_t)std::max(1, occ) * e->sm_count);
        CK(cudaEventRecord(ev_a, st));
        kern<<<grid, kReadThreads, dyn, st>>>(P);
        CK(cudaEventRecord(ev_b, st));
    } else {
        // ---- reverse scans: the block-staging kernel --------------------------------------------------------------------
        cudaFuncAttributes attr;
        CK(cudaFuncGetAttributes(&attr, k_scan));
        P.warp_scratch = 0;
        uint32_t fixed_dyn = (4 + (uint32_t)runs.size()) * ((P.KS + 8 + 15) & ~15u) + kScanWarps * P.warp_scratch;
        uint32_t max_blk = 0, max_rec = 0;
        for (auto &r : runs) { max_blk = std::max(max_blk, r->info.max_block_size); max_rec = std::max(max_rec, r->info.max_block_records); }
        uint64_t one = (((uint64_t)max_blk + 15) & ~15ull) + 32 + (uint64_t)max_rec * (P.KS + kScanRecExtra);
        uint64_t want = std::max<uint64_t>(one * std::max<size_t>(1, runs.size()) + 4096, 48 * 1024);
        if (n == 1) want = std::max<uint64_t>(want, 160 * 1024);
        uint64_t max_dyn = (uint64_t)e->max_smem_optin - attr.sharedSizeBytes - 256;
        uint64_t dyn = std::min<uint64_t>(max_dyn, fixed_dyn + want);
        if (dyn < fixed_dyn + one * std::max<size_t>(1, runs.size()) + 64) {
            cleanup();
            set_error("scan: blocks too large for shared memory");
            return PGS_NOT_SUPPORTED;
        }
        dyn &= ~127ull;
        P.pool_bytes = (uint32_t)(dyn - fixed_dyn);
        int occ = 0; // resident CTAs per SM for this dynamic shared-memory size (registers count too)
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan, (int)kScanThreads, (size_t)dyn));
        uint32_t grid = std::min<uint32_t>(n, (uint32_t)std::max(1, occ) * e->sm_count);
        CK(cudaEventRecord(ev_a, st));
        k_scan<<<grid, kScanThreads, dyn, st>>>(P);
        CK(cudaEventRecord(ev_b, st));
        if (phase_timing) {
            unsigned long long h[16] = {0};
            cudaMemcpyAsync(h, d_err + 16, sizeof h, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            static const char *names[13] = {"init", "choose", "stage", "farbound", "decode1", "decode2", "window", "rank", "visible", "loop", "emit", "advance", "result"};
            unsigned long long tot = 0;
            for (int i = 0; i < 13; i++) tot += h[i];
            fprintf(stderr, "[k_scan phases] requests=%u grid=%u dyn=%llu", n, grid, (unsigned long long)dyn);
            for (int i = 0; i < 13; i++) fprintf(stderr, " %s=%.1f%%", names[i], tot ? 100.0 * (double)h[i] / (double)tot : 0.0);
            fprintf(stderr, " cycles/request=%.0f\n", n ? (double)tot / n : 0.0);
        }
    }
    e->launches++;
    uint32_t herr = 0;
    if (n == 1) {
        CK(cudaMemcpyAsync(results, d_res, sizeof(pgs_scan_result), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (!herr) {
            if (results[0].arena_used > arena_cap || results[0].n_kvs > kv_cap) { cleanup(); return PGS_INCOMPLETE; }
            if (results[0].arena_used) CK(cudaMemcpyAsync(arena, d_arena, results[0].arena_used, cudaMemcpyDeviceToHost, st));
            if (results[0].n_kvs) CK(cudaMemcpyAsync(kvs, d_kvs, sizeof(pgs_kv) * results[0].n_kvs, cudaMemcpyDeviceToHost, st));
            if (results[0].iter_valid && resume && resume_stride) CK(cudaMemcpyAsync(resume, d_resume, results[0].resume_len, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        if (arena_base) { arena_base[0] = 0; arena_base[1] = results[0].arena_used; }
        if (kv_base) { kv_base[0] = 0; kv_base[1] = results[0].n_kvs; }
    } else {
        CK(cudaMallocAsync(&d_abase, sizeof(unsigned long long) * (n + 1), st));
        CK(cudaMallocAsync(&d_kbase, sizeof(uint32_t) * (n + 1), st));
        k_pack_offsets<<<1, 1024, 0, st>>>(d_res, n, d_abase, d_kbase);
        std::vector<unsigned long long> ab(n + 1);
        std::vector<uint32_t> kb(n + 1);
        CK(cudaMemcpyAsync(ab.data(), d_abase, sizeof(unsigned long long) * (n + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(kb.data(), d_kbase, sizeof(uint32_t) * (n + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(results, d_res, sizeof(pgs_scan_result) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        e->launches++;
        if (!herr) {
            if (ab[n] > arena_cap || kb[n] > kv_cap) { cleanup(); set_error("scan_many: output arena too small"); return PGS_INCOMPLETE; }
            CK(cudaMallocAsync(&d_parena, ab[n] + 16, st));
            CK(cudaMallocAsync(&d_pkvs, sizeof(pgs_kv) * ((size_t)kb[n] + 1), st));
            k_pack_copy<<<std::min<uint32_t>(n, 8 * e->sm_count), 128, 0, st>>>(d_res, n, d_arena, P.arena_stride, d_kvs, kv_stride, d_abase,
                                                                              d_kbase, d_parena, d_pkvs);
            e->launches++;
            if (ab[n]) CK(cudaMemcpyAsync(arena, d_parena, ab[n], cudaMemcpyDeviceToHost, st));
            if (kb[n]) CK(cudaMemcpyAsync(kvs, d_pkvs, sizeof(pgs_kv) * kb[n], cudaMemcpyDeviceToHost, st));
            if (resume && resume_stride) CK(cudaMemcpyAsync(resume, d_resume, (size_t)P.resume_stride * n, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
        if (arena_base) for (uint32_t i = 0; i <= n; i++) arena_base[i] = ab[i];
        if (kv_base) for (uint32_t i = 0; i <= n; i++) kv_base[i] = kb[i];
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_a, ev_b);
    set_last_read_stats(ms, 0, 0);
    cleanup();
#undef CK
    if (herr) {
        set_error("scan kernel failed with status %u", herr);
        // PGS_ABORTED = a request's output did not fit its arena / kv slice: the caller may retry with more room
        return herr == PGS_CORRUPTION ? PGS_CORRUPTION : (herr == PGS_NOT_SUPPORTED ? PGS_NOT_SUPPORTED : (herr == PGS_ABORTED ? PGS_ABORTED : PGS_IO_ERROR));
    }
    return PGS_OK;
}

} // namespace pgs

using namespace pgs;

// one get launch over the keys of one partition (parts[0], key_part null) or of several partitions of one engine
static int32_t get_batch_impl(pgs_partition *const *parts, uint32_t n_parts, const uint8_t *keys, const uint32_t *key_off, const uint32_t *key_part,
                              uint32_t n, uint32_t now, uint8_t *arena, uint64_t arena_cap, pgs_get_result *results, uint64_t *arena_used)
{
    Partition &part = parts[0]->p;
    Engine *e = part.eng;
    if (arena_used) *arena_used = 0;
    if (n == 0) return PGS_OK;
    std::vector<std::shared_ptr<Run>> runs; // every run a key may touch stays alive until the launch is done
    GetParams P{};
    std::vector<RunDev> packed;
    std::vector<uint32_t> begin;
    if (!key_part) {
        int32_t rc = snapshot_runs(part, runs, P.rr, P.KS);
        if (rc != PGS_OK) return rc;
    } else {
        P.rr.n = 0;
        P.KS = 8;
        begin.push_back(0);
        for (uint32_t p = 0; p < n_parts; p++) {
            Partition &pp = parts[p]->p;
            if (pp.eng != e || pp.data_version != part.data_version) { set_error("get_batch_multi: partitions of different engines / data versions"); return PGS_INVALID_ARGUMENT; }
            std::vector<std::shared_ptr<Run>> rs;
            ReadRuns rr;
            uint32_t ks = 0;
            int32_t rc = snapshot_runs(pp, rs, rr, ks);
            if (rc != PGS_OK) return rc;
            for (uint32_t i = 0; i < rr.n; i++) packed.push_back(rr.runs[i]);
            begin.push_back((uint32_t)packed.size());
            runs.insert(runs.end(), rs.begin(), rs.end());
            P.KS = std::max(P.KS, ks);
        }
        for (uint32_t i = 0; i < n; i++)
            if (key_part[i] >= n_parts) { set_error("get_batch_multi: key %u names partition slot %u of %u", i, key_part[i], n_parts); return PGS_INVALID_ARGUMENT; }
        if (!packed.empty()) P.rr.runs[0] = packed[0]; // a valid dummy for idle groups
    }
    if (key_part ? packed.empty() : P.rr.n == 0) {
        for (uint32_t i = 0; i < n; i++) { memset(&results[i], 0, sizeof results[i]); results[i].status = PGS_NOT_FOUND; }
        return PGS_OK;
    }
    PGS_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = e->read_stream();
    uint64_t key_bytes = key_off[n];
    uint8_t *d_keys = nullptr, *d_arena = nullptr;
    uint32_t *d_off = nullptr, *d_err = nullptr;
    pgs_get_result *d_res = nullptr;
    unsigned long long *d_cur = nullptr;
    RunDev *d_multi = nullptr;
    uint32_t *d_begin = nullptr, *d_part = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    auto cleanup = [&]() {
        cudaFreeAsync(d_keys, st); cudaFreeAsync(d_arena, st); cudaFreeAsync(d_off, st); cudaFreeAsync(d_err, st);
        cudaFreeAsync(d_res, st); cudaFreeAsync(d_cur, st);
        if (d_multi) cudaFreeAsync(d_multi, st);
        if (d_begin) cudaFreeAsync(d_begin, st);
        if (d_part) cudaFreeAsync(d_part, st);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
    };
#define CK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return cuda_fail(_e, #expr); } } while (0)
    CK(cudaEventCreate(&ev_a));
    CK(cudaEventCreate(&ev_b));
    CK(cudaMallocAsync(&d_keys, key_bytes + 16, st));
    CK(cudaMallocAsync(&d_off, sizeof(uint32_t) * (n + 1), st));
    CK(cudaMallocAsync(&d_res, sizeof(pgs_get_result) * n, st));
    CK(cudaMallocAsync(&d_arena, arena_cap + 16, st));
    CK(cudaMallocAsync(&d_cur, 32, st));
    CK(cudaMallocAsync(&d_err, 16, st));
    CK(cudaMemcpyAsync(d_keys, keys, key_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_off, key_off, sizeof(uint32_t) * (n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_cur, 0, 32, st));
    CK(cudaMemsetAsync(d_err, 0, 16, st));
    if (key_part) {
        CK(cudaMallocAsync(&d_multi, sizeof(RunDev) * packed.size(), st));
        CK(cudaMallocAsync(&d_begin, sizeof(uint32_t) * begin.size(), st));
        CK(cudaMallocAsync(&d_part, sizeof(uint32_t) * n, st));
        CK(cudaMemcpyAsync(d_multi, packed.data(), sizeof(RunDev) * packed.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_begin, begin.data(), sizeof(uint32_t) * begin.size(), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_part, key_part, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
        P.multi_runs = d_multi; P.multi_begin = d_begin; P.key_part = d_part;
    }
    P.keys = d_keys; P.key_off = d_off; P.n = n; P.now = now; P.data_version = part.data_version;
    P.results = d_res; P.arena = d_arena; P.arena_cap = arena_cap; P.arena_cursor = d_cur; P.error = d_err; P.ticket = d_err + 1;
    constexpr uint32_t G = 8;
    P.KS = (P.KS + 3) & ~3u;
    P.KSW = (P.KS + 8) / 4 + 1;
    P.group_smem = (uint32_t)((sizeof(CurState) + 2 * P.KSW * 4 + 15) & ~(size_t)15);
    const uint32_t dyn = kMaxReadRuns * (uint32_t)sizeof(RunDev) + (kReadThreads / G) * P.group_smem;
    if (dyn > (uint32_t)e->max_smem_optin) { cleanup(); return PGS_NOT_SUPPORTED; }
    int occ = 0;
    auto kern = key_part ? k_get<G, true> : k_get<G, false>;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)kReadThreads, (size_t)dyn));
    const uint32_t per_cta = kReadThreads / G;
    const uint32_t grid = std::min<uint32_t>((n + per_cta - 1) / per_cta, (uint32_t)std::max(1, occ) * e->sm_count);
    CK(cudaEventRecord(ev_a, st));
    kern<<<grid, kReadThreads, dyn, st>>>(P);
    CK(cudaEventRecord(ev_b, st));
    e->launches++;
    uint32_t herr = 0;
    unsigned long long cur3[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(results, d_res, sizeof(pgs_get_result) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cur3, d_cur, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&herr, d_err, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev_a, ev_b);
    set_last_read_stats(ms, cur3[1], cur3[2]);
    const unsigned long long used = cur3[0];
    if (arena_used) *arena_used = used;
    if (!herr && used) {
        CK(cudaMemcpyAsync(arena, d_arena, std::min<unsigned long long>(used, arena_cap), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    cleanup();
#undef CK
    if (herr) { set_error("get kernel failed with status %u", herr); return herr == PGS_CORRUPTION ? PGS_CORRUPTION : PGS_IO_ERROR; }
    return used > arena_cap ? PGS_INCOMPLETE : PGS_OK;
}

extern "C" int32_t pgs_get_batch(pgs_partition *ph, const uint8_t *keys, const uint32_t *key_off, uint32_t n, uint32_t now, uint8_t *arena,
                                 uint64_t arena_cap, pgs_get_result *results, uint64_t *arena_used)
{
    if (!ph || (n && (!keys || !key_off || !results))) return PGS_INVALID_ARGUMENT;
    return get_batch_impl(&ph, 1, keys, key_off, nullptr, n, now, arena, arena_cap, results, arena_used);
}

extern "C" int32_t pgs_get_batch_multi(pgs_partition *const *parts, uint32_t n_parts, const uint8_t *keys, const uint32_t *key_off,
                                       const uint32_t *key_part, uint32_t n, uint32_t now, uint8_t *arena, uint64_t arena_cap,
                                       pgs_get_result *results, uint64_t *arena_used)
{
    if (!parts || !n_parts || (n && (!keys || !key_off || !key_part || !results))) return PGS_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < n_parts; p++)
        if (!parts[p]) return PGS_INVALID_ARGUMENT;
    return get_batch_impl(parts, n_parts, keys, key_off, key_part, n, now, arena, arena_cap, results, arena_used);
}

extern "C" int32_t pgs_range_scan(pgs_partition *ph, const pgs_scan_request *req, uint32_t now, uint8_t *arena,
                                  uint64_t arena_cap, pgs_kv *kvs, uint32_t kv_cap, uint8_t *resume_key,
                                  uint32_t resume_cap, pgs_scan_result *out)
{
    if (!ph || !req || !out) return PGS_INVALID_ARGUMENT;
    return scan_many(ph->p, req, 1, now, arena_cap, kv_cap, arena, arena_cap, kvs, kv_cap, resume_key, resume_cap, out, nullptr, nullptr, nullptr);
}

extern "C" int32_t pgs_range_scan_many(pgs_partition *ph, const pgs_scan_request *reqs, uint32_t n, uint32_t now,
                                       uint64_t arena_stride, uint32_t kv_stride, uint8_t *arena, uint64_t arena_cap,
                                       pgs_kv *kvs, uint64_t kv_cap, uint8_t *resume_keys, uint32_t resume_stride,
                                       pgs_scan_result *results, uint64_t *arena_base, uint32_t *kv_base)
{
    if (!ph || (n && (!reqs || !results))) return PGS_INVALID_ARGUMENT;
    return scan_many(ph->p, reqs, n, now, arena_stride, kv_stride, arena, arena_cap, kvs, kv_cap, resume_keys, resume_stride, results,
                     arena_base, kv_base, nullptr);
}

extern "C" int32_t pgs_range_scan_many_multi(pgs_partition *const *parts, uint32_t n_parts, const pgs_scan_request *reqs, const uint32_t *req_part,
                                             uint32_t n, uint32_t now, uint64_t arena_stride, uint32_t kv_stride, uint8_t *arena,
                                             uint64_t arena_cap, pgs_kv *kvs, uint64_t kv_cap, uint8_t *resume_keys, uint32_t resume_stride,
                                             pgs_scan_result *results, uint64_t *arena_base, uint32_t *kv_base)
{
    if (!parts || !n_parts || (n && (!reqs || !req_part || !results))) return PGS_INVALID_ARGUMENT;
    for (uint32_t p = 0; p < n_parts; p++)
        if (!parts[p]) return PGS_INVALID_ARGUMENT;
    return scan_many_multi(parts, n_parts, reqs, req_part, n, now, arena_stride, kv_stride, arena, arena_cap, kvs, kv_cap, resume_keys,
                           resume_stride, results, arena_base, kv_base);
}
