// batcher.cpp — C ABI of the request-batching front end (include/pegasus_b200.h §9): a Coalescer (batcher.h) whose launch is
// pgs_get_batch_multi.  Nothing here computes an answer on the host.
#include <memory>

#include "batcher.h"

using namespace pgs;

struct pgs_batcher {
    std::vector<pgs_partition *> parts;
    struct Flush {
        pgs_batcher *b;
        void operator()(std::vector<GetItem *> &items) const { flush_gets(b->parts.data(), (uint32_t)b->parts.size(), items, pgs_get_batch_multi); }
    };
    std::unique_ptr<Coalescer<GetItem, Flush>> gets;
};

extern "C" {

int32_t pgs_batcher_open(pgs_partition *const *parts, uint32_t n_parts, uint32_t max_batch, uint32_t max_wait_us, pgs_batcher **out)
{
    if (!parts || !n_parts || !out) return PGS_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < n_parts; i++)
        if (!parts[i]) return PGS_INVALID_ARGUMENT;
    auto *b = new pgs_batcher;
    b->parts.assign(parts, parts + n_parts);
    b->gets.reset(new Coalescer<GetItem, pgs_batcher::Flush>(max_batch ? max_batch : 4096, max_wait_us, pgs_batcher::Flush{b}));
    *out = b;
    return PGS_OK;
}

void pgs_batcher_close(pgs_batcher *b) { delete b; }

int32_t pgs_batcher_get(pgs_batcher *b, uint32_t part_slot, const uint8_t *key, uint32_t key_len, uint32_t now, uint8_t *value,
                        uint32_t value_cap, pgs_get_result *result)
{
    if (!b || !key || !result || (value_cap && !value) || part_slot >= b->parts.size()) return PGS_INVALID_ARGUMENT;
    GetItem it{};
    it.slot = part_slot; it.key = key; it.key_len = key_len; it.now = now; it.value = value; it.value_cap = value_cap;
    b->gets->run(it);
    *result = it.result;
    return it.rc;
}

void pgs_batcher_stats(pgs_batcher *b, uint64_t *requests, uint64_t *launches)
{
    if (b) b->gets->stats(requests, launches);
}

} // extern "C"
