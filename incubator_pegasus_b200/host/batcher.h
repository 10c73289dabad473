// batcher.h — request-batching front end for point reads: the rrdb read handlers run on a pool of host
// threads (THREAD_POOL_LOCAL_APP, src/server/config.ini:140-150), one blocking call per RPC; the engine wants thousands of
// keys per launch.  A Coalescer lets the calls that arrive within a short window share one pgs_get_batch_multi launch.
//
// No dispatcher thread: the first caller of a window leads it -- it waits until the window is full or its time is up, takes
// everything queued, runs the launch for all of them and wakes them; callers that arrive while a launch is in flight open the
// next window with a leader of their own, so launches of consecutive windows overlap.  Host-only code (no CUDA headers): the
// launch is a callable with pgs_get_batch_multi's signature, which is what tests/test_batcher.py instantiates with a stand-in
// to exercise the windows, the marshalling and the error paths on the CPU; the product (batcher.cpp) passes pgs_get_batch_multi.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pegasus_b200.h"

namespace pgs {

// Waiters of one window share one launch.  Flush: void(std::vector<Item *> &), called without the lock held, must fill every item.
template <class Item, class Flush>
class Coalescer
{
public:
    Coalescer(size_t max_batch, uint32_t max_wait_us, Flush flush)
        : max_batch_(max_batch ? max_batch : 1), max_wait_(std::chrono::microseconds(max_wait_us)), flush_(std::move(flush))
    {
    }
    // blocks until `it` has been handled by some window's launch
    void run(Item &it)
    {
        std::unique_lock<std::mutex> lk(mu_);
        Waiter w{&it, false};
        queue_.push_back(&w);
        requests_++;
        if (has_leader_) { // follow: the leader of this window takes the item along
            if (queue_.size() >= max_batch_) full_.notify_one();
            done_.wait(lk, [&] { return w.done; });
            return;
        }
        has_leader_ = true; // lead this window
        if (queue_.size() < max_batch_ && max_wait_.count() > 0)
            full_.wait_for(lk, max_wait_, [&] { return queue_.size() >= max_batch_; });
        std::vector<Waiter *> mine;
        mine.swap(queue_);
        has_leader_ = false; // the next arrival opens the next window while this one is in flight
        launches_++;
        lk.unlock();
        std::vector<Item *> items;
        items.reserve(mine.size());
        for (Waiter *x : mine) items.push_back(x->item);
        flush_(items);
        lk.lock();
        for (Waiter *x : mine) x->done = true;
        done_.notify_all();
    }
    void stats(uint64_t *requests, uint64_t *launches)
    {
        std::lock_guard<std::mutex> g(mu_);
        if (requests) *requests = requests_;
        if (launches) *launches = launches_;
    }

private:
    struct Waiter {
        Item *item;
        bool done;
    };
    const size_t max_batch_;
    const std::chrono::microseconds max_wait_;
    Flush flush_;
    std::mutex mu_;
    std::condition_variable full_, done_;
    std::vector<Waiter *> queue_;
    bool has_leader_ = false;
    uint64_t requests_ = 0, launches_ = 0;
};

// one on_get-shaped request: the value comes back in the caller's buffer (value_len = the whole length even when it did not fit)
struct GetItem {
    uint32_t slot;
    const uint8_t *key;
    uint32_t key_len, now;
    uint8_t *value;
    uint32_t value_cap;
    pgs_get_result result; // status / expire_ts / expired / value_len as pgs_get_batch; value_off is meaningless to the caller
    int32_t rc;            // the launch's own status (PGS_OK, or the engine's failure for the whole window)
};

// Marshals a window into pgs_get_batch_multi-shaped calls (one per distinct `now`: the TTL check needs the request's own clock;
// a window spans microseconds, so that is one call in practice) and scatters results and values back.  Call = the launch.
template <class Call>
void flush_gets(pgs_partition *const *parts, uint32_t n_parts, std::vector<GetItem *> &items, Call &&call)
{
    std::vector<char> taken(items.size(), 0);
    for (size_t first = 0; first < items.size(); first++) {
        if (taken[first]) continue;
        const uint32_t now = items[first]->now;
        std::vector<size_t> idx;
        for (size_t i = first; i < items.size(); i++)
            if (!taken[i] && items[i]->now == now) { taken[i] = 1; idx.push_back(i); }
        std::vector<uint8_t> keys;
        std::vector<uint32_t> off(idx.size() + 1, 0), slot(idx.size());
        for (size_t j = 0; j < idx.size(); j++) {
            const GetItem &g = *items[idx[j]];
            keys.insert(keys.end(), g.key, g.key + g.key_len);
            off[j + 1] = (uint32_t)keys.size();
            slot[j] = g.slot;
        }
        keys.resize(keys.size() + 16); // the engine reads whole words
        std::vector<pgs_get_result> res(idx.size());
        uint64_t cap = 0, used = 0;
        for (size_t j = 0; j < idx.size(); j++) cap += items[idx[j]]->value_cap;
        cap = cap < 4096 ? 4096 : cap;
        std::vector<uint8_t> arena;
        int32_t rc = PGS_OK;
        for (int attempt = 0; attempt < 2; attempt++) { // a window whose values outgrow the callers' buffers is read once more
            arena.resize(cap);
            used = 0;
            rc = call(parts, n_parts, keys.data(), off.data(), slot.data(), (uint32_t)idx.size(), now, arena.data(), cap, res.data(), &used);
            if (rc != PGS_INCOMPLETE || used <= cap) break;
            cap = used;
        }
        for (size_t j = 0; j < idx.size(); j++) {
            GetItem &g = *items[idx[j]];
            memset(&g.result, 0, sizeof g.result);
            g.rc = (rc == PGS_OK || rc == PGS_INCOMPLETE) ? PGS_OK : rc;
            if (g.rc != PGS_OK) { g.result.status = rc; continue; } // the launch failed: every request of the window sees it
            g.result = res[j]; // NOT_FOUND / expired / INCOMPLETE (the arena was still too small) as the engine reported them
            if (res[j].status == PGS_OK) {
                if (res[j].value_len > g.value_cap) g.result.status = PGS_INCOMPLETE; // the caller's buffer is too small: value_len tells the need
                else if (res[j].value_len) memcpy(g.value, arena.data() + res[j].value_off, res[j].value_len);
            }
        }
    }
}

} // namespace pgs
