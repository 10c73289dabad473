// CPU test of the request-batching front end's host logic (incubator_pegasus_b200/host/batcher.h): the Coalescer's windows and
// flush_gets' marshalling, with a stand-in for pgs_get_batch_multi that keeps its contract (values packed into the arena, a
// record that does not fit gets PGS_INCOMPLETE and *arena_used is the total need).  Built and run by tests/test_batcher.py.
#include <atomic>
#include <cassert>
#include <cstdio>
#include <map>
#include <thread>

#include "../../incubator_pegasus_b200/host/batcher.h"

using namespace pgs;

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "%s:%d: CHECK(%s) failed\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

struct FakeEngine { // slot -> key -> (value, expire_ts)
    std::vector<std::map<std::string, std::pair<std::string, uint32_t>>> parts;
    std::atomic<uint64_t> calls{0}, keys{0}, max_batch{0};
    int32_t fail_with = PGS_OK;
    std::mutex mu;
    std::vector<uint32_t> nows;
    int32_t operator()(pgs_partition *const *, uint32_t n_parts, const uint8_t *kb, const uint32_t *off, const uint32_t *slot, uint32_t n,
                       uint32_t now, uint8_t *arena, uint64_t cap, pgs_get_result *res, uint64_t *used)
    {
        calls++; keys += n;
        uint64_t m = max_batch.load();
        while (n > m && !max_batch.compare_exchange_weak(m, n)) {}
        { std::lock_guard<std::mutex> g(mu); nows.push_back(now); }
        std::this_thread::sleep_for(std::chrono::microseconds(300)); // a launch takes a while: followers pile up behind it
        if (fail_with != PGS_OK) return fail_with;
        CHECK(n_parts == parts.size());
        uint64_t cur = 0;
        for (uint32_t i = 0; i < n; i++) {
            memset(&res[i], 0, sizeof res[i]);
            CHECK(slot[i] < parts.size());
            std::string k((const char *)kb + off[i], off[i + 1] - off[i]);
            auto f = parts[slot[i]].find(k);
            if (f == parts[slot[i]].end()) { res[i].status = PGS_NOT_FOUND; continue; }
            const uint32_t ets = f->second.second;
            res[i].expire_ts = ets;
            if (ets && ets <= now) { res[i].status = PGS_NOT_FOUND; res[i].expired = 1; continue; }
            const std::string &v = f->second.first;
            res[i].value_len = (uint32_t)v.size();
            if (cur + v.size() <= cap) { memcpy(arena + cur, v.data(), v.size()); res[i].value_off = cur; res[i].status = PGS_OK; }
            else res[i].status = PGS_INCOMPLETE;
            cur += v.size();
        }
        *used = cur;
        return cur > cap ? PGS_INCOMPLETE : PGS_OK;
    }
};

static pgs_partition *fake_parts[3] = {(pgs_partition *)0x10, (pgs_partition *)0x20, (pgs_partition *)0x30};

struct Front {
    FakeEngine eng;
    struct Flush { Front *f; void operator()(std::vector<GetItem *> &it) const { flush_gets(fake_parts, 3, it, f->eng); } };
    Coalescer<GetItem, Flush> co;
    Front(size_t max_batch, uint32_t wait_us) : co(max_batch, wait_us, Flush{this}) { eng.parts.resize(3); }
    int32_t get(uint32_t slot, const std::string &key, uint32_t now, std::string &value, uint32_t cap, pgs_get_result &r)
    {
        value.assign(cap, '\0');
        GetItem it{};
        it.slot = slot; it.key = (const uint8_t *)key.data(); it.key_len = (uint32_t)key.size(); it.now = now;
        it.value = (uint8_t *)&value[0]; it.value_cap = cap;
        co.run(it);
        r = it.result;
        if (r.status == PGS_OK) value.resize(r.value_len);
        return it.rc;
    }
};

static std::string val_of(uint32_t slot, uint32_t i) { return std::string(10 + (i * 7 + slot) % 90, (char)('a' + (i + slot) % 26)) + std::to_string(i); }

static void fill(Front &f, uint32_t n)
{
    for (uint32_t s = 0; s < 3; s++)
        for (uint32_t i = 0; i < n; i++) f.eng.parts[s]["k" + std::to_string(i)] = {val_of(s, i), i % 5 == 0 ? 50u : 0u}; // every fifth expires at 50
}

static void test_concurrent_windows()
{
    Front f(16, 2000);
    fill(f, 200);
    const int T = 24, PER = 60;
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++)
        th.emplace_back([&, t] {
            for (int j = 0; j < PER; j++) {
                const uint32_t slot = (t + j) % 3, i = (uint32_t)(t * 131 + j * 17) % 260; // some keys do not exist
                std::string v;
                pgs_get_result r;
                const int32_t rc = f.get(slot, "k" + std::to_string(i), 100, v, 256, r);
                bool ok = rc == PGS_OK;
                if (i >= 200) ok = ok && r.status == PGS_NOT_FOUND && !r.expired;
                else if (i % 5 == 0) ok = ok && r.status == PGS_NOT_FOUND && r.expired && r.expire_ts == 50;
                else ok = ok && r.status == PGS_OK && v == val_of(slot, i);
                if (!ok) bad++;
            }
        });
    for (auto &x : th) x.join();
    uint64_t rq = 0, ln = 0;
    f.co.stats(&rq, &ln);
    CHECK(bad == 0);
    CHECK(rq == (uint64_t)T * PER && f.eng.keys == rq && f.eng.calls == ln);
    CHECK(ln * 3 < rq);            // windows were shared: far fewer launches than requests
    CHECK(f.eng.max_batch >= 8);
    printf("concurrent: %llu requests, %llu launches, largest window %llu\n", (unsigned long long)rq, (unsigned long long)ln,
           (unsigned long long)f.eng.max_batch.load());
}

static void test_no_wait_is_one_launch_per_call()
{
    Front f(16, 0);
    fill(f, 10);
    for (int i = 0; i < 7; i++) {
        std::string v;
        pgs_get_result r;
        CHECK(f.get(1, "k" + std::to_string(i + 1), 10, v, 128, r) == PGS_OK);
        CHECK(r.status == PGS_OK && v == val_of(1, i + 1));
    }
    uint64_t rq, ln;
    f.co.stats(&rq, &ln);
    CHECK(rq == 7 && ln == 7);
}

static void test_a_full_window_leaves_early()
{
    Front f(4, 5000000); // five seconds: only the size limit can end the window in time
    fill(f, 10);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++)
        th.emplace_back([&, t] { std::string v; pgs_get_result r; CHECK(f.get(0, "k" + std::to_string(t + 1), 10, v, 128, r) == PGS_OK && r.status == PGS_OK); });
    for (auto &x : th) x.join();
    CHECK(std::chrono::steady_clock::now() - t0 < std::chrono::seconds(3));
    CHECK(f.eng.calls == 1 && f.eng.max_batch == 4);
}

static void test_requests_keep_their_own_clock()
{
    Front f(8, 300000);
    fill(f, 10);
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int t = 0; t < 8; t++)
        th.emplace_back([&, t] {
            std::string v;
            pgs_get_result r;
            const uint32_t now = t % 2 ? 40 : 60; // k5 expires at 50
            f.get(2, "k5", now, v, 128, r);
            if (now == 40 ? !(r.status == PGS_OK && v == val_of(2, 5)) : !(r.status == PGS_NOT_FOUND && r.expired)) bad++;
        });
    for (auto &x : th) x.join();
    CHECK(bad == 0);
    for (uint32_t n : f.eng.nows) CHECK(n == 40 || n == 60);
    CHECK(f.eng.calls >= 2); // one launch per clock value at least
}

static void test_buffers_too_small()
{
    Front f(2, 0);
    f.eng.parts[0]["big"] = {std::string(10000, 'x'), 0};
    f.eng.parts[0]["small"] = {"s", 0};
    std::string v;
    pgs_get_result r;
    CHECK(f.get(0, "big", 1, v, 100, r) == PGS_OK);       // arena of the window (4096) too small: read once more with the need,
    CHECK(r.status == PGS_INCOMPLETE && r.value_len == 10000); // then the caller's own buffer is the limit and learns the length
    CHECK(f.eng.calls == 2);
    CHECK(f.get(0, "big", 1, v, 10000, r) == PGS_OK && r.status == PGS_OK && v == std::string(10000, 'x'));
    CHECK(f.get(0, "small", 1, v, 0, r) == PGS_OK && r.status == PGS_INCOMPLETE && r.value_len == 1);
}

static void test_a_failed_launch_reaches_every_caller()
{
    Front f(4, 200000);
    fill(f, 10);
    f.eng.fail_with = PGS_IO_ERROR;
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++)
        th.emplace_back([&, t] { std::string v; pgs_get_result r; if (f.get(0, "k1", 1, v, 64, r) != PGS_IO_ERROR || r.status != PGS_IO_ERROR) bad++; });
    for (auto &x : th) x.join();
    CHECK(bad == 0);
}

int main()
{
    test_no_wait_is_one_launch_per_call();
    test_a_full_window_leaves_early();
    test_requests_keep_their_own_clock();
    test_buffers_too_small();
    test_a_failed_launch_reaches_every_caller();
    test_concurrent_windows();
    printf("OK\n");
    return 0;
}
