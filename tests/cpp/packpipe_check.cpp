// packpipe_check.cpp -- CPU check of the pack threads' ring discipline (csrc/packpipe.h) against a simulated link: a "DMA"
// thread reads each issued chunk's ring slot some time AFTER it was issued (as the copy engine does), expands it with the
// device kernel's arithmetic and only then retires the chunk.  Every literal must arrive intact whatever the relative speed
// of packers and link, with and without the as-it-is fallback; compiled with hostpack.cpp by tests/test_boundary.py.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>

#include "packpipe.h"

using namespace alll;

static void unpack(const uint8_t *lo3, const uint8_t *hi, uint32_t *out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const uint8_t *b = lo3 + 3 * i;
        out[i] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)((hi[i / 8] >> (i % 8)) & 1u) << 24);
    }
}

static int run(uint64_t m, uint32_t k, uint64_t chunk_rows, uint32_t n_threads, int link_us, bool may_go_raw, uint64_t seed)
{
    std::vector<uint32_t> src(m * k), staging(m * k, 0xFFFFFFFFu);
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
    for (auto &v : src) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s >> 11) & 0x1FFFFFFu; }
    std::vector<uint64_t> cut;
    for (uint64_t c = 0; c < m; c += chunk_rows) cut.push_back(c);
    cut.push_back(m);
    const size_t n_chunks = cut.size() - 1;
    uint64_t widest = 0;
    for (size_t i = 0; i < n_chunks; i++) widest = std::max(widest, cut[i + 1] - cut[i]);
    const size_t hi_off = (3 * widest * k + 63) / 64 * 64, slot_bytes = hi_off + ((widest * k + 7) / 8 + 63) / 64 * 64;
    uint8_t *ring = (uint8_t *)aligned_alloc(64, slot_bytes * PACK_SLOTS);
    std::memset(ring, 0xEE, slot_bytes * PACK_SLOTS);

    // the simulated link: issued chunks complete in order, each link_us after the previous one (or after its issue)
    struct Job { uint32_t chunk; bool packed; };
    std::deque<Job> q;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint32_t> retired{0};
    bool done = false;
    std::thread link([&] {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !q.empty() || done; });
                if (q.empty()) return;
                j = q.front();
                q.pop_front();
            }
            if (link_us) std::this_thread::sleep_for(std::chrono::microseconds(link_us));
            const size_t n_l = (cut[j.chunk + 1] - cut[j.chunk]) * k;
            uint32_t *dst = staging.data() + cut[j.chunk] * k;
            if (j.packed) {
                const uint8_t *slot = ring + (size_t)(j.chunk % PACK_SLOTS) * slot_bytes;
                unpack(slot, slot + hi_off, dst, n_l);
            } else {
                std::memcpy(dst, src.data() + cut[j.chunk] * k, n_l * 4);
            }
            retired.fetch_add(1, std::memory_order_release);
        }
    });
    uint32_t n_packed = 0, n_raw = 0;
    {
        PackPipe pipe(src.data(), k, cut, ring, slot_bytes, hi_off, n_threads);
        for (uint32_t ci = 0; ci < n_chunks; ci++) {
            const bool packed = pipe.wait_chunk(ci, may_go_raw, [&] { return retired.load(std::memory_order_acquire); });
            (packed ? n_packed : n_raw)++;
            { std::lock_guard<std::mutex> lk(mu); q.push_back(Job{ci, packed}); }
            cv.notify_all();
        }
        { std::lock_guard<std::mutex> lk(mu); done = true; }
        cv.notify_all();
        link.join();
        if (pipe.or_acc.load() >> 25) { printf("OR word has bits above 24\n"); return 1; }
    }
    free(ring);
    if (!may_go_raw && n_raw) { printf("a chunk went as it is although that was not allowed\n"); return 1; }
    for (size_t i = 0; i < src.size(); i++)
        if (staging[i] != src[i]) {
            printf("m=%llu k=%u chunk=%llu threads=%u link=%dus raw=%d: literal %zu arrived as %x, was %x (%u packed, %u as they are)\n",
                   (unsigned long long)m, k, (unsigned long long)chunk_rows, n_threads, link_us, (int)may_go_raw, i, staging[i], src[i], n_packed, n_raw);
            return 1;
        }
    printf("m=%llu k=%u chunks=%zu threads=%u link=%dus may_go_raw=%d: ok (%u packed, %u as they are)\n", (unsigned long long)m, k, n_chunks, n_threads,
           link_us, (int)may_go_raw, n_packed, n_raw);
    return 0;
}

int main()
{
    int bad = 0;
    uint64_t seed = 1;
    for (int link_us : {0, 200, 3000})
        for (int raw = 0; raw < 2; raw++)
            for (uint32_t threads : {1u, 3u, 8u}) {
                bad |= run(600000, 8, 65536, threads, link_us, raw != 0, seed++);       // 10 chunks of 4 units: the ring wraps twice
                bad |= run(200003, 7, 16384 * 3, threads, link_us, raw != 0, seed++);   // ragged tail (literal count not a multiple of 8)
            }
    bad |= run(100, 3, 65536, 4, 0, true, seed++);                                       // one tiny chunk
    bad |= run(16384 * 32, 5, 16384, 6, 50, true, seed++);                               // 32 chunks of one unit
    if (!bad) printf("packpipe ok\n");
    return bad;
}
