// packpipe.h -- host side of the packed host-to-device transport: the pack threads and their ring discipline.
// Plain C++ (no CUDA), so that tests/cpp/packpipe_check.cpp can drive it on a CPU against a simulated link.
#pragma once

#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

namespace alll {

constexpr int PACK_SLOTS = 4;                     // chunks in the page-locked ring (and in its device mirror)
constexpr uint64_t PACK_UNIT_ROWS = 16384;        // clauses a pack thread takes at a time (a multiple of 32)

// hostpack.cpp
uint32_t host_pack25(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi);

// A large host-buffer upload is bound by the PCIe transfer of its 32-bit literal words (cfg4: 23 of 27 ms).  With
// n_vars <= 2^24 a literal has 25 significant bits, so the host threads of the upload re-pack every chunk (hostpack.cpp:
// 3 bytes + 1 bit per literal) into a page-locked ring while earlier chunks are on the link, and unpack25_kernel expands
// the chunk into the staging buffer the layout passes read -- 0.78 of the bytes cross the link, and a pageable caller
// buffer is read by all host threads instead of by the driver's single staging copy.  The main thread stays the only one
// that talks to CUDA: it sends chunk i packed as soon as its units are done, or -- from page-locked memory -- as it is
// whenever the link has run dry before that (the packers skip what is left of that chunk), so a host that cannot pack
// at the link's rate falls back chunk by chunk instead of slowing the upload down.
struct PackPipe {
    struct Unit { uint32_t chunk; uint64_t r0, r1; };
    const uint32_t *src;
    const uint32_t k;
    const std::vector<uint64_t> &cut;
    uint8_t *const ring;
    const size_t slot_bytes, hi_off;
    std::vector<Unit> units;
    std::atomic<size_t> next{0};
    std::atomic<uint32_t> released{0};                    // chunks whose copy has completed: chunk j may be packed iff j < released + PACK_SLOTS
    std::atomic<int> quit{0};
    std::unique_ptr<std::atomic<uint32_t>[]> left;        // per chunk: units still to be packed (or skipped)
    std::unique_ptr<std::atomic<uint8_t>[]> raw;          // per chunk: 1 = the main thread sends it as it is
    std::atomic<uint32_t> or_acc{0};                      // OR of every packed literal (bits above 24 = not representable)
    std::vector<std::thread> th;
    std::thread lead;

    PackPipe(const uint32_t *src_, uint32_t k_, const std::vector<uint64_t> &cut_, uint8_t *ring_, size_t slot_bytes_, size_t hi_off_, uint32_t n_threads)
        : src(src_), k(k_), cut(cut_), ring(ring_), slot_bytes(slot_bytes_), hi_off(hi_off_)
    {
        const size_t n_chunks = cut.size() - 1;
        left.reset(new std::atomic<uint32_t>[n_chunks]);
        raw.reset(new std::atomic<uint8_t>[n_chunks]);
        for (size_t c = 0; c < n_chunks; c++) {
            uint32_t cnt = 0;
            for (uint64_t r = cut[c]; r < cut[c + 1]; r += PACK_UNIT_ROWS, ++cnt) units.push_back(Unit{(uint32_t)c, r, std::min(r + PACK_UNIT_ROWS, cut[c + 1])});
            left[c].store(cnt);
            raw[c].store(0);
        }
        // (the first pack thread starts the others: the caller gets back to issuing chunk 0 after one thread creation, not n)
        th.reserve(n_threads);
        lead = std::thread([this, n_threads] {
            for (uint32_t t = 1; t < n_threads; t++) th.emplace_back([this] { work(); });
            work();
        });
    }
    ~PackPipe()
    {
        quit.store(1);
        lead.join();                                      // (th is complete once the lead thread has ended)
        for (auto &x : th) x.join();
    }
    PackPipe(const PackPipe &) = delete;
    PackPipe &operator=(const PackPipe &) = delete;

    // Main thread, when it needs chunk ci (chunks are issued in order): true = every unit of the chunk is packed in its
    // ring slot, false = the chunk goes as it is (only if may_go_raw: the link ran dry before the packers were done; they
    // skip what is left of it).  poll(): chunks whose copy has completed so far (that frees their ring slots).
    template <class Poll> bool wait_chunk(uint32_t ci, bool may_go_raw, Poll &&poll)
    {
        for (;;) {
            const uint32_t retired = poll();
            released.store(retired, std::memory_order_release);
            if (left[ci].load(std::memory_order_acquire) == 0) return true;
            if (may_go_raw && retired == ci) { raw[ci].store(1, std::memory_order_release); return false; }
            std::this_thread::yield();
        }
    }

    void work()
    {
        for (;;) {
            const size_t u = next.fetch_add(1);
            if (u >= units.size() || quit.load()) return;
            const Unit &un = units[u];
            // the ring slot is free once the copy of the chunk that used it last has completed AND that chunk's packers have
            // all left it (a chunk sent as it is may still have a packer finishing a unit nobody will read)
            for (;;) {
                if (quit.load()) return;
                if (raw[un.chunk].load(std::memory_order_acquire)) break;
                if (un.chunk < released.load(std::memory_order_acquire) + (uint32_t)PACK_SLOTS &&
                    (un.chunk < (uint32_t)PACK_SLOTS || left[un.chunk - PACK_SLOTS].load(std::memory_order_acquire) == 0))
                    break;
                std::this_thread::yield();
            }
            if (!raw[un.chunk].load(std::memory_order_acquire)) {
                uint8_t *slot = ring + (size_t)(un.chunk % PACK_SLOTS) * slot_bytes;
                const size_t l0 = (size_t)(un.r0 - cut[un.chunk]) * k, n = (size_t)(un.r1 - un.r0) * k;
                const uint32_t o = host_pack25(src + un.r0 * k, n, slot + 3 * l0, slot + hi_off + l0 / 8);
                or_acc.fetch_or(o, std::memory_order_relaxed);
            }
            left[un.chunk].fetch_sub(1, std::memory_order_release);
        }
    }
};


} // namespace alll
