// multi.cu -- several GPUs behind one call from one process (include/alll_b200.h: alll_multi_*).
//
// The reference scales inside ONE process through its constructor argument / CLI flag -p (SATInstance.h:51-56,259;
// example/main.cpp:56-61,76-84).  This layer gives the drop-in headers and the CLI the same shape over a list of GPUs:
// no process per GPU, no torch.distributed.  It owns one alll_handle per device slot and adds nothing to the device
// code: the clause-range sharded solve is alll_solve_p2p's fused NVLink exchange with the peers' exchange regions
// addressed directly (cudaDeviceEnablePeerAccess) instead of through CUDA IPC mappings; with one GPU per slot one host
// thread enqueues every GPU's persistent solve kernel and then waits for all of them.
//
// Host code only; every clause evaluation, independent-set decision and resample happens in the kernels behind the
// per-device handles.  No CPU fallback.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/alll_b200.h"
#include "alll_host.h"

using namespace alll;

namespace {
// One long-lived host thread per device slot (slot 0 is the caller's thread): the per-slot calls of an upload or a
// batch solve run side by side without creating threads per call -- a fresh thread pays the CUDA runtime's per-thread
// set-up on its first call, milliseconds that would otherwise land inside every end-to-end step.
struct SlotWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> job;
    bool has_job = false, quit = false;
    void loop()
    {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has_job || quit; });
            if (quit) return;
            lk.unlock();
            job();
            lk.lock();
            has_job = false;
            cv.notify_all();
        }
    }
    void submit(std::function<void()> f)
    {
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(f);
        has_job = true;
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !has_job; });
    }
};
} // namespace

struct alll_multi {
    std::vector<std::unique_ptr<SlotWorker>> workers;   // [slot - 1]
    std::vector<alll_handle> h;
    std::vector<int> dev;
    bool distinct = true;        // every slot has a GPU of its own -> persistent kernels, one launcher thread
    std::string err;
    uint32_t flags = 0;
    // instance
    bool has_instance = false, sharded = false;
    uint64_t n_vars = 0, m = 0, cap_records = 0;
    uint32_t k = 0;
    std::vector<uint64_t> lo, hi;
    uint32_t epoch = 0;
    // batch
    bool has_batch = false;
    uint32_t b_n_inst = 0;
    uint64_t b_n_vars = 0;
    std::vector<uint32_t> b_lo, b_hi;
    bool flag_ready = false;
};

namespace {

thread_local std::string g_multi_create_error;

int mfail(alll_multi_handle mh, int status, const std::string &msg)
{
    if (mh) mh->err = msg; else g_multi_create_error = msg;
    return status;
}

// (no C++ exception crosses the C boundary: see ALLL_GUARD in capi.cu)
int mguard_fail(alll_multi_handle mh, const char *what) noexcept
{
    try { return mfail(mh, ALLL_CUDA_ERROR, std::string("host-side failure: ") + what); } catch (...) { return ALLL_CUDA_ERROR; }
}
#define ALLL_MGUARD(mh)                                                                \
    catch (const std::exception &e__) { return mguard_fail((mh), e__.what()); }        \
    catch (...) { return mguard_fail((mh), "unknown exception"); }

// Runs fn(slot) for every slot, concurrently (one host thread per slot; slot 0 on the caller's thread).  Returns the
// first non-OK status (and records that slot's error text).
int for_all_slots(alll_multi_handle mh, uint32_t n, const std::function<int(uint32_t)> &fn, bool allow_max_rounds = false)
{
    std::vector<int> rc(n, ALLL_OK);
    for (uint32_t r = 1; r < n; r++) mh->workers[r - 1]->submit([&rc, &fn, r] { rc[r] = fn(r); });
    rc[0] = fn(0);
    for (uint32_t r = 1; r < n; r++) mh->workers[r - 1]->wait();
    for (uint32_t r = 0; r < n; r++)
        if (rc[r] != ALLL_OK && !(allow_max_rounds && rc[r] == ALLL_MAX_ROUNDS))
            return mfail(mh, rc[r], "device slot " + std::to_string(r) + ": " + alll_last_error(mh->h[r]));
    return ALLL_OK;
}

// contiguous, balanced ranges (sizes differ by at most one): the same split as sharded.py:partition
void split(uint64_t total, uint32_t parts, std::vector<uint64_t> &lo, std::vector<uint64_t> &hi)
{
    lo.assign(parts, 0); hi.assign(parts, 0);
    const uint64_t base = total / parts, rem = total % parts;
    uint64_t at = 0;
    for (uint32_t r = 0; r < parts; r++) {
        lo[r] = at;
        at += base + (r < rem ? 1 : 0);
        hi[r] = at;
    }
}

uint32_t slots_in_use(alll_multi_handle mh) { return mh->sharded ? (uint32_t)mh->h.size() : 1u; }

// ALLL_TRACE_HOST=1: host wall time of the stages of an upload / solve on stderr
struct HostTrace {
    const bool on = getenv("ALLL_TRACE_HOST") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[alll host] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

#define MNEED_INSTANCE()                                                                       \
    do {                                                                                       \
        if (!mh) return ALLL_BAD_ARG;                                                          \
        if (!mh->has_instance) return mfail(mh, ALLL_NO_INSTANCE, "no instance uploaded");    \
    } while (0)

#define MCALL(slot, call)                                                                      \
    do {                                                                                       \
        const int rc__ = (call);                                                               \
        if (rc__ != ALLL_OK)                                                                   \
            return mfail(mh, rc__, "device slot " + std::to_string(slot) + ": " + alll_last_error(mh->h[slot])); \
    } while (0)

} // namespace

extern "C" {

const char *alll_multi_last_error(alll_multi_handle mh) { return mh ? mh->err.c_str() : g_multi_create_error.c_str(); }

int alll_multi_create(const int32_t *devices, uint32_t n_devices, const alll_config *cfg, alll_multi_handle *out)
try {
    if (!out) return mfail(nullptr, ALLL_BAD_ARG, "out == NULL");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > MAX_SHARDS) return mfail(nullptr, ALLL_BAD_ARG, "need 1.." + std::to_string(MAX_SHARDS) + " devices");
    int n_dev = 0;
    const cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return mfail(nullptr, ALLL_CUDA_ERROR, std::string("no CUDA device: the solver has no CPU fallback (") + cudaGetErrorString(e) + ")");
    alll_multi *mh = new alll_multi;
    mh->flags = cfg ? cfg->flags : 0u;
    for (uint32_t r = 0; r < n_devices; r++) {
        if (devices[r] < 0 || devices[r] >= n_dev) { delete mh; return mfail(nullptr, ALLL_BAD_ARG, "device ordinal out of range"); }
        for (uint32_t q = 0; q < r; q++) mh->distinct = mh->distinct && devices[q] != devices[r];
        mh->dev.push_back(devices[r]);
    }
    // every GPU stores into every other GPU's exchange region and reads its flags: peer access both ways
    for (uint32_t a = 0; a < n_devices; a++)
        for (uint32_t b = 0; b < n_devices; b++) {
            if (mh->dev[a] == mh->dev[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, mh->dev[a], mh->dev[b]);
            if (!can) { delete mh; return mfail(nullptr, ALLL_CUDA_ERROR, "GPUs " + std::to_string(mh->dev[a]) + " and " + std::to_string(mh->dev[b]) + " have no peer access"); }
            cudaSetDevice(mh->dev[a]);
            const cudaError_t pe = cudaDeviceEnablePeerAccess(mh->dev[b], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
                delete mh;
                return mfail(nullptr, ALLL_CUDA_ERROR, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe));
            }
            cudaGetLastError();
        }
    for (uint32_t r = 0; r < n_devices; r++) {
        alll_config c{};
        if (cfg) c = *cfg;
        c.device = mh->dev[r];
        // one persistent kernel per rank needs every rank's kernel resident at once: only with a GPU per slot
        if (mh->distinct && n_devices > 1) c.flags |= ALLL_FLAG_P2P_PERSISTENT; else c.flags &= ~ALLL_FLAG_P2P_PERSISTENT;
        alll_handle h = nullptr;
        const int rc = alll_create(&c, &h);
        if (rc != ALLL_OK) {
            const std::string msg = alll_last_error(nullptr);
            for (alll_handle x : mh->h) alll_destroy(x);
            delete mh;
            return mfail(nullptr, rc, msg);
        }
        if (n_devices > 1) internal_h2d_pack_default_off(h);
        mh->h.push_back(h);
    }
    for (uint32_t r = 1; r < n_devices; r++) {
        mh->workers.emplace_back(new SlotWorker);
        SlotWorker *w = mh->workers.back().get();
        w->th = std::thread([w] { w->loop(); });
    }
    *out = mh;
    return ALLL_OK;
}
ALLL_MGUARD(nullptr)

int alll_multi_destroy(alll_multi_handle mh)
try {
    if (!mh) return ALLL_OK;
    for (auto &w : mh->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        w->th.join();
    }
    // borrowed winner words first (they point into slot 0's allocation)
    for (size_t r = mh->h.size(); r-- > 0;) alll_destroy(mh->h[r]);
    delete mh;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_upload_fixedk(alll_multi_handle mh, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit)
try {
    return alll_multi_upload_fixedk_streamed(mh, n_vars, m, k, lit, nullptr, nullptr);
}
ALLL_MGUARD(mh)

int alll_multi_upload_fixedk_streamed(alll_multi_handle mh, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit,
                                      alll_filled_fn filled, void *user)
try {
    if (!mh) return ALLL_BAD_ARG;
    mh->has_instance = false;
    const uint32_t n = (uint32_t)mh->h.size();
    mh->n_vars = n_vars; mh->m = m; mh->k = k;
    // the fused exchange carries fixed-size records {id, k literals}, k <= 8; tiny instances gain nothing from sharding
    mh->sharded = n > 1 && k >= 1 && k <= 8 && (m >= (uint64_t)n * 4096 || ((mh->flags & ALLL_FLAG_FORCE_SHARDING) && m >= n));
    if (!mh->sharded) {
        MCALL(0, internal_upload_fixedk_streamed(mh->h[0], n_vars, m, k, lit, filled, user, 0));
        MCALL(0, alll_set_id_base(mh->h[0], 0));
        mh->has_instance = true;
        return ALLL_OK;
    }
    HostTrace tr;
    split(m, n, mh->lo, mh->hi);
    // records one rank may publish per round: a random start violates about m_r / 2^k clauses; a quarter of the range
    // holds that for every k >= 3 with a wide margin (an adversarial start that violates more ends in ALLL_CAPACITY)
    const uint64_t widest = mh->hi[0] - mh->lo[0];
    mh->cap_records = (k >= 3 ? widest / 4 : widest) + 8192;
    // every device uploads ONLY its own clause range from the caller's buffer: N host->device copies side by side
    if (int rc = for_all_slots(mh, n, [&](uint32_t r) {
            // (streamed: device r's chunks become ready as the producer's fill position passes them; `filled` is called from
            // every slot thread)
            if (int e = internal_upload_fixedk_streamed(mh->h[r], n_vars, mh->hi[r] - mh->lo[r], k, lit + mh->lo[r] * k, filled, user, mh->lo[r])) return e;
            if (int e = alll_set_id_base(mh->h[r], mh->lo[r])) return e;
            return internal_p2p_create_local(mh->h[r], n, r, mh->cap_records);
        }))
        return rc;
    tr.mark("upload + region (all slots)");
    void *regions[MAX_SHARDS] = {};
    for (uint32_t r = 0; r < n; r++) regions[r] = internal_p2p_region(mh->h[r]);
    for (uint32_t r = 0; r < n; r++) MCALL(r, internal_p2p_connect_ptrs(mh->h[r], regions));
    tr.mark("link");
    mh->has_instance = true;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_upload_csr(alll_multi_handle mh, uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit)
try {
    if (!mh) return ALLL_BAD_ARG;
    if (!off) return mfail(mh, ALLL_BAD_ARG, "off == NULL");
    // uniform width k <= 8 is what the sharded exchange carries; everything else goes to the first device alone
    if (mh->h.size() > 1) {
        const uint64_t k0 = m ? off[1] - off[0] : 0;
        if (k0 >= 1 && k0 <= 8 && internal_csr_is_uniform(off, m))       // (all host threads; one device: alll_upload_csr scans itself)
            return alll_multi_upload_fixedk(mh, n_vars, m, (uint32_t)k0, lit + off[0]);
    }
    mh->has_instance = false;
    mh->sharded = false;
    mh->n_vars = n_vars; mh->m = m; mh->k = 0;
    MCALL(0, alll_upload_csr(mh->h[0], n_vars, m, off, lit));
    mh->has_instance = true;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_set_assignment(alll_multi_handle mh, const uint8_t *bools)
try {
    MNEED_INSTANCE();
    return for_all_slots(mh, slots_in_use(mh), [&](uint32_t r) { return alll_set_assignment(mh->h[r], bools); });
}
ALLL_MGUARD(mh)

int alll_multi_get_assignment(alll_multi_handle mh, uint8_t *bools)
try {
    MNEED_INSTANCE();
    MCALL(0, alll_get_assignment(mh->h[0], bools));          // the replicas are bit-identical
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_randomize(alll_multi_handle mh, uint64_t seed)
try {
    MNEED_INSTANCE();
    for (uint32_t r = 0; r < slots_in_use(mh); r++) MCALL(r, alll_randomize(mh->h[r], seed));
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_verify(alll_multi_handle mh, int *valid)
try {
    MNEED_INSTANCE();
    const uint32_t n = slots_in_use(mh);
    std::vector<int> ok(n, 0);
    if (int rc = for_all_slots(mh, n, [&](uint32_t r) { return alll_verify(mh->h[r], &ok[r]); })) return rc;
    if (valid) *valid = std::all_of(ok.begin(), ok.end(), [](int v) { return v != 0; }) ? 1 : 0;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_solve(alll_multi_handle mh, uint64_t seed, uint64_t max_rounds, alll_stats *stats)
try {
    MNEED_INSTANCE();
    if (!stats) return mfail(mh, ALLL_BAD_ARG, "stats == NULL");
    if (!mh->sharded) {
        const int rc = alll_solve(mh->h[0], seed, max_rounds, stats);
        if (rc != ALLL_OK && rc != ALLL_MAX_ROUNDS) return mfail(mh, rc, std::string("device slot 0: ") + alll_last_error(mh->h[0]));
        return rc;
    }
    const uint32_t n = (uint32_t)mh->h.size();
    HostTrace tr;
    mh->epoch++;
    std::vector<alll_stats> st(n);
    bool persistent = mh->distinct;
    for (uint32_t r = 0; r < n && persistent; r++) persistent = internal_p2p_persistent_possible(mh->h[r]);
    if (persistent) {
        // enqueue every GPU's persistent solve kernel, then collect.  (The kernels wait for each other's round flags on
        // the device; nothing here blocks until all of them are running.)
        // (enqueued from the slot threads side by side: launched one after the other from one thread, the last GPU's
        // kernel starts ~10 us per GPU after the first, and the first round of every GPU waits for it)
        std::vector<uint64_t> l0(n, 0);
        if (int rc = for_all_slots(mh, n, [&](uint32_t r) { return internal_solve_p2p_begin(mh->h[r], seed, max_rounds, mh->epoch, &l0[r]); }))
            return rc;
        tr.mark("solve: kernels enqueued");
        int worst = ALLL_OK;
        std::string msg;
        for (uint32_t r = 0; r < n; r++) {
            const int rc = internal_solve_p2p_end(mh->h[r], mh->m, l0[r], &st[r]);
            if (rc != ALLL_OK && rc != ALLL_MAX_ROUNDS && worst == ALLL_OK) { worst = rc; msg = "device slot " + std::to_string(r) + ": " + alll_last_error(mh->h[r]); }
        }
        tr.mark("solve: collected");
        if (worst != ALLL_OK) return mfail(mh, worst, msg);
    } else {
        // slots share GPUs (or the persistent kernel does not fit): one kernel per phase, one host thread per slot
        if (int rc = for_all_slots(mh, n, [&](uint32_t r) { return alll_solve_p2p(mh->h[r], seed, max_rounds, mh->m, mh->epoch, &st[r]); }, true))
            return rc;
    }
    *stats = st[0];                                          // identical on every rank (replicated trajectory)...
    stats->n_clause_evals = 0;
    stats->n_kernel_launches = 0;
    for (uint32_t r = 0; r < n; r++) {                       // ...except what is per device
        stats->solve_ms = std::max(stats->solve_ms, st[r].solve_ms);
        stats->sweep_ms = std::max(stats->sweep_ms, st[r].sweep_ms);
        stats->between_sweeps_ms = std::max(stats->between_sweeps_ms, st[r].between_sweeps_ms);
        stats->n_kernel_launches += st[r].n_kernel_launches;
        if (st[r].status != st[0].status || st[r].n_iterations != st[0].n_iterations || st[r].n_resamples != st[0].n_resamples)
            return mfail(mh, ALLL_CUDA_ERROR, "replicas diverged: device slots report different statistics");
    }
    if (st[0].n_incremental_rounds) {
        // incremental rounds evaluate only the clauses next to resampled variables: every rank counted its own range
        const uint64_t full = st[0].n_iterations - st[0].n_incremental_rounds;
        uint64_t evals = mh->m * full;
        for (uint32_t r = 0; r < n; r++) evals += st[r].n_clause_evals - mh->m * full;
        stats->n_clause_evals = evals;
    } else {
        stats->n_clause_evals = mh->m * st[0].n_iterations;
    }
    return stats->status;
}
ALLL_MGUARD(mh)

int alll_multi_info(alll_multi_handle mh, uint64_t info[4])
try {
    MNEED_INSTANCE();
    if (!info) return ALLL_BAD_ARG;
    info[0] = slots_in_use(mh);
    info[1] = mh->sharded ? 1 : 0;
    info[2] = mh->sharded ? mh->hi[0] - mh->lo[0] : mh->m;
    info[3] = mh->sharded ? mh->cap_records : 0;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_device_handle(alll_multi_handle mh, uint32_t i, alll_handle *out)
try {
    if (!mh || !out || i >= mh->h.size()) return ALLL_BAD_ARG;
    *out = mh->h[i];
    return ALLL_OK;
}
ALLL_MGUARD(mh)

// ---- batched small instances / seed portfolio over the device list -------------------------------------------------

int alll_multi_batch_upload(alll_multi_handle mh, uint32_t n_instances, uint64_t n_vars, uint32_t k, const uint64_t *clause_off,
                            const uint32_t *lit)
try {
    if (!mh) return ALLL_BAD_ARG;
    mh->has_batch = false;
    if (n_instances == 0 || !clause_off) return mfail(mh, ALLL_BAD_ARG, "no instances");
    const uint32_t n = (uint32_t)mh->h.size();
    mh->b_n_inst = n_instances; mh->b_n_vars = n_vars;
    if (n_instances == 1) {
        // one instance: the multi-GPU seed portfolio -- every device holds the instance
        mh->b_lo.assign(n, 0); mh->b_hi.assign(n, 1);
        if (int rc = for_all_slots(mh, n, [&](uint32_t r) { return alll_batch_upload(mh->h[r], 1, n_vars, k, clause_off, lit); })) return rc;
    } else {
        std::vector<uint64_t> lo, hi;
        split(n_instances, n, lo, hi);
        mh->b_lo.assign(lo.begin(), lo.end()); mh->b_hi.assign(hi.begin(), hi.end());
        // alll_batch_upload reads rows clause_off[0] .. clause_off[n_instances] of lit: a block is the same call on a
        // window of the offset array, nothing is copied on the host
        if (int rc = for_all_slots(mh, n, [&](uint32_t r) {
                if (hi[r] == lo[r]) return (int)ALLL_OK;
                return alll_batch_upload(mh->h[r], (uint32_t)(hi[r] - lo[r]), n_vars, k, clause_off + lo[r], lit);
            }))
            return rc;
    }
    if (!mh->flag_ready && n > 1) {                         // ONE first-SAT word for all devices, owned by slot 0
        uint8_t ipc[64];
        MCALL(0, alll_flag_create(mh->h[0], ipc));
        for (uint32_t r = 1; r < n; r++) MCALL(r, internal_flag_attach(mh->h[r], internal_flag_ptr(mh->h[0])));
        mh->flag_ready = true;
    }
    mh->has_batch = true;
    return ALLL_OK;
}
ALLL_MGUARD(mh)

int alll_multi_batch_solve(alll_multi_handle mh, uint32_t n_jobs, const uint64_t *seeds, uint64_t max_rounds, int portfolio,
                           uint8_t *assignments, alll_batch_stats *stats, int32_t *winner, double *device_ms)
try {
    if (!mh) return ALLL_BAD_ARG;
    if (!mh->has_batch) return mfail(mh, ALLL_NO_INSTANCE, "no batch uploaded");
    if (n_jobs == 0 || !seeds || !stats) return mfail(mh, ALLL_BAD_ARG, "n_jobs / seeds / stats");
    const uint32_t n = (uint32_t)mh->h.size();
    std::vector<uint64_t> lo, hi;
    if (portfolio) {
        if (mh->b_n_inst != 1) return mfail(mh, ALLL_BAD_ARG, "a multi-GPU portfolio needs the single instance uploaded (n_instances == 1)");
        split(n_jobs, n, lo, hi);
        if (n > 1) MCALL(0, alll_flag_reset(mh->h[0]));
    } else {
        if (n_jobs != mh->b_n_inst) return mfail(mh, ALLL_BAD_ARG, "n_jobs must equal the number of uploaded instances");
        if (mh->b_n_inst == 1) { lo.assign(n, 0); hi.assign(n, 0); hi[0] = 1; }
        else { lo.assign(mh->b_lo.begin(), mh->b_lo.end()); hi.assign(mh->b_hi.begin(), mh->b_hi.end()); }
    }
    std::vector<int32_t> win(n, -1);
    std::vector<double> ms(n, 0.0);
    const int mode = portfolio ? (n > 1 ? 2 : 1) : 0;
    if (int rc = for_all_slots(mh, n, [&](uint32_t r) {
            const uint32_t cnt = (uint32_t)(hi[r] - lo[r]);
            if (cnt == 0) return (int)ALLL_OK;
            if (int e = alll_batch_set_job_base(mh->h[r], portfolio ? (uint32_t)lo[r] : 0u)) return e;
            return alll_batch_solve(mh->h[r], cnt, seeds + lo[r], max_rounds, mode,
                                    assignments ? assignments + lo[r] * mh->b_n_vars : nullptr, stats + lo[r], &win[r], &ms[r]);
        }))
        return rc;
    if (device_ms) *device_ms = *std::max_element(ms.begin(), ms.end());
    if (winner) {
        *winner = win[0];
        if (mode == 2) {                                        // the one shared word, read after every device has finished
            int64_t w = -1;
            MCALL(0, alll_flag_read(mh->h[0], &w));
            *winner = (int32_t)w;
        }
    }
    return ALLL_OK;
}
ALLL_MGUARD(mh)

} // extern "C"
