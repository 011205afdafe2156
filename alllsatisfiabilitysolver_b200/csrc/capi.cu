// capi.cu -- the extern "C" boundary (include/alll_b200.h) and the host-side round loop.
//
// Host control flow replaces SATInstance::solve -> parallel_solve (SATInstance.h:60-66, :217-320).  There is
// no CPU compute path in this file: every clause evaluation, independent-set decision and resample happens
// in the kernels of persist.cu / sweep.cu / mis.cu / layout.cu.
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/alll_b200.h"
#include "alll_host.h"
#include "csr_body.cuh"
#include "packpipe.h"

using namespace alll;

namespace {

thread_local std::string g_create_error = "";

constexpr uint32_t DEFAULT_SWEEP_SMEM = 192u * 1024u;
constexpr int MAX_TIMED_ROUNDS = 256;
constexpr uint64_t URECORD_CAP = 16384;  // violated-clause records are written for violated sets up to this size
constexpr int ROUNDS_IN_FLIGHT = 3;      // rounds the host enqueues ahead of the last one it has seen retire
constexpr uint64_t PACK_MIN_LITERALS = 32ull << 20;   // automatic mode: uploads of at least 128 MB of literals

} // namespace

struct alll_solver {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of a host-buffer upload, chunk by chunk ahead of the layout kernels
    cudaEvent_t ev_chunk[33] = {};       // [i]: chunk i has arrived; [32]: the layout stream is done with the staging buffer
    int sm_count = 0;
    int clock_khz = 0;                   // SM clock (kHz) for turning a time-out into clock64 ticks
    uint32_t smem_budget = DEFAULT_SWEEP_SMEM;
    uint32_t flags = 0;
    uint32_t tune = 0;                   // TUNE_* measurement knobs (environment ALLL_TUNE, read at alll_create)
    std::string err;
    uint64_t launches = 0;

    // instance
    bool has_instance = false;
    uint64_t n_vars = 0, m = 0, m_pad = 0;
    uint32_t k = 0;                      // 0 = CSR
    uint32_t kmax = 0;                   // widest clause
    uint32_t n_words_alloc = 0, bucket_words = 0, n_buckets = 1, n_tiles = 0;
    uint32_t n_segs = 1;                 // bucket segments of the slot order: upload chunks x buckets (BucketSeg)
    bool resident_all = true;
    uint32_t min_resident = 0;           // measured by the bucketing pass; selects the sweep specialisation
    uint32_t resident_cap = RESIDENT_CAP; // literals per clause placed as bucket-resident (compile-time choice: 2 / 3 / 4 were measured, profiles/)
    uint32_t *d_planes = nullptr, *d_orig_id = nullptr;
    uint32_t *d_packed = nullptr;        // packed eager planes [4][m_pad] (alll_device.cuh: EagerPack), streamed by the sweep when packed_on
    bool packed_on = false;
    bool rows8_on = false;               // d_rows holds the row-major copy [m_pad][8] (5 <= k <= 8): the sweep's tail fetch and the independent-set gather read it
    BucketSeg *d_segs = nullptr;
    BucketSeg *d_sweep_segs = nullptr;   // the non-empty segments in sweep order (bucket-major over the chunk-major slot order)
    uint32_t n_sweep_segs = 1;
    SweepRun *d_runs = nullptr;          // per-CTA run lists of the sweep (SweepRun), cut for sweep_grid CTAs
    uint32_t *d_run_begin = nullptr;
    uint64_t *d_off = nullptr;
    uint32_t *d_csr_lit = nullptr;       // [csr_l_pad]: the literal array, padded to a multiple of 128 (>= 1 padding position)
    uint64_t n_lit = 0, csr_l_pad = 0;
    uint32_t *d_csr_start = nullptr, *d_csr_rank = nullptr;   // clause-start bits, clause rank of every 128-literal chunk (csr.cu)
    uint32_t csr_staged_words = 0;       // assignment words staged in shared memory by the CSR sweep (0: lookups through L2)
    uint32_t *d_bits = nullptr;
    unsigned long long *d_claim = nullptr;
    bool persistent_ok = false;          // the persistent solve kernel fits this instance and device
    uint32_t *d_urec = nullptr;          // records of the violated clauses of the current round, written by the sweep
    uint32_t urec_cap = 0;
    uint32_t *d_viol = nullptr, *d_s = nullptr, *d_ids_out = nullptr;
    uint8_t *d_state = nullptr, *d_bools = nullptr;
    uint8_t *h_bools = nullptr;          // pinned staging for the 1-byte-per-variable boundary (grow-only)
    size_t h_bools_cap = 0;
    uint8_t *h_stage = nullptr;          // pinned staging for rows the library builds itself (padded ragged input; grow-only)
    size_t h_stage_cap = 0;
    // packed H2D transport (PackPipe below): page-locked ring of PACK_SLOTS packed chunks, its device mirror, one event per
    // slot (the device slot has been unpacked), the mode (ALLL_H2D_PACK: -1 automatic, 0 off, 1 always), the last upload's record
    uint8_t *h_pack = nullptr, *d_pack = nullptr;
    size_t h_pack_cap = 0;
    cudaEvent_t ev_slot[PACK_SLOTS] = {};
    int h2d_pack = -1;
    uint64_t up_link_bytes = 0, up_packed_chunks = 0, up_raw_chunks = 0, up_pack_threads = 0;
    uint8_t *d_width = nullptr, *d_width_in = nullptr;   // padded planes: true clause widths by slot / by caller id
    bool use_width = false;
    Counters *d_ctr = nullptr;
    Counters *h_ctr = nullptr;           // pinned
    RoundNote *h_ring = nullptr;         // pinned [ROUNDS_IN_FLIGHT]: written by the MIS kernels, polled by the round loop
    unsigned long long seq = 0;          // last sequence number handed to a round
    cudaEvent_t ev_round[ROUNDS_IN_FLIGHT] = {};
    uint32_t sweep_grid = 1, mis_grid = 1, csr_grid = 1;
    std::vector<cudaEvent_t> ev;         // 2 * MAX_TIMED_ROUNDS + 2

    // Device buffers are kept across uploads and only grow: re-uploading an instance of the same shape (the
    // end-to-end path) then costs no cudaMalloc/cudaFree, which dominate a 1.3 GB upload otherwise.
    std::map<void **, size_t> caps;
    bool use_orig_id = false;
    uint32_t id_base = 0;                // first global clause id of this clause range (sharded mode)
    // sharded mode: dense copy of the gathered violated records
    uint32_t *d_sh_planes = nullptr, *d_sh_ids = nullptr, *d_sh_iota = nullptr, *d_sh_s = nullptr;
    uint8_t *d_sh_state = nullptr;
    // incremental re-evaluation (ALLL_FLAG_INCREMENTAL, fixed-width layout)
    bool incr_ready = false;
    uint32_t incr_stride = 0, incr_max_vars = 0;
    uint32_t *d_occ_off = nullptr, *d_occ = nullptr, *d_rows = nullptr, *d_visited = nullptr, *d_incr_tmp = nullptr;
    uint64_t visited_words = 0;
    // sharded P2P mode: our exchange region, the peers' mappings, the device-resident link table
    uint8_t *d_p2p_region = nullptr;
    size_t p2p_region_bytes = 0;
    void *p2p_peer[MAX_SHARDS] = {};
    uint8_t p2p_peer_ipc[MAX_SHARDS][64] = {};   // the IPC handle each open peer mapping came from
    uint8_t p2p_ipc[64] = {};                    // IPC handle of our own region (valid while the allocation lives)
    bool p2p_ipc_valid = false;
    P2PLink *d_p2p_link = nullptr;
    uint32_t p2p_world = 0, p2p_rank = 0;
    uint64_t p2p_cap = 0;
    bool p2p_ready = false;
    // batched small instances
    bool has_batch = false;
    uint32_t b_n_inst = 0, b_n_vars = 0, b_n_words = 0, b_k = 0, b_m_max = 0;
    uint64_t b_m_pad = 0;
    uint32_t *d_b_planes = nullptr, *d_b_off = nullptr, *d_b_m = nullptr, *d_b_bits = nullptr, *d_b_lit = nullptr;
    uint64_t *d_b_src_off = nullptr, *d_b_seeds = nullptr;
    BatchJobStats *d_b_stats = nullptr;
    uint8_t *d_b_bytes = nullptr;
    int *d_b_winner = nullptr;
    uint32_t *d_b_retry = nullptr;      // small batch kernel -> large batch kernel: count + job ids
    // multi-GPU portfolio: one winner word for all ranks, owned by one GPU and peer-mapped (CUDA IPC) by the others
    int *d_flag = nullptr;               // the word as this process addresses it
    bool flag_owner = false;
    bool flag_borrowed = false;          // the word belongs to another handle of this process (multi.cu): never freed / closed here
    uint32_t b_job_base = 0;
    uint8_t *d_tmp_bkt = nullptr;
    uint32_t *d_tmp_cnt = nullptr, *d_tmp_err = nullptr, *d_stage = nullptr;
    // enumerated clauses (alll_upload_generator): nothing stored but the violated records of the current round
    bool gen_mode = false;
    alll_gen_launch_fn gen_launch = nullptr;
    void *gen_user = nullptr;
    BuiltinGenerator *gen_builtin = nullptr;   // owned (alll_upload_builtin_generator)
    uint32_t *d_gen_rec = nullptr;
    uint64_t gen_cap = 0;
};

namespace {

int fail(alll_handle h, int status, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return status;
}

// No C++ exception may cross the C boundary (a caller in C, Go or Python cannot catch it, and unwinding through foreign
// frames is undefined): every extern "C" function that returns a status is a function-try-block that ends in ALLL_GUARD.
int guard_fail(alll_handle h, const char *what) noexcept
{
    try { return fail(h, ALLL_CUDA_ERROR, std::string("host-side failure: ") + what); } catch (...) { return ALLL_CUDA_ERROR; }
}
#define ALLL_GUARD(h)                                                                                   \
    catch (const std::exception &e__) { return guard_fail((h), e__.what()); }                           \
    catch (...) { return guard_fail((h), "unknown exception"); }

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(h, ALLL_CUDA_ERROR, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

template <typename T> void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

// Forgets the current instance; its buffers stay pooled for the next upload.
void free_instance(alll_handle h)
{
    h->has_instance = false;
    h->packed_on = false;
    h->rows8_on = false;
    h->use_orig_id = false;
    h->use_width = false;
    h->incr_ready = false;
    h->gen_mode = false;
    h->gen_launch = nullptr;
    h->gen_user = nullptr;
    if (h->gen_builtin) { builtin_generator_destroy(h->gen_builtin); h->gen_builtin = nullptr; }
}

void release_buffers(alll_handle h)
{
    dfree(h->d_planes); dfree(h->d_packed); dfree(h->d_orig_id); dfree(h->d_segs); dfree(h->d_sweep_segs); dfree(h->d_runs); dfree(h->d_run_begin); dfree(h->d_off); dfree(h->d_csr_lit); dfree(h->d_csr_start); dfree(h->d_csr_rank);
    dfree(h->d_bits); dfree(h->d_claim); dfree(h->d_urec); dfree(h->d_viol); dfree(h->d_s); dfree(h->d_ids_out);
    dfree(h->d_state); dfree(h->d_bools); dfree(h->d_width); dfree(h->d_width_in); dfree(h->d_tmp_bkt); dfree(h->d_tmp_cnt); dfree(h->d_tmp_err); dfree(h->d_stage); dfree(h->d_pack);
    dfree(h->d_sh_planes); dfree(h->d_sh_ids); dfree(h->d_sh_iota); dfree(h->d_sh_s); dfree(h->d_sh_state);
    dfree(h->d_b_planes); dfree(h->d_b_off); dfree(h->d_b_m); dfree(h->d_b_bits); dfree(h->d_b_lit); dfree(h->d_b_src_off);
    dfree(h->d_b_seeds); dfree(h->d_b_stats); dfree(h->d_b_bytes); dfree(h->d_b_winner); dfree(h->d_b_retry);
    h->has_batch = false;
    for (uint32_t q = 0; q < MAX_SHARDS; q++)
        if (h->p2p_peer[q]) { cudaIpcCloseMemHandle(h->p2p_peer[q]); h->p2p_peer[q] = nullptr; }
    dfree(h->d_p2p_region); dfree(h->d_p2p_link);
    h->p2p_region_bytes = 0; h->p2p_ipc_valid = false;
    dfree(h->d_occ_off); dfree(h->d_occ); dfree(h->d_rows); dfree(h->d_visited); dfree(h->d_incr_tmp);
    dfree(h->d_gen_rec);
    if (h->d_flag) { if (h->flag_owner) cudaFree(h->d_flag); else if (!h->flag_borrowed) cudaIpcCloseMemHandle(h->d_flag); h->d_flag = nullptr; h->flag_borrowed = false; }
    free_instance(h);
    h->incr_ready = false;
    h->p2p_ready = false;
    h->caps.clear();
    h->has_instance = false;
}

// Grow-only allocation of the buffer behind *slot (a field of the handle).
template <typename T> int pool_alloc(alll_handle h, T **slot, size_t bytes)
{
    void **key = reinterpret_cast<void **>(slot);
    bytes = std::max<size_t>(bytes, 16);
    auto it = h->caps.find(key);
    if (*slot && it != h->caps.end() && it->second >= bytes) return ALLL_OK;
    if (*slot) { cudaFree(*slot); *slot = nullptr; }
    CK(cudaMalloc(slot, bytes));
    h->caps[key] = bytes;
    return ALLL_OK;
}
#define POOL(slot, bytes) do { if (int rc__ = pool_alloc(h, &(slot), (bytes))) return rc__; } while (0)

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Host-side passes over the caller's arrays (width scan, ragged padding) are split over the host's threads.
inline uint32_t host_threads_for(uint64_t items, uint64_t min_per_thread)
{
    const uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint32_t>(hw, 32u), items / std::max<uint64_t>(min_per_thread, 1)));
}
template <class F> void parallel_ranges(uint64_t items, uint32_t nt, F &&fn)     // fn(thread, begin, end)
{
    if (nt <= 1) { fn(0u, (uint64_t)0, items); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (uint32_t t = 1; t < nt; t++) th.emplace_back([&, t] { fn(t, items * t / nt, items * (t + 1) / nt); });
    fn(0u, (uint64_t)0, items / nt);
    for (auto &x : th) x.join();
}

// flags bits 16..23: L2 prefetch distance in tiles; 0 = default (2, measured best on B200), 0xFF = off
inline uint32_t prefetch_distance(uint32_t flags)
{
    const uint32_t d = (flags >> 16) & 0xFFu;
    return d == 0 ? 2u : (d == 0xFFu ? 0u : d);
}

MisScratch mis_scratch(alll_handle h, bool with_records)
{
    MisScratch sc{};
    sc.claim = h->d_claim;
    // (incremental rounds produce the violated list without records)
    const bool rec = with_records && h->urec_cap != 0 && !h->incr_ready;
    sc.urec = rec ? h->d_urec : nullptr; sc.urec_cap = rec ? h->urec_cap : 0u;
    return sc;
}

ClauseView clause_view(alll_handle h)
{
    ClauseView cv;
    cv.planes = h->d_planes; cv.m_pad = h->m_pad; cv.k = h->k;
    cv.off = h->d_off; cv.csr_lit = h->d_csr_lit; cv.orig_id = h->use_orig_id ? h->d_orig_id : nullptr;
    cv.id_base = h->id_base;
    cv.width_arr = h->use_width ? h->d_width : nullptr;
    cv.rec = h->gen_mode ? h->d_gen_rec : nullptr;
    cv.rows = (h->incr_ready || h->rows8_on) ? h->d_rows : nullptr;          // (incremental mode builds the same rows for any k)
    cv.row_stride = h->incr_ready ? h->incr_stride : 8u;
    return cv;
}

// Buffers every instance needs regardless of the clause layout.
// list_cap: the largest violated set the per-round lists must hold (m for stored clauses).
int alloc_common(alll_handle h, uint64_t list_cap)
{
    const uint64_t m1 = std::max<uint64_t>(list_cap, 1), n1 = std::max<uint64_t>(h->n_vars, 1);
    h->urec_cap = 0;                                       // (set by the layouts whose sweep writes records)
    h->persistent_ok = false;
    POOL(h->d_bits, (size_t)std::max<uint32_t>(h->n_words_alloc, 4) * 4);
    CK(cudaMemsetAsync(h->d_bits, 0, (size_t)std::max<uint32_t>(h->n_words_alloc, 4) * 4, h->stream));
    POOL(h->d_claim, 2 * n1 * 8);                          // claim[v][2]: even / odd Luby steps of variable v side by side
    CK(launch_fill_u64(h->d_claim, 2 * n1, CLAIM_FREE, h->stream)); h->launches++;
    POOL(h->d_viol, m1 * 4);
    POOL(h->d_s, m1 * 4);
    POOL(h->d_ids_out, m1 * 4);
    POOL(h->d_state, m1);
    POOL(h->d_bools, n1);
    CK(launch_reset_counters(h->d_ctr, 1, h->stream)); h->launches++;
    CK(mis_configure(h->device, h->kmax, &h->mis_grid));
    return ALLL_OK;
}

int check_sizes(alll_handle h, uint64_t n_vars, uint64_t m)
{
    if (n_vars == 0 || n_vars > (1ull << 31)) return fail(h, ALLL_BAD_ARG, "n_vars must be in [1, 2^31] (literals are uint32 2*var+neg)");
    if (m > 0xFFFFFFFFull - 2 * TILE) return fail(h, ALLL_BAD_ARG, "m must be below 2^32 (clause ids are uint32)");
    return ALLL_OK;
}

// d_width_in (may be NULL): true width of every clause (by caller id) when the rows are padded to k literals with
// copies of their first literal (ragged input on the plane layout; a repeated literal never changes a clause's value).
// host_lit != NULL: d_lit is the staging buffer of a host-buffer upload that has NOT been filled yet -- the copy is issued
// here, in chunks on the copy stream, and the first layout pass (transpose, or bucket count) runs chunk by chunk behind it,
// so that pass hides under the PCIe transfer instead of following it.
// filled != NULL (host-buffer upload only): the producer is still filling host_lit; filled(user, filled_base + c) blocks
// until rows [0, c) of host_lit are there (non-zero return: the producer gave up -> ALLL_BAD_ARG).
int upload_fixedk_device_impl(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *d_lit,
                              const uint8_t *d_width_in = nullptr, const uint32_t *host_lit = nullptr,
                              alll_filled_fn filled = nullptr, void *filled_user = nullptr, uint64_t filled_base = 0)
{
    free_instance(h);
    if (int rc = check_sizes(h, n_vars, m)) return rc;
    if (k < 1 || k > MAX_K) return fail(h, ALLL_BAD_ARG, "k must be in [1, 32] for the fixed-width layout");
    h->n_vars = n_vars; h->m = m; h->k = k; h->kmax = k;

    const uint32_t n_words4 = (uint32_t)align_up((n_vars + 31) / 32, 4);
    const uint32_t budget_words = h->smem_budget / 16 * 4;
    if (n_words4 <= budget_words || (h->flags & ALLL_FLAG_NO_BUCKETING)) {
        h->n_buckets = 1;
        h->bucket_words = std::min(n_words4, budget_words);
        h->resident_all = h->bucket_words >= n_words4;
    } else {
        uint32_t nb = (n_words4 + budget_words - 1) / budget_words;
        if (nb > MAX_BUCKETS) { nb = MAX_BUCKETS; h->bucket_words = budget_words; }
        else h->bucket_words = (uint32_t)align_up((n_words4 + nb - 1) / nb, 4);
        h->n_buckets = nb;
        h->resident_all = false;
    }
    h->n_words_alloc = std::max<uint32_t>(n_words4, h->n_buckets * h->bucket_words);

    POOL(h->d_tmp_err, 16);
    uint32_t *d_err = h->d_tmp_err;                  // [0] error flags, [1] min resident-placed literals per clause, [2] tile cursor of the bucketing pass
    const uint32_t err_init[4] = {0u, 0xFFFFFFFFu, 0u, 0u};
    CK(cudaMemcpyAsync(d_err, err_init, 16, cudaMemcpyHostToDevice, h->stream));
    std::vector<BucketSeg> segs(h->n_buckets);
    h->n_segs = h->n_buckets;
    uint32_t tiles_used = 0;
    bool fused_pack = false, fused_rows = false;     // written by the bucket scatter already
    h->min_resident = 0;
    h->resident_cap = RESIDENT_CAP;

    // clause ranges of the first pass: one, or (host-buffer upload) one per H2D chunk of about 64 MB
    std::vector<uint64_t> cut{0, m};
    if (host_lit && m) {
        const uint64_t per = bucket_pass_clauses_per_cta();
        uint64_t chunk = std::max<uint64_t>(per, ((64ull << 20) / (4ull * k)) / per * per);
        if (const char *e = getenv("ALLL_H2D_CHUNK_ROWS"))               // test knob: small chunks, so that small instances exercise the chunked paths
            if (const uint64_t v = strtoull(e, nullptr, 0)) chunk = align_up(v, per);
        if ((m + chunk - 1) / chunk > 32) chunk = align_up((m + 31) / 32, per);
        cut.clear();
        for (uint64_t c0 = 0; c0 < m; c0 += chunk) cut.push_back(c0);
        cut.push_back(m);
        CK(cudaEventRecord(h->ev_chunk[32], h->stream));                // the staging buffer may still be read by earlier work
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_chunk[32], 0));
    }
    // Packed transport (PackPipe above): complete host buffers of literals that fit 25 bits; automatic for large uploads.
    std::unique_ptr<PackPipe> pipe;
    bool src_pinned = false;
    size_t pack_slot_bytes = 0, pack_hi_off = 0;
    uint32_t retired = 0;                            // chunks whose copy is known to have completed
    h->up_link_bytes = h->up_packed_chunks = h->up_raw_chunks = h->up_pack_threads = 0;
    if (host_lit && m && !filled && 2 * n_vars <= (1ull << 25) && h->h2d_pack != 0 && (h->h2d_pack == 1 || m * k >= PACK_MIN_LITERALS)) {
        cudaPointerAttributes attr{};
        src_pinned = cudaPointerGetAttributes(&attr, host_lit) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        uint64_t widest = 0;
        for (size_t i = 0; i + 1 < cut.size(); i++) widest = std::max(widest, cut[i + 1] - cut[i]);
        pack_hi_off = (size_t)align_up(3 * widest * k, 64);
        pack_slot_bytes = pack_hi_off + (size_t)align_up((widest * k + 7) / 8, 64);
        const size_t need = pack_slot_bytes * PACK_SLOTS;
        bool ring_ok = true;
        if (h->h_pack_cap < need) {
            if (h->h_pack) { cudaFreeHost(h->h_pack); h->h_pack = nullptr; h->h_pack_cap = 0; }
            // (no page-locked memory to spare: the upload goes as it is -- the transport is an optimisation, not a requirement)
            if (cudaMallocHost(&h->h_pack, need) == cudaSuccess) h->h_pack_cap = need;
            else { h->h_pack = nullptr; cudaGetLastError(); ring_ok = false; }
        }
        if (ring_ok) {
            POOL(h->d_pack, need);
            // (the cores this process may run on: a rank bound to its GPU's NUMA node packs with that node's cores)
            uint32_t hw = std::max(1u, std::thread::hardware_concurrency());
            cpu_set_t cs;
            if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) hw = (uint32_t)CPU_COUNT(&cs);
            const uint32_t nt = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint32_t>(hw, 24u), (m + PACK_UNIT_ROWS - 1) / PACK_UNIT_ROWS));
            h->up_pack_threads = nt;
            pipe.reset(new PackPipe(host_lit, k, cut, h->h_pack, pack_slot_bytes, pack_hi_off, nt));
        }
    }
    bool producer_gave_up = false;
    // chunk i: (streamed upload: wait until the producer has filled it,) enqueue its copy, make the layout stream wait for it
    auto chunk_ready = [&](size_t i) -> cudaError_t {
        if (!(host_lit && m)) return cudaSuccess;
        if (filled && filled(filled_user, filled_base + cut[i + 1]) != 0) { producer_gave_up = true; return cudaErrorUnknown; }
        if (pipe) {
            const uint32_t ci = (uint32_t)i;
            // (from page-locked memory the chunk goes as it is when the link has run dry before it is packed)
            const bool packed = pipe->wait_chunk(ci, src_pinned, [&] {
                while (retired < ci && cudaEventQuery(h->ev_chunk[retired]) == cudaSuccess) ++retired;
                return retired;
            });
            cudaGetLastError();                      // (cudaErrorNotReady of the polls)
            if (packed) {
                const uint32_t slot = ci % PACK_SLOTS;
                const uint64_t n_l = (cut[i + 1] - cut[i]) * k;
                uint8_t *hs = h->h_pack + (size_t)slot * pack_slot_bytes, *ds = h->d_pack + (size_t)slot * pack_slot_bytes;
                cudaError_t e = cudaSuccess;
                // (the device slot's previous content has been expanded: ev_slot is recorded behind its unpack kernel)
                if (ci >= (uint32_t)PACK_SLOTS) e = cudaStreamWaitEvent(h->copy_stream, h->ev_slot[slot], 0);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ds, hs, 3 * n_l, cudaMemcpyHostToDevice, h->copy_stream);
                if (e == cudaSuccess) e = cudaMemcpyAsync(ds + pack_hi_off, hs + pack_hi_off, (n_l + 7) / 8, cudaMemcpyHostToDevice, h->copy_stream);
                if (e == cudaSuccess) e = cudaEventRecord(h->ev_chunk[i], h->copy_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(h->stream, h->ev_chunk[i], 0);
                if (e == cudaSuccess) e = launch_unpack25(ds, ds + pack_hi_off, const_cast<uint32_t *>(d_lit) + cut[i] * k, n_l, h->stream);
                if (e == cudaSuccess) e = cudaEventRecord(h->ev_slot[slot], h->stream);
                h->launches++;
                h->up_packed_chunks++;
                h->up_link_bytes += 3 * n_l + (n_l + 7) / 8;
                return e;
            }
            h->up_raw_chunks++;
        }
        h->up_link_bytes += (cut[i + 1] - cut[i]) * k * 4;
        cudaError_t e = cudaMemcpyAsync(const_cast<uint32_t *>(d_lit) + cut[i] * k, host_lit + cut[i] * k, (cut[i + 1] - cut[i]) * k * 4,
                                        cudaMemcpyHostToDevice, h->copy_stream);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(h->ev_chunk[i], h->copy_stream);
        if (e != cudaSuccess) return e;
        return cudaStreamWaitEvent(h->stream, h->ev_chunk[i], 0);
    };
#define CK_CHUNK(i)                                                                                     \
    do {                                                                                                \
        const cudaError_t e_ = chunk_ready(i);                                                          \
        if (producer_gave_up) { cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream); free_instance(h);  \
                                return fail(h, ALLL_BAD_ARG, "streamed upload: the producer of the host buffer gave up"); } \
        CK(e_);                                                                                         \
    } while (0)

    if (h->n_buckets == 1) {
        h->m_pad = align_up(m, TILE);
        // padding slots are never evaluated (masked by slot_end), so the planes need no clearing
        if (h->m_pad) POOL(h->d_planes, h->m_pad * k * 4);
        for (size_t i = 0; i + 1 < cut.size(); i++) {
            CK_CHUNK(i);
            CK(launch_transpose(d_lit, cut[i], cut[i + 1], k, n_vars, h->d_planes, h->m_pad, d_err, h->stream)); h->launches++;
        }
        segs[0] = BucketSeg{0u, (uint32_t)m, 0u, 0u};
        if (d_width_in && m) {
            POOL(h->d_width, h->m_pad);
            CK(cudaMemcpyAsync(h->d_width, d_width_in, m, cudaMemcpyDeviceToDevice, h->stream));
        }
    } else {
        // Bucketing, chunk by chunk behind the H2D copy: count -> scan (device) -> scatter per chunk, no host round trip
        // in between.  Slot order: chunk-major, bucket by bucket inside a chunk, every (chunk, bucket) segment starting on
        // a sweep-tile boundary -- so the planes are allocated for the worst-case padding up front and m_pad is their stride.
        const uint32_t bucket_vars = h->bucket_words * 32u, nb = h->n_buckets;
        const uint32_t n_cta = bucket_pass_ctas(m);
        const uint32_t n_chunks = (uint32_t)cut.size() - 1;
        h->n_segs = nb * n_chunks;
        const uint64_t m_pad_max = align_up(m, TILE) + (uint64_t)h->n_segs * TILE;
        if (m_pad_max >= 0xFFFFFFFFull) return fail(h, ALLL_BAD_ARG, "m too large for the bucketed layout (slot numbers are uint32)");
        h->m_pad = m_pad_max;
        POOL(h->d_tmp_bkt, std::max<uint64_t>(m, 1));
        POOL(h->d_tmp_cnt, (size_t)nb * std::max<uint32_t>(n_cta, 1) * 4);
        POOL(h->d_segs, sizeof(BucketSeg) * h->n_segs);
        POOL(h->d_planes, h->m_pad * k * 4);
        POOL(h->d_orig_id, h->m_pad * 4);
        if (d_width_in) POOL(h->d_width, h->m_pad);
        h->use_orig_id = true;
        uint8_t *d_bkt = h->d_tmp_bkt;
        uint32_t *d_cnt = h->d_tmp_cnt;
        // The sweep's packed eager planes and tail rows are written by the scatter itself where it can (k > buckets: every
        // clause has two literals in the bucket that holds most of its variables -- pigeonhole); otherwise by passes below.
        const bool want_pack = !(h->flags & ALLL_FLAG_NO_PACKING) && !(h->tune & TUNE_NO_PACKED_PLANES) && h->resident_cap == 3 &&
                               2ull * bucket_vars <= (1ull << 22) && m > 0 && k >= EAGER_PLANES && k <= 8;
        const bool want_rows = !(h->flags & ALLL_FLAG_NO_PACKING) && !(h->tune & TUNE_NO_TAIL_ROWS) && m > 0 && k >= EAGER_PLANES && k <= 8;
        const bool fuses = bucket_scatter_fuses(k, d_width_in != nullptr);
        fused_pack = fuses && want_pack && k > nb && n_vars <= (1ull << 27);
        fused_rows = fuses && want_rows;
        if (fused_pack) POOL(h->d_packed, h->m_pad * 4 * 4);
        if (fused_rows) POOL(h->d_rows, h->m_pad * 8 * 4);
        for (size_t i = 0; i + 1 < cut.size() && m; i++) {
            CK_CHUNK(i);
            CK(launch_bucket_count(d_lit, m, cut[i], cut[i + 1], k, n_vars, bucket_vars, nb, d_bkt, d_cnt, d_err, h->stream));
            CK(launch_bucket_scan(d_cnt, m, cut[i], cut[i + 1], nb, h->d_segs + i * nb, d_err + 2, h->stream));
            CK(launch_bucket_scatter(d_lit, m, cut[i], cut[i + 1], k, bucket_vars, nb, d_bkt, d_cnt, h->d_planes, h->m_pad, h->d_orig_id,
                                     d_err + 1, h->resident_cap, d_width_in, d_width_in ? h->d_width : nullptr,
                                     fused_pack ? h->d_packed : nullptr, fused_rows ? reinterpret_cast<uint4 *>(h->d_rows) : nullptr, h->stream));
            h->launches += 3;
        }
        if (m == 0) {                                    // no chunk ran: empty segments
            segs.assign(h->n_segs, BucketSeg{0u, 0u, 0u, 0u});
            for (uint32_t i = 0; i < h->n_segs; i++) segs[i].bucket = i % nb;
            CK(cudaMemcpyAsync(h->d_segs, segs.data(), sizeof(BucketSeg) * h->n_segs, cudaMemcpyHostToDevice, h->stream));
        }
    }
    if (pipe) {
        // every chunk has been issued, so every unit is packed or skipped; a literal above 25 bits lost its top bits on the
        // way (the device-side range check saw the truncated value): same answer as the plain path gives
        const uint32_t ora = pipe->or_acc.load();
        pipe.reset();
        if (ora >> 25) {
            cudaStreamSynchronize(h->copy_stream); cudaStreamSynchronize(h->stream);
            free_instance(h);
            return fail(h, ALLL_BAD_ARG, "a literal references a variable >= n_vars");
        }
    }
    if (h->n_buckets == 1) {
        tiles_used = (uint32_t)(h->m_pad / TILE);
        POOL(h->d_segs, sizeof(BucketSeg));
        CK(cudaMemcpyAsync(h->d_segs, segs.data(), sizeof(BucketSeg), cudaMemcpyHostToDevice, h->stream));
    }

    uint32_t err_out[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(err_out, d_err, 16, cudaMemcpyDeviceToHost, h->stream));
    if (h->n_buckets > 1 && m) {
        segs.resize(h->n_segs);
        CK(cudaMemcpyAsync(segs.data(), h->d_segs, sizeof(BucketSeg) * h->n_segs, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (h->n_buckets > 1) tiles_used = err_out[2];
    h->n_tiles = tiles_used;
    std::vector<BucketSeg> sw;                            // sweep order: bucket by bucket over the chunk-major slot order, empty segments dropped
    {
        if (h->n_buckets > 1 && m) {
            const uint32_t nb = h->n_buckets, n_chunks = h->n_segs / nb;
            uint32_t at = 0;
            for (uint32_t bkt = 0; bkt < nb; bkt++)
                for (uint32_t c = 0; c < n_chunks; c++) {
                    const BucketSeg &g = segs[c * nb + bkt];
                    const uint32_t t_end = (uint32_t)(((uint64_t)g.slot_end + TILE - 1) / TILE);
                    if (t_end <= g.tile_begin) continue;
                    sw.push_back(BucketSeg{at, g.slot_end, bkt, g.tile_begin});
                    at += t_end - g.tile_begin;
                }
            if (at != tiles_used) return fail(h, ALLL_CUDA_ERROR, "internal: bucket segments do not cover the tiles the scatter used");
        } else if (h->n_buckets == 1) {
            sw.push_back(segs[0]);
        }
        if (sw.empty()) sw.push_back(BucketSeg{0u, 0u, 0u, 0u});
        h->n_sweep_segs = (uint32_t)sw.size();
        POOL(h->d_sweep_segs, sizeof(BucketSeg) * sw.size());
        CK(cudaMemcpyAsync(h->d_sweep_segs, sw.data(), sizeof(BucketSeg) * sw.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));             // (`sw` is pageable and goes out of scope)
    }
    if (err_out[0]) { free_instance(h); return fail(h, ALLL_BAD_ARG, "a literal references a variable >= n_vars"); }
    if (h->n_buckets > 1 && m > 0 && err_out[1] != 0xFFFFFFFFu) h->min_resident = err_out[1];

    h->use_width = d_width_in != nullptr && m > 0;
    {
        // Packed eager planes: the five literals the sweep streams per clause in one 128-bit word (EagerPack).  Needs a
        // bucketed layout whose clauses all have a bucket-resident leading literal (two for the wider global fields).
        const uint32_t rb = std::min<uint32_t>(h->min_resident, 2u);
        const uint64_t glob_limit = rb >= 2 ? (1ull << 27) : (1ull << 25);
        h->packed_on = h->n_buckets > 1 && m > 0 && k >= EAGER_PLANES && k <= 8 && rb >= 1 && h->resident_cap == 3 &&
                       2ull * h->bucket_words * 32u <= (1ull << 22) && n_vars <= glob_limit && !(h->flags & ALLL_FLAG_NO_PACKING) && !(h->tune & TUNE_NO_PACKED_PLANES);
        if (fused_pack && rb < 2) return fail(h, ALLL_CUDA_ERROR, "internal: packed planes written for two resident literals, but a clause has fewer");
        if (h->packed_on && !fused_pack) {
            POOL(h->d_packed, h->m_pad * 4 * 4);
            CK(launch_pack_eager(h->d_planes, h->m_pad, h->d_segs, h->n_segs, h->bucket_words * 32u, rb, h->d_packed, h->stream));
            h->launches++;
        }
    }
    h->rows8_on = m > 0 && k >= EAGER_PLANES && k <= 8 && !(h->flags & ALLL_FLAG_NO_PACKING) && !(h->tune & TUNE_NO_TAIL_ROWS);
    if (h->rows8_on && !fused_rows) {
        POOL(h->d_rows, h->m_pad * 8 * 4);
        CK(launch_rows8(h->d_planes, h->m_pad, k, reinterpret_cast<uint4 *>(h->d_rows), h->stream)); h->launches++;
    }
    if (int rc = alloc_common(h, h->m)) return rc;
    if (k >= 1 && k <= 8 && !h->use_width && m > 0) {
        // violated-clause records for the independent-set kernels (sweep_body.cuh:write_records), for violated sets of up to
        // URECORD_CAP clauses.  Larger sets read the literal planes instead: writing their records costs the sweep more
        // (scattered reads at the tail of the kernel: +54 us at |U| = 156 k) than it saves the gather (31 us there).
        const uint64_t cap = std::min<uint64_t>(m, URECORD_CAP);
        POOL(h->d_urec, cap * (k + 1) * 4);
        h->urec_cap = (uint32_t)cap;
    }
    SweepParams sp{};
    sp.bucket_words = h->bucket_words; sp.k = k; sp.min_resident = h->min_resident; sp.resident_cap = h->resident_cap; sp.eager = (h->flags >> 8) & 0xFFu; sp.prefetch_tiles = prefetch_distance(h->flags);
    sp.packed = h->packed_on ? h->d_packed : nullptr;
    sp.rows8 = (h->rows8_on && h->k > EAGER_PLANES) ? reinterpret_cast<const uint4 *>(h->d_rows) : nullptr;
    CK(configure_sweep_planes(sp, h->resident_all));
    h->sweep_grid = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)h->sm_count, h->n_tiles));
    {
        // per-CTA run lists: CTA c sweeps tiles [c * n_tiles / grid, (c + 1) * n_tiles / grid) of the sweep order, cut at segment ends
        std::vector<SweepRun> runs;
        std::vector<uint32_t> run_begin(h->sweep_grid + 1, 0u);
        size_t sg = 0;
        for (uint32_t c = 0; c < h->sweep_grid; c++) {
            run_begin[c] = (uint32_t)runs.size();
            uint32_t t = (uint32_t)(((uint64_t)c * h->n_tiles) / h->sweep_grid);
            const uint32_t t1 = (uint32_t)(((uint64_t)(c + 1) * h->n_tiles) / h->sweep_grid);
            while (t < t1) {
                while (sg + 1 < sw.size() && sw[sg + 1].tile_begin <= t) ++sg;
                const uint32_t seg_end = sg + 1 < sw.size() ? sw[sg + 1].tile_begin : h->n_tiles;
                const uint32_t e = std::min(t1, seg_end), delta = sw[sg].phys_tile - sw[sg].tile_begin;
                runs.push_back(SweepRun{t + delta, e + delta, sw[sg].slot_end, sw[sg].bucket});
                t = e;
            }
        }
        run_begin[h->sweep_grid] = (uint32_t)runs.size();
        POOL(h->d_runs, sizeof(SweepRun) * std::max<size_t>(runs.size(), 1));
        POOL(h->d_run_begin, sizeof(uint32_t) * run_begin.size());
        if (!runs.empty()) CK(cudaMemcpyAsync(h->d_runs, runs.data(), sizeof(SweepRun) * runs.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->d_run_begin, run_begin.data(), sizeof(uint32_t) * run_begin.size(), cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));             // (pageable sources going out of scope)
    }
    {
        int ok = 0, coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
        if (coop && m > 0 && !h->use_width) CK(configure_solve_persistent(sp, h->resident_all, h->kmax, &ok));
        h->persistent_ok = ok != 0;
    }
    if ((h->flags & ALLL_FLAG_INCREMENTAL) && m > 0) {
        // occurrence lists + row-major copy for incremental re-evaluation (see incremental.cu)
        const uint64_t n_lit = m * k;
        if (n_lit >= 0xFFFFFFF0ull) return fail(h, ALLL_BAD_ARG, "incremental mode needs fewer than 2^32 literals");
        h->incr_stride = (k + 3u) & ~3u;
        const uint32_t nb = (uint32_t)((n_vars + 1023) / 1024);
        POOL(h->d_occ_off, (n_vars + 1) * 4);
        POOL(h->d_occ, n_lit * 4);
        POOL(h->d_rows, h->m_pad * h->incr_stride * 4);
        h->visited_words = (h->m_pad + 31) / 32;
        POOL(h->d_visited, h->visited_words * 4);
        POOL(h->d_incr_tmp, (n_vars + nb + 4) * 4);                 // cursors | block sums | total
        CK(cudaMemsetAsync(h->d_visited, 0, h->visited_words * 4, h->stream));
        CK(launch_incr_build(h->d_planes, h->m_pad, k, h->incr_stride, h->d_segs, h->n_segs, n_vars, h->d_occ_off,
                             h->d_incr_tmp, h->d_incr_tmp + n_vars, h->d_rows, h->d_occ, h->d_incr_tmp + n_vars + nb,
                             h->use_width ? h->d_width : nullptr, h->stream));
        h->launches += 5;
        // next round is incremental when the resampled variables' occurrence lists cover <= m / divisor clauses
        const uint32_t div_log2 = (h->flags >> 24) & 0xFu;
        const uint64_t divisor = div_log2 ? (1ull << div_log2) : 8ull;
        const double avg_occ = (double)n_lit / (double)n_vars;
        h->incr_max_vars = (uint32_t)std::max<double>(1.0, (double)m / ((double)divisor * avg_occ));
        h->incr_ready = true;
    }
    CK(cudaStreamSynchronize(h->stream));
    h->has_instance = true;
    return ALLL_OK;
}

CsrSweepParams csr_params(alll_handle h)
{
    CsrSweepParams cp{};
    cp.lit = h->d_csr_lit; cp.start = h->d_csr_start; cp.chunk_rank = h->d_csr_rank;
    cp.n_chunks = (uint32_t)(h->csr_l_pad / CSR_CHUNK); cp.m = (uint32_t)h->m;
    cp.bits = h->d_bits; cp.n_words = h->n_words_alloc; cp.staged_words = h->csr_staged_words;
    cp.viol = h->d_viol; cp.ctr = h->d_ctr;
    return cp;
}

// records: leave {id, literals} records next to the violated list (an independent-set phase follows and will read them)
SweepParams sweep_params(alll_handle h, uint32_t p2p_parity, uint32_t p2p_tag, uint32_t round, bool records)
{
    SweepParams sp{};
    sp.planes = h->d_planes; sp.m_pad = h->m_pad; sp.bits = h->d_bits; sp.n_words = h->n_words_alloc;
    sp.bucket_words = h->bucket_words; sp.n_segs = h->n_sweep_segs; sp.n_tiles = h->n_tiles;
    sp.runs = h->d_runs; sp.run_begin = h->d_run_begin; sp.run_grid = h->sweep_grid;
    sp.segs = h->d_sweep_segs; sp.viol = h->d_viol; sp.ctr = h->d_ctr; sp.k = h->k; sp.min_resident = h->min_resident;
    sp.resident_cap = h->resident_cap; sp.eager = (h->flags >> 8) & 0xFFu; sp.prefetch_tiles = prefetch_distance(h->flags);
    sp.round = round;
    sp.tune = h->tune;
    sp.packed = h->packed_on ? h->d_packed : nullptr;
    sp.rows8 = (h->rows8_on && h->k > EAGER_PLANES) ? reinterpret_cast<const uint4 *>(h->d_rows) : nullptr;
    sp.orig_id = h->use_orig_id ? h->d_orig_id : nullptr; sp.id_base = h->id_base;
    if (p2p_tag) {
        sp.p2p = h->d_p2p_link; sp.p2p_parity = p2p_parity; sp.p2p_tag = p2p_tag; sp.p2p_epoch = p2p_tag >> 20;
    } else if (records && h->urec_cap && !h->incr_ready) {      // (incremental rounds produce no records, so the independent set reads rows[] instead)
        sp.urec = h->d_urec; sp.urec_cap = h->urec_cap;
    }
    return sp;
}

// Enqueues one sweep.  Invariant: ctr->n_viol == 0 on entry (kept by the MIS kernel / reset kernel).
// p2p_tag != 0: sharded P2P mode -- violated records are stored into every GPU's exchange region.
int enqueue_sweep(alll_handle h, uint32_t p2p_parity = 0, uint32_t p2p_tag = 0, uint32_t round = 0xFFFFFFFFu, bool records = false)
{
    if (h->gen_mode) {
        if (h->m == 0) return ALLL_OK;
        alll_gen_sweep_args a{};
        a.bits = h->d_bits; a.m = h->m; a.k = h->k; a.grid_hint = (uint32_t)h->sm_count * 8u;
        a.records = h->d_gen_rec; a.cap = h->gen_cap;
        a.n_violated = &h->d_ctr->n_viol; a.skip = &h->d_ctr->done;
        const int e = h->gen_launch(h->gen_user, &a, (void *)h->stream);
        if (e != 0) return fail(h, ALLL_CUDA_ERROR, std::string("generator sweep launch: ") + cudaGetErrorString((cudaError_t)e));
        h->launches++;
        return ALLL_OK;
    }
    if (h->k) {
        if (h->n_tiles == 0 && !p2p_tag) return ALLL_OK;      // (a P2P rank without clauses still has to publish its round)
        const SweepParams sp = sweep_params(h, p2p_parity, p2p_tag, round, records);
        CK(launch_sweep_planes(sp, h->resident_all, h->sweep_grid, h->stream));
    } else {
        if (h->m == 0) return ALLL_OK;
        CK(launch_sweep_csr(csr_params(h), h->csr_grid, h->stream));
    }
    h->launches++;
    return ALLL_OK;
}

int enqueue_mis_resample(alll_handle h, uint64_t seed, uint32_t round, bool with_grid = true, RoundNote *note = nullptr,
                         unsigned long long seq = 0, bool allow_incremental = false, bool records = true)
{
    CK(launch_mis_resample_args(clause_view(h), h->kmax, h->gen_mode ? nullptr : h->d_viol, h->d_state, h->d_s, mis_scratch(h, records),
                                h->n_vars, h->d_bits, h->d_ctr, seed, round, h->mis_grid, with_grid, note, seq, nullptr, 0u, 0u,
                                (allow_incremental && h->incr_ready) ? h->incr_max_vars : 0u,
                                h->gen_mode ? (uint32_t)h->gen_cap : 0u, h->stream));
    h->launches += with_grid ? 2 : 1;    // cluster kernel (+ cooperative grid kernel)
    return ALLL_OK;
}

// ALLL_TRACE: %globaltimer stamps written by the kernels (us relative to the sweep entry of each round)
void print_phases(const Counters &c, uint64_t rounds)
{
    if (c.dbg_step[0][0]) {
        for (int s = 0; s < 16 && c.dbg_step[s][0]; s++) {
            const unsigned long long *d = c.dbg_step[s];
            fprintf(stderr, "[alll steps] round 0 step %d: decide %.2f | reduce+publish %.2f | barrier %.2f us%s\n", s,
                    (double)(d[1] - d[0]) * 1e-3, (double)(d[2] - d[1]) * 1e-3, (double)(d[3] - d[2]) * 1e-3,
                    s + 1 < 16 && c.dbg_step[s + 1][0] ? "" : "  (last)");
        }
    }
    const uint64_t nr = std::min<uint64_t>(rounds, DBG_ROUNDS);
    if (nr > 1 && c.dbg_cta[0] >= c.dbg[1][0]) {        // when every CTA left the sweep body of round 1 (us after block 0 entered it)
        std::vector<std::pair<double, int>> t;
        for (int b = 0; b < 256 && c.dbg_cta[b] >= c.dbg[1][0]; b++) t.emplace_back((double)(c.dbg_cta[b] - c.dbg[1][0]) * 1e-3, b);
        std::sort(t.begin(), t.end());
        fprintf(stderr, "[alll tail] round 1, %zu CTAs: sweep body done min %.1f | median %.1f | p90 %.1f | max %.1f us; slowest:", t.size(), t.front().first,
                t[t.size() / 2].first, t[t.size() * 9 / 10].first, t.back().first);
        for (size_t i = t.size() > 8 ? t.size() - 8 : 0; i < t.size(); i++) fprintf(stderr, " cta %d %.1f", t[i].second, t[i].first);
        fprintf(stderr, "; fastest:");
        for (size_t i = 0; i < 4 && i < t.size(); i++) fprintf(stderr, " cta %d %.1f", t[i].second, t[i].first);
        fprintf(stderr, "\n");
    }
    for (uint64_t r = 0; r < nr; r++) {
        const unsigned long long *d = c.dbg[r];
        auto us = [&](int i) { return d[i] >= d[0] ? (double)(d[i] - d[0]) * 1e-3 : -1.0; };
        const double gap = r ? (double)(d[0] - c.dbg[r - 1][6]) * 1e-3 : 0.0;
        fprintf(stderr, "[alll phases] round %llu: path=%llu steps=%llu | prev round end -> sweep entry %.1f | mis entry %.1f "
                        "|U| known %.1f gather %.1f steps %.1f resample %.1f finished %.1f (us after sweep entry)\n",
                (unsigned long long)r, d[7] & 0xFF, d[7] >> 8, gap, us(1), us(2), us(3), us(4), us(5), us(6));
        const unsigned long long *x = c.dbg_x[r];
        if (x[0] >= d[0] && x[3] >= x[0])        // sharded persistent solve: the exchange between sweep and independent set
            fprintf(stderr, "[alll exchange] round %llu: own sweep done %.1f | fenced + ticket %.1f | last CTA stored the flags %.1f | all ranks seen %.1f "
                            "(us after sweep entry; epoch ns of sweep entry %llu)\n",
                    (unsigned long long)r, (double)(x[0] - d[0]) * 1e-3, (double)(x[1] - d[0]) * 1e-3, (double)(x[2] - d[0]) * 1e-3,
                    (double)(x[3] - d[0]) * 1e-3, d[0]);
    }
}

int fetch_counters(alll_handle h)
{
    CK(cudaMemcpyAsync(h->h_ctr, h->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}

// device slots -> caller ids -> host buffer
int copy_ids_out(alll_handle h, const uint32_t *d_slots, uint64_t n, uint32_t *out, uint64_t cap)
{
    const uint64_t n_copy = std::min(n, cap);
    if (!out || n_copy == 0) return ALLL_OK;
    CK(launch_map_ids(clause_view(h), d_slots, (uint32_t)n_copy, h->d_ids_out, h->stream)); h->launches++;
    CK(cudaMemcpyAsync(out, h->d_ids_out, n_copy * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}

// Exchange region of this rank (sharded P2P mode): header + records[2 parities][world][cap][k+1].  The allocation is
// kept across uploads of the same shape and only grows (a 1.28 GB / 8-GPU exchange region is 0.7 GB: cudaMalloc and a
// full clear per upload would cost more than the upload); it is its own allocation because it may be exported over CUDA
// IPC.  A retained region keeps its header: round flags are tagged with the solve's epoch, which the caller never
// reuses on a link, so stale flags cannot match -- and peers may already be storing into it while we are still here.
int p2p_region_setup(alll_handle h, uint32_t world, uint32_t rank, uint64_t cap_records, bool *fresh)
{
    if (!h) return ALLL_BAD_ARG;
    if (!h->has_instance) return fail(h, ALLL_NO_INSTANCE, "no instance uploaded");
    CK(cudaSetDevice(h->device));
    if (!h->k || h->k > 8 || h->use_width || h->gen_mode) return fail(h, ALLL_BAD_ARG, "P2P sharding needs stored clauses of uniform width k <= 8");
    if (world < 1 || world > MAX_SHARDS || rank >= world) return fail(h, ALLL_BAD_ARG, "bad world / rank");
    if (cap_records == 0) return fail(h, ALLL_BAD_ARG, "cap_records == 0");
    h->p2p_ready = false;
    const size_t bytes = P2P_HEADER_BYTES + (size_t)2 * world * cap_records * (h->k + 1) * 4;
    *fresh = false;
    if (h->d_p2p_region && h->p2p_region_bytes < bytes) {
        // (peers' mappings of the old allocation die with it; they re-open on their next alll_p2p_connect)
        cudaFree(h->d_p2p_region);
        h->d_p2p_region = nullptr; h->p2p_region_bytes = 0; h->p2p_ipc_valid = false;
    }
    if (!h->d_p2p_region) {
        CK(cudaMalloc(&h->d_p2p_region, bytes));
        CK(cudaMemset(h->d_p2p_region, 0, P2P_HEADER_BYTES));
        CK(cudaDeviceSynchronize());                       // (the clear is ordered on the default stream, our kernels are not)
        h->p2p_region_bytes = bytes;
        *fresh = true;
    }
    h->p2p_world = world; h->p2p_rank = rank; h->p2p_cap = cap_records;
    return ALLL_OK;
}

// Builds the device-resident link table of the sharded P2P mode from every rank's exchange-region base address (as THIS
// process addresses it: CUDA IPC mappings for peers in other processes, plain device pointers for peers in this one).
int p2p_link_up(alll_handle h, void *const *bases)
{
    P2PLink link{};
    link.world = h->p2p_world; link.rank = h->p2p_rank; link.k = h->k; link.cap = h->p2p_cap;
    {   // bounded wait for a peer's round: long enough for a peer that starts late (module load, a busy host), short
        // enough to turn a dead peer into an error instead of a hang
        double ms = 20000.0;
        if (const char *e = getenv("ALLL_P2P_TIMEOUT_MS")) { const double v = atof(e); if (v > 0.0) ms = v; }
        const int khz = h->clock_khz > 0 ? h->clock_khz : 2000000;     // (cached at alll_create: the attribute query costs milliseconds)
        link.timeout_cycles = (long long)(ms * (double)khz);
    }
    for (uint32_t q = 0; q < h->p2p_world; q++) {
        uint8_t *base = static_cast<uint8_t *>(bases[q]);
        if (!base) return fail(h, ALLL_BAD_ARG, "missing exchange region of a peer");
        link.hdr[q] = reinterpret_cast<P2PHeader *>(base);
        link.rec[q] = reinterpret_cast<uint32_t *>(base + P2P_HEADER_BYTES);
    }
    POOL(h->d_p2p_link, sizeof(P2PLink));
    CK(cudaMemcpy(h->d_p2p_link, &link, sizeof(P2PLink), cudaMemcpyHostToDevice));
    const uint64_t total_cap = (uint64_t)h->p2p_world * h->p2p_cap;
    POOL(h->d_sh_s, total_cap * 4);
    POOL(h->d_sh_state, total_cap);
    h->p2p_ready = true;
    return ALLL_OK;
}

// ---- sharded P2P solve as one persistent kernel per rank, split into enqueue / collect so that ONE host thread can start
// the kernels of all ranks of a single-process multi-GPU solve (multi.cu) before it waits for any of them ----
bool p2p_persistent_possible(alll_handle h)
{
    return (h->flags & ALLL_FLAG_P2P_PERSISTENT) && h->persistent_ok && h->k && h->n_tiles && h->p2p_ready;
}

int p2p_persistent_begin(alll_handle h, uint64_t seed, uint64_t max_rounds, uint32_t epoch)
{
    CK(cudaSetDevice(h->device));
    if (max_rounds == 0) max_rounds = 1;
    max_rounds = std::min<uint64_t>(max_rounds, (1u << 20) - 2);       // the round lives in 20 bits of the tag
    CK(launch_reset_counters(h->d_ctr, 1, h->stream)); h->launches++;
    CK(cudaEventRecord(h->ev[2 * MAX_TIMED_ROUNDS], h->stream));
    SweepParams sp = sweep_params(h, 0u, 1u, 0u, false);                    // (tag != 0 selects the P2P form; the kernel derives parity / tag per round)
    sp.p2p_epoch = epoch & 0xFFFu;
    ClauseView pcv{};
    pcv.k = h->k;
    IncrParams ip{};
    if (h->incr_ready)
        ip = IncrParams{h->d_sh_s, h->d_rows, h->incr_stride, h->k, h->d_occ_off, h->d_occ, h->d_visited, h->d_bits, h->d_viol, h->d_ctr};
    CK(launch_solve_persistent(sp, h->resident_all, h->sweep_grid, pcv, h->k, h->d_sh_state, h->d_sh_s, mis_scratch(h, false),
                               h->n_vars, seed, (uint32_t)max_rounds, epoch, h->incr_ready ? &ip : nullptr,
                               (uint32_t)h->visited_words, h->incr_ready ? h->incr_max_vars : 0u, h->stream));
    h->launches++;
    CK(cudaEventRecord(h->ev[2 * MAX_TIMED_ROUNDS + 1], h->stream));
    return ALLL_OK;
}

int p2p_persistent_end(alll_handle h, uint64_t m_global, uint64_t launches0, alll_stats *stats)
{
    CK(cudaSetDevice(h->device));
    if (int rc = fetch_counters(h)) return rc;
    float pms = 0.f;
    CK(cudaEventElapsedTime(&pms, h->ev[2 * MAX_TIMED_ROUNDS], h->ev[2 * MAX_TIMED_ROUNDS + 1]));
    const Counters c = *h->h_ctr;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    if (getenv("ALLL_TRACE")) print_phases(c, c.n_iterations);
    if (c.p2p_error || c.done == 2)
        return fail(h, c.p2p_error == 1 ? ALLL_CAPACITY : ALLL_CUDA_ERROR,
                    c.p2p_error == 1 ? "P2P exchange region too small for a round's violated records"
                                     : c.p2p_error == 3 ? "P2P exchange: a peer aborted the solve"
                                                        : "P2P exchange: a peer did not publish its round in time");
    stats->n_iterations = c.n_iterations;
    stats->n_resamples = c.n_resamples;
    stats->sum_mis_size = c.sum_mis;
    stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;
    stats->n_clause_evals = m_global * (c.n_iterations - c.n_incr_rounds) + c.n_evals_incr;   // (n_evals_incr: this rank's share)
    stats->n_incremental_rounds = c.n_incr_rounds;
    stats->n_luby_steps = c.n_luby_steps;
    stats->n_kernel_launches = h->launches - launches0;
    stats->solve_ms = pms;
    stats->sweep_ms = (double)c.t_sweep_ns * 1e-6;
    stats->between_sweeps_ms = (double)c.t_mis_ns * 1e-6;
    stats->status = c.done == 1 ? ALLL_OK : ALLL_MAX_ROUNDS;
    return stats->status;
}

#define NEED_INSTANCE()                                                                   \
    do {                                                                                  \
        if (!h) return ALLL_BAD_ARG;                                                      \
        if (!h->has_instance) return fail(h, ALLL_NO_INSTANCE, "no instance uploaded");   \
        CK(cudaSetDevice(h->device));                                                     \
    } while (0)

} // namespace

extern "C" {

int alll_abi_version(void) { return ALLL_ABI_VERSION; }

int alll_device_count(int32_t *n)
try {
    if (!n) return ALLL_BAD_ARG;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); c = 0; }
    *n = c;
    return ALLL_OK;
}
ALLL_GUARD(nullptr)

int alll_host_alloc(uint64_t bytes, void **out)
try {
    if (!out) return ALLL_BAD_ARG;
    *out = nullptr;
    if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return ALLL_CUDA_ERROR; }
    return ALLL_OK;
}
ALLL_GUARD(nullptr)

int alll_host_free(void *p)
try {
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return ALLL_CUDA_ERROR; }
    return ALLL_OK;
}
ALLL_GUARD(nullptr)

const char *alll_last_error(alll_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int alll_create(const alll_config *cfg, alll_handle *out)
try {
    alll_handle h = nullptr;   // CK() routes the message to g_create_error while h == NULL
    if (!out) return fail(nullptr, ALLL_BAD_ARG, "out == NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, ALLL_CUDA_ERROR, std::string("no CUDA device: the solver has no CPU fallback (") +
                                                  cudaGetErrorString(e) + ")");
    int device = cfg ? cfg->device : -1;
    if (device < 0) CK(cudaGetDevice(&device));
    if (device >= n_dev) return fail(nullptr, ALLL_BAD_ARG, "device ordinal out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, ALLL_CUDA_ERROR, "device is not sm_100-class; this library carries only sm_100a code");
    alll_solver *s = new alll_solver;
    s->device = device;
    s->sm_count = prop.multiProcessorCount;
    s->clock_khz = prop.clockRate;
    s->flags = cfg ? cfg->flags : 0;
    if (const char *e = getenv("ALLL_TUNE")) s->tune = (uint32_t)strtoul(e, nullptr, 0);
    if (const char *e = getenv("ALLL_H2D_PACK")) s->h2d_pack = atoi(e) > 0 ? 1 : 0;
    const uint32_t wbuf_bytes = (SWEEP_THREADS / 32) * (WBUF + QBUF) * 4;
    const uint32_t max_bits = (uint32_t)prop.sharedMemPerBlockOptin - wbuf_bytes - 1024;
    s->smem_budget = (cfg && cfg->sweep_smem_bytes) ? cfg->sweep_smem_bytes : DEFAULT_SWEEP_SMEM;
    s->smem_budget = std::max<uint32_t>(16, std::min(s->smem_budget, max_bits));
    h = s;
    auto bail = [&](int rc) { std::string m = h->err; alll_destroy(h); g_create_error = m; return rc; };
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaStreamCreate failed"));
    if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaStreamCreate failed"));
    for (auto &ev : s->ev_chunk)
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaEventCreate failed"));
    for (auto &ev : s->ev_slot)
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaEventCreate failed"));
    if (cudaMalloc(&s->d_ctr, sizeof(Counters)) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaMalloc(counters) failed"));
    if (cudaMemset(s->d_ctr, 0, sizeof(Counters)) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaMemset(counters) failed"));
    if (cudaMallocHost(&s->h_ctr, sizeof(Counters)) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaMallocHost failed"));
    if (cudaMallocHost(&s->h_ring, sizeof(RoundNote) * ROUNDS_IN_FLIGHT) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaMallocHost failed"));
    // page-locked memory is handed out uncleared and is recycled between handles: a stale sequence number left by an
    // earlier handle's first round would look like OUR first round having retired (the host round loops poll `seq`)
    std::memset(s->h_ring, 0, sizeof(RoundNote) * ROUNDS_IN_FLIGHT);
    std::memset(s->h_ctr, 0, sizeof(Counters));
    for (auto &ev : s->ev_round)
        if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaEventCreate failed"));
    s->ev.resize(2 * MAX_TIMED_ROUNDS + 2);
    for (auto &ev : s->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(h, ALLL_CUDA_ERROR, "cudaEventCreate failed"));
    *out = s;
    return ALLL_OK;
}
ALLL_GUARD(nullptr)

int alll_destroy(alll_handle h)
try {
    if (!h) return ALLL_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    release_buffers(h);
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->d_ctr) cudaFree(h->d_ctr);
    if (h->h_ctr) cudaFreeHost(h->h_ctr);
    if (h->h_ring) cudaFreeHost(h->h_ring);
    if (h->h_bools) cudaFreeHost(h->h_bools);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->h_pack) cudaFreeHost(h->h_pack);
    for (auto &ev : h->ev_round) if (ev) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t e : h->ev_chunk) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_slot) if (e) cudaEventDestroy(e);
    delete h;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_upload_fixedk_device(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *d_lit)
try {
    if (!h) return ALLL_BAD_ARG;
    if (m && !d_lit) return fail(h, ALLL_BAD_ARG, "d_lit == NULL");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());   // the caller's buffer may have been produced on another stream
    return upload_fixedk_device_impl(h, n_vars, m, k, d_lit);
}
ALLL_GUARD(h)

static int upload_fixedk_host(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit, alll_filled_fn filled,
                              void *user, uint64_t filled_base);

int alll_upload_fixedk(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit)
try {
    return upload_fixedk_host(h, n_vars, m, k, lit, nullptr, nullptr, 0);
}
ALLL_GUARD(h)

// alll_upload_fixedk with the host buffer still being produced: see include/alll_b200.h.  filled_base: position of lit's row 0
// in the producer's numbering (multi.cu hands every device its own clause range of one buffer).
static int upload_fixedk_host(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit, alll_filled_fn filled,
                              void *user, uint64_t filled_base)
{
    if (!h) return ALLL_BAD_ARG;
    if (m && !lit) return fail(h, ALLL_BAD_ARG, "lit == NULL");
    if (k < 1 || k > MAX_K) return fail(h, ALLL_BAD_ARG, "k must be in [1, 32] for the fixed-width layout");
    CK(cudaSetDevice(h->device));
    free_instance(h);
    const size_t bytes = (size_t)std::max<uint64_t>(m * k, 1) * 4;
    POOL(h->d_stage, bytes);
    const int rc = upload_fixedk_device_impl(h, n_vars, m, k, h->d_stage, nullptr, lit, filled, user, filled_base);
    cudaStreamSynchronize(h->copy_stream);            // `lit` is the caller's again when we return, on every path
    cudaStreamSynchronize(h->stream);
    if (bytes > (4ull << 30)) {          // do not sit on a very large staging buffer
        dfree(h->d_stage);
        h->caps.erase(reinterpret_cast<void **>(&h->d_stage));
    }
    return rc;
}

int alll_upload_fixedk_streamed(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit, alll_filled_fn filled, void *user)
try {
    return upload_fixedk_host(h, n_vars, m, k, lit, filled, user, 0);
}
ALLL_GUARD(h)

int alll_upload_csr(alll_handle h, uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit)
try {
    if (!h) return ALLL_BAD_ARG;
    if (!off || (m && off[m] > off[0] && !lit)) return fail(h, ALLL_BAD_ARG, "off/lit == NULL");
    CK(cudaSetDevice(h->device));
    // width scan: refuse empty clauses (Clause.h:35-45 makes them unsatisfiable), route uniform width to planes.
    // One pass over m+1 offsets, split over the host's threads (320 MB at m = 40 M: ~40 ms on one core).
    bool uniform = m > 0;
    const uint64_t k0 = m ? off[1] - off[0] : 0;
    uint64_t kmax = 0;
    {
        const uint32_t nt = host_threads_for(m, 1u << 18);
        std::vector<uint64_t> t_kmax(nt, 0), t_bad(nt, ~0ull), t_empty(nt, ~0ull);
        std::vector<uint8_t> t_uniform(nt, 1);
        parallel_ranges(m, nt, [&](uint32_t t, uint64_t c0, uint64_t c1) {
            uint64_t km = 0;
            bool uni = true;
            for (uint64_t c = c0; c < c1; c++) {
                if (off[c + 1] < off[c]) { t_bad[t] = c; return; }
                const uint64_t w = off[c + 1] - off[c];
                if (w == 0) { t_empty[t] = c; return; }
                uni &= (w == k0);
                km = std::max(km, w);
            }
            t_kmax[t] = km; t_uniform[t] = uni;
        });
        for (uint32_t t = 0; t < nt; t++) {                  // first offender in clause order, like the serial scan
            if (t_bad[t] != ~0ull) return fail(h, ALLL_BAD_ARG, "offsets must be non-decreasing");
            if (t_empty[t] != ~0ull) return fail(h, ALLL_EMPTY_CLAUSE, "clause " + std::to_string(t_empty[t]) + " is empty and can never be satisfied");
            uniform = uniform && t_uniform[t];
            kmax = std::max(kmax, t_kmax[t]);
        }
    }
    if (uniform && k0 <= MAX_K && !(h->flags & ALLL_FLAG_FORCE_CSR)) return alll_upload_fixedk(h, n_vars, m, (uint32_t)k0, lit + off[0]);
    const uint64_t n_lit_in = m ? off[m] - off[0] : 0;
    if (m > 0 && kmax <= MAX_K && m * kmax <= 2 * n_lit_in + 1024 && !(h->flags & ALLL_FLAG_FORCE_CSR)) {
        // ragged input with modest spread: pad every clause to the widest one with copies of its first literal and
        // use the plane layout (fast sweep); true widths are kept for the independent-set / resample accounting
        // (padded rows + widths are built by all host threads straight into a page-locked staging buffer of the handle:
        // the copy then runs at the PCIe rate instead of the pageable path's fifth of it)
        const size_t pad_bytes = (size_t)m * kmax * 4, need = pad_bytes + m;
        if (h->h_stage_cap < need) {
            if (h->h_stage) { cudaFreeHost(h->h_stage); h->h_stage = nullptr; h->h_stage_cap = 0; }
            CK(cudaMallocHost(&h->h_stage, need));
            h->h_stage_cap = need;
        }
        uint32_t *padded = reinterpret_cast<uint32_t *>(h->h_stage);
        uint8_t *widths = h->h_stage + pad_bytes;
        parallel_ranges(m, host_threads_for(m, 1u << 16), [&](uint32_t, uint64_t c0, uint64_t c1) {
            for (uint64_t c = c0; c < c1; c++) {
                const uint64_t w = off[c + 1] - off[c];
                widths[c] = (uint8_t)w;
                const uint32_t *src = lit + off[c];
                uint32_t *dst = padded + c * kmax;
                for (uint64_t j = 0; j < kmax; j++) dst[j] = src[j < w ? j : 0];
            }
        });
        free_instance(h);
        POOL(h->d_stage, pad_bytes);
        POOL(h->d_width_in, m);
        CK(cudaMemcpyAsync(h->d_width_in, widths, m, cudaMemcpyHostToDevice, h->stream));
        const int rc = upload_fixedk_device_impl(h, n_vars, m, (uint32_t)kmax, h->d_stage, h->d_width_in, padded);   // chunked H2D behind the layout pass
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
        return rc;
    }

    free_instance(h);
    if (int rc = check_sizes(h, n_vars, m)) return rc;
    if (kmax > 0xFFFFFFFFull) return fail(h, ALLL_BAD_ARG, "clause too wide");
    h->n_vars = n_vars; h->m = m; h->k = 0; h->kmax = (uint32_t)kmax; h->m_pad = m; h->n_buckets = 1; h->n_tiles = 0;
    h->n_words_alloc = (uint32_t)align_up((n_vars + 31) / 32, 4);
    h->bucket_words = 0; h->resident_all = false;
    h->n_lit = m ? off[m] - off[0] : 0;
    std::vector<uint64_t> off0(m + 1);
    for (uint64_t c = 0; c <= m; c++) off0[c] = off[c] - off[0];
    POOL(h->d_off, (m + 1) * 8);
    CK(cudaMemcpyAsync(h->d_off, off0.data(), (m + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    // device layout of the warp-cooperative sweep (csr_body.cuh): literals padded to whole 128-literal chunks with at
    // least one padding position, one start bit per position, the clause rank of every chunk
    h->csr_l_pad = align_up(h->n_lit + 1, CSR_CHUNK);
    if (h->csr_l_pad / CSR_CHUNK > 0xFFFFFFF0ull) return fail(h, ALLL_BAD_ARG, "too many literals for the CSR layout");
    POOL(h->d_csr_lit, h->csr_l_pad * 4);
    if (h->n_lit) CK(cudaMemcpyAsync(h->d_csr_lit, lit + off[0], h->n_lit * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_csr_lit + h->n_lit, 0, (h->csr_l_pad - h->n_lit) * 4, h->stream));
    POOL(h->d_csr_start, h->csr_l_pad / 8);
    POOL(h->d_csr_rank, (h->csr_l_pad / CSR_CHUNK + 1) * 4);
    CK(launch_csr_build(h->d_off, m, h->n_lit, h->csr_l_pad, h->d_csr_start, h->d_csr_rank, h->stream)); h->launches += 2;
    POOL(h->d_tmp_err, 8);
    uint32_t *d_err = h->d_tmp_err;
    CK(cudaMemsetAsync(d_err, 0, 4, h->stream));
    CK(launch_validate_csr(h->d_csr_lit, h->n_lit, n_vars, d_err, h->stream)); h->launches++;
    uint32_t err_flags = 0;
    CK(cudaMemcpyAsync(&err_flags, d_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (err_flags) { free_instance(h); return fail(h, ALLL_BAD_ARG, "a literal references a variable >= n_vars"); }
    if (int rc = alloc_common(h, h->m)) return rc;
    {
        // stage the whole assignment in shared memory when it fits next to what the independent-set phases need
        h->csr_staged_words = (size_t)h->n_words_alloc * 4 <= h->smem_budget ? h->n_words_alloc : 0u;
        const CsrSweepParams cp = csr_params(h);
        int per_sm = 1;
        CK(configure_sweep_csr(cp, &per_sm));
        h->sweep_grid = (uint32_t)std::max(1, h->sm_count);                       // persistent CSR solve: one CTA per SM
        h->csr_grid = (uint32_t)std::max(1, h->sm_count * std::min(per_sm, 4));   // stand-alone sweep: as many as fit
        int ok = 0, coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
        if (coop && m > 0) CK(configure_solve_persistent_csr(cp, h->kmax, &ok));
        h->persistent_ok = ok != 0;
    }
    CK(cudaStreamSynchronize(h->stream));
    h->has_instance = true;
    return ALLL_OK;
}
ALLL_GUARD(h)

// The caller's bool array is pageable memory: stage it through a pinned buffer (one fast memcpy + one DMA instead of
// the driver's chunked pageable path).
// ---- enumerated clauses (SATInstance.h:70-153): nothing is stored, the caller's kernel launcher is the instance ----

static int upload_generator_impl(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, alll_gen_launch_fn launch, void *user,
                                 uint64_t cap_records)
{
    if (int rc = check_sizes(h, n_vars, m)) return rc;
    if (k < 1 || k > MAX_K) return fail(h, ALLL_BAD_ARG, "k must be in [1, 32]");
    if (!launch) return fail(h, ALLL_BAD_ARG, "launch == NULL");
    h->n_vars = n_vars; h->m = m; h->k = k; h->kmax = k;
    h->n_words_alloc = (uint32_t)align_up((n_vars + 31) / 32, 4);
    h->n_buckets = 1; h->bucket_words = 0; h->resident_all = false; h->m_pad = 0; h->n_tiles = 0; h->min_resident = 0;
    h->gen_cap = std::min<uint64_t>(cap_records ? cap_records : m, std::max<uint64_t>(m, 1));
    h->gen_cap = std::min<uint64_t>(std::max<uint64_t>(h->gen_cap, 1), 0xFFFFFFF0ull);
    POOL(h->d_gen_rec, h->gen_cap * (k + 1) * 4);
    h->gen_mode = true; h->gen_launch = launch; h->gen_user = user;
    if (int rc = alloc_common(h, h->gen_cap)) { free_instance(h); return rc; }
    CK(cudaStreamSynchronize(h->stream));
    h->has_instance = true;
    return ALLL_OK;
}

int alll_upload_generator(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, alll_gen_launch_fn launch, void *user,
                          uint64_t cap_records)
try {
    if (!h) return ALLL_BAD_ARG;
    CK(cudaSetDevice(h->device));
    free_instance(h);
    return upload_generator_impl(h, n_vars, m, k, launch, user, cap_records);
}
ALLL_GUARD(h)

int alll_upload_builtin_generator(alll_handle h, uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed,
                                  uint32_t d, uint64_t cap_records)
try {
    if (!h) return ALLL_BAD_ARG;
    CK(cudaSetDevice(h->device));
    free_instance(h);
    BuiltinGenerator *g = nullptr;
    if (const char *e = builtin_generator_create(kind, n_vars, m, k, seed, d, &g)) return fail(h, ALLL_BAD_ARG, e);
    const int rc = upload_generator_impl(h, n_vars, m, k, builtin_generator_launch, g, cap_records);
    if (rc != ALLL_OK) { builtin_generator_destroy(g); return rc; }
    h->gen_builtin = g;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_builtin_generator_clause(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                  uint64_t index, uint32_t *lits)
try {
    if (!lits) return ALLL_BAD_ARG;
    return builtin_generator_clause(kind, n_vars, m, k, seed, d, index, lits) ? ALLL_BAD_ARG : ALLL_OK;
}
ALLL_GUARD(nullptr)

// true when the caller's buffer is page-locked (cudaMallocHost / cudaHostRegister): the copy engine can then use it
// directly and the pinned staging hop (a 10 MB memcpy at n = 10 M) is skipped
static bool caller_buffer_is_pinned(const void *p)
{
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_pinned_bools(alll_handle h)
{
    if (h->h_bools_cap >= h->n_vars) return ALLL_OK;
    if (h->h_bools) { cudaFreeHost(h->h_bools); h->h_bools = nullptr; h->h_bools_cap = 0; }
    CK(cudaMallocHost(&h->h_bools, h->n_vars));
    h->h_bools_cap = h->n_vars;
    return ALLL_OK;
}

int alll_set_assignment(alll_handle h, const uint8_t *bools)
try {
    NEED_INSTANCE();
    if (!bools) return fail(h, ALLL_BAD_ARG, "bools == NULL");
    const uint8_t *src = bools;
    if (!caller_buffer_is_pinned(bools)) {
        if (int rc = ensure_pinned_bools(h)) return rc;
        std::memcpy(h->h_bools, bools, h->n_vars);
        src = h->h_bools;
    }
    CK(cudaMemcpyAsync(h->d_bools, src, h->n_vars, cudaMemcpyHostToDevice, h->stream));
    CK(launch_pack_bits(h->d_bools, h->n_vars, h->d_bits, h->n_words_alloc, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_get_assignment(alll_handle h, uint8_t *bools)
try {
    NEED_INSTANCE();
    if (!bools) return fail(h, ALLL_BAD_ARG, "bools == NULL");
    const bool direct = caller_buffer_is_pinned(bools);
    if (!direct)
        if (int rc = ensure_pinned_bools(h)) return rc;
    CK(launch_unpack_bits(h->d_bits, h->n_vars, h->d_bools, h->stream)); h->launches++;
    CK(cudaMemcpyAsync(direct ? bools : h->h_bools, h->d_bools, h->n_vars, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!direct) std::memcpy(bools, h->h_bools, h->n_vars);
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_randomize(alll_handle h, uint64_t seed)
try {
    NEED_INSTANCE();
    CK(launch_randomize(h->d_bits, h->n_vars, h->n_words_alloc, seed, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_eval(alll_handle h, uint32_t *ids, uint64_t cap, uint64_t *n_violated)
try {
    NEED_INSTANCE();
    if (int rc = enqueue_sweep(h)) return rc;
    if (int rc = fetch_counters(h)) return rc;
    const uint64_t n = h->h_ctr->n_viol;
    if (n_violated) *n_violated = n;
    const bool overflow = h->gen_mode && n > h->gen_cap;
    if (!overflow)
        if (int rc = copy_ids_out(h, h->gen_mode ? nullptr : h->d_viol, n, ids, cap)) return rc;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    if (overflow) return fail(h, ALLL_CAPACITY, "violated set exceeds cap_records of the enumerated instance");
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_verify(alll_handle h, int *valid)
try {
    uint64_t n = 0;
    const int rc = alll_eval(h, nullptr, 0, &n);
    if (rc == ALLL_OK && valid) *valid = n == 0;
    return rc;
}
ALLL_GUARD(h)

int alll_round(alll_handle h, uint64_t seed, uint32_t round, uint32_t *u_ids, uint64_t u_cap, uint64_t *n_u,
               uint32_t *s_ids, uint64_t s_cap, uint64_t *n_s, uint64_t *n_resampled)
try {
    NEED_INSTANCE();
    if (int rc = enqueue_sweep(h, 0u, 0u, 0xFFFFFFFFu, true)) return rc;
    if (int rc = enqueue_mis_resample(h, seed, round)) return rc;
    if (int rc = fetch_counters(h)) return rc;
    const Counters &c = *h->h_ctr;
    if (c.p2p_error) {                                        // enumerated clauses: the records of this round did not fit
        CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
        CK(cudaStreamSynchronize(h->stream));
        return fail(h, ALLL_CAPACITY, "violated set exceeds cap_records of the enumerated instance");
    }
    if (n_u) *n_u = c.last_n_viol;
    if (n_s) *n_s = c.last_n_s;
    if (n_resampled) *n_resampled = c.last_resampled;
    if (int rc = copy_ids_out(h, h->gen_mode ? nullptr : h->d_viol, c.last_n_viol, u_ids, u_cap)) return rc;
    if (int rc = copy_ids_out(h, h->d_s, c.last_n_s, s_ids, s_cap)) return rc;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_solve(alll_handle h, uint64_t seed, uint64_t max_rounds, alll_stats *stats)
try {
    NEED_INSTANCE();
    if (!stats) return fail(h, ALLL_BAD_ARG, "stats == NULL");
    std::memset(stats, 0, sizeof(*stats));
    const uint64_t launches0 = h->launches;
    CK(launch_reset_counters(h->d_ctr, 1, h->stream)); h->launches++;
    cudaEvent_t ev_begin = h->ev[2 * MAX_TIMED_ROUNDS];
    CK(cudaEventRecord(ev_begin, h->stream));
    const bool trace = getenv("ALLL_TRACE") != nullptr;

    if (h->persistent_ok && h->k && h->n_tiles && !h->gen_mode && !(h->flags & ALLL_FLAG_HOST_ROUND_LOOP)) {
        // The whole round loop in one cooperative launch (persist.cu: solve_persistent_kernel).
        if (max_rounds == 0) max_rounds = 1;
        const uint32_t cap = (uint32_t)std::min<uint64_t>(max_rounds, 0xFFFFFFFFull);
        const SweepParams sp = sweep_params(h, 0u, 0u, 0u, true);
        IncrParams ip{};
        if (h->incr_ready)
            ip = IncrParams{h->d_s, h->d_rows, h->incr_stride, h->k, h->d_occ_off, h->d_occ, h->d_visited, h->d_bits, h->d_viol, h->d_ctr};
        CK(launch_solve_persistent(sp, h->resident_all, h->sweep_grid, clause_view(h), h->kmax, h->d_state, h->d_s,
                                   mis_scratch(h, true), h->n_vars, seed, cap, 0u, h->incr_ready ? &ip : nullptr,
                                   (uint32_t)h->visited_words, h->incr_max_vars, h->stream));
        h->launches++;
        cudaEvent_t ev_end = h->ev[2 * MAX_TIMED_ROUNDS + 1];
        CK(cudaEventRecord(ev_end, h->stream));
        if (int rc = fetch_counters(h)) return rc;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
        CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;      // clears `done` for the single-step calls
        CK(cudaStreamSynchronize(h->stream));
        const Counters &c = *h->h_ctr;
        if (trace) print_phases(c, c.n_iterations);
        const int status = c.done ? ALLL_OK : ALLL_MAX_ROUNDS;
        stats->n_iterations = c.n_iterations;
        stats->n_resamples = c.n_resamples;
        stats->sum_mis_size = c.sum_mis;
        stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;     // SATInstance.h:317
        stats->n_clause_evals = h->m * (c.n_iterations - c.n_incr_rounds) + c.n_evals_incr;   // clauses actually evaluated
        stats->n_incremental_rounds = c.n_incr_rounds;
        stats->n_luby_steps = c.n_luby_steps;
        stats->n_kernel_launches = h->launches - launches0;
        stats->solve_ms = ms;
        stats->sweep_ms = (double)c.t_sweep_ns * 1e-6;            // as block 0 saw it (%globaltimer): sweep + its grid barrier
        stats->between_sweeps_ms = (double)c.t_mis_ns * 1e-6;
        stats->status = status;
        return status;
    }

    if (h->persistent_ok && !h->k && h->m && !h->gen_mode && !(h->flags & ALLL_FLAG_HOST_ROUND_LOOP)) {
        // CSR instance: the whole round loop in one cooperative launch as well (csr.cu: solve_persistent_csr_kernel)
        if (max_rounds == 0) max_rounds = 1;
        const uint32_t cap = (uint32_t)std::min<uint64_t>(max_rounds, 0xFFFFFFFFull);
        CK(launch_solve_persistent_csr(csr_params(h), h->sweep_grid, clause_view(h), h->kmax, h->d_state, h->d_s, mis_scratch(h, false),
                                       h->n_vars, seed, cap, h->stream));
        h->launches++;
        cudaEvent_t ev_end = h->ev[2 * MAX_TIMED_ROUNDS + 1];
        CK(cudaEventRecord(ev_end, h->stream));
        if (int rc = fetch_counters(h)) return rc;
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
        CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
        CK(cudaStreamSynchronize(h->stream));
        const Counters &c = *h->h_ctr;
        if (trace) print_phases(c, c.n_iterations);
        const int status = c.done ? ALLL_OK : ALLL_MAX_ROUNDS;
        stats->n_iterations = c.n_iterations;
        stats->n_resamples = c.n_resamples;
        stats->sum_mis_size = c.sum_mis;
        stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;     // SATInstance.h:317
        stats->n_clause_evals = h->m * c.n_iterations;
        stats->n_luby_steps = c.n_luby_steps;
        stats->n_kernel_launches = h->launches - launches0;
        stats->solve_ms = ms;
        stats->sweep_ms = (double)c.t_sweep_ns * 1e-6;
        stats->between_sweeps_ms = (double)c.t_mis_ns * 1e-6;
        stats->status = status;
        return status;
    }

    // Pipelined round loop (replaces the per-round host control of SATInstance.h:260-311): up to
    // ROUNDS_IN_FLIGHT rounds are enqueued ahead of the last round whose counters the host has seen.  The
    // MIS kernel that finds the violated set empty raises ctr->done; kernels enqueued behind it return at
    // entry, so the device never idles waiting for the host and nothing runs past the terminal sweep.
    int status = ALLL_MAX_ROUNDS;
    uint64_t issued = 0, retired = 0;
    const unsigned long long seq0 = h->seq;
    uint64_t last_seen_u = h->m;          // |U| of the newest retired round: violated sets shrink, so once it fits one
                                          // cluster the cooperative grid kernel is no longer enqueued
    cudaEvent_t ev_last = ev_begin;
    if (max_rounds == 0) max_rounds = 1;             // the loop body always runs once (SATInstance.h:260-261)
    while (retired < max_rounds) {
        while (issued < max_rounds && issued - retired < (uint64_t)ROUNDS_IN_FLIGHT) {
            const bool time_this = issued < (uint64_t)MAX_TIMED_ROUNDS;
            if (time_this) CK(cudaEventRecord(h->ev[2 * issued], h->stream));
            // records for the independent set only while the violated set is expected to fit them (see persist.cu)
            const bool records = last_seen_u <= 2ull * h->urec_cap;
            if (int rc = enqueue_sweep(h, 0u, 0u, (uint32_t)issued, records)) return rc;
            if (time_this) CK(cudaEventRecord(h->ev[2 * issued + 1], h->stream));
            if (h->incr_ready && issued > 0) {
                // incremental mode: the device decided at the end of the previous round which of the two kernels
                // produces this round's violated set; the other one returns at entry
                const uint32_t grid = (uint32_t)h->sm_count * 8;
                CK(launch_incr_eval(h->d_s, h->d_rows, h->incr_stride, h->k, h->d_occ_off, h->d_occ, h->d_visited,
                                    h->visited_words, h->d_bits, h->d_viol, h->d_ctr, grid, h->stream));
                h->launches++;
            }
            const int slot = (int)(issued % ROUNDS_IN_FLIGHT);
            if (int rc = enqueue_mis_resample(h, seed, (uint32_t)issued, last_seen_u > MIS_CLUSTER_MAX_U, &h->h_ring[slot],
                                              seq0 + issued + 1, true, records))
                return rc;
            CK(cudaEventRecord(h->ev_round[slot], h->stream));      // timing only: marks the end of this round on the device
            issued++;
        }
        const int slot = (int)(retired % ROUNDS_IN_FLIGHT);
        {   // wait for the MIS kernel of round `retired` to announce itself in pinned memory
            volatile unsigned long long *seq = &h->h_ring[slot].seq;
            uint32_t spins = 0;
            while (*seq != seq0 + retired + 1) {
                if ((++spins & 0x3FFu) == 0) {
                    const cudaError_t q = cudaStreamQuery(h->stream);
                    if (q == cudaSuccess && *seq != seq0 + retired + 1)
                        return fail(h, ALLL_CUDA_ERROR, "round finished without announcing itself");
                    if (q != cudaSuccess && q != cudaErrorNotReady)
                        return fail(h, ALLL_CUDA_ERROR, std::string("round loop: ") + cudaGetErrorString(q));
                }
            }
        }
        ev_last = h->ev_round[slot];
        if (trace && retired < (uint64_t)MAX_TIMED_ROUNDS) {
            // ALLL_TRACE=1: per-round device times on stderr (sweep kernel | sweep end -> MIS kernels done)
            float t_sweep = 0.f, t_mis = 0.f;
            cudaEventSynchronize(h->ev_round[slot]);
            cudaEventElapsedTime(&t_sweep, h->ev[2 * retired], h->ev[2 * retired + 1]);
            cudaEventElapsedTime(&t_mis, h->ev[2 * retired + 1], h->ev_round[slot]);
            fprintf(stderr, "[alll trace] round %llu: |U|=%u |S|=%u sweep=%.1f us mis=%.1f us\n",
                    (unsigned long long)retired, h->h_ring[slot].n_viol, h->h_ring[slot].n_s, t_sweep * 1e3, t_mis * 1e3);
        }
        retired++;
        last_seen_u = h->h_ring[slot].n_viol;
        if (last_seen_u == 0) { status = ALLL_OK; break; }                    // SATInstance.h:285-287
        if (last_seen_u == 0xFFFFFFFFu) { status = ALLL_CAPACITY; break; }    // enumerated clauses: records did not fit
    }
    const uint64_t useful_rounds = retired;              // rounds whose sweep actually ran (incl. the terminal one)
    h->seq = seq0 + issued;
    CK(cudaStreamSynchronize(h->stream));                 // drain the speculative (no-op) rounds
    float ms = 0.f;
    if (useful_rounds) CK(cudaEventElapsedTime(&ms, ev_begin, ev_last));
    double sweep_ms = 0.0, between_ms = 0.0;
    const int timed = (int)std::min<uint64_t>(useful_rounds, MAX_TIMED_ROUNDS);
    for (int i = 0; i < timed; i++) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]));
        sweep_ms += t;
        if (i + 1 < timed) {
            CK(cudaEventElapsedTime(&t, h->ev[2 * i + 1], h->ev[2 * i + 2]));
            between_ms += t;
        }
    }
    if (int rc = fetch_counters(h)) return rc;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;      // clears `done` for the single-step calls
    CK(cudaStreamSynchronize(h->stream));
    const Counters &c = *h->h_ctr;
    if (trace) print_phases(c, useful_rounds);
    stats->n_iterations = c.n_iterations;
    stats->n_resamples = c.n_resamples;
    stats->sum_mis_size = c.sum_mis;
    stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;     // SATInstance.h:317
    stats->n_clause_evals = h->m * (c.n_iterations - c.n_incr_rounds) + c.n_evals_incr;   // clauses actually evaluated
    stats->n_incremental_rounds = c.n_incr_rounds;
    stats->n_luby_steps = c.n_luby_steps;
    stats->n_kernel_launches = h->launches - launches0;
    stats->solve_ms = ms;
    stats->sweep_ms = timed ? sweep_ms * ((double)c.n_iterations / timed) : 0.0;
    stats->between_sweeps_ms = timed > 1 ? between_ms * ((double)(c.n_iterations - 1) / (timed - 1)) : 0.0;
    stats->status = status;
    if (status == ALLL_CAPACITY) return fail(h, ALLL_CAPACITY, "violated set exceeds cap_records of the enumerated instance");
    return status;
}
ALLL_GUARD(h)

// ---- clause-range sharded mode ----------------------------------------------------

int alll_set_id_base(alll_handle h, uint64_t id_base)
try {
    if (!h) return ALLL_BAD_ARG;
    if (id_base > 0xFFFFFFFFull) return fail(h, ALLL_BAD_ARG, "id_base must fit 32 bits");
    h->id_base = (uint32_t)id_base;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_shard_sweep(alll_handle h, uint32_t *d_records, uint64_t cap_records, uint64_t *n_local)
try {
    NEED_INSTANCE();
    if (!h->k || h->use_width || h->gen_mode) return fail(h, ALLL_BAD_ARG, "sharded mode needs stored clauses of uniform width");
    if (!d_records && cap_records) return fail(h, ALLL_BAD_ARG, "d_records == NULL");
    if (int rc = enqueue_sweep(h)) return rc;
    if (cap_records) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>((cap_records + 255) / 256, (uint64_t)h->sm_count * 8);
        CK(launch_export_records(clause_view(h), h->d_viol, h->d_ctr, d_records, cap_records, std::max(grid, 1u), h->stream));
        h->launches++;
    }
    if (int rc = fetch_counters(h)) return rc;
    const uint64_t n = h->h_ctr->n_viol;
    if (n_local) *n_local = n;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
    if (n > cap_records) return fail(h, ALLL_CAPACITY, "record buffer too small for the local violated set");
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_shard_round(alll_handle h, const uint32_t *d_records, const uint64_t *counts, uint32_t n_blocks,
                     uint64_t block_cap, uint64_t seed, uint32_t round, uint64_t *n_total, uint64_t *n_s,
                     uint64_t *n_resampled)
try {
    NEED_INSTANCE();
    if (!h->k || h->gen_mode) return fail(h, ALLL_BAD_ARG, "sharded mode needs the fixed-width layout");
    if (!counts || n_blocks == 0 || n_blocks > MAX_SHARDS) return fail(h, ALLL_BAD_ARG, "bad shard count");
    uint32_t prefix[MAX_SHARDS + 1];
    uint64_t total = 0;
    for (uint32_t b = 0; b < n_blocks; b++) {
        if (counts[b] > block_cap) return fail(h, ALLL_BAD_ARG, "count exceeds block capacity");
        prefix[b] = (uint32_t)total;
        total += counts[b];
    }
    if (total > 0xFFFFFFF0ull) return fail(h, ALLL_BAD_ARG, "violated set too large");
    prefix[n_blocks] = (uint32_t)total;
    if (total && !d_records) return fail(h, ALLL_BAD_ARG, "d_records == NULL");
    const uint64_t cap = std::max<uint64_t>(align_up(total, 1024), 1024);
    POOL(h->d_sh_planes, cap * h->k * 4);
    POOL(h->d_sh_ids, cap * 4);
    POOL(h->d_sh_iota, cap * 4);
    POOL(h->d_sh_s, cap * 4);
    POOL(h->d_sh_state, cap);
    const uint32_t grid = (uint32_t)std::min<uint64_t>((total + 255) / 256 + 1, (uint64_t)h->sm_count * 8);
    CK(launch_repack_records(d_records, block_cap, h->k, n_blocks, prefix, h->d_sh_planes, cap, h->d_sh_ids,
                             h->d_sh_iota, h->d_ctr, grid, h->stream));
    h->launches++;
    ClauseView cv{};
    cv.planes = h->d_sh_planes; cv.m_pad = cap; cv.k = h->k; cv.orig_id = h->d_sh_ids; cv.id_base = 0;
    CK(launch_mis_resample_args(cv, h->k, h->d_sh_iota, h->d_sh_state, h->d_sh_s, mis_scratch(h, false), h->n_vars, h->d_bits,
                                h->d_ctr, seed, round, h->mis_grid, total > MIS_CLUSTER_MAX_U, nullptr, 0ull, nullptr, 0u, 0u, 0u, 0u, h->stream));
    h->launches += total > MIS_CLUSTER_MAX_U ? 2 : 1;
    if (int rc = fetch_counters(h)) return rc;
    const Counters &c = *h->h_ctr;
    if (n_total) *n_total = total;
    if (n_s) *n_s = c.last_n_s;
    if (n_resampled) *n_resampled = c.last_resampled;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;      // clears `done` after a terminal round
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_get_stats(alll_handle h, alll_stats *stats)
try {
    NEED_INSTANCE();
    if (!stats) return fail(h, ALLL_BAD_ARG, "stats == NULL");
    if (int rc = fetch_counters(h)) return rc;
    const Counters &c = *h->h_ctr;
    std::memset(stats, 0, sizeof(*stats));
    stats->n_iterations = c.n_iterations;
    stats->n_resamples = c.n_resamples;
    stats->sum_mis_size = c.sum_mis;
    stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;
    stats->n_clause_evals = h->m * c.n_iterations;
    stats->n_luby_steps = c.n_luby_steps;
    stats->n_kernel_launches = h->launches;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_reset_stats(alll_handle h)
try {
    NEED_INSTANCE();
    CK(launch_reset_counters(h->d_ctr, 1, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}
ALLL_GUARD(h)

// ---- sharded mode with the exchange fused into the kernels (NVLink P2P stores, CUDA IPC mappings) ---------

int alll_p2p_create(alll_handle h, uint32_t world, uint32_t rank, uint64_t cap_records, uint8_t handle_out[64])
try {
    if (!handle_out) return h ? fail(h, ALLL_BAD_ARG, "handle_out == NULL") : ALLL_BAD_ARG;
    bool fresh = false;
    if (int rc = p2p_region_setup(h, world, rank, cap_records, &fresh)) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (fresh || !h->p2p_ipc_valid) {
        cudaIpcMemHandle_t ipc;
        CK(cudaIpcGetMemHandle(&ipc, h->d_p2p_region));
        std::memcpy(h->p2p_ipc, &ipc, 64);
        h->p2p_ipc_valid = true;
    }
    std::memcpy(handle_out, h->p2p_ipc, 64);
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_p2p_connect(alll_handle h, const uint8_t *handles)
try {
    NEED_INSTANCE();
    if (!h->d_p2p_region || !handles) return fail(h, ALLL_BAD_ARG, "call alll_p2p_create first");
    void *bases[MAX_SHARDS] = {};
    for (uint32_t q = 0; q < h->p2p_world; q++) {
        if (q == h->p2p_rank) { bases[q] = h->d_p2p_region; continue; }
        // a peer that kept its region (same allocation => same IPC handle) keeps our mapping of it: opening and closing
        // IPC mappings costs milliseconds, which an end-to-end step that re-uploads the instance would pay every time
        if (h->p2p_peer[q] && std::memcmp(h->p2p_peer_ipc[q], handles + (size_t)q * 64, 64) == 0) { bases[q] = h->p2p_peer[q]; continue; }
        if (h->p2p_peer[q]) { cudaIpcCloseMemHandle(h->p2p_peer[q]); h->p2p_peer[q] = nullptr; }
        cudaIpcMemHandle_t ipc;
        std::memcpy(&ipc, handles + (size_t)q * 64, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_peer[q] = ptr;
        std::memcpy(h->p2p_peer_ipc[q], handles + (size_t)q * 64, 64);
        bases[q] = ptr;
    }
    for (uint32_t q = h->p2p_world; q < MAX_SHARDS; q++)
        if (h->p2p_peer[q]) { cudaIpcCloseMemHandle(h->p2p_peer[q]); h->p2p_peer[q] = nullptr; }
    return p2p_link_up(h, bases);
}
ALLL_GUARD(h)

int alll_solve_p2p(alll_handle h, uint64_t seed, uint64_t max_rounds, uint64_t m_global, uint32_t epoch, alll_stats *stats)
try {
    NEED_INSTANCE();
    if (!stats) return fail(h, ALLL_BAD_ARG, "stats == NULL");
    if (!h->p2p_ready) return fail(h, ALLL_BAD_ARG, "call alll_p2p_create / alll_p2p_connect first");
    std::memset(stats, 0, sizeof(*stats));
    const uint64_t launches0 = h->launches;
    if (max_rounds == 0) max_rounds = 1;
    max_rounds = std::min<uint64_t>(max_rounds, (1u << 20) - 2);       // the round lives in 20 bits of the tag
    if (p2p_persistent_possible(h)) {
        // every rank: the whole sharded solve in one cooperative launch (persist.cu: solve_persistent_kernel, p2p branch)
        if (int rc = p2p_persistent_begin(h, seed, max_rounds, epoch)) return rc;
        return p2p_persistent_end(h, m_global, launches0, stats);
    }
    CK(launch_reset_counters(h->d_ctr, 1, h->stream)); h->launches++;
    cudaEvent_t ev_begin = h->ev[2 * MAX_TIMED_ROUNDS];
    CK(cudaEventRecord(ev_begin, h->stream));
    int status = ALLL_MAX_ROUNDS;
    uint64_t issued = 0, retired = 0;
    const unsigned long long seq0 = h->seq;
    uint64_t last_seen_u = m_global;
    cudaEvent_t ev_last = ev_begin;
    ClauseView cv{};                                      // unused by the kernels in P2P mode except k
    cv.k = h->k;
    bool failed = false;
    while (retired < max_rounds && !failed) {
        while (issued < max_rounds && issued - retired < (uint64_t)ROUNDS_IN_FLIGHT) {
            const uint32_t parity = (uint32_t)(issued & 1u);
            const uint32_t tag = ((epoch & 0xFFFu) << 20) | (uint32_t)(issued + 1);
            const bool time_this = issued < (uint64_t)MAX_TIMED_ROUNDS;
            if (time_this) CK(cudaEventRecord(h->ev[2 * issued], h->stream));
            if (int rc = enqueue_sweep(h, parity, tag)) return rc;
            if (time_this) CK(cudaEventRecord(h->ev[2 * issued + 1], h->stream));
            const int slot = (int)(issued % ROUNDS_IN_FLIGHT);
            const bool with_grid = last_seen_u > MIS_CLUSTER_MAX_U;
            CK(launch_mis_resample_args(cv, h->k, nullptr, h->d_sh_state, h->d_sh_s, mis_scratch(h, false), h->n_vars, h->d_bits,
                                        h->d_ctr, seed, (uint32_t)issued, h->mis_grid, with_grid, &h->h_ring[slot],
                                        seq0 + issued + 1, h->d_p2p_link, parity, tag, 0u, 0u, h->stream));
            h->launches += with_grid ? 2 : 1;
            CK(cudaEventRecord(h->ev_round[slot], h->stream));
            issued++;
        }
        const int slot = (int)(retired % ROUNDS_IN_FLIGHT);
        {
            volatile unsigned long long *seq = &h->h_ring[slot].seq;
            uint32_t spins = 0;
            while (*seq != seq0 + retired + 1) {
                if ((++spins & 0x3FFu) == 0) {
                    const cudaError_t q = cudaStreamQuery(h->stream);
                    if (q == cudaSuccess && *seq != seq0 + retired + 1) { failed = true; break; }   // kernels returned on `done` after an error
                    if (q != cudaSuccess && q != cudaErrorNotReady)
                        return fail(h, ALLL_CUDA_ERROR, std::string("p2p round loop: ") + cudaGetErrorString(q));
                }
            }
        }
        if (failed) break;
        ev_last = h->ev_round[slot];
        retired++;
        last_seen_u = h->h_ring[slot].n_viol;
        if (last_seen_u == 0xFFFFFFFFu) { failed = true; break; }
        if (last_seen_u == 0) { status = ALLL_OK; break; }
    }
    h->seq = seq0 + issued;
    CK(cudaStreamSynchronize(h->stream));
    if (int rc = fetch_counters(h)) return rc;
    const Counters c = *h->h_ctr;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    if (failed || c.p2p_error)
        return fail(h, c.p2p_error == 1 ? ALLL_CAPACITY : ALLL_CUDA_ERROR,
                    c.p2p_error == 1 ? "P2P exchange region too small for a round's violated records"
                                     : c.p2p_error == 3 ? "P2P exchange: a peer aborted the solve"
                                                        : "P2P exchange: a peer did not publish its round in time");
    float ms = 0.f;
    if (retired) CK(cudaEventElapsedTime(&ms, ev_begin, ev_last));
    double sweep_ms = 0.0, between_ms = 0.0;
    const int timed = (int)std::min<uint64_t>(retired, MAX_TIMED_ROUNDS);
    for (int i = 0; i < timed; i++) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]));
        sweep_ms += t;
        if (i + 1 < timed) { CK(cudaEventElapsedTime(&t, h->ev[2 * i + 1], h->ev[2 * i + 2])); between_ms += t; }
    }
    stats->n_iterations = c.n_iterations;
    stats->n_resamples = c.n_resamples;
    stats->sum_mis_size = c.sum_mis;
    stats->avg_mis_size = c.n_iterations ? c.sum_mis / c.n_iterations : 0;
    stats->n_clause_evals = m_global * c.n_iterations;
    stats->n_luby_steps = c.n_luby_steps;
    stats->n_kernel_launches = h->launches - launches0;
    stats->solve_ms = ms;
    stats->sweep_ms = timed ? sweep_ms * ((double)c.n_iterations / timed) : 0.0;
    stats->between_sweeps_ms = timed > 1 ? between_ms * ((double)(c.n_iterations - 1) / (timed - 1)) : 0.0;
    stats->status = status;
    return status;
}
ALLL_GUARD(h)

// ---- batched small instances / seed portfolio -------------------------------------------------------------

int alll_batch_upload(alll_handle h, uint32_t n_instances, uint64_t n_vars, uint32_t k, const uint64_t *clause_off,
                      const uint32_t *lit)
try {
    if (!h) return ALLL_BAD_ARG;
    CK(cudaSetDevice(h->device));
    h->has_batch = false;
    if (n_instances == 0 || !clause_off) return fail(h, ALLL_BAD_ARG, "no instances");
    if (k < 1 || k > MAX_K) return fail(h, ALLL_BAD_ARG, "k must be in [1, 32]");
    if (n_vars == 0 || n_vars > (1u << 24)) return fail(h, ALLL_BAD_ARG, "n_vars out of range for the batched path");
    const uint64_t total = clause_off[n_instances] - clause_off[0];
    if (total && !lit) return fail(h, ALLL_BAD_ARG, "lit == NULL");
    std::vector<uint32_t> off(n_instances + 1), mm(n_instances);
    std::vector<uint64_t> src(n_instances + 1);
    uint64_t pos = 0;
    uint32_t m_max = 0;
    for (uint32_t i = 0; i < n_instances; i++) {
        if (clause_off[i + 1] < clause_off[i]) return fail(h, ALLL_BAD_ARG, "offsets must be non-decreasing");
        const uint64_t m = clause_off[i + 1] - clause_off[i];
        if (m > 0x3FFFFFFFull || pos > 0xF0000000ull) return fail(h, ALLL_BAD_ARG, "batch too large");
        off[i] = (uint32_t)pos;
        mm[i] = (uint32_t)m;
        src[i] = clause_off[i] - clause_off[0];
        m_max = std::max(m_max, (uint32_t)m);
        pos = align_up(pos + m, 4);
    }
    off[n_instances] = (uint32_t)pos;
    src[n_instances] = total;
    const uint64_t m_pad = std::max<uint64_t>(align_up(pos, 4), 4);
    const uint32_t n_words = (uint32_t)((n_vars + 31) / 32);
    int max_smem = 0;
    CK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    if (batch_smem_bytes((uint32_t)n_vars, n_words, m_max) + 64 > (size_t)max_smem)
        return fail(h, ALLL_BAD_ARG, "instance too large for the one-CTA-per-instance path (use alll_upload_* + alll_solve)");
    POOL(h->d_b_lit, std::max<uint64_t>(total * k, 1) * 4);
    POOL(h->d_b_planes, m_pad * k * 4);
    POOL(h->d_b_off, (n_instances + 1) * 4);
    POOL(h->d_b_m, n_instances * 4);
    POOL(h->d_b_src_off, (n_instances + 1) * 8);
    POOL(h->d_tmp_err, 8);
    if (total) CK(cudaMemcpyAsync(h->d_b_lit, lit + clause_off[0] * k, total * k * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_b_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_b_m, mm.data(), mm.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_b_src_off, src.data(), src.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->d_tmp_err, 0, 8, h->stream));
    CK(cudaMemsetAsync(h->d_b_planes, 0, m_pad * k * 4, h->stream));
    CK(launch_batch_transpose(h->d_b_lit, h->d_b_src_off, h->d_b_off, n_instances, k, (uint32_t)n_vars, h->d_b_planes, m_pad,
                              h->d_tmp_err, h->stream));
    h->launches++;
    uint32_t err = 0;
    CK(cudaMemcpyAsync(&err, h->d_tmp_err, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (err) return fail(h, ALLL_BAD_ARG, "a literal references a variable >= n_vars");
    h->b_n_inst = n_instances; h->b_n_vars = (uint32_t)n_vars; h->b_n_words = n_words; h->b_k = k; h->b_m_max = m_max;
    h->b_m_pad = m_pad;
    h->has_batch = true;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_batch_solve(alll_handle h, uint32_t n_jobs, const uint64_t *seeds, uint64_t max_rounds, int portfolio,
                     uint8_t *assignments, alll_batch_stats *stats, int32_t *winner, double *device_ms)
try {
    if (!h) return ALLL_BAD_ARG;
    if (!h->has_batch) return fail(h, ALLL_NO_INSTANCE, "no batch uploaded");
    CK(cudaSetDevice(h->device));
    if (n_jobs == 0 || !seeds || !stats) return fail(h, ALLL_BAD_ARG, "n_jobs / seeds / stats");
    const bool shared = portfolio == 2;
    if (shared && !h->d_flag) return fail(h, ALLL_BAD_ARG, "portfolio == 2 needs alll_flag_create / alll_flag_open first");
    if ((uint64_t)h->b_job_base + n_jobs > 0x7FFFFFFFull) return fail(h, ALLL_BAD_ARG, "job ids must fit 31 bits");
    int *d_winner_word = nullptr;
    if (!portfolio && n_jobs != h->b_n_inst) return fail(h, ALLL_BAD_ARG, "n_jobs must equal the number of uploaded instances");
    static_assert(sizeof(alll_batch_stats) == sizeof(BatchJobStats), "ABI mirror");
    POOL(h->d_b_seeds, (size_t)n_jobs * 8);
    POOL(h->d_b_stats, (size_t)n_jobs * sizeof(BatchJobStats));
    POOL(h->d_b_bits, (size_t)n_jobs * h->b_n_words * 4);
    POOL(h->d_b_winner, 4);
    POOL(h->d_b_retry, ((size_t)n_jobs + 1) * 4);
    const int minus1 = -1;
    CK(cudaMemcpyAsync(h->d_b_seeds, seeds, (size_t)n_jobs * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_b_winner, &minus1, 4, cudaMemcpyHostToDevice, h->stream));
    d_winner_word = shared ? h->d_flag : h->d_b_winner;      // (the shared word is reset by its owner between portfolios)
    CK(cudaMemsetAsync(h->d_b_bits, 0, (size_t)n_jobs * h->b_n_words * 4, h->stream));
    cudaEvent_t e0 = h->ev[0], e1 = h->ev[1];
    CK(cudaEventRecord(e0, h->stream));
    int n_launched = 1;
    CK(launch_batch_solve(h->d_b_planes, h->b_m_pad, h->d_b_off, h->d_b_m, h->b_n_inst, h->b_n_vars, h->b_n_words, h->b_k,
                          h->b_m_max, h->d_b_seeds, max_rounds, h->d_b_bits, h->d_b_stats, portfolio ? 1 : 0, d_winner_word,
                          (int)h->b_job_base, shared ? 1 : 0, n_jobs, h->d_b_retry, &n_launched, h->stream));
    h->launches += n_launched;
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaMemcpyAsync(stats, h->d_b_stats, (size_t)n_jobs * sizeof(BatchJobStats), cudaMemcpyDeviceToHost, h->stream));
    int w = -1;
    CK(cudaMemcpyAsync(&w, d_winner_word, 4, cudaMemcpyDeviceToHost, h->stream));
    if (assignments) {
        const uint64_t total = (uint64_t)n_jobs * h->b_n_vars;
        POOL(h->d_b_bytes, total);
        CK(launch_batch_unpack(h->d_b_bits, h->b_n_words, h->b_n_vars, total, h->d_b_bytes, h->stream)); h->launches++;
        CK(cudaMemcpyAsync(assignments, h->d_b_bytes, total, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (device_ms) *device_ms = ms;
    if (winner) *winner = w;
    return ALLL_OK;
}
ALLL_GUARD(h)

// ---- multi-GPU portfolio: the first-SAT word shared by all ranks ------------------------------

int alll_flag_create(alll_handle h, uint8_t handle_out[64])
try {
    if (!h || !handle_out) return ALLL_BAD_ARG;
    CK(cudaSetDevice(h->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (h->d_flag) { if (h->flag_owner) cudaFree(h->d_flag); else if (!h->flag_borrowed) cudaIpcCloseMemHandle(h->d_flag); h->d_flag = nullptr; h->flag_borrowed = false; }
    CK(cudaMalloc(&h->d_flag, 256));
    h->flag_owner = true;
    CK(cudaMemset(h->d_flag, 0xFF, 256));                    // -1: open
    cudaIpcMemHandle_t ipc;
    CK(cudaIpcGetMemHandle(&ipc, h->d_flag));
    std::memcpy(handle_out, &ipc, 64);
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_flag_open(alll_handle h, const uint8_t *handle)
try {
    if (!h || !handle) return ALLL_BAD_ARG;
    CK(cudaSetDevice(h->device));
    if (h->d_flag) { if (h->flag_owner) cudaFree(h->d_flag); else if (!h->flag_borrowed) cudaIpcCloseMemHandle(h->d_flag); h->d_flag = nullptr; h->flag_borrowed = false; }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, handle, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->d_flag = static_cast<int *>(p);
    h->flag_owner = false;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_flag_reset(alll_handle h)
try {
    if (!h) return ALLL_BAD_ARG;
    if (!h->d_flag || !h->flag_owner) return fail(h, ALLL_BAD_ARG, "only the rank that created the flag resets it");
    CK(cudaSetDevice(h->device));
    CK(cudaMemsetAsync(h->d_flag, 0xFF, 4, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_flag_read(alll_handle h, int64_t *value)
try {
    if (!h || !value) return ALLL_BAD_ARG;
    if (!h->d_flag) return fail(h, ALLL_BAD_ARG, "no flag");
    CK(cudaSetDevice(h->device));
    int w = -1;
    CK(cudaMemcpyAsync(&w, h->d_flag, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *value = w;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_batch_set_job_base(alll_handle h, uint32_t job_base)
try {
    if (!h) return ALLL_BAD_ARG;
    h->b_job_base = job_base;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_time_sweep(alll_handle h, uint32_t reps, double *ms_per_sweep, uint64_t *n_violated)
try {
    NEED_INSTANCE();
    if (reps == 0) return fail(h, ALLL_BAD_ARG, "reps == 0");
    double total = 0.0;
    uint32_t done = 0;
    while (done < reps) {
        const uint32_t batch = std::min<uint32_t>(reps - done, MAX_TIMED_ROUNDS);
        for (uint32_t i = 0; i < batch; i++) {
            CK(cudaEventRecord(h->ev[2 * i], h->stream));
            if (int rc = enqueue_sweep(h)) return rc;
            CK(cudaEventRecord(h->ev[2 * i + 1], h->stream));
            if (i + 1 < batch || done + batch < reps) { CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++; }
        }
        CK(cudaStreamSynchronize(h->stream));
        for (uint32_t i = 0; i < batch; i++) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]));
            total += t;
        }
        done += batch;
    }
    if (int rc = fetch_counters(h)) return rc;
    if (n_violated) *n_violated = h->h_ctr->n_viol;
    CK(launch_reset_counters(h->d_ctr, 0, h->stream)); h->launches++;
    CK(cudaStreamSynchronize(h->stream));
    if (ms_per_sweep) *ms_per_sweep = total / reps;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_launch_count(alll_handle h, uint64_t *n)
try {
    if (!h || !n) return ALLL_BAD_ARG;
    *n = h->launches;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_layout_info(alll_handle h, uint64_t info[6])
try {
    if (!h || !info) return ALLL_BAD_ARG;
    if (!h->has_instance) return fail(h, ALLL_NO_INSTANCE, "no instance uploaded");
    info[0] = h->m;
    info[1] = h->k;
    info[2] = h->n_buckets;
    info[3] = h->m_pad;
    info[4] = h->k ? h->m_pad * h->k * 4 : h->csr_l_pad * 4 + h->csr_l_pad / 8 + (h->csr_l_pad / CSR_CHUNK + 1) * 4;   // bytes one sweep reads
    info[5] = h->k ? sweep_planes_smem_bytes(h->bucket_words) : sweep_csr_smem_bytes(h->csr_staged_words, SWEEP_THREADS);
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_sweep_info(alll_handle h, uint64_t info[4])
try {
    if (!h || !info) return ALLL_BAD_ARG;
    if (!h->has_instance) return fail(h, ALLL_NO_INSTANCE, "no instance uploaded");
    const uint32_t eager = h->k ? std::min<uint32_t>(h->k, EAGER_PLANES) : 0u;
    info[0] = h->packed_on ? 1 : 0;
    info[1] = h->packed_on ? std::min<uint32_t>(h->min_resident, 2u) : 0u;
    info[2] = h->packed_on ? 16u : 4u * eager;
    info[3] = h->min_resident;
    return ALLL_OK;
}
ALLL_GUARD(h)

int alll_upload_info(alll_handle h, uint64_t info[4])
try {
    if (!h || !info) return ALLL_BAD_ARG;
    info[0] = h->up_link_bytes;
    info[1] = h->up_packed_chunks;
    info[2] = h->up_raw_chunks;
    info[3] = h->up_pack_threads;
    return ALLL_OK;
}
ALLL_GUARD(h)

} // extern "C"

// ---- internal entry points for the single-process multi-GPU layer (multi.cu); not part of the C ABI -----------------
namespace alll {

void *internal_p2p_region(alll_handle h) { return h ? h->d_p2p_region : nullptr; }

int internal_upload_fixedk_streamed(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit, alll_filled_fn filled,
                                    void *user, uint64_t filled_base)
{
    return upload_fixedk_host(h, n_vars, m, k, lit, filled, user, filled_base);
}

// alll_p2p_create without the CUDA IPC export: the peers live in this process and address the region directly
int internal_p2p_create_local(alll_handle h, uint32_t world, uint32_t rank, uint64_t cap_records)
{
    bool fresh = false;
    return p2p_region_setup(h, world, rank, cap_records, &fresh);
}

// true iff every clause of the CSR has the width of the first one (all host threads)
bool internal_csr_is_uniform(const uint64_t *off, uint64_t m)
{
    if (m == 0) return false;
    const uint64_t k0 = off[1] - off[0];
    const uint32_t nt = host_threads_for(m, 1u << 18);
    std::vector<uint8_t> ok(nt, 1);
    parallel_ranges(m, nt, [&](uint32_t t, uint64_t c0, uint64_t c1) {
        bool u = true;
        for (uint64_t c = c0; c < c1 && u; c++) u = off[c + 1] - off[c] == k0;
        ok[t] = u;
    });
    return std::all_of(ok.begin(), ok.end(), [](uint8_t v) { return v != 0; });
}

int internal_p2p_connect_ptrs(alll_handle h, void *const *regions)
{
    if (!h) return ALLL_BAD_ARG;
    if (!h->has_instance) return fail(h, ALLL_NO_INSTANCE, "no instance uploaded");
    if (!h->d_p2p_region || !regions) return fail(h, ALLL_BAD_ARG, "call alll_p2p_create first");
    CK(cudaSetDevice(h->device));
    return p2p_link_up(h, regions);
}

bool internal_p2p_persistent_possible(alll_handle h) { return h && p2p_persistent_possible(h); }

int internal_solve_p2p_begin(alll_handle h, uint64_t seed, uint64_t max_rounds, uint32_t epoch, uint64_t *launches0)
{
    if (!h) return ALLL_BAD_ARG;
    if (!h->has_instance || !h->p2p_ready) return fail(h, ALLL_NO_INSTANCE, "no sharded instance");
    *launches0 = h->launches;
    return p2p_persistent_begin(h, seed, max_rounds, epoch);
}

int internal_solve_p2p_end(alll_handle h, uint64_t m_global, uint64_t launches0, alll_stats *stats)
{
    std::memset(stats, 0, sizeof(*stats));
    return p2p_persistent_end(h, m_global, launches0, stats);
}

int *internal_flag_ptr(alll_handle h) { return h ? h->d_flag : nullptr; }

// multi.cu: several device slots upload side by side from one host -- their pack threads would oversubscribe its cores, so
// the packed transport stays off there unless ALLL_H2D_PACK=1 asks for it
void internal_h2d_pack_default_off(alll_handle h) { if (h && h->h2d_pack < 0) h->h2d_pack = 0; }

// Same-process peer: use the winner word another handle of this process created (plain device pointer, peer access on).
int internal_flag_attach(alll_handle h, int *word)
{
    if (!h || !word) return ALLL_BAD_ARG;
    if (h->d_flag) { if (h->flag_owner) cudaFree(h->d_flag); else if (!h->flag_borrowed) cudaIpcCloseMemHandle(h->d_flag); h->d_flag = nullptr; }
    h->d_flag = word;
    h->flag_owner = false;
    h->flag_borrowed = true;
    return ALLL_OK;
}

int internal_device(alll_handle h) { return h ? h->device : -1; }

} // namespace alll

