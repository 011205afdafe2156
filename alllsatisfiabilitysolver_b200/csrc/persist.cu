// persist.cu -- the whole solve in one cooperative launch: solve_persistent_kernel and its launchers.
// Device code of the phases: sweep_body.cuh (clause evaluation + compaction), mis_body.cuh (independent set + resample),
// incr_body.cuh (incremental re-evaluation).
#include "incr_body.cuh"
#include "mis_body.cuh"
#include "sweep_body.cuh"

namespace alll {

// ---- the whole solve in one launch ------------------------------------------------------------------------
// Replaces the round loop of parallel_solve (SATInstance.h:260-311) for the plane layout: sweep -> grid barrier ->
// independent set + resample -> grid barrier, repeated on the device until a sweep finds no violated clause.
// Why one kernel: the independent-set phases are a few microseconds of work but, launched as kernels of their own
// behind a sweep that has just streamed > 1 GB through L2, they spend 20-100 us per round fetching their code cold
// from DRAM (every phase cost about 0.6 us per 128-byte line of instructions it touched, whatever the size of U --
// profiles/r01_mis_phases.md).  A persistent kernel keeps that code in the SMs' instruction caches from the second
// round on, and launch gaps, event records and the host round trip disappear as well.
// One CTA per SM (cooperative launch).  |U| is accumulated in one of two counters selected by round parity: the
// one for round r+1 is cleared during the independent-set phase of round r, when nobody adds to it.
template <int K, int RB, int RC, int E, bool PK>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) solve_persistent_kernel(const SweepParams sp, const MisParams mp_arg,
                                                                           const uint32_t max_rounds, const uint32_t epoch,
                                                                           const IncrParams ip, const uint32_t visited_words)
{
    // The independent-set bodies are out-of-line functions: they get the parameter block through a pointer, and a
    // pointer to kernel parameters would force a per-thread local-memory copy.  One copy per CTA in shared memory
    // (thread 0 also keeps the per-round exchange parity / tag of the sharded mode up to date in it).
    __shared__ MisParams s_mp;
    __shared__ uint32_t s_prefix[MAX_SHARDS + 1];
    if (threadIdx.x == 0) s_mp = mp_arg;
    __syncthreads();
    const MisParams &mp = s_mp;
    // Nothing but `round` stays in registers across the sweep body: the streaming loop is latency-bound at the register
    // limit of a 512-thread CTA, and every value kept live around it (timers, barrier handle, thread indices) costs it
    // scheduling freedom -- the same body ran 0.185 ms as a kernel of its own and 0.203 ms in here before this was moved
    // to shared memory / recomputed (profiles/r02_pipelined_layout.md).
    __shared__ unsigned long long s_t0;      // block 0: %globaltimer at sweep entry
    __shared__ uint32_t s_prev_n_u;          // |U| of the previous round (m / 2^K before the first)
    Counters *const c = sp.ctr;
#define ALLL_LEAD (blockIdx.x == 0 && threadIdx.x == 0)
    if (threadIdx.x == 0) {
        s_prev_n_u = (uint32_t)min((uint64_t)0xFFFFFFFFu, ((uint64_t)sp.n_tiles * TILE) >> K);
        g_remote_dirty = 0u;
    }
    __syncthreads();
    for (uint32_t round = 0; round < max_rounds; ++round) {
        const uint32_t par = round & 1u, tag = ((epoch & 0xFFFu) << 20) | (round + 1u);
        if (ALLL_LEAD) {
            s_t0 = global_ns();
            if (round < DBG_ROUNDS) c->dbg[round][0] = s_t0;
        }
        // records next to the violated list only while the violated set is expected to fit them (the previous round's
        // |U|, or m / 2^K before the first round): writing the first urec_cap records of a larger set is wasted work
        const bool rec_on = (uint64_t)s_prev_n_u <= 2ull * sp.urec_cap;
        if (threadIdx.x == 0) { s_mp.p2p_parity = par; s_mp.p2p_tag = tag; s_mp.urec_cap = rec_on ? sp.urec_cap : 0u; }
        // incremental mode (ip.rows != NULL): the round that just ended decided whether this round's violated set comes
        // from the occurrence lists of the variables it resampled (same set as the sweep's, incremental.cu) or from a sweep
        const bool incremental = ip.rows != nullptr && round > 0 && gm::ld_cg(&c->incr_next) != 0;
        if (incremental) incr_eval_body(ip, gm::ld_cg(&c->last_n_s), &c->n_viol_pp[par], IncrP2P{sp.p2p, s_prefix, par, sp.orig_id, sp.id_base, c, sp.p2p_epoch + 1u});
        else sweep_planes_body<K, RB, RC, E, false, PK>(sp, &c->n_viol_pp[par], par, rec_on);
        if (threadIdx.x == 0 && round == 1u && blockIdx.x < 256u) c->dbg_cta[blockIdx.x] = global_ns();
        GridBarrier bar{cg::this_grid()};
        const bool lead = ALLL_LEAD;
        const bool p2p = sp.p2p != nullptr;  // clause-range sharded solve: every GPU runs this kernel on its range
        const uint32_t first = blockIdx.x * SWEEP_THREADS + threadIdx.x, stride = gridDim.x * SWEEP_THREADS;
        unsigned long long t1 = 0;
        uint32_t n_u;
        if (p2p) {
            // Fused exchange: the violated records went straight into every GPU's region during the sweep.  Every CTA orders
            // its record stores (NVLink) before its ticket; the CTA that draws the last ticket stores this rank's count +
            // arrival flag into every GPU, our own included -- so the wait for all flags of the round below is also the
            // grid barrier of this GPU (no grid.sync() on the exchange's critical path).
            if (lead && round < DBG_ROUNDS) c->dbg_x[round][0] = global_ns();
            __syncthreads();
            if (threadIdx.x == 0) {
                if (g_remote_dirty) {                    // (only CTAs that stored records this round pay the system-scope fence)
                    fence_acq_rel_sys();
                    g_remote_dirty = 0u;
                }
                const unsigned int t = atomicAdd(&c->cta_done, 1u);
                if (lead && round < DBG_ROUNDS) c->dbg_x[round][1] = global_ns();
                if (t == gridDim.x - 1) {
                    c->cta_done = 0;                     // next round's tickets are drawn behind two grid barriers
                    fence_acq_rel_sys();                 // the other CTAs' tickets (and what they fenced) happen before our flag
                    const P2PLink &L = *sp.p2p;
                    const unsigned long long word = ((unsigned long long)tag << 32) | gm::ld_cg(&c->n_viol_pp[par]);
                    for (uint32_t q = 0; q < L.world; q++) *(volatile unsigned long long *)&L.hdr[q]->cf[par][L.rank] = word;
                    if (round < DBG_ROUNDS) c->dbg_x[round][2] = global_ns();
                }
            }
            n_u = p2p_wait(mp, s_prefix);
            if (lead && round < DBG_ROUNDS) c->dbg_x[round][3] = global_ns();
        } else {
            bar.sync();
            n_u = gm::ld_cg(&c->n_viol_pp[par]);
        }
        if (lead) {
            t1 = global_ns();
            c->t_sweep_ns += t1 - s_t0;
            if (round < DBG_ROUNDS) { c->dbg[round][1] = t1; c->dbg[round][2] = t1; }
            c->n_viol_pp[par ^ 1u] = 0;
        }
        if (incremental)                                 // the first-visit bits of this round: nobody reads them before the next one
            for (uint32_t i = first; i < visited_words; i += stride) ip.visited[i] = 0u;
        if (n_u == 0xFFFFFFFFu) {                        // a peer overflowed its exchange area, aborted or never arrived: stop
            if (lead) {                                  // (p2p_wait left the reason: 2 = time-out, 3 = a peer aborted)
                c->p2p_error = c->p2p_error ? c->p2p_error : 2;
                c->done = 2;
            }
            break;
        }
        if (n_u == 0) {                                  // SATInstance.h:285-287; the terminal sweep counts (:261)
            if (lead) {
                gm::red_add(&c->n_iterations, 1ull);
                if (incremental) c->n_incr_rounds += 1;
                c->last_n_viol = 0;
                c->last_n_s = 0;
                c->last_resampled = 0;
                c->done = 1;
            }
            break;
        }
        if (n_u <= SMALL_U && (uint64_t)n_u * mp.kmax <= HSLOTS / 2 && mp.small_ok) {
            if (blockIdx.x == 0) {
                mis_small_body(mp, round, s_prefix, n_u);
                if (threadIdx.x == 0) finish_round(mp, round, n_u, 0u);
            }
        } else {
            if ((uint64_t)n_u <= (uint64_t)stride * mp.cache_items) mis_resample_body<GridBarrier, true>(mp, round, bar, s_prefix, first, stride, n_u);
            else mis_resample_body<GridBarrier, false>(mp, round, bar, s_prefix, first, stride, n_u);
            bar.sync();
            if (lead) finish_round(mp, round, n_u, 2u);
        }
        // While the round's last barrier gathers the CTAs (in a small round 147 of them have been idle since the sweep):
        // pull the first tiles of this CTA's next sweep from HBM into L2, so the sweep's first loads do not start cold.
        if (threadIdx.x == 32 && sp.prefetch_tiles != 0 && !(sp.tune & TUNE_NO_NEXT_SWEEP_PREFETCH)) {
            const uint32_t r0 = __ldg(sp.run_begin + blockIdx.x), r1 = __ldg(sp.run_begin + blockIdx.x + 1);
            if (r0 < r1) {
                const uint4 run = __ldg(reinterpret_cast<const uint4 *>(sp.runs) + r0);
                const uint32_t *const stream = PK ? sp.packed : sp.planes;
                constexpr int NS = PK ? 4 : E;
                for (uint32_t t = run.x; t < run.y && t < run.x + 2u + sp.prefetch_tiles; t++)
#pragma unroll
                    for (int j = 0; j < NS; j++) tma_prefetch_l2(stream + (uint64_t)j * sp.m_pad + (uint64_t)t * TILE, TILE * 4);
            }
        }
        bar.sync();                                      // new assignment visible to every SM before it is staged again
        if (threadIdx.x == 0) s_prev_n_u = n_u;          // (read again after the __syncthreads() inside the next sweep's staging... and here:)
        __syncthreads();
        if (lead) c->t_mis_ns += global_ns() - t1;
    }
#undef ALLL_LEAD
}

// ---- launchers ------------------------------------------------------------------------

namespace {
struct PersistOp {
    const SweepParams &p;
    uint32_t grid;
    size_t smem;
    cudaStream_t s;
    const MisParams *mp;          // NULL: configure (shared-memory opt-in + occupancy) instead of launching
    uint32_t max_rounds, epoch;
    const IncrParams *ip;
    uint32_t visited_words;
    int *max_ctas_per_sm;
    template <int K, int RB, int RC, int E, bool PK> cudaError_t run()
    {
        if (mp == nullptr) {
            cudaError_t e = cudaFuncSetAttribute(solve_persistent_kernel<K, RB, RC, E, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_ctas_per_sm, solve_persistent_kernel<K, RB, RC, E, PK>, SWEEP_THREADS, smem);
        }
        void *args[] = {(void *)&p, (void *)mp, (void *)&max_rounds, (void *)&epoch, (void *)ip, (void *)&visited_words};
        return cudaLaunchCooperativeKernel((const void *)solve_persistent_kernel<K, RB, RC, E, PK>, dim3(grid), dim3(SWEEP_THREADS), args, smem, s);
    }
};
} // namespace

// ---- persistent solve kernel: shared memory = the sweep's, or what the independent-set phases need if that is more
static size_t persistent_smem_bytes(uint32_t bucket_words, uint32_t kmax)
{
    const size_t small_words = mis_small_words(SWEEP_THREADS, kmax);
    const size_t one_item = (size_t)SWEEP_THREADS * mis_cache_words(kmax);
    size_t b = sweep_smem_bytes_for(bucket_words);
    if (small_words * 4 <= 200u * 1024u) b = b > small_words * 4 ? b : small_words * 4;
    else if (one_item * 4 <= 200u * 1024u) b = b > one_item * 4 ? b : one_item * 4;
    return b;
}

static void persistent_fill(MisParams &mp, size_t smem)
{
    mp.cache_items = (uint32_t)((smem / 4 / SWEEP_THREADS) / mis_cache_words(mp.kmax));
    mp.small_ok = mis_small_words(SWEEP_THREADS, mp.kmax) * 4 <= smem ? 1u : 0u;
}

// ok_out: 1 when the instance can be solved by the persistent kernel on this device (k <= 8, one CTA per SM fits)
cudaError_t configure_solve_persistent(const SweepParams &p, bool resident_all, uint32_t kmax, int *ok_out)
{
    *ok_out = 0;
    if (p.k == 0 || p.k > 8) return cudaSuccess;
    int per_sm = 0;
    PersistOp op{p, 0u, persistent_smem_bytes(p.bucket_words, kmax), 0, nullptr, 0u, 0u, nullptr, 0u, &per_sm};
    const cudaError_t e = dispatch_variant(p, resident_all, op);
    if (e != cudaSuccess) return e;
    *ok_out = per_sm >= 1;
    return cudaSuccess;
}

cudaError_t launch_solve_persistent(const SweepParams &p, bool resident_all, uint32_t grid, const ClauseView &cv, uint32_t kmax,
                                    uint8_t *state, uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint64_t seed,
                                    uint32_t max_rounds, uint32_t epoch, const IncrParams *incr, uint32_t visited_words,
                                    uint32_t incr_max_vars, cudaStream_t s)
{
    if (p.runs == nullptr || grid != p.run_grid) return cudaErrorInvalidValue;      // the run lists are cut for one grid size
    const size_t smem = persistent_smem_bytes(p.bucket_words, kmax);
    MisParams mp{};
    mp.cv = cv; mp.viol = p.p2p ? nullptr : p.viol; mp.state = state; mp.s_slots = s_slots;
    mp.p2p = p.p2p;                                    // sharded: U = the record blocks in our exchange region
    mp.claim = sc.claim;
    mp.n_vars = n_vars; mp.bits = const_cast<uint32_t *>(p.bits); mp.ctr = p.ctr; mp.seed = seed; mp.kmax = kmax;
    mp.urec = sc.urec; mp.urec_cap = sc.urec_cap;
    mp.tune = p.tune;
    persistent_fill(mp, smem);
    mp.incr_max_vars = incr ? incr_max_vars : 0u;
    const IncrParams no_incr{};
    PersistOp op{p, grid, smem, s, &mp, max_rounds, epoch, incr ? incr : &no_incr, incr ? visited_words : 0u, nullptr};
    return dispatch_variant(p, resident_all, op);
}

} // namespace alll
