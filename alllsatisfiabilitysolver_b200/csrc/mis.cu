// mis.cu -- K3 (maximal independent set of the violated clauses) + K4 (resample).
//
// Replaces populate_mis_parallel (SATInstance.h:391-451; greedy, O(|S|*|U|*k^2), one OpenMP fork/join per
// picked clause) and resample_clauses (SATInstance.h:340-365).
//
// K3: fixed-priority Luby.  Every violated clause c gets the key (Philox priority(seed, round, id), id).
//     Every undecided clause atomicMin-s its key into the claim word of every variable it touches; a clause
//     that owns all its claims joins S and marks its variables TAKEN; a clause that sees a TAKEN variable
//     drops out; the rest claim again for the next step (in the variable's other claim word).
//     Iterated to a fixed point this is exactly the greedy independent set in ascending key order
//     (what oracle/alll_oracle.c:alll_oracle_priority_mis computes sequentially) -- independent and
//     maximal like the reference's set (SATInstance.h:415-447), and a pure function of (seed, round, U):
//     no dependence on thread schedule, compaction order or shard count.
//     Keys carry a step tag in the top bits that decreases every step, so claims of clauses that lost
//     or dropped out in earlier steps are undercut without a reset pass.
// K4: every variable of every clause in S is redrawn from Philox(seed, RESAMPLE, round, var); bits are
//     written with atomicOr/atomicAnd on the packed word (idempotent, so a variable occurring twice in a
//     clause is harmless).  The same pass restores the claim words of everything U touched (clean-after-use).
//
// Where the claim words live: claim[v][2] -- the words of a variable for even and odd Luby steps share one 16-byte
// pair, so whatever a round does to a variable touches a single 32-byte sector.  After a sweep has streamed the
// literals through L2 those sectors are cold in DRAM; the first touch of a round is a fire-and-forget reduction
// (red.min), every later access hits L2.  (Measured and rejected, profiles/r01_mis_phases.md: a compact
// open-addressing table over the touched variables -- denser sectors, so cheaper steps, but the atomicCAS inserts
// return values from cold lines and cost more than they saved at every size of U.)
//
// Latency structure: literals, priority, id and state of every violated clause are held in shared memory; all claim
// reads of a step are issued together, and deciding step s is fused with claiming for step s+1: one barrier and one
// memory round trip per Luby step however wide the clauses are.  The bodies are out-of-line functions with explicit
// global-space memory operations (mis_body.cuh) shared with the persistent solve kernel of persist.cu, where they run
// between sweeps without kernel boundaries; Philox is one out-of-line function and literal loops are only unrolled
// where loads must overlap (the first version compiled to 160 KB per kernel, all of it fetched cold after a sweep).
//
// Two kernels are enqueued per round and decide on the device which one acts (the host does not know
// |U|: rounds are enqueued speculatively, see capi.cu):
//   mis_cluster_kernel : |U| <= 8192 -- one thread-block cluster of 8 x 1024 threads, one clause per
//                        thread, phases separated by the hardware cluster barrier; |U| <= 512 is done by
//                        CTA 0 alone with the claims in a shared-memory hash table (mis_small_body);
//   mis_grid_kernel    : larger |U| -- cooperative launch, phases separated by grid-wide barriers.
#include "mis_body.cuh"

namespace cg = cooperative_groups;

namespace alll {

// First MIS kernel of a round: owns the terminal case (|U| == 0) and violated sets that fit one cluster.
__global__ void __cluster_dims__(CL_SIZE, 1, 1) __launch_bounds__(CL_THREADS) mis_cluster_kernel(const MisParams p_arg, const uint32_t round)
{
    __shared__ MisParams s_p;            // the bodies take the parameter block by reference: one copy per CTA, not one per thread
    if (threadIdx.x == 0) s_p = p_arg;
    __syncthreads();
    const MisParams &p = s_p;
    __shared__ uint32_t s_prefix[MAX_SHARDS + 1];
    const unsigned long long t_entry = global_ns();
    if (ld_u32(&p.ctr->done)) return;             // speculative round behind the terminal one
    // |U|: left by the sweep kernel that ran before us, or (sharded P2P mode) the sum over all ranks' record blocks
    const uint32_t n_u = p.p2p ? p2p_wait(p, s_prefix) : ld_u32(&p.ctr->n_viol);
    if (blockIdx.x == 0 && threadIdx.x == 0 && round < DBG_ROUNDS) {
        p.ctr->dbg[round][1] = t_entry;
        p.ctr->dbg[round][2] = global_ns();
    }
    if (n_u == 0xFFFFFFFFu || (p.u_cap && n_u > p.u_cap)) {   // a peer overflowed or never arrived / records did not fit: stop the solve
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.ctr->p2p_error = p.ctr->p2p_error ? p.ctr->p2p_error : (n_u == 0xFFFFFFFFu ? 2 : 1);
            p.ctr->n_viol = 0;
            p.ctr->done = 1;
            announce(p, 0xFFFFFFFFu, 0u);
        }
        return;
    }
    if (n_u == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            gm::red_add(&p.ctr->n_iterations, 1ull);  // the terminal all-satisfied sweep counts (SATInstance.h:261,285-287)
            if (ld_u32(&p.ctr->incr_next)) p.ctr->n_incr_rounds += 1;
            p.ctr->last_n_viol = 0;
            p.ctr->last_n_s = 0;
            p.ctr->last_resampled = 0;
            p.ctr->done = 1;
            p.ctr->n_viol = 0;
            p.ctr->handled_tag = p.p2p_tag;
            announce(p, 0u, 0u);
        }
        return;
    }
    if (n_u > CLUSTER_U && p.grid_follows) return;  // the grid kernel behind us takes it (else: strided, slower, still exact)
    if (n_u <= SMALL_U && (uint64_t)n_u * p.kmax <= HSLOTS / 2 && p.cache_items) {
        if (blockIdx.x != 0) return;              // uniform over the cluster: nobody waits on a cluster barrier below
        mis_small_body(p, round, s_prefix, n_u);
        if (threadIdx.x == 0) finish_round(p, round, n_u, 0u);
        return;
    }
    ClusterBarrier bar;
    if (n_u <= CLUSTER_U * p.cache_items) mis_resample_body<ClusterBarrier, true>(p, round, bar, s_prefix, blockIdx.x * CL_THREADS + threadIdx.x, CLUSTER_U, n_u);
    else mis_resample_body<ClusterBarrier, false>(p, round, bar, s_prefix, blockIdx.x * CL_THREADS + threadIdx.x, CLUSTER_U, n_u);
    bar.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) finish_round(p, round, n_u, 1u);
}

// Second MIS kernel of a round (cooperative launch): violated sets too large for one cluster.
__global__ void __launch_bounds__(GRID_THREADS) mis_grid_kernel(const MisParams p_arg, const uint32_t round)
{
    __shared__ MisParams s_p;
    if (threadIdx.x == 0) s_p = p_arg;
    __syncthreads();
    const MisParams &p = s_p;
    __shared__ uint32_t s_prefix[MAX_SHARDS + 1];
    const unsigned long long t_entry = global_ns();
    if (ld_u32(&p.ctr->done)) return;
    if (p.p2p && ld_u32(&p.ctr->handled_tag) == p.p2p_tag) return;     // the cluster kernel already did this round
    const uint32_t n_u = p.p2p ? p2p_wait(p, s_prefix) : ld_u32(&p.ctr->n_viol);  // 0 when the cluster kernel handled it
    if (n_u <= CLUSTER_U || n_u == 0xFFFFFFFFu) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && round < DBG_ROUNDS) {
        p.ctr->dbg[round][1] = t_entry;
        p.ctr->dbg[round][2] = global_ns();
    }
    GridBarrier bar{cg::this_grid()};
    const uint32_t stride = gridDim.x * GRID_THREADS;
    if ((uint64_t)n_u <= (uint64_t)stride * p.cache_items) mis_resample_body<GridBarrier, true>(p, round, bar, s_prefix, blockIdx.x * GRID_THREADS + threadIdx.x, stride, n_u);
    else mis_resample_body<GridBarrier, false>(p, round, bar, s_prefix, blockIdx.x * GRID_THREADS + threadIdx.x, stride, n_u);
    bar.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) finish_round(p, round, n_u, 2u);
}

// Per-round scratch reset + (optionally) totals reset.
__global__ void reset_counters_kernel(Counters *c, int reset_totals)
{
    c->n_viol = 0;
    c->n_s = 0;
    c->n_resampled_round = 0;
    c->last_n_viol = 0;
    c->last_n_s = 0;
    c->last_resampled = 0;
    c->done = 0;
    c->cta_done = 0;
    c->p2p_error = 0;
    c->incr_next = 0;
    c->n_viol_pp[0] = 0;
    c->n_viol_pp[1] = 0;
    c->luby_bar[0] = 0;
    c->luby_bar[1] = 0;
    c->t_sweep_ns = 0;
    c->t_mis_ns = 0;
    if (reset_totals) {
        for (int s = 0; s < 16; s++) c->dbg_step[s][0] = 0;
        for (int r = 0; r < 32; r++) c->dbg_x[r][0] = c->dbg_x[r][3] = 0;
        c->dbg_cta[0] = 0;
        c->n_incr_rounds = 0;
        c->n_evals_incr = 0;
        c->n_iterations = 0;
        c->sum_mis = 0;
        c->n_resamples = 0;
        c->n_luby_steps = 0;
    }
}

// slots -> caller clause ids (for alll_eval / alll_round outputs)
__global__ void map_ids_kernel(const ClauseView cv, const uint32_t *slots, uint32_t n, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cv.id(slots ? slots[i] : i);
}

// ---- host side --------------------------------------------------------------------------------------

// per cached clause: kmax literals, priority, id, width | state
static uint32_t grid_cache_items(uint32_t kmax) { return GRID_SMEM_WORDS_PER_THREAD / mis_cache_words(kmax); }
static size_t grid_smem_bytes(uint32_t kmax) { return (size_t)grid_cache_items(kmax) * mis_cache_words(kmax) * GRID_THREADS * 4; }
// the cluster kernel caches its one clause per thread whenever that fits the opt-in shared memory
static uint32_t cluster_cache_items(uint32_t kmax) { return mis_small_words(CL_THREADS, kmax) * 4 <= 212u * 1024u ? 1u : 0u; }
static size_t cluster_smem_bytes(uint32_t kmax)
{
    if (!cluster_cache_items(kmax)) return 0;
    return mis_small_words(CL_THREADS, kmax) * 4;        // mis_small_body's layout; mis_resample_body's one clause per thread fits in it
}

// Called once per upload on the handle's device: shared-memory opt-in + cooperative grid size.
cudaError_t mis_configure(int device, uint32_t kmax, uint32_t *grid_out)
{
    cudaError_t e = cudaFuncSetAttribute(mis_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cluster_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mis_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mis_grid_kernel, GRID_THREADS, grid_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm > 4) per_sm = 4;          // 4 x 256 threads per SM is plenty for an atomics-bound kernel
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid_out = (uint32_t)(per_sm * sms);
    return cudaSuccess;
}

// Enqueues the MIS kernels of one round: the cluster kernel, and (with_grid) the cooperative grid kernel behind it.
// Callers that know the violated set is small skip the grid kernel: a cooperative launch is not free even as a no-op.
cudaError_t launch_mis_resample_args(const ClauseView &cv, uint32_t kmax, const uint32_t *viol, uint8_t *state,
                                     uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint32_t *bits,
                                     Counters *ctr, uint64_t seed, uint32_t round, uint32_t grid, bool with_grid,
                                     RoundNote *note, unsigned long long seq, const P2PLink *p2p, uint32_t p2p_parity,
                                     uint32_t p2p_tag, uint32_t incr_max_vars, uint32_t u_cap, cudaStream_t s)
{
    MisParams p{cv, viol, state, s_slots, sc.claim, n_vars, bits, ctr, seed, kmax,
                cluster_cache_items(kmax), cluster_cache_items(kmax), with_grid ? 1u : 0u, note, seq, sc.urec, sc.urec_cap, p2p, p2p_parity, p2p_tag,
                incr_max_vars, u_cap};
    mis_cluster_kernel<<<CL_SIZE, CL_THREADS, cluster_smem_bytes(kmax), s>>>(p, round);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !with_grid) return e;
    p.cache_items = grid_cache_items(kmax);
    void *args[] = {(void *)&p, (void *)&round};
    return cudaLaunchCooperativeKernel((const void *)mis_grid_kernel, dim3(grid), dim3(GRID_THREADS), args,
                                       grid_smem_bytes(kmax), s);
}

cudaError_t launch_reset_counters(Counters *c, int reset_totals, cudaStream_t s)
{
    reset_counters_kernel<<<1, 1, 0, s>>>(c, reset_totals);
    return cudaGetLastError();
}

cudaError_t launch_map_ids(const ClauseView &cv, const uint32_t *slots, uint32_t n, uint32_t *out, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    map_ids_kernel<<<(n + 255) / 256, 256, 0, s>>>(cv, slots, n, out);
    return cudaGetLastError();
}

} // namespace alll
