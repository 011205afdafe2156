// csr.cu -- variable-width clauses on the device: layout build, the warp-cooperative CSR sweep as a kernel of its own
// and the whole solve of a CSR instance in one cooperative launch.  Device code of the sweep: csr_body.cuh.
//
// Replaces, for input that stays in CSR form, the same reference sites as sweep.cu / persist.cu: the violated sweep
// SATInstance.h:273-280 -> Clause::is_not_satisfied (Clause.h:34-46) over clauses of arbitrary width (Clause.h:20-28),
// and the round loop of parallel_solve (SATInstance.h:260-311).
#include "csr_body.cuh"
#include "mis_body.cuh"

namespace alll {

// ---- layout build (once per upload) -------------------------------------------------------------------------------
// start bits: one atomicOr per clause; padding positions [n_lit, l_pad) are starts of their own (ids >= m, never emitted)
__global__ void __launch_bounds__(256) csr_mark_starts_kernel(const uint64_t *__restrict__ off, uint64_t m, uint64_t n_lit,
                                                               uint64_t l_pad, uint32_t *start)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const uint64_t p = off[i];
        atomicOr(&start[p >> 5], 1u << (p & 31u));
    } else if (i - m < l_pad - n_lit) {
        const uint64_t p = n_lit + (i - m);
        atomicOr(&start[p >> 5], 1u << (p & 31u));
    }
}

// chunk_rank[c] = number of clauses starting before position 128 * c = lower_bound(off[0..m), 128 * c); clauses behind
// the real ones (padding starts) continue the count.
__global__ void __launch_bounds__(256) csr_chunk_rank_kernel(const uint64_t *__restrict__ off, uint64_t m, uint64_t n_lit,
                                                              uint32_t n_chunks, uint32_t *chunk_rank)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    const uint64_t target = (uint64_t)c * CSR_CHUNK;
    uint64_t r;
    if (target >= n_lit) r = m + (target - n_lit);
    else {
        uint64_t lo = 0, hi = m;                       // first clause with off >= target
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (off[mid] < target) lo = mid + 1; else hi = mid;
        }
        r = lo;
    }
    chunk_rank[c] = (uint32_t)r;
}

cudaError_t launch_csr_build(const uint64_t *off, uint64_t m, uint64_t n_lit, uint64_t l_pad, uint32_t *start, uint32_t *chunk_rank,
                             cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(start, 0, l_pad / 8, s);
    if (e != cudaSuccess) return e;
    const uint64_t items = m + (l_pad - n_lit);
    csr_mark_starts_kernel<<<(uint32_t)((items + 255) / 256), 256, 0, s>>>(off, m, n_lit, l_pad, start);
    const uint32_t n_chunks = (uint32_t)(l_pad / CSR_CHUNK);
    csr_chunk_rank_kernel<<<(n_chunks + 1 + 255) / 256, 256, 0, s>>>(off, m, n_lit, n_chunks, chunk_rank);
    return cudaGetLastError();
}

// ---- the sweep as a kernel of its own (alll_eval / alll_round / host round loop / alll_time_sweep) ---------------
template <bool STAGED>
__global__ void __launch_bounds__(SWEEP_THREADS, STAGED ? 1 : 2) sweep_csr_warp_kernel(const CsrSweepParams p)
{
    if (__ldcg(&p.ctr->done)) return;
    sweep_csr_body<STAGED>(p, &p.ctr->n_viol);
}

// ctas_per_sm: how many CTAs of the stand-alone sweep fit one SM (the L2-lookup variant is latency-bound: more resident
// warps hide more lookups; the staged variant owns the shared memory and runs one CTA per SM)
cudaError_t configure_sweep_csr(const CsrSweepParams &p, int *ctas_per_sm)
{
    const size_t smem = sweep_csr_smem_bytes(p.staged_words, SWEEP_THREADS);
    cudaError_t e = p.staged_words ? cudaFuncSetAttribute(sweep_csr_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                   : cudaFuncSetAttribute(sweep_csr_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = p.staged_words ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, sweep_csr_warp_kernel<true>, SWEEP_THREADS, smem)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, sweep_csr_warp_kernel<false>, SWEEP_THREADS, smem);
    if (e == cudaSuccess && *ctas_per_sm < 1) *ctas_per_sm = 1;
    return e;
}

cudaError_t launch_sweep_csr(const CsrSweepParams &p, uint32_t grid, cudaStream_t s)
{
    const size_t smem = sweep_csr_smem_bytes(p.staged_words, SWEEP_THREADS);
    if (p.staged_words) sweep_csr_warp_kernel<true><<<grid, SWEEP_THREADS, smem, s>>>(p);
    else sweep_csr_warp_kernel<false><<<grid, SWEEP_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

// ---- the whole solve of a CSR instance in one cooperative launch --------------------------------------------------
// Same structure as solve_persistent_kernel (persist.cu): sweep -> grid barrier -> |U| == 0 ? done : independent set +
// resample (CTA 0 alone for small |U|, all CTAs otherwise) -> grid barrier -> next round.  The independent-set bodies read
// the clauses through the CSR view (off[] / lit[]); violated "slots" are clause ids.
template <bool STAGED>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) solve_persistent_csr_kernel(const CsrSweepParams sp, const MisParams mp_arg,
                                                                               const uint32_t max_rounds)
{
    __shared__ MisParams s_mp;
    if (threadIdx.x == 0) s_mp = mp_arg;
    __syncthreads();
    const MisParams &mp = s_mp;
    GridBarrier bar{cg::this_grid()};
    Counters *const c = sp.ctr;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const uint32_t first = blockIdx.x * SWEEP_THREADS + threadIdx.x, stride = gridDim.x * SWEEP_THREADS;
    unsigned long long t_sweep = 0, t_mis = 0;
    for (uint32_t round = 0; round < max_rounds; ++round) {
        const uint32_t par = round & 1u;
        unsigned long long t0 = 0, t1 = 0;
        if (lead) {
            t0 = global_ns();
            if (round < DBG_ROUNDS) c->dbg[round][0] = t0;
        }
        sweep_csr_body<STAGED>(sp, &c->n_viol_pp[par]);
        bar.sync();
        const uint32_t n_u = gm::ld_cg(&c->n_viol_pp[par]);
        if (lead) {
            t1 = global_ns();
            t_sweep += t1 - t0;
            if (round < DBG_ROUNDS) { c->dbg[round][1] = t1; c->dbg[round][2] = t1; }
            c->n_viol_pp[par ^ 1u] = 0;
        }
        if (n_u == 0) {                                  // SATInstance.h:285-287; the terminal sweep counts (:261)
            if (lead) {
                gm::red_add(&c->n_iterations, 1ull);
                c->last_n_viol = 0;
                c->last_n_s = 0;
                c->last_resampled = 0;
                c->done = 1;
            }
            break;
        }
        __syncthreads();                                 // the sweep's shared memory (staged assignment) is reused below
        if (n_u <= SMALL_U && (uint64_t)n_u * mp.kmax <= HSLOTS / 2 && mp.small_ok) {
            if (blockIdx.x == 0) {
                mis_small_body(mp, round, nullptr, n_u);
                if (threadIdx.x == 0) finish_round(mp, round, n_u, 0u);
            }
        } else {
            if (mp.cache_items && (uint64_t)n_u <= (uint64_t)stride * mp.cache_items) mis_resample_body<GridBarrier, true>(mp, round, bar, nullptr, first, stride, n_u);
            else mis_resample_body<GridBarrier, false>(mp, round, bar, nullptr, first, stride, n_u);
            bar.sync();
            if (lead) finish_round(mp, round, n_u, 2u);
        }
        bar.sync();                                      // new assignment visible to every SM before it is staged again
        if (lead) t_mis += global_ns() - t1;
    }
    if (lead) {
        c->t_sweep_ns = t_sweep;
        c->t_mis_ns = t_mis;
    }
}

static size_t persistent_csr_smem_bytes(uint32_t staged_words, uint32_t kmax)
{
    const size_t small_words = mis_small_words(SWEEP_THREADS, kmax);
    const size_t one_item = (size_t)SWEEP_THREADS * mis_cache_words(kmax);
    size_t b = sweep_csr_smem_bytes(staged_words, SWEEP_THREADS);
    if (small_words * 4 <= 200u * 1024u) b = b > small_words * 4 ? b : small_words * 4;
    else if (one_item * 4 <= 200u * 1024u) b = b > one_item * 4 ? b : one_item * 4;
    return b;
}

// ok_out: 1 when the persistent CSR kernel fits this device (one CTA per SM) and the clauses are narrow enough for the
// shared-memory clause cache of the independent-set phases
cudaError_t configure_solve_persistent_csr(const CsrSweepParams &p, uint32_t kmax, int *ok_out)
{
    *ok_out = 0;
    if (kmax == 0 || kmax > 64) return cudaSuccess;
    const size_t smem = persistent_csr_smem_bytes(p.staged_words, kmax);
    int per_sm = 0;
    cudaError_t e;
    if (p.staged_words) {
        e = cudaFuncSetAttribute(solve_persistent_csr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_persistent_csr_kernel<true>, SWEEP_THREADS, smem);
    } else {
        e = cudaFuncSetAttribute(solve_persistent_csr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, solve_persistent_csr_kernel<false>, SWEEP_THREADS, smem);
    }
    if (e != cudaSuccess) return e;
    *ok_out = per_sm >= 1;
    return cudaSuccess;
}

cudaError_t launch_solve_persistent_csr(const CsrSweepParams &p, uint32_t grid, const ClauseView &cv, uint32_t kmax, uint8_t *state,
                                        uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint64_t seed, uint32_t max_rounds,
                                        cudaStream_t s)
{
    const size_t smem = persistent_csr_smem_bytes(p.staged_words, kmax);
    MisParams mp{};
    mp.cv = cv; mp.viol = p.viol; mp.state = state; mp.s_slots = s_slots;
    mp.claim = sc.claim;
    mp.n_vars = n_vars; mp.bits = const_cast<uint32_t *>(p.bits); mp.ctr = p.ctr; mp.seed = seed; mp.kmax = kmax;
    mp.cache_items = (uint32_t)((smem / 4 / SWEEP_THREADS) / mis_cache_words(kmax));
    mp.small_ok = mis_small_words(SWEEP_THREADS, kmax) * 4 <= smem ? 1u : 0u;
    void *args[] = {(void *)&p, (void *)&mp, (void *)&max_rounds};
    return p.staged_words
               ? cudaLaunchCooperativeKernel((const void *)solve_persistent_csr_kernel<true>, dim3(grid), dim3(SWEEP_THREADS), args, smem, s)
               : cudaLaunchCooperativeKernel((const void *)solve_persistent_csr_kernel<false>, dim3(grid), dim3(SWEEP_THREADS), args, smem, s);
}

} // namespace alll
