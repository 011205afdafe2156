// batch.cu -- many small independent instances, one CTA each, ALL rounds inside one launch
// (BASELINE config 5: 8,192 x 5-SAT n=10k; SURVEY.md section 8e "batched instances" and "portfolio").
//
// The whole solver state of an instance lives in shared memory: bit-packed assignment, 64-bit claim word per
// variable, violated list and per-entry state.  Only the literals stream from global memory (L2-resident after
// the first round: 120 KB per instance).  The round specification is exactly the one of the large-instance
// path (alll_device.cuh / oracle): same Philox streams, same (priority, id) greedy independent set, same
// Statistics semantics (SATInstance.h:261,291,317,363) -- an instance solved here ends with the same assignment
// and counters as alll_solve or the oracle would give for the same seed.
//
// Two kernels, one specification:
//  * batch_solve_small_kernel<K> (k = 3..8, the case BASELINE config 5 is): 256-thread CTAs with ~36 KB of shared memory,
//    so 5-6 instances are resident per SM and their latency chains overlap.  All k literal planes of a group of four
//    clauses are requested before the first lookup (one L2 round trip per sweep step instead of k); a violated clause
//    drops its literals, still in registers, into a shared-memory record, so the independent set and the resample never
//    touch global memory.  Claims are 16-bit indices into the violated list, compared through the 64-bit
//    (priority, id) keys of the entries (atomicCAS loop), cleared at the end of every Luby step.
//    A job whose violated set outgrows the BS_UCAP records appends itself to a retry list and leaves.
//  * batch_solve_kernel (any k, any |U|): 1024-thread CTAs, 64-bit claim word per variable, literals re-read from
//    global memory.  Runs the whole batch when the small kernel is not eligible, otherwise -- as a few persistent CTAs
//    behind the small kernel on the same stream -- the jobs of the retry list, from scratch (a job is a pure function of
//    its seed).
//
// Portfolio mode: every CTA works on instance 0 with its own seed; the first CTA that reaches an empty violated
// set claims the device-wide `winner` word (atomicCAS) and the others stop at their next round boundary.
#include <cstdlib>

#include <mutex>

#include "alll_device.cuh"

namespace alll {

constexpr uint32_t BATCH_THREADS = 1024;

struct BatchParams {
    const uint32_t *planes;        // [k][m_pad] literal planes of the whole batch
    uint64_t m_pad;
    const uint32_t *inst_off;      // [n_instances + 1] first slot of every instance (multiples of 4); slots [off, off + m_i)
    const uint32_t *inst_m;        // [n_instances] clauses per instance
    uint32_t n_instances;          // real instances stored
    uint32_t n_vars, n_words;      // per instance
    uint32_t k;
    uint32_t m_max;                // widest instance (sizes the shared-memory lists)
    const uint64_t *seeds;         // [n_jobs]
    uint64_t max_rounds;
    uint32_t *bits_out;            // [n_jobs][n_words]
    BatchJobStats *stats;          // [n_jobs]
    int portfolio;                 // 1: all jobs solve instance 0, first to finish wins
    int *winner;                   // portfolio: job id of the winner, -1 while open
    int job_base;                  // id written to *winner = job_base + job (distinct per GPU in a multi-GPU portfolio)
    int shared;                    // *winner is one word shared by several GPUs (peer-mapped): system-scope atomics
    uint32_t *retry;               // retry[0] = number of jobs the small kernel gave up on, retry[1..] = their ids (NULL: none)
    int from_retry;                // batch_solve_kernel: work through the retry list instead of job = blockIdx.x
};

extern __shared__ __align__(16) uint32_t b_smem[];

// ---- clause evaluation against the bit-packed assignment in shared memory (both kernels)
// sb = 32-bit shared-window address of the assignment words, held in a register that nvcc cannot see through (callers
// re-make it with an empty asm volatile after the barrier that ends a resample): the lookup is shift, mask, LDS, funnel
// shift, one LOP3 -- predicated on the clause having seen no true literal yet (a dead lane touches no bank), no branch.
// With `if`s around bits[v >> 5] nvcc emits BSSY/BRA/BSYNC per lookup and rebuilds the window base (S2UR SR_CgaCtaId, UMOV,
// ULEA) every time: 17 instead of 8 instructions per literal (profiles/r01_batch_small.md).
__device__ __forceinline__ uint32_t bs_lds32(uint32_t byte_addr)
{
    uint32_t w;
    asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(byte_addr));
    return w;
}
__device__ __forceinline__ uint32_t lane_of(const uint4 &v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }
// a[q] bit 0: clause q of the group is still unsatisfied; one literal plane vector L
__device__ __forceinline__ void bs_eval4(const uint4 &L, uint32_t (&a)[4], uint32_t sb)
{
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t l = lane_of(L, q);
        const uint32_t w = a[q] ? bs_lds32(sb + ((l >> 6) << 2)) : 0u;
        a[q] &= ~(__funnelshift_r(w, 0u, l >> 1) ^ l);            // bit (v & 31) of the word, xor the negation flag
    }
}

// End of a job: portfolio winner claim, assignment and statistics out.  Called by every thread of the CTA.
__device__ __forceinline__ void batch_finish(const BatchParams &p, uint32_t job, int status, const uint32_t *bits,
                                             uint64_t n_iter, uint64_t n_res, uint64_t sum_mis)
{
    const uint32_t tid = threadIdx.x;
    __syncthreads();
    bool publish = true;
    if (p.portfolio) {
        __shared__ int s_won;
        if (tid == 0) {
            const int id = p.job_base + (int)job;
            s_won = (status == 0) ? ((p.shared ? atomicCAS_system(p.winner, -1, id) : atomicCAS(p.winner, -1, id)) == -1) : 0;
        }
        __syncthreads();
        publish = s_won != 0;
        if (status == 0 && !publish) status = BATCH_PREEMPTED; // finished, but somebody else was first
    }
    if (publish)
        for (uint32_t w = tid; w < p.n_words; w += blockDim.x) p.bits_out[(uint64_t)job * p.n_words + w] = bits[w];
    if (tid == 0) {
        BatchJobStats st;
        st.n_iterations = n_iter;
        st.n_resamples = n_res;
        st.sum_mis_size = sum_mis;
        st.status = status;
        st.reserved = 0;
        p.stats[job] = st;
    }
}

static __device__ __forceinline__ void batch_job_large(const BatchParams &p, const uint32_t job)
{
    const uint32_t inst = p.portfolio ? 0u : job;
    const uint32_t off = p.inst_off[inst], m = p.inst_m[inst];
    const uint64_t seed = p.seeds[job];
    const uint32_t tid = threadIdx.x;

    // shared-memory carve-up (all sizes are per instance)
    unsigned long long *claim = reinterpret_cast<unsigned long long *>(b_smem);                  // [n_vars]
    uint32_t *bits = b_smem + 2ull * p.n_vars;                                                    // [n_words]
    uint32_t *ulist = bits + ((p.n_words + 3u) & ~3u);                                            // [m_max]
    uint8_t *state = reinterpret_cast<uint8_t *>(ulist + p.m_max);                                // [m_max]
    __shared__ unsigned int s_nu, s_live, s_ns, s_stop;
    __shared__ unsigned long long s_res;

    for (uint32_t v = tid; v < p.n_vars; v += BATCH_THREADS) claim[v] = CLAIM_FREE;
    // initial assignment: Philox INIT stream, one call per 128 variables (same as randomize_kernel)
    for (uint32_t g = tid; g * 4 < p.n_words; g += BATCH_THREADS) {
        const Philox o = philox4x32_10(g, 0u, STREAM_INIT, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t out[4] = {o.x, o.y, o.z, o.w};
        for (uint32_t i = 0; i < 4 && g * 4 + i < p.n_words; i++) {
            const uint32_t w = g * 4 + i, base = w * 32;
            uint32_t word = out[i];
            if (p.n_vars - base < 32) word &= (1u << (p.n_vars - base)) - 1u;
            bits[w] = word;
        }
    }
    uint64_t n_iter = 0, n_res = 0, sum_mis = 0;
    int status = 1;                       // ALLL_MAX_ROUNDS until proven otherwise
    const uint64_t max_rounds = p.max_rounds ? p.max_rounds : 1;
    const uint32_t sbits = (uint32_t)__cvta_generic_to_shared(bits);

    for (uint64_t round = 0; round < max_rounds; round++) {
        if (tid == 0) { s_nu = 0; s_ns = 0; s_res = 0; s_stop = p.portfolio && *(volatile int *)p.winner >= 0; }
        __syncthreads();
        if (s_stop) { status = BATCH_PREEMPTED; break; }        // somebody else finished: give up (decided by one thread: uniform)

        // ---- K1+K2: sweep this instance's clauses, 4 per thread and plane, two planes in flight; stops at the first pair of
        // planes after which none of the four clauses is still unsatisfied
        uint32_t sb = sbits;
        asm volatile("" : "+r"(sb));
        for (uint32_t c0 = tid * 4; c0 < m; c0 += BATCH_THREADS * 4) {
            const uint32_t *src = p.planes + off + c0;
            uint32_t a[4];
#pragma unroll
            for (int q = 0; q < 4; q++) a[q] = c0 + q < m ? 1u : 0u;
            for (uint32_t j = 0; j < p.k && (a[0] | a[1] | a[2] | a[3]); j += 2) {
                const bool two = j + 1 < p.k;
                const uint4 L0 = ld_stream_v4(src + (uint64_t)j * p.m_pad);
                const uint4 L1 = two ? ld_stream_v4(src + (uint64_t)(j + 1) * p.m_pad) : make_uint4(0u, 0u, 0u, 0u);
                bs_eval4(L0, a, sb);
                if (two) bs_eval4(L1, a, sb);
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (a[q]) ulist[atomicAdd(&s_nu, 1u)] = c0 + q;
        }
        __syncthreads();
        const uint32_t n_u = s_nu;
        n_iter++;                                              // SATInstance.h:261: the terminal sweep counts
        if (n_u == 0) { status = 0; break; }                   // SATInstance.h:285-287

        // ---- K3: fixed-priority Luby on shared-memory claims (ids are the instance-local clause indices)
        for (uint32_t i = tid; i < n_u; i += BATCH_THREADS) state[i] = 0;
        uint32_t step = 0;
        for (;;) {
            if (tid == 0) s_live = 0;
            __syncthreads();
            for (uint32_t i = tid; i < n_u; i += BATCH_THREADS) {
                if (state[i]) continue;
                const uint32_t c = ulist[i];
                bool taken = false;
                for (uint32_t j = 0; j < p.k; j++)
                    taken |= claim[p.planes[(uint64_t)j * p.m_pad + off + c] >> 1] == CLAIM_TAKEN;
                if (taken) { state[i] = 2; continue; }
                const unsigned long long key = claim_key(step, clause_priority(seed, (uint32_t)round, c), c);
                for (uint32_t j = 0; j < p.k; j++) atomicMin(&claim[p.planes[(uint64_t)j * p.m_pad + off + c] >> 1], key);
                atomicAdd(&s_live, 1u);
            }
            __syncthreads();
            if (s_live == 0) break;
            for (uint32_t i = tid; i < n_u; i += BATCH_THREADS) {
                if (state[i]) continue;
                const uint32_t c = ulist[i];
                const unsigned long long key = claim_key(step, clause_priority(seed, (uint32_t)round, c), c);
                bool win = true;
                for (uint32_t j = 0; j < p.k; j++) win &= claim[p.planes[(uint64_t)j * p.m_pad + off + c] >> 1] == key;
                if (!win) continue;
                state[i] = 1;
                atomicAdd(&s_ns, 1u);
            }
            __syncthreads();
            // winners mark their variables only now: a loser that shares a variable read its claim above
            for (uint32_t i = tid; i < n_u; i += BATCH_THREADS)
                if (state[i] == 1) {
                    const uint32_t c = ulist[i];
                    for (uint32_t j = 0; j < p.k; j++) claim[p.planes[(uint64_t)j * p.m_pad + off + c] >> 1] = CLAIM_TAKEN;
                }
            step++;
            if (step % TAGS == 0) {       // tag wrap: clear stale claims of the survivors (see mis.cu)
                __syncthreads();
                for (uint32_t i = tid; i < n_u; i += BATCH_THREADS)
                    if (!state[i]) {
                        const uint32_t c = ulist[i];
                        for (uint32_t j = 0; j < p.k; j++) {
                            const uint32_t v = p.planes[(uint64_t)j * p.m_pad + off + c] >> 1;
                            if (claim[v] != CLAIM_TAKEN) claim[v] = CLAIM_FREE;
                        }
                    }
            }
        }
        // ---- K4 + claim reset (two passes: every reset must land before any later round reads claims)
        for (uint32_t i = tid; i < n_u; i += BATCH_THREADS) {
            const uint32_t c = ulist[i];
            for (uint32_t j = 0; j < p.k; j++) {
                const uint32_t v = p.planes[(uint64_t)j * p.m_pad + off + c] >> 1;
                claim[v] = CLAIM_FREE;
                if (state[i] == 1) {
                    const uint32_t mask = 1u << (v & 31u);
                    if (random_bit(seed, STREAM_RESAMPLE, (uint32_t)round, v)) atomicOr(&bits[v >> 5], mask);
                    else atomicAnd(&bits[v >> 5], ~mask);
                }
            }
            if (state[i] == 1) atomicAdd(&s_res, (unsigned long long)p.k);     // SATInstance.h:363
        }
        __syncthreads();
        sum_mis += s_ns;                                       // SATInstance.h:291
        n_res += s_res;
        __syncthreads();
    }

    batch_finish(p, job, status, bits, n_iter, n_res, sum_mis);
    __syncthreads();                                           // (retry loop: the next job re-initialises the shared state)
}

__global__ void __launch_bounds__(BATCH_THREADS) batch_solve_kernel(const BatchParams p)
{
    if (!p.from_retry) { batch_job_large(p, blockIdx.x); return; }
    const uint32_t n = *(volatile uint32_t *)p.retry;          // written by the small kernel, complete before this launch starts
    for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) batch_job_large(p, p.retry[1 + r]);
}

// ---- the small kernel ---------------------------------------------------------------------------------------------
constexpr uint32_t BS_THREADS = 256;
constexpr uint32_t BS_UCAP = 512;                 // violated-clause records per job (cfg5: |U| ~ 190 in round 0)
constexpr unsigned short C16_FREE = 0xFFFFu, C16_TAKEN = 0xFFFEu;
static_assert(BS_UCAP < C16_TAKEN, "claims are 16-bit indices into the violated list");

// shared memory: bits | keys[UCAP] (64-bit) | literal records [K][UCAP] | state[UCAP] (bytes) | claims[n_vars] (16-bit)
__host__ __device__ constexpr uint32_t bs_align4(uint32_t x) { return (x + 3u) & ~3u; }
size_t batch_small_smem_bytes(uint32_t n_vars, uint32_t n_words, uint32_t k)
{
    return (size_t)bs_align4(n_words) * 4 + (size_t)BS_UCAP * 8 + (size_t)k * BS_UCAP * 4 + BS_UCAP + (size_t)bs_align4((n_vars + 1) / 2) * 4;
}
// The small kernel takes the batch when this holds.
static bool batch_small_eligible(uint32_t n_vars, uint32_t n_words, uint32_t k)
{
    return k >= 3 && k <= 8 && batch_small_smem_bytes(n_vars, n_words, k) <= 100u * 1024u;      // >= 2 CTAs per SM
}


// 5 CTAs per SM (47 registers at K = 5): measured best on B200 -- 4 per SM with 64 registers 0.94 ms, 6 per SM with 40
// registers 1.04 ms (the literals of 6 x 148 resident jobs no longer stay in L2), two groups of four clauses in flight per
// thread 0.92 ms at 4 per SM; this shape 0.84 ms (profiles/r01_batch_small.md).
template <uint32_t K>
__global__ void __launch_bounds__(BS_THREADS, 5) batch_solve_small_kernel(const BatchParams p)
{
    const uint32_t job = blockIdx.x;
    const uint32_t inst = p.portfolio ? 0u : job;
    const uint32_t off = p.inst_off[inst], m = p.inst_m[inst];
    const uint64_t seed = p.seeds[job];
    const uint32_t tid = threadIdx.x;

    uint32_t *bits = b_smem;                                                                      // [n_words]
    unsigned long long *key = reinterpret_cast<unsigned long long *>(bits + bs_align4(p.n_words)); // [UCAP] (priority << 32) | id
    uint32_t *rec = reinterpret_cast<uint32_t *>(key + BS_UCAP);                                  // [K][UCAP] literals of the violated clauses
    uint8_t *state = reinterpret_cast<uint8_t *>(rec + K * BS_UCAP);                              // [UCAP] 0 undecided, 1 in S, 2 dropped
    unsigned short *claim = reinterpret_cast<unsigned short *>(state + BS_UCAP);                  // [n_vars] index of the best claimant
    __shared__ unsigned int s_nu, s_ns, s_stop;

    {
        uint32_t *c32 = reinterpret_cast<uint32_t *>(claim);
        for (uint32_t w = tid; w < (p.n_vars + 1) / 2; w += BS_THREADS) c32[w] = 0xFFFFFFFFu;     // C16_FREE twice
    }
    for (uint32_t g = tid; g * 4 < p.n_words; g += BS_THREADS) {                                  // as randomize_kernel
        const Philox o = philox4x32_10(g, 0u, STREAM_INIT, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint32_t out[4] = {o.x, o.y, o.z, o.w};
        for (uint32_t i = 0; i < 4 && g * 4 + i < p.n_words; i++) {
            const uint32_t w = g * 4 + i, base = w * 32;
            uint32_t word = out[i];
            if (p.n_vars - base < 32) word &= (1u << (p.n_vars - base)) - 1u;
            bits[w] = word;
        }
    }
    uint64_t n_iter = 0, n_res = 0, sum_mis = 0;
    int status = 1;                       // ALLL_MAX_ROUNDS until proven otherwise
    const uint64_t max_rounds = p.max_rounds ? p.max_rounds : 1;
    const uint32_t *lit0 = p.planes + off;
    const uint32_t sbits = (uint32_t)__cvta_generic_to_shared(bits);

    for (uint64_t round = 0; round < max_rounds; round++) {
        if (tid == 0) { s_nu = 0; s_ns = 0; s_stop = p.portfolio && *(volatile int *)p.winner >= 0; }
        __syncthreads();
        if (s_stop) { status = BATCH_PREEMPTED; break; }
        uint32_t sb = sbits;
        asm volatile("" : "+r"(sb));      // opaque register copy of the window address: nvcc otherwise rebuilds it (S2R, MOV, LEA) per
                                          // lookup; re-made after the barrier so that no lookup is hoisted above the resample of the last round

        // ---- K1+K2: four clauses per thread and step; all K plane loads of the step are in flight together
        auto eval_group = [&](const uint4 (&L)[K], uint32_t c0) {
            uint32_t a[4];                                     // bit 0: clause q has seen no true literal yet
#pragma unroll
            for (int q = 0; q < 4; q++) a[q] = c0 + q < m ? 1u : 0u;
#pragma unroll
            for (uint32_t j = 0; j < K; j++) bs_eval4(L[j], a, sb);
            if (a[0] | a[1] | a[2] | a[3]) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (a[q]) {
                        const uint32_t idx = atomicAdd(&s_nu, 1u);
                        if (idx < BS_UCAP) {
                            key[idx] = c0 + q;
#pragma unroll
                            for (uint32_t j = 0; j < K; j++) rec[j * BS_UCAP + idx] = lane_of(L[j], q);
                        }
                    }
            }
        };
        for (uint32_t c0 = tid * 4; c0 < m; c0 += BS_THREADS * 4) {
            uint4 L[K];
#pragma unroll
            for (uint32_t j = 0; j < K; j++) L[j] = ld_stream_v4(lit0 + (uint64_t)j * p.m_pad + c0);
            eval_group(L, c0);
        }
        __syncthreads();
        const uint32_t n_u = s_nu;
        if (n_u > BS_UCAP) {                                  // more violated clauses than records: the large kernel redoes this job
            if (tid == 0) p.retry[1 + atomicAdd(p.retry, 1u)] = job;
            return;
        }
        n_iter++;                                              // SATInstance.h:261: the terminal sweep counts
        if (n_u == 0) { status = 0; break; }                   // SATInstance.h:285-287

        // ---- K3: fixed-priority Luby steps; S = greedy independent set in ascending (priority, id) order
        for (uint32_t i = tid; i < n_u; i += BS_THREADS) {
            const uint32_t c = (uint32_t)key[i];
            key[i] = ((unsigned long long)clause_priority(seed, (uint32_t)round, c) << 32) | c;
            state[i] = 0;
        }
        __syncthreads();
        for (;;) {
            int claimed = 0;
            for (uint32_t i = tid; i < n_u; i += BS_THREADS) {
                if (state[i]) continue;
                uint32_t var[K];
                bool taken = false;
#pragma unroll
                for (uint32_t j = 0; j < K; j++) { var[j] = rec[j * BS_UCAP + i] >> 1; taken |= claim[var[j]] == C16_TAKEN; }
                if (taken) { state[i] = 2; continue; }
                const unsigned long long mine = key[i];
#pragma unroll
                for (uint32_t j = 0; j < K; j++) {
                    unsigned short cur = *(volatile unsigned short *)&claim[var[j]];
                    for (;;) {
                        if (cur != C16_FREE && key[cur] <= mine) break;          // a better claim (or my own) stands
                        const unsigned short old = atomicCAS(&claim[var[j]], cur, (unsigned short)i);
                        if (old == cur) break;
                        cur = old;
                    }
                }
                claimed = 1;
            }
            if (!__syncthreads_or(claimed)) break;                               // nobody undecided
            for (uint32_t i = tid; i < n_u; i += BS_THREADS) {
                if (state[i]) continue;
                uint32_t var[K];
                bool win = true;
#pragma unroll
                for (uint32_t j = 0; j < K; j++) { var[j] = rec[j * BS_UCAP + i] >> 1; win &= claim[var[j]] == (unsigned short)i; }
                // a winner owns all its variables and marks them; a loser withdraws the claims it holds.  Either way only
                // slots holding i change, and everybody else compares slots against their own index: no order matters.
                if (win) {
                    state[i] = 1;
                    atomicAdd(&s_ns, 1u);
#pragma unroll
                    for (uint32_t j = 0; j < K; j++) claim[var[j]] = C16_TAKEN;
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < K; j++)
                        if (claim[var[j]] == (unsigned short)i) claim[var[j]] = C16_FREE;
                }
            }
            __syncthreads();
        }
        // ---- K4 + claim reset, one (clause, literal) pair per thread and pass
        for (uint32_t t = tid; t < n_u * K; t += BS_THREADS) {
            const uint32_t i = t / K, j = t - i * K;
            const uint32_t v = rec[j * BS_UCAP + i] >> 1;
            claim[v] = C16_FREE;
            if (state[i] == 1) {
                const uint32_t mask = 1u << (v & 31u);
                if (random_bit(seed, STREAM_RESAMPLE, (uint32_t)round, v)) atomicOr(&bits[v >> 5], mask);
                else atomicAnd(&bits[v >> 5], ~mask);
            }
        }
        __syncthreads();
        sum_mis += s_ns;                                       // SATInstance.h:291
        n_res += (uint64_t)s_ns * K;                           // SATInstance.h:363
        __syncthreads();
    }
    batch_finish(p, job, status, bits, n_iter, n_res, sum_mis);
}

// row-major [total][k] -> planes with every instance's first slot aligned to 4
__global__ void __launch_bounds__(256) batch_transpose_kernel(const uint32_t *__restrict__ lit, const uint64_t *__restrict__ src_off,
                                                               const uint32_t *__restrict__ inst_off, uint32_t n_instances,
                                                               uint32_t k, uint32_t n_vars, uint32_t *__restrict__ planes,
                                                               uint64_t m_pad, uint32_t *err)
{
    const uint32_t inst = blockIdx.y;
    const uint64_t lo = src_off[inst], m = src_off[inst + 1] - lo;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < m; c += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t bad = 0;
        for (uint32_t j = 0; j < k; j++) {
            const uint32_t l = lit[(lo + c) * k + j];
            bad |= (l >> 1) >= n_vars;
            planes[(uint64_t)j * m_pad + inst_off[inst] + c] = l;
        }
        if (bad) atomicOr(err, 1u);
    }
}

// bit-packed [n_jobs][n_words] -> bytes [n_jobs][n_vars]
__global__ void __launch_bounds__(256) batch_unpack_kernel(const uint32_t *__restrict__ bits, uint32_t n_words, uint32_t n_vars,
                                                            uint64_t total, uint8_t *__restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t job = i / n_vars;
    const uint32_t v = (uint32_t)(i % n_vars);
    out[i] = (bits[job * n_words + (v >> 5)] >> (v & 31u)) & 1u;
}

size_t batch_smem_bytes(uint32_t n_vars, uint32_t n_words, uint32_t m_max)
{
    return (size_t)n_vars * 8 + (size_t)((n_words + 3u) & ~3u) * 4 + (size_t)m_max * 4 + (((size_t)m_max + 15) & ~(size_t)15);
}

template <uint32_t K>
static cudaError_t launch_small(const BatchParams &p, size_t smem, uint32_t n_jobs, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(batch_solve_small_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    batch_solve_small_kernel<K><<<n_jobs, BS_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

// retry: n_jobs + 1 words of device scratch (NULL: large kernel only).  *n_launches: kernels enqueued.
cudaError_t launch_batch_solve(const uint32_t *planes, uint64_t m_pad, const uint32_t *inst_off, const uint32_t *inst_m,
                               uint32_t n_instances, uint32_t n_vars, uint32_t n_words, uint32_t k, uint32_t m_max,
                               const uint64_t *seeds, uint64_t max_rounds, uint32_t *bits_out, BatchJobStats *stats,
                               int portfolio, int *winner, int job_base, int shared, uint32_t n_jobs, uint32_t *retry,
                               int *n_launches, cudaStream_t s)
{
    static const bool large_only = getenv("ALLL_BATCH_LARGE_ONLY") != nullptr;        // measurement / test knob
    // alll_multi_batch_solve enqueues the device slots from one host thread each: the function attributes and the
    // (first: module-loading) launches of the same kernels are serialised here -- microseconds of enqueue, not the solve
    static std::mutex enqueue_mu;
    std::lock_guard<std::mutex> lock(enqueue_mu);
    (void)cudaGetLastError();                       // a stale non-sticky error of this thread must not be taken for the launch's
    const size_t smem = batch_smem_bytes(n_vars, n_words, m_max);
    cudaError_t e = cudaFuncSetAttribute(batch_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    BatchParams p{planes, m_pad, inst_off, inst_m, n_instances, n_vars, n_words, k, m_max, seeds, max_rounds, bits_out, stats,
                  portfolio, winner, job_base, shared, retry, 0};
    const size_t small = batch_small_smem_bytes(n_vars, n_words, k);
    const bool use_small = retry && !large_only && batch_small_eligible(n_vars, n_words, k);
    if (n_launches) *n_launches = use_small ? 2 : 1;
    if (!use_small) {
        batch_solve_kernel<<<n_jobs, BATCH_THREADS, smem, s>>>(p);
        return cudaGetLastError();
    }
    e = cudaMemsetAsync(retry, 0, 4, s);
    if (e != cudaSuccess) return e;
    switch (k) {
    case 3: e = launch_small<3>(p, small, n_jobs, s); break;
    case 4: e = launch_small<4>(p, small, n_jobs, s); break;
    case 5: e = launch_small<5>(p, small, n_jobs, s); break;
    case 6: e = launch_small<6>(p, small, n_jobs, s); break;
    case 7: e = launch_small<7>(p, small, n_jobs, s); break;
    default: e = launch_small<8>(p, small, n_jobs, s); break;
    }
    if (e != cudaSuccess) return e;
    p.from_retry = 1;                                           // the jobs that outgrew the records, if any
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t grid = n_jobs < (uint32_t)sms ? n_jobs : (uint32_t)sms;
    batch_solve_kernel<<<grid, BATCH_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_batch_transpose(const uint32_t *lit, const uint64_t *src_off, const uint32_t *inst_off, uint32_t n_instances,
                                   uint32_t k, uint32_t n_vars, uint32_t *planes, uint64_t m_pad, uint32_t *err, cudaStream_t s)
{
    if (n_instances == 0) return cudaSuccess;
    batch_transpose_kernel<<<dim3(8, n_instances), 256, 0, s>>>(lit, src_off, inst_off, n_instances, k, n_vars, planes, m_pad, err);
    return cudaGetLastError();
}

cudaError_t launch_batch_unpack(const uint32_t *bits, uint32_t n_words, uint32_t n_vars, uint64_t total, uint8_t *out, cudaStream_t s)
{
    if (total == 0) return cudaSuccess;
    batch_unpack_kernel<<<(uint32_t)((total + 255) / 256), 256, 0, s>>>(bits, n_words, n_vars, total, out);
    return cudaGetLastError();
}

} // namespace alll
