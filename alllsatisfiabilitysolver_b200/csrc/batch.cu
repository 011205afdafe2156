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
// Portfolio mode: every CTA works on instance 0 with its own seed; the first CTA that reaches an empty violated
// set claims the device-wide `winner` word (atomicCAS) and the others stop at their next round boundary.
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
};

extern __shared__ __align__(16) uint32_t b_smem[];

__global__ void __launch_bounds__(BATCH_THREADS) batch_solve_kernel(const BatchParams p)
{
    const uint32_t job = blockIdx.x;
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

    for (uint64_t round = 0; round < max_rounds; round++) {
        if (tid == 0) { s_nu = 0; s_ns = 0; s_res = 0; s_stop = p.portfolio && *(volatile int *)p.winner >= 0; }
        __syncthreads();
        if (s_stop) { status = BATCH_PREEMPTED; break; }        // somebody else finished: give up (decided by one thread: uniform)

        // ---- K1+K2: sweep this instance's clauses, 4 per thread and plane
        for (uint32_t c0 = tid * 4; c0 < m; c0 += BATCH_THREADS * 4) {
            uint32_t alive = (c0 + 0 < m ? 1u : 0u) | (c0 + 1 < m ? 2u : 0u) | (c0 + 2 < m ? 4u : 0u) | (c0 + 3 < m ? 8u : 0u);
            for (uint32_t j = 0; j < p.k && alive; j++) {
                const uint4 L = *reinterpret_cast<const uint4 *>(p.planes + (uint64_t)j * p.m_pad + off + c0);
                const uint32_t l[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if ((alive >> q) & 1u) {
                        const uint32_t v = l[q] >> 1;
                        if (((bits[v >> 5] >> (v & 31u)) ^ l[q]) & 1u) alive &= ~(1u << q);
                    }
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if ((alive >> q) & 1u) ulist[atomicAdd(&s_nu, 1u)] = c0 + q;
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
        for (uint32_t w = tid; w < p.n_words; w += BATCH_THREADS) p.bits_out[(uint64_t)job * p.n_words + w] = bits[w];
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

cudaError_t launch_batch_solve(const uint32_t *planes, uint64_t m_pad, const uint32_t *inst_off, const uint32_t *inst_m,
                               uint32_t n_instances, uint32_t n_vars, uint32_t n_words, uint32_t k, uint32_t m_max,
                               const uint64_t *seeds, uint64_t max_rounds, uint32_t *bits_out, BatchJobStats *stats,
                               int portfolio, int *winner, int job_base, int shared, uint32_t n_jobs, cudaStream_t s)
{
    const size_t smem = batch_smem_bytes(n_vars, n_words, m_max);
    cudaError_t e = cudaFuncSetAttribute(batch_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    BatchParams p{planes, m_pad, inst_off, inst_m, n_instances, n_vars, n_words, k, m_max, seeds, max_rounds, bits_out, stats,
                  portfolio, winner, job_base, shared};
    batch_solve_kernel<<<n_jobs, BATCH_THREADS, smem, s>>>(p);
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
