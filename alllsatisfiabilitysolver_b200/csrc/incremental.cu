// incremental.cu -- incremental re-evaluation (no reference counterpart: the reference
// re-sweeps all m clauses every round, SATInstance.h:273-280).
//
// After a resample round only clauses containing a resampled variable can change status, and because the
// independent set S is MAXIMAL every clause of the old violated set contains one too.  So the next violated set is
//     U' = { c in occ(vars(S)) : c violated },
// where occ(v) lists the clauses containing v.  The set is the same as the full sweep's, hence (the MIS being a
// function of the set only) the whole trajectory, the statistics and the final assignment are bit-identical with
// or without this mode.  It pays once |S| is small: a round then touches k*d*|S| clauses instead of m.
//
// Extra HBM state (built once per upload, only when the mode is enabled):
//   occ_off[n+1], occ[L]   variable -> clause slots (CSR, any polarity)
//   rows[m_pad][stride]    literals row-major by slot (stride = k rounded up to 4): one 16/32-byte fetch per clause
//   visited[m_pad/32]      per-round "already evaluated" bits (cleared after every incremental round)
#include "incr_body.cuh"

namespace alll {

// ---- build ------------------------------------------------------------------------------------------------

// one thread per slot: count occurrences, write the row-major copy (padding literals repeat literal 0: neutral)
__global__ void __launch_bounds__(256) incr_count_rows_kernel(const uint32_t *__restrict__ planes, uint64_t m_pad, uint32_t k,
                                                               uint32_t stride, const BucketSeg *__restrict__ segs,
                                                               uint32_t n_buckets, uint32_t *__restrict__ occ_cnt,
                                                               uint32_t *__restrict__ rows, const uint8_t *__restrict__ width)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m_pad) return;
    // valid slot?  (bucket segments are padded to tile boundaries)
    uint32_t b = 0;
    b = find_segment(segs, n_buckets, (uint32_t)(p / TILE));
    const bool valid = p < segs[b].slot_end;
    uint32_t first = 0;
    for (uint32_t j = 0; j < stride; j++) {
        uint32_t l = first;
        if (j < k) {
            l = planes[(uint64_t)j * m_pad + p];
            if (j == 0) first = l;
            if (valid && (!width || j < width[p])) atomicAdd(&occ_cnt[l >> 1], 1u);      // padding literals are not occurrences
        }
        rows[p * stride + j] = valid ? l : 0u;
    }
}

// exclusive scan of n counters, three passes (block sums -> scan of block sums -> add), 1024 items per block
constexpr uint32_t SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK) scan_block_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                uint32_t *__restrict__ out, uint32_t *__restrict__ block_sums)
{
    __shared__ uint32_t warp_sums[32];
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    uint32_t x = v;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u);
    if (i < n) out[i] = incl - v;                                   // exclusive within the block
    if (threadIdx.x == SCAN_BLOCK - 1) block_sums[blockIdx.x] = incl;
}

// single block: exclusive scan of the block sums in place (n_blocks <= a few 10^4: sequential chunks of 1024)
__global__ void __launch_bounds__(SCAN_BLOCK) scan_sums_kernel(uint32_t *sums, uint32_t n_blocks, uint32_t *total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? sums[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= (uint32_t)o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= (uint32_t)o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (i < n_blocks) sums[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_add_kernel(uint32_t *__restrict__ out, uint64_t n,
                                                              const uint32_t *__restrict__ block_sums, uint32_t *__restrict__ cursor)
{
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) {
        const uint32_t v = out[i] + block_sums[blockIdx.x];
        out[i] = v;
        cursor[i] = v;                                              // fill cursors start at the list heads
    }
}

__global__ void __launch_bounds__(256) incr_fill_kernel(const uint32_t *__restrict__ planes, uint64_t m_pad, uint32_t k,
                                                         const BucketSeg *__restrict__ segs, uint32_t n_buckets,
                                                         uint32_t *__restrict__ cursor, uint32_t *__restrict__ occ,
                                                         const uint8_t *__restrict__ width)
{
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= m_pad) return;
    uint32_t b = 0;
    b = find_segment(segs, n_buckets, (uint32_t)(p / TILE));
    if (p >= segs[b].slot_end) return;
    const uint32_t kw = width ? width[p] : k;
    for (uint32_t j = 0; j < kw; j++) {
        const uint32_t v = planes[(uint64_t)j * m_pad + p] >> 1;
        occ[atomicAdd(&cursor[v], 1u)] = (uint32_t)p;
    }
}

// ---- per round ----------------------------------------------------------------------------------------------

// One warp per (clause of S, literal): see incr_eval_body.
__global__ void __launch_bounds__(256) incr_eval_kernel(const IncrParams p)
{
    if (__ldcg(&p.ctr->done) || !__ldcg(&p.ctr->incr_next)) return;   // this round is a full sweep (or none at all)
    incr_eval_body(p, __ldcg(&p.ctr->last_n_s), &p.ctr->n_viol, IncrP2P{});
}

// ---- launchers ----------------------------------------------------------------------------------------------
static inline uint32_t blocks_for(uint64_t n, uint32_t t) { return (uint32_t)((n + t - 1) / t); }

cudaError_t launch_incr_build(const uint32_t *planes, uint64_t m_pad, uint32_t k, uint32_t stride, const BucketSeg *segs,
                              uint32_t n_buckets, uint64_t n_vars, uint32_t *occ_off /*[n_vars+1]*/, uint32_t *cursor /*[n_vars]*/,
                              uint32_t *block_sums, uint32_t *rows, uint32_t *occ, uint32_t *d_total, const uint8_t *width, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(cursor, 0, n_vars * 4, s);       // cursor doubles as the occurrence counter
    if (e != cudaSuccess) return e;
    if (m_pad) incr_count_rows_kernel<<<blocks_for(m_pad, 256), 256, 0, s>>>(planes, m_pad, k, stride, segs, n_buckets, cursor, rows, width);
    const uint32_t nb = blocks_for(n_vars, SCAN_BLOCK);
    scan_block_kernel<<<nb, SCAN_BLOCK, 0, s>>>(cursor, n_vars, occ_off, block_sums);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, s>>>(block_sums, nb, d_total);
    scan_add_kernel<<<nb, SCAN_BLOCK, 0, s>>>(occ_off, n_vars, block_sums, cursor);
    e = cudaMemcpyAsync(occ_off + n_vars, d_total, 4, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
    if (m_pad) incr_fill_kernel<<<blocks_for(m_pad, 256), 256, 0, s>>>(planes, m_pad, k, segs, n_buckets, cursor, occ, width);
    return cudaGetLastError();
}

cudaError_t launch_incr_eval(const uint32_t *s_slots, const uint32_t *rows, uint32_t stride, uint32_t k,
                             const uint32_t *occ_off, const uint32_t *occ, uint32_t *visited, uint64_t visited_words,
                             const uint32_t *bits, uint32_t *viol, Counters *ctr, uint32_t grid, cudaStream_t s)
{
    IncrParams p{s_slots, rows, stride, k, occ_off, occ, visited, bits, viol, ctr};
    incr_eval_kernel<<<grid, 256, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the bitmap must be clean for the next incremental round; clearing it unconditionally costs ~1 us per 5 MB
    return cudaMemsetAsync(visited, 0, visited_words * 4, s);
}

} // namespace alll
