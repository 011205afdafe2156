// layout.cu -- one-time (per upload) kernels that turn the caller's flattened clauses into the HBM layout
// the sweep streams, plus assignment pack/unpack/randomise.
//
// Replaces the reference's heap object graph (Clause.h:17-28: vector<tV>* per clause, >= 3 heap blocks each)
// and bool vars[n] (VariablesArray.h:18-35) by: k literal-major uint32 planes [k][m_pad] and a bit-packed
// assignment.  When the assignment does not fit the sweep's shared-memory budget, clauses are stably
// counting-sorted into variable-range buckets (the bucket holding most of their variables) and each
// clause's literals are reordered bucket-resident first; orig_id[] maps slots back to caller clause ids.
#include "alll_device.cuh"

namespace alll {

constexpr uint32_t LAYOUT_THREADS = 1024;   // clauses per CTA in the bucketing passes (one per thread)

// error flags (bit-ored into *err)
constexpr uint32_t ERR_LITERAL_RANGE = 1u;

// ---- no bucketing: transpose row-major [m][k] into planes, validating literals -------------------
// clauses [c0, c1): the host-buffer upload runs this chunk by chunk behind the H2D copy of each chunk
__global__ void __launch_bounds__(256) transpose_kernel(const uint32_t *__restrict__ lit, uint64_t c0, uint64_t c1, uint32_t k,
                                                         uint64_t n_vars, uint32_t *__restrict__ planes,
                                                         uint64_t m_pad, uint32_t *err)
{
    const uint64_t c = c0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    uint32_t bad = 0;
    for (uint32_t j = 0; j < k; j++) {
        const uint32_t l = lit[c * k + j];
        bad |= ((uint64_t)(l >> 1) >= n_vars);
        planes[(uint64_t)j * m_pad + c] = l;
    }
    if (bad) atomicOr(err, ERR_LITERAL_RANGE);
}

__global__ void __launch_bounds__(256) validate_csr_kernel(const uint32_t *__restrict__ lit, uint64_t n_lit,
                                                            uint64_t n_vars, uint32_t *err)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lit && (uint64_t)(lit[i] >> 1) >= n_vars) atomicOr(err, ERR_LITERAL_RANGE);
}

// ---- bucketing pass 1: bucket of every clause + per-CTA histogram --------------------------------
// bucket(c) = the variable-range bucket that holds most of c's variables (ties: lowest bucket).
// This launch covers the logical CTAs [cta0, cta0 + gridDim.x) of the n_cta the whole pass has (chunked behind the H2D copy).
__global__ void __launch_bounds__(LAYOUT_THREADS) bucket_count_kernel(const uint32_t *__restrict__ lit, uint64_t m,
                                                                       uint32_t k, uint64_t n_vars,
                                                                       uint32_t bucket_vars, uint32_t n_buckets,
                                                                       uint8_t *__restrict__ bkt,
                                                                       uint32_t *__restrict__ cta_counts, uint32_t *err,
                                                                       uint32_t cta0, uint32_t n_cta)
{
    __shared__ uint32_t hist[MAX_BUCKETS];
    for (uint32_t i = threadIdx.x; i < n_buckets; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const uint32_t cta = cta0 + blockIdx.x;
    const uint64_t c = (uint64_t)cta * LAYOUT_THREADS + threadIdx.x;
    if (c < m) {
        uint32_t b_of[MAX_K];
        uint32_t bad = 0;
        for (uint32_t j = 0; j < k; j++) {
            const uint32_t v = lit[c * k + j] >> 1;
            bad |= ((uint64_t)v >= n_vars);
            b_of[j] = v / bucket_vars;          // may be >= n_buckets when n_buckets was clamped: never resident
        }
        if (bad) atomicOr(err, ERR_LITERAL_RANGE);
        uint32_t best = 0, best_cnt = 0;
        for (uint32_t j = 0; j < k; j++) {
            if (b_of[j] >= n_buckets) continue;
            uint32_t cnt = 0;
            for (uint32_t i = 0; i < k; i++) cnt += (b_of[i] == b_of[j]);
            if (cnt > best_cnt || (cnt == best_cnt && b_of[j] < best)) { best_cnt = cnt; best = b_of[j]; }
        }
        bkt[c] = (uint8_t)best;
        atomicAdd(&hist[best], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_buckets; i += blockDim.x)
        cta_counts[(uint64_t)i * n_cta + cta] = hist[i];
}

// ---- bucketing pass 2: stable scatter into planes, bucket-resident literals first -----------------
// cta_base[b * n_cta + cta] = first slot for this CTA's clauses of bucket b (exclusive scan done on the host).
__global__ void __launch_bounds__(LAYOUT_THREADS) bucket_scatter_kernel(const uint32_t *__restrict__ lit, uint64_t m,
                                                                         uint32_t k, uint32_t bucket_vars,
                                                                         uint32_t n_buckets,
                                                                         const uint8_t *__restrict__ bkt,
                                                                         const uint32_t *__restrict__ cta_base,
                                                                         uint32_t *__restrict__ planes, uint64_t m_pad,
                                                                         uint32_t *__restrict__ orig_id,
                                                                         uint32_t *__restrict__ min_resident, uint32_t resident_cap,
                                                                         const uint8_t *__restrict__ width_in,
                                                                         uint8_t *__restrict__ width_out, uint32_t cta0, uint32_t n_cta)
{
    __shared__ uint16_t warp_cnt[(LAYOUT_THREADS / 32) * MAX_BUCKETS];   // [warp][bucket], then exclusive over warps
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < (LAYOUT_THREADS / 32) * n_buckets; i += blockDim.x) warp_cnt[i] = 0;
    __syncthreads();
    const uint32_t cta = cta0 + blockIdx.x;              // logical CTA of the whole pass (this launch covers one upload chunk)
    const uint64_t c = (uint64_t)cta * LAYOUT_THREADS + threadIdx.x;
    const bool active = c < m;
    const uint32_t act = __ballot_sync(0xffffffffu, active);
    uint32_t b = 0, rank = 0;
    if (active) {
        b = bkt[c];
        const uint32_t peers = __match_any_sync(act, b);
        rank = __popc(peers & ((1u << lane) - 1u));
        if (rank == 0) warp_cnt[warp * n_buckets + b] = (uint16_t)__popc(peers);
    }
    __syncthreads();
    for (uint32_t bb = threadIdx.x; bb < n_buckets; bb += blockDim.x) {
        uint32_t run = 0;
        for (uint32_t w = 0; w < LAYOUT_THREADS / 32; w++) {
            const uint32_t t = warp_cnt[w * n_buckets + bb];
            warp_cnt[w * n_buckets + bb] = (uint16_t)run;
            run += t;
        }
    }
    __syncthreads();
    if (!active) return;
    const uint64_t dst = (uint64_t)cta_base[(uint64_t)b * n_cta + cta] + warp_cnt[warp * n_buckets + b] + rank;
    orig_id[dst] = (uint32_t)c;
    if (width_in) width_out[dst] = width_in[c];
    const uint32_t lo = b * bucket_vars;
    // the first (at most resident_cap) bucket-resident literals go to the leading planes, original order kept;
    // everything else follows (a resident literal beyond the cap is simply looked up through L2 like the rest).
    // Padded rows (width_in): only the TRUE literals [0, w) are reordered; the pad copies of literal 0 stay in planes
    // [w, k), because the independent set, the resample and the incremental row build read exactly the first w planes
    // of a slot and must see every variable of the clause there (never a pad copy in place of a true literal).
    const uint32_t w = width_in ? width_in[c] : k;
    uint32_t j_out = 0, placed = 0;
    uint32_t front_mask = 0;
    for (uint32_t j = 0; j < w && placed < resident_cap; j++) {
        const uint32_t l = lit[c * k + j];
        if ((l >> 1) - lo < bucket_vars) {
            planes[(uint64_t)(j_out++) * m_pad + dst] = l;
            front_mask |= 1u << j;
            placed++;
        }
    }
    for (uint32_t j = 0; j < k; j++)
        if (!((front_mask >> j) & 1u)) planes[(uint64_t)(j_out++) * m_pad + dst] = lit[c * k + j];
    // one atomicMin per warp
    const uint32_t wmin = __reduce_min_sync(act, placed);
    if (__popc(act & ((1u << lane) - 1u)) == 0) atomicMin(min_resident, wmin);
}

// ---- assignment ------------------------------------------------------------------------------
// bools (1 byte per variable, VariablesArray.h:21) -> packed words; one word per thread.
__global__ void __launch_bounds__(256) pack_bits_kernel(const uint8_t *__restrict__ bools, uint64_t n_vars,
                                                         uint32_t *__restrict__ bits, uint32_t n_words_alloc)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words_alloc) return;
    uint32_t word = 0;
    const uint64_t base = (uint64_t)w * 32;
    for (uint32_t i = 0; i < 32 && base + i < n_vars; i++) word |= (bools[base + i] != 0 ? 1u : 0u) << i;
    bits[w] = word;
}

__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint32_t *__restrict__ bits, uint64_t n_vars,
                                                           uint8_t *__restrict__ bools)
{
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n_vars) bools[v] = (bits[v >> 5] >> (v & 31u)) & 1u;
}

// bit v = Philox(ctr={v>>7, 0, INIT, 0}, key=seed) word (v>>5)&3 bit v&31; one Philox call per 4 words.
__global__ void __launch_bounds__(256) randomize_kernel(uint32_t *__restrict__ bits, uint64_t n_vars,
                                                         uint32_t n_words_alloc, uint64_t seed)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;       // group of 128 variables
    if ((uint64_t)g * 4 >= n_words_alloc) return;
    const Philox o = philox4x32_10(g, 0u, STREAM_INIT, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t out[4] = {o.x, o.y, o.z, o.w};
    for (uint32_t i = 0; i < 4; i++) {
        const uint32_t w = g * 4 + i;
        if (w >= n_words_alloc) break;
        const uint64_t base = (uint64_t)w * 32;
        uint32_t word = out[i];
        if (base >= n_vars) word = 0;
        else if (n_vars - base < 32) word &= (1u << (uint32_t)(n_vars - base)) - 1u;
        bits[w] = word;
    }
}

__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long *p, uint64_t n, unsigned long long value)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = value;
}

// ---- launchers ---------------------------------------------------------------------------------
static inline uint32_t blocks_for(uint64_t n, uint32_t threads) { return (uint32_t)((n + threads - 1) / threads); }

// clauses [c0, c1) of the instance
cudaError_t launch_transpose(const uint32_t *lit, uint64_t c0, uint64_t c1, uint32_t k, uint64_t n_vars, uint32_t *planes,
                             uint64_t m_pad, uint32_t *err, cudaStream_t s)
{
    if (c1 <= c0) return cudaSuccess;
    transpose_kernel<<<blocks_for(c1 - c0, 256), 256, 0, s>>>(lit, c0, c1, k, n_vars, planes, m_pad, err);
    return cudaGetLastError();
}

cudaError_t launch_validate_csr(const uint32_t *lit, uint64_t n_lit, uint64_t n_vars, uint32_t *err, cudaStream_t s)
{
    if (n_lit == 0) return cudaSuccess;
    validate_csr_kernel<<<blocks_for(n_lit, 256), 256, 0, s>>>(lit, n_lit, n_vars, err);
    return cudaGetLastError();
}

uint32_t bucket_pass_ctas(uint64_t m) { return blocks_for(m, LAYOUT_THREADS); }

// clauses [c0, c1) of the m the whole pass covers; c0 must be a multiple of bucket_pass_clauses_per_cta()
uint32_t bucket_pass_clauses_per_cta() { return LAYOUT_THREADS; }
cudaError_t launch_bucket_count(const uint32_t *lit, uint64_t m, uint64_t c0, uint64_t c1, uint32_t k, uint64_t n_vars,
                                uint32_t bucket_vars, uint32_t n_buckets, uint8_t *bkt, uint32_t *cta_counts, uint32_t *err,
                                cudaStream_t s)
{
    if (c1 <= c0) return cudaSuccess;
    bucket_count_kernel<<<blocks_for(c1 - c0, LAYOUT_THREADS), LAYOUT_THREADS, 0, s>>>(
        lit, m, k, n_vars, bucket_vars, n_buckets, bkt, cta_counts, err, (uint32_t)(c0 / LAYOUT_THREADS), bucket_pass_ctas(m));
    return cudaGetLastError();
}

// Same result as bucket_scatter_kernel for k <= 8 without per-clause widths, with the CTA's 1024 clauses staged in shared
// memory in destination order first: thread t of the write phase owns the t-th clause of that order, so consecutive threads
// write consecutive slots of a bucket segment (whole sectors) instead of each warp scattering 4-5 word fragments over all
// segments.  Dynamic shared memory: (k + 1) * 1024 words.
extern __shared__ uint32_t stage_smem[];
__global__ void __launch_bounds__(LAYOUT_THREADS) bucket_scatter_staged_kernel(const uint32_t *__restrict__ lit, uint64_t m,
                                                                                uint32_t k, uint32_t bucket_vars,
                                                                                uint32_t n_buckets,
                                                                                const uint8_t *__restrict__ bkt,
                                                                                const uint32_t *__restrict__ cta_base,
                                                                                uint32_t *__restrict__ planes, uint64_t m_pad,
                                                                                uint32_t *__restrict__ orig_id,
                                                                                uint32_t *__restrict__ min_resident, uint32_t resident_cap,
                                                                                uint32_t cta0, uint32_t n_cta,
                                                                                uint32_t *__restrict__ packed, uint4 *__restrict__ rows)
{
    __shared__ uint16_t warp_cnt[(LAYOUT_THREADS / 32) * MAX_BUCKETS];   // [warp][bucket], then exclusive over warps
    __shared__ uint32_t bucket_off[MAX_BUCKETS + 1];                     // CTA-local start of every bucket's run
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < (LAYOUT_THREADS / 32) * n_buckets; i += blockDim.x) warp_cnt[i] = 0;
    __syncthreads();
    const uint32_t cta = cta0 + blockIdx.x;              // logical CTA of the whole pass (this launch covers one upload chunk)
    const uint64_t c = (uint64_t)cta * LAYOUT_THREADS + threadIdx.x;
    const bool active = c < m;
    const uint32_t act = __ballot_sync(0xffffffffu, active);
    uint32_t b = 0, rank = 0;
    uint32_t l[8];
    if (active) {
        b = bkt[c];
        const uint32_t peers = __match_any_sync(act, b);
        rank = __popc(peers & ((1u << lane) - 1u));
        if (rank == 0) warp_cnt[warp * n_buckets + b] = (uint16_t)__popc(peers);
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) l[j] = j < k ? lit[c * k + j] : 0u;
    }
    __syncthreads();
    for (uint32_t bb = threadIdx.x; bb < n_buckets; bb += blockDim.x) {
        uint32_t run = 0;
        for (uint32_t w = 0; w < LAYOUT_THREADS / 32; w++) {
            const uint32_t t = warp_cnt[w * n_buckets + bb];
            warp_cnt[w * n_buckets + bb] = (uint16_t)run;
            run += t;
        }
        bucket_off[bb + 1] = run;                       // totals for now
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        bucket_off[0] = 0;
        for (uint32_t bb = 0; bb < n_buckets; bb++) { run += bucket_off[bb + 1]; bucket_off[bb + 1] = run; }
    }
    __syncthreads();
    uint32_t placed = 0;
    if (active) {
        const uint32_t loc = bucket_off[b] + warp_cnt[warp * n_buckets + b] + rank;
        const uint32_t lo = b * bucket_vars;
        // the first (at most resident_cap) bucket-resident literals go to the leading planes, original order kept
        uint32_t j_out = 0, front_mask = 0;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)
            if (j < k && placed < resident_cap && (l[j] >> 1) - lo < bucket_vars) {
                stage_smem[(j_out++) * LAYOUT_THREADS + loc] = l[j];
                front_mask |= 1u << j;
                placed++;
            }
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)
            if (j < k && !((front_mask >> j) & 1u)) stage_smem[(j_out++) * LAYOUT_THREADS + loc] = l[j];
        stage_smem[k * LAYOUT_THREADS + loc] = (uint32_t)c;
        const uint32_t wmin = __reduce_min_sync(act, placed);           // one atomicMin per warp
        if (__popc(act & ((1u << lane) - 1u)) == 0) atomicMin(min_resident, wmin);
    }
    __syncthreads();
    const uint32_t t = threadIdx.x;
    if (t >= bucket_off[n_buckets]) return;
    uint32_t bb = 0;
    while (bucket_off[bb + 1] <= t) ++bb;
    const uint64_t dst = (uint64_t)cta_base[(uint64_t)bb * n_cta + cta] + (t - bucket_off[bb]);
    for (uint32_t j = 0; j < k; j++) planes[(uint64_t)j * m_pad + dst] = stage_smem[j * LAYOUT_THREADS + t];
    orig_id[dst] = stage_smem[k * LAYOUT_THREADS + t];
    // The sweep's own copies of the clause, written while its literals are at hand instead of in passes of their own:
    // packed eager planes (caller guarantees two bucket-resident leading literals for every clause: k > n_buckets) ...
    if (packed != nullptr) {
        uint32_t l[EagerPack<2>::N], w[EagerPack<2>::WORDS];
#pragma unroll
        for (int j = 0; j < EagerPack<2>::N; j++) l[j] = stage_smem[j * LAYOUT_THREADS + t];
        l[0] -= 2u * bb * bucket_vars;
        l[1] -= 2u * bb * bucket_vars;
        EagerPack<2>::encode(l, w);
#pragma unroll
        for (int i = 0; i < EagerPack<2>::WORDS; i++) packed[(uint64_t)i * m_pad + dst] = w[i];
    }
    // ... and the row-major copy (32 bytes per clause, literals beyond k zero): its second half is the sweep's tail row, the
    // whole row is what the independent-set gather reads -- one sector per violated clause instead of one per literal plane
    if (rows != nullptr) {
        uint32_t l[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) l[j] = j < k ? stage_smem[j * LAYOUT_THREADS + t] : 0u;
        rows[2 * dst] = make_uint4(l[0], l[1], l[2], l[3]);
        rows[2 * dst + 1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
}

// ---- packed eager planes (alll_device.cuh: EagerPack) -- one pass over the finished planes ------------------------
template <int RB>
__global__ void __launch_bounds__(256) pack_eager_kernel(const uint32_t *__restrict__ planes, uint64_t m_pad,
                                                          const BucketSeg *__restrict__ segs, uint32_t n_buckets /* segments */,
                                                          uint32_t bucket_vars, uint32_t *__restrict__ packed)
{
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= m_pad) return;
    const uint32_t tile = (uint32_t)(slot / TILE);
    const uint32_t b = find_segment(segs, n_buckets, tile);
    uint32_t w[EagerPack<RB>::WORDS] = {0u, 0u, 0u, 0u};
    if (slot < segs[b].slot_end) {                       // (padding slots are never evaluated: all-zero words)
        uint32_t l[EagerPack<RB>::N];
#pragma unroll
        for (int j = 0; j < EagerPack<RB>::N; j++) l[j] = planes[(uint64_t)j * m_pad + slot];
#pragma unroll
        for (int j = 0; j < RB; j++) l[j] -= 2u * segs[b].bucket * bucket_vars;
        EagerPack<RB>::encode(l, w);
    }
#pragma unroll
    for (int i = 0; i < EagerPack<RB>::WORDS; i++) packed[(uint64_t)i * m_pad + slot] = w[i];
}

// ---- row-major copy for 5 <= k <= 8: rows[slot] = literals 0..7 (zero beyond k), 32 bytes -- as a pass of its own where
// the bucket scatter does not write it
__global__ void __launch_bounds__(256) rows8_kernel(const uint32_t *__restrict__ planes, uint64_t m_pad, uint32_t k,
                                                     uint4 *__restrict__ rows)
{
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= m_pad) return;
    uint32_t l[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (uint32_t j = 0; j < k && j < 8; j++) l[j] = planes[(uint64_t)j * m_pad + slot];
    rows[2 * slot] = make_uint4(l[0], l[1], l[2], l[3]);
    rows[2 * slot + 1] = make_uint4(l[4], l[5], l[6], l[7]);
}

cudaError_t launch_rows8(const uint32_t *planes, uint64_t m_pad, uint32_t k, uint4 *rows, cudaStream_t s)
{
    if (m_pad == 0) return cudaSuccess;
    rows8_kernel<<<(uint32_t)((m_pad + 255) / 256), 256, 0, s>>>(planes, m_pad, k, rows);
    return cudaGetLastError();
}

cudaError_t launch_pack_eager(const uint32_t *planes, uint64_t m_pad, const BucketSeg *segs, uint32_t n_buckets,
                              uint32_t bucket_vars, uint32_t rb, uint32_t *packed, cudaStream_t s)
{
    if (m_pad == 0) return cudaSuccess;
    const uint32_t grid = (uint32_t)((m_pad + 255) / 256);
    if (rb >= 2) pack_eager_kernel<2><<<grid, 256, 0, s>>>(planes, m_pad, segs, n_buckets, bucket_vars, packed);
    else pack_eager_kernel<1><<<grid, 256, 0, s>>>(planes, m_pad, segs, n_buckets, bucket_vars, packed);
    return cudaGetLastError();
}

// ---- between the two bucketing passes of one upload chunk: exclusive scan on the device ----------------------------
// cta_counts[b * n_cta + cta] for the chunk's CTAs [cta0, cta1) -> first slot of that CTA's clauses of bucket b; the chunk's
// n_buckets segments (each starting on a sweep-tile boundary at *tile_cursor, which moves on) -> segs_out[0 .. n_buckets).
// One CTA: warp w scans buckets w, w + 32, ...  No host round trip, so the scatter of a chunk follows its count directly
// and both run behind the H2D copy of the next chunk.
__global__ void __launch_bounds__(1024) bucket_scan_kernel(uint32_t *__restrict__ cta_counts, uint32_t n_cta, uint32_t cta0,
                                                            uint32_t cta1, uint32_t n_buckets, BucketSeg *__restrict__ segs_out,
                                                            uint32_t *__restrict__ tile_cursor)
{
    __shared__ uint32_t tot[MAX_BUCKETS], base[MAX_BUCKETS];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t b = warp; b < n_buckets; b += 32) {
        uint32_t run = 0;
        for (uint32_t i0 = cta0; i0 < cta1; i0 += 32) {
            const uint32_t i = i0 + lane;
            const uint32_t t = i < cta1 ? cta_counts[(uint64_t)b * n_cta + i] : 0u;
            uint32_t inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)lane >= o) inc += x;
            }
            if (i < cta1) cta_counts[(uint64_t)b * n_cta + i] = run + inc - t;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) tot[b] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t pos = (uint64_t)*tile_cursor * TILE;
        for (uint32_t b = 0; b < n_buckets; b++) {
            base[b] = (uint32_t)pos;
            segs_out[b] = BucketSeg{(uint32_t)(pos / TILE), (uint32_t)(pos + tot[b]), b, (uint32_t)(pos / TILE)};
            pos = (pos + tot[b] + TILE - 1) / TILE * TILE;
        }
        *tile_cursor = (uint32_t)(pos / TILE);
    }
    __syncthreads();
    const uint32_t span = cta1 - cta0;
    for (uint64_t x = threadIdx.x; x < (uint64_t)n_buckets * span; x += blockDim.x) {
        const uint32_t b = (uint32_t)(x / span), i = cta0 + (uint32_t)(x % span);
        cta_counts[(uint64_t)b * n_cta + i] += base[b];
    }
}

cudaError_t launch_bucket_scan(uint32_t *cta_counts, uint64_t m, uint64_t c0, uint64_t c1, uint32_t n_buckets, BucketSeg *segs_out,
                               uint32_t *tile_cursor, cudaStream_t s)
{
    bucket_scan_kernel<<<1, 1024, 0, s>>>(cta_counts, bucket_pass_ctas(m), (uint32_t)(c0 / LAYOUT_THREADS),
                                          (uint32_t)((c1 + LAYOUT_THREADS - 1) / LAYOUT_THREADS), n_buckets, segs_out, tile_cursor);
    return cudaGetLastError();
}

// true when launch_bucket_scatter can write the packed eager planes / tail rows itself (the shared-memory staged kernel)
bool bucket_scatter_fuses(uint32_t k, bool with_widths) { return k >= EAGER_PLANES && k <= 8 && !with_widths; }

// clauses [c0, c1) of the m the whole pass covers (c0 a multiple of bucket_pass_clauses_per_cta())
cudaError_t launch_bucket_scatter(const uint32_t *lit, uint64_t m, uint64_t c0, uint64_t c1, uint32_t k, uint32_t bucket_vars,
                                  uint32_t n_buckets, const uint8_t *bkt, const uint32_t *cta_base, uint32_t *planes, uint64_t m_pad,
                                  uint32_t *orig_id, uint32_t *min_resident, uint32_t resident_cap, const uint8_t *width_in,
                                  uint8_t *width_out, uint32_t *packed, uint4 *rows, cudaStream_t s)
{
    if (c1 <= c0) return cudaSuccess;
    if ((packed || rows) && !bucket_scatter_fuses(k, width_in != nullptr)) return cudaErrorInvalidValue;
    const uint32_t grid = blocks_for(c1 - c0, LAYOUT_THREADS), cta0 = (uint32_t)(c0 / LAYOUT_THREADS), n_cta = bucket_pass_ctas(m);
    if (k <= 8 && width_in == nullptr) {
        // static (17 KB) + dynamic shared memory exceed the 48 KB default: opt in (per device, cheap, idempotent)
        const cudaError_t e = cudaFuncSetAttribute(bucket_scatter_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)((k + 1) * LAYOUT_THREADS * 4));
        if (e != cudaSuccess) return e;
        bucket_scatter_staged_kernel<<<grid, LAYOUT_THREADS, (size_t)(k + 1) * LAYOUT_THREADS * 4, s>>>(
            lit, m, k, bucket_vars, n_buckets, bkt, cta_base, planes, m_pad, orig_id, min_resident, resident_cap, cta0, n_cta, packed, rows);
    } else
        bucket_scatter_kernel<<<grid, LAYOUT_THREADS, 0, s>>>(lit, m, k, bucket_vars, n_buckets, bkt, cta_base, planes, m_pad, orig_id,
                                                               min_resident, resident_cap, width_in, width_out, cta0, n_cta);
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(const uint8_t *bools, uint64_t n_vars, uint32_t *bits, uint32_t n_words_alloc, cudaStream_t s)
{
    pack_bits_kernel<<<blocks_for(n_words_alloc, 256), 256, 0, s>>>(bools, n_vars, bits, n_words_alloc);
    return cudaGetLastError();
}

cudaError_t launch_unpack_bits(const uint32_t *bits, uint64_t n_vars, uint8_t *bools, cudaStream_t s)
{
    if (n_vars == 0) return cudaSuccess;
    unpack_bits_kernel<<<blocks_for(n_vars, 256), 256, 0, s>>>(bits, n_vars, bools);
    return cudaGetLastError();
}

cudaError_t launch_randomize(uint32_t *bits, uint64_t n_vars, uint32_t n_words_alloc, uint64_t seed, cudaStream_t s)
{
    randomize_kernel<<<blocks_for((n_words_alloc + 3) / 4, 256), 256, 0, s>>>(bits, n_vars, n_words_alloc, seed);
    return cudaGetLastError();
}

// Device side of the packed host-to-device transport (hostpack.cpp): lo3[3 n] + hi[ceil(n / 8)] -> n 32-bit literals.
// One thread per 8 literals: 24 + 1 bytes in (three 8-byte loads), two 16-byte stores out; the last, partial group byte-wise.
__global__ void __launch_bounds__(256) unpack25_kernel(const uint8_t *__restrict__ lo3, const uint8_t *__restrict__ hi,
                                                      uint32_t *__restrict__ out, uint64_t n)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, i = g * 8;
    if (i >= n) return;
    const uint32_t hb = hi[g];
    if (i + 8 <= n) {
        const uint2 *p = reinterpret_cast<const uint2 *>(lo3 + 3 * i);
        const uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        const uint32_t w0 = a.x, w1 = a.y, w2 = b.x, w3 = b.y, w4 = c.x, w5 = c.y;
        uint4 o0, o1;
        o0.x = (w0 & 0xFFFFFFu) | ((hb & 1u) << 24);
        o0.y = (w0 >> 24) | ((w1 & 0xFFFFu) << 8) | (((hb >> 1) & 1u) << 24);
        o0.z = (w1 >> 16) | ((w2 & 0xFFu) << 16) | (((hb >> 2) & 1u) << 24);
        o0.w = (w2 >> 8) | (((hb >> 3) & 1u) << 24);
        o1.x = (w3 & 0xFFFFFFu) | (((hb >> 4) & 1u) << 24);
        o1.y = (w3 >> 24) | ((w4 & 0xFFFFu) << 8) | (((hb >> 5) & 1u) << 24);
        o1.z = (w4 >> 16) | ((w5 & 0xFFu) << 16) | (((hb >> 6) & 1u) << 24);
        o1.w = (w5 >> 8) | (((hb >> 7) & 1u) << 24);
        uint4 *q = reinterpret_cast<uint4 *>(out + i);
        q[0] = o0;
        q[1] = o1;
    } else {
        for (uint32_t j = 0; i + j < n; j++) {
            const uint8_t *b = lo3 + 3 * (i + j);
            out[i + j] = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | (((hb >> j) & 1u) << 24);
        }
    }
}

// lo3 8-byte aligned, out 16-byte aligned
cudaError_t launch_unpack25(const uint8_t *lo3, const uint8_t *hi, uint32_t *out, uint64_t n, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    unpack25_kernel<<<blocks_for((n + 7) / 8, 256), 256, 0, s>>>(lo3, hi, out, n);
    return cudaGetLastError();
}

cudaError_t launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long value, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    uint32_t grid = blocks_for(n, 256);
    if (grid > 148 * 16) grid = 148 * 16;
    fill_u64_kernel<<<grid, 256, 0, s>>>(p, n, value);
    return cudaGetLastError();
}

} // namespace alll
