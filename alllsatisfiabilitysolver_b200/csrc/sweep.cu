// sweep.cu -- K1+K2: violated-clause sweep with fused warp-ballot compaction (sm_100a).
//
// Replaces the reference's O1 site, SATInstance.h:273-280 -> Clause::is_not_satisfied (Clause.h:34-46):
// a clause is violated iff every literal is false; literal l is true iff bit(var=l>>1) != (l&1).
//
// Data path (DESIGN.md "Sweep kernel"):
//   * literals: k literal-major planes planes[j][slot], streamed once per sweep with 128-bit
//     ld.global.nc.L1::no_allocate loads -- 4 consecutive clause slots per thread per plane, fully
//     coalesced for every k (7-SAT rows are 28 B; planes need no padding);
//   * assignment: bit-packed, staged in shared memory.  When it does not fit (n > ~1.5 M variables)
//     the clauses were bucketed at upload by variable range so that >= 2 literals of every clause
//     (placed in the first planes) hit the bucket staged by the CTA; remaining literals are only
//     looked up (from L2) for clauses still unsatisfied after the resident ones -- ~0.3 L2 sectors
//     per clause instead of ~2;
//   * evaluation is lazy per clause (expected 2 lookups), level-major so the 4 clauses of a
//     thread keep 4 independent lookups in flight;
//   * violated slots are compacted per warp with __ballot_sync/__popc into a shared-memory staging
//     buffer and flushed with one global atomicAdd per >= 32 entries.
#include "alll_device.cuh"

namespace alll {

namespace {

struct WarpCompactor {
    uint32_t *wbuf;      // this warp's staging buffer (WBUF entries)
    uint32_t *viol;
    Counters *ctr;
    uint32_t count;      // warp-uniform
    uint32_t lane;

    __device__ __forceinline__ void flush()
    {
        __syncwarp();
        unsigned int g = 0;
        if (lane == 0) g = atomicAdd(&ctr->n_viol, count);
        g = __shfl_sync(0xffffffffu, g, 0);
        for (uint32_t i = lane; i < count; i += 32) viol[g + i] = wbuf[i];
        __syncwarp();
        count = 0;
    }

    // vmask: bit q set <=> clause slot (slot0 + q) is violated.  Must be called by the whole warp.
    __device__ __forceinline__ void push4(uint32_t vmask, uint32_t slot0)
    {
        if (!__any_sync(0xffffffffu, vmask != 0)) return;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool mine = (vmask >> q) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            if (bal) {
                if (mine) wbuf[count + __popc(bal & lt)] = slot0 + q;
                count += __popc(bal);
                if (count >= 32) flush();
            }
        }
    }
};

// true iff literal l is TRUE under the assignment
template <bool RESIDENT_ALL>
__device__ __forceinline__ uint32_t literal_true(uint32_t l, const uint32_t *sbits, const uint32_t *gbits,
                                                 uint32_t vbase, uint32_t bucket_vars)
{
    const uint32_t v = l >> 1;
    uint32_t w;
    if (RESIDENT_ALL) {
        w = sbits[v >> 5];
    } else {
        const uint32_t rel = v - vbase;                 // wraps to a huge value when v < vbase
        w = (rel < bucket_vars) ? sbits[rel >> 5] : __ldg(gbits + (v >> 5));
    }
    return ((w >> (v & 31u)) ^ l) & 1u;
}

} // namespace

// K > 0: compile-time clause width (all planes loaded up front, 8 x 128-bit loads in flight per thread for K=8).
// K == 0: run-time width p.k (planes loaded lazily, level by level).
template <int K, bool RESIDENT_ALL>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_planes_kernel(const SweepParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *sbits = smem;
    const uint32_t lane = threadIdx.x & 31u;
    WarpCompactor comp{smem + p.bucket_words + (threadIdx.x >> 5) * WBUF, p.viol, p.ctr, 0u, lane};

    const uint32_t t0 = (uint32_t)(((uint64_t)blockIdx.x * p.n_tiles) / gridDim.x);
    const uint32_t t1 = (uint32_t)(((uint64_t)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    if (t0 >= t1) return;

    uint32_t b = 0;
    while (b + 1 < p.n_buckets && p.segs[b + 1].tile_begin <= t0) ++b;
    uint32_t bucket_tile_end = (b + 1 < p.n_buckets) ? p.segs[b + 1].tile_begin : p.n_tiles;
    uint32_t slot_end = p.segs[b].slot_end;
    uint32_t loaded = 0xFFFFFFFFu;
    const uint32_t bucket_vars = p.bucket_words * 32u;

    for (uint32_t tile = t0; tile < t1; ++tile) {
        while (tile >= bucket_tile_end) {
            ++b;
            bucket_tile_end = (b + 1 < p.n_buckets) ? p.segs[b + 1].tile_begin : p.n_tiles;
            slot_end = p.segs[b].slot_end;
        }
        if (b != loaded) {
            __syncthreads();                      // everyone is done with the previous bucket's bits
            const uint4 *src = reinterpret_cast<const uint4 *>(p.bits + (uint64_t)b * p.bucket_words);
            uint4 *dst = reinterpret_cast<uint4 *>(sbits);
            for (uint32_t i = threadIdx.x; i < p.bucket_words / 4; i += SWEEP_THREADS) dst[i] = __ldg(src + i);
            __syncthreads();
            loaded = b;
        }
        const uint32_t vbase = b * bucket_vars;
        const uint32_t slot0 = tile * TILE + threadIdx.x * CLAUSES_PER_THREAD;
        const uint32_t *src = p.planes + slot0;

        uint32_t alive = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) alive |= (slot0 + q < slot_end) ? (1u << q) : 0u;

        if (K > 0) {
            uint4 L[K > 0 ? K : 1];
#pragma unroll
            for (int j = 0; j < K; j++) L[j] = ld_stream_v4(src + (uint64_t)j * p.m_pad);
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (alive & 1u) alive &= ~(literal_true<RESIDENT_ALL>(L[j].x, sbits, p.bits, vbase, bucket_vars) << 0);
                if (alive & 2u) alive &= ~(literal_true<RESIDENT_ALL>(L[j].y, sbits, p.bits, vbase, bucket_vars) << 1);
                if (alive & 4u) alive &= ~(literal_true<RESIDENT_ALL>(L[j].z, sbits, p.bits, vbase, bucket_vars) << 2);
                if (alive & 8u) alive &= ~(literal_true<RESIDENT_ALL>(L[j].w, sbits, p.bits, vbase, bucket_vars) << 3);
            }
        } else {
            for (uint32_t j = 0; j < p.k && alive; j++) {
                const uint4 Lj = ld_stream_v4(src + (uint64_t)j * p.m_pad);
                if (alive & 1u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.x, sbits, p.bits, vbase, bucket_vars) << 0);
                if (alive & 2u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.y, sbits, p.bits, vbase, bucket_vars) << 1);
                if (alive & 4u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.z, sbits, p.bits, vbase, bucket_vars) << 2);
                if (alive & 8u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.w, sbits, p.bits, vbase, bucket_vars) << 3);
            }
        }
        comp.push4(alive, slot0);
    }
    if (comp.count) comp.flush();
}

// Variable-width fallback (general DIMACS input): one clause per thread over CSR, assignment words
// gathered through L1/L2.  Not the roofline-graded path.
__global__ void __launch_bounds__(256) sweep_csr_kernel(const uint64_t *__restrict__ off, const uint32_t *__restrict__ lit,
                                                         uint64_t m, const uint32_t *__restrict__ bits,
                                                         uint32_t *viol, Counters *ctr)
{
    __shared__ uint32_t wbuf_all[8 * WBUF];
    const uint32_t lane = threadIdx.x & 31u;
    WarpCompactor comp{wbuf_all + (threadIdx.x >> 5) * WBUF, viol, ctr, 0u, lane};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t m_round = (m + 31) / 32 * 32;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < m_round; c += stride) {
        uint32_t violated = 0;
        if (c < m) {
            violated = 1;
            const uint64_t e = off[c + 1];
            for (uint64_t j = off[c]; j < e; j++) {
                const uint32_t l = __ldg(lit + j);
                const uint32_t v = l >> 1;
                if (((__ldg(bits + (v >> 5)) >> (v & 31u)) ^ l) & 1u) { violated = 0; break; }
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, violated);
        if (bal) {
            if (violated) comp.wbuf[comp.count + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)c;
            comp.count += __popc(bal);
            if (comp.count >= 32) comp.flush();
        }
    }
    if (comp.count) comp.flush();
}

// ---- launchers ------------------------------------------------------------------------

template <int K, bool R>
static cudaError_t launch_planes(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, bool configure_only)
{
    if (configure_only)   // function attributes are per device: the handle configures its kernel once at upload
        return cudaFuncSetAttribute(sweep_planes_kernel<K, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sweep_planes_kernel<K, R><<<grid, SWEEP_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

template <bool R>
static cudaError_t dispatch_k(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, bool cfg)
{
    switch (p.k) {
    case 1: return launch_planes<1, R>(p, grid, smem, s, cfg);
    case 2: return launch_planes<2, R>(p, grid, smem, s, cfg);
    case 3: return launch_planes<3, R>(p, grid, smem, s, cfg);
    case 4: return launch_planes<4, R>(p, grid, smem, s, cfg);
    case 5: return launch_planes<5, R>(p, grid, smem, s, cfg);
    case 6: return launch_planes<6, R>(p, grid, smem, s, cfg);
    case 7: return launch_planes<7, R>(p, grid, smem, s, cfg);
    case 8: return launch_planes<8, R>(p, grid, smem, s, cfg);
    default: return launch_planes<0, R>(p, grid, smem, s, cfg);
    }
}

size_t sweep_planes_smem_bytes(uint32_t bucket_words)
{
    return (size_t)bucket_words * 4 + (SWEEP_THREADS / 32) * WBUF * 4;
}

cudaError_t configure_sweep_planes(const SweepParams &p, bool resident_all)
{
    const size_t smem = sweep_planes_smem_bytes(p.bucket_words);
    return resident_all ? dispatch_k<true>(p, 0, smem, 0, true) : dispatch_k<false>(p, 0, smem, 0, true);
}

cudaError_t launch_sweep_planes(const SweepParams &p, bool resident_all, uint32_t grid, cudaStream_t s)
{
    const size_t smem = sweep_planes_smem_bytes(p.bucket_words);
    return resident_all ? dispatch_k<true>(p, grid, smem, s, false) : dispatch_k<false>(p, grid, smem, s, false);
}

cudaError_t launch_sweep_csr(const uint64_t *off, const uint32_t *lit, uint64_t m, const uint32_t *bits,
                             uint32_t *viol, Counters *ctr, uint32_t grid, cudaStream_t s)
{
    sweep_csr_kernel<<<grid, 256, 0, s>>>(off, lit, m, bits, viol, ctr);
    return cudaGetLastError();
}

} // namespace alll
