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
//   * only the first E = min(k, 5) planes are streamed.  Evaluation over them is lazy per clause and
//     branch-free; their non-resident lookups are issued in one batch (one L2 round trip); the next
//     tile's literals are already in flight (register double buffering).  The 2^-E fraction of clauses
//     still unsatisfied is parked per warp and finished densely, one clause per lane, fetching the
//     tail literals only then -- the reference's early exit (Clause.h:42-44) applied to HBM traffic;
//   * violated slots are compacted per warp with __ballot_sync/__popc into a shared-memory staging
//     buffer and flushed with one global atomicAdd per >= 32 entries.
#include "sweep_body.cuh"

namespace alll {

template <int K, int RB, int RC, int E, bool PK>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_planes_kernel(const SweepParams p)
{
    if (__ldcg(&p.ctr->done) || __ldcg(&p.ctr->incr_next)) return;   // behind the terminal round / this round is incremental
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.round < DBG_ROUNDS) p.ctr->dbg[p.round][0] = global_ns();
    sweep_planes_body<K, RB, RC, E, true, PK>(p, &p.ctr->n_viol, p.p2p_parity, true);
}

// Run-time clause width (k > 8): planes are loaded lazily level by level; no prefetch.
template <bool RESIDENT_ALL>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_planes_generic_kernel(const SweepParams p)
{
    if (__ldcg(&p.ctr->done) || __ldcg(&p.ctr->incr_next)) return;
    const uint32_t lane = threadIdx.x & 31u;
    WarpCompactor comp{p.bucket_words + (threadIdx.x >> 5) * WBUF, p.viol, p.ctr, &p.ctr->n_viol, 0u, false, 0u, lane, nullptr};

    const uint32_t t0 = (uint32_t)(((uint64_t)blockIdx.x * p.n_tiles) / gridDim.x);
    const uint32_t t1 = (uint32_t)(((uint64_t)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    if (t0 >= t1) return;
    TileCursor cur;
    cur.init(p, t0);
    const uint32_t bucket_vars = p.bucket_words * 32u;

    for (uint32_t tile = t0; tile < t1; ++tile) {
        cur.enter(p, tile);
        const uint32_t vbase = cur.bucket * bucket_vars;
        const uint32_t slot0 = cur.phys * TILE + threadIdx.x * CLAUSES_PER_THREAD;
        const uint32_t *src = p.planes + slot0;
        uint32_t alive = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) alive |= (slot0 + q < cur.slot_end) ? (1u << q) : 0u;
        for (uint32_t j = 0; j < p.k && alive; j++) {
            const uint4 Lj = ld_stream_v4(src + (uint64_t)j * p.m_pad);
            if (alive & 1u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.x, p.bits, vbase, bucket_vars) << 0);
            if (alive & 2u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.y, p.bits, vbase, bucket_vars) << 1);
            if (alive & 4u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.z, p.bits, vbase, bucket_vars) << 2);
            if (alive & 8u) alive &= ~(literal_true<RESIDENT_ALL>(Lj.w, p.bits, vbase, bucket_vars) << 3);
        }
        comp.push4(alive, slot0);
    }
    if (comp.count) comp.flush();
}

// (Variable-width clauses that stay in CSR form: csr.cu / csr_body.cuh -- the warp-cooperative sweep.)

// ---- launchers ------------------------------------------------------------------------

namespace {
struct SweepOp {                 // launch or (configure) opt in to the shared memory of one sweep_planes_kernel variant
    const SweepParams &p;
    uint32_t grid;
    size_t smem;
    cudaStream_t s;
    bool configure;
    template <int K, int RB, int RC, int E, bool PK> cudaError_t run()
    {
        if (configure)   // function attributes are per device: the handle configures its kernel once at upload
            return cudaFuncSetAttribute(sweep_planes_kernel<K, RB, RC, E, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        sweep_planes_kernel<K, RB, RC, E, PK><<<grid, SWEEP_THREADS, smem, s>>>(p);
        return cudaGetLastError();
    }
};
} // namespace

template <bool R>
static cudaError_t launch_generic(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, bool configure)
{
    if (configure)
        return cudaFuncSetAttribute(sweep_planes_generic_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sweep_planes_generic_kernel<R><<<grid, SWEEP_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

static cudaError_t sweep_op(const SweepParams &p, bool resident_all, uint32_t grid, cudaStream_t s, bool configure)
{
    const size_t smem = sweep_smem_bytes_for(p.bucket_words);
    if (p.k > 8) return resident_all ? launch_generic<true>(p, grid, smem, s, configure) : launch_generic<false>(p, grid, smem, s, configure);
    if (!configure && (p.runs == nullptr || grid != p.run_grid)) return cudaErrorInvalidValue;   // the run lists are cut for one grid size
    SweepOp op{p, grid, smem, s, configure};
    return dispatch_variant(p, resident_all, op);
}

size_t sweep_planes_smem_bytes(uint32_t bucket_words) { return sweep_smem_bytes_for(bucket_words); }

cudaError_t configure_sweep_planes(const SweepParams &p, bool resident_all) { return sweep_op(p, resident_all, 0, 0, true); }

cudaError_t launch_sweep_planes(const SweepParams &p, bool resident_all, uint32_t grid, cudaStream_t s)
{
    return sweep_op(p, resident_all, grid, s, false);
}

} // namespace alll
