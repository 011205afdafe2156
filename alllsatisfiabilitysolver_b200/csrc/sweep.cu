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
#include "alll_device.cuh"
#include "incr_body.cuh"
#include "mis_body.cuh"

namespace alll {

// All shared-memory traffic indexes this array directly (never through a stored pointer): a generic
// pointer would make the compiler rebuild the shared-window base (S2R SR_CgaCtaId + LEA) at every lookup.
extern __shared__ __align__(16) uint32_t g_smem[];

namespace {

// Next to the violated list the sweep leaves one record {caller id, k literals} per violated clause (for the first
// urec_cap of them): the independent-set kernel that runs next then reads each clause with one contiguous access
// instead of chasing k literal planes through cold DRAM on its critical path.  Here the k + 1 scattered reads overlap
// with the streaming of the other warps.  Out of line: it runs once per >= 32 violated clauses and must not cost the
// streaming loop registers.
// (Arguments by value: taking the address of the kernel's parameter block would move it to local memory, and the
// streaming loop would then read its parameters through L1/L2 instead of the constant bank -- measured 0.22 -> 0.30 ms.)
__device__ __noinline__ void write_records(uint32_t *__restrict__ urec, const uint32_t *__restrict__ planes, uint64_t m_pad,
                                           const uint32_t *__restrict__ orig_id, uint32_t id_base, uint32_t k, uint32_t wbuf,
                                           uint32_t g, uint32_t count, uint32_t lane)
{
    // One (clause, word) pair per lane and pass; k <= 8, count <= 63 => at most 18 passes.  All scattered reads are issued
    // before the first store (one DRAM round trip per flush instead of one per pass), the stores of a pass are consecutive.
    const uint32_t w = k + 1, total = count * w;
    uint32_t *out = urec + (uint64_t)g * w;
    for (uint32_t t0 = 0; t0 < total; t0 += 32 * 9) {
        uint32_t val[9];
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const uint32_t t = t0 + q * 32 + lane;
            val[q] = 0;
            if (t < total) {
                const uint32_t i = t / w, j = t - i * w;
                const uint32_t slot = g_smem[wbuf + i];
                val[q] = j == 0 ? (orig_id ? __ldg(orig_id + slot) : slot) + id_base : __ldg(planes + (uint64_t)(j - 1) * m_pad + slot);
            }
        }
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const uint32_t t = t0 + q * 32 + lane;
            if (t < total) out[t] = val[q];
        }
    }
}

struct WarpCompactor {
    uint32_t wbuf;       // index in g_smem of this warp's staging buffer (WBUF entries)
    uint32_t *viol;
    Counters *ctr;
    unsigned int *n_viol;   // where |U| is accumulated (ctr->n_viol, or the round-parity counter of the persistent solve kernel)
    uint32_t p2p_parity; // sharded P2P mode: which of the two record areas this round uses
    bool rec_on;         // write records next to the violated list this round
    uint32_t count;      // warp-uniform
    uint32_t lane;
    const SweepParams *sp;   // non-NULL with sp->p2p set: sharded P2P mode

    __device__ __forceinline__ void flush()
    {
        __syncwarp();
        unsigned int g = 0;
        if (lane == 0) g = atomicAdd(n_viol, count);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (sp != nullptr && sp->p2p != nullptr) {
            // fused compute + collective: the violated clauses go straight into every GPU's receive slot for this
            // rank and round (NVLink P2P stores), as records {global id, k literals}
            const P2PLink &L = *sp->p2p;
            if ((uint64_t)g + count > L.cap) {
                if (lane == 0) { ctr->p2p_error = 1; for (uint32_t q = 0; q < L.world; q++) L.hdr[q]->abort = 1; }
            } else {
                const uint32_t w = L.k + 1;
                const uint64_t base = (((uint64_t)p2p_parity * L.world + L.rank) * L.cap + g) * w;
                for (uint32_t i = lane; i < count; i += 32) {
                    const uint32_t slot = g_smem[wbuf + i];
                    for (uint32_t j = 0; j < w; j++) {
                        const uint32_t word = j == 0 ? (sp->orig_id ? sp->orig_id[slot] : slot) + sp->id_base
                                                     : sp->planes[(uint64_t)(j - 1) * sp->m_pad + slot];
                        for (uint32_t q = 0; q < L.world; q++) L.rec[q][base + (uint64_t)i * w + j] = word;
                    }
                }
            }
        } else {
            for (uint32_t i = lane; i < count; i += 32) viol[g + i] = g_smem[wbuf + i];
            if (rec_on && sp != nullptr && sp->urec != nullptr && (uint64_t)g + count <= sp->urec_cap)
                write_records(sp->urec, sp->planes, sp->m_pad, sp->orig_id, sp->id_base, sp->k, wbuf, g, count, lane);
        }
        __syncwarp();
        count = 0;
    }

    // One candidate per lane.  Must be called by the whole warp.
    __device__ __forceinline__ void push1(bool mine, uint32_t slot)
    {
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (!bal) return;
        if (mine) g_smem[wbuf + count + __popc(bal & ((1u << lane) - 1u))] = slot;
        count += __popc(bal);
        if (count >= 32) flush();
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
                if (mine) g_smem[wbuf + count + __popc(bal & lt)] = slot0 + q;
                count += __popc(bal);
                if (count >= 32) flush();
            }
        }
    }
};

// Sharded P2P mode, end of the sweep kernel: every CTA orders its record stores before its ticket; the CTA that
// draws the last ticket publishes this rank's count and arrival flag on every GPU.
__device__ __forceinline__ void p2p_publish(const SweepParams &p)
{
    if (p.p2p == nullptr) return;
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence_system();
    const unsigned int t = atomicAdd(&p.ctr->cta_done, 1u);
    if (t != gridDim.x - 1) return;
    p.ctr->cta_done = 0;                                   // ready for the next launch (stream-ordered)
    __threadfence_system();
    const P2PLink &L = *p.p2p;
    const unsigned int total = __ldcg(&p.ctr->n_viol);
    for (uint32_t q = 0; q < L.world; q++) L.hdr[q]->count[p.p2p_parity][L.rank] = total;
    __threadfence_system();
    for (uint32_t q = 0; q < L.world; q++) *(volatile unsigned int *)&L.hdr[q]->flag[p.p2p_parity][L.rank] = p.p2p_tag;
}

// true iff literal l is TRUE under the assignment
template <bool RESIDENT_ALL>
__device__ __forceinline__ uint32_t literal_true(uint32_t l, const uint32_t *gbits, uint32_t vbase, uint32_t bucket_vars)
{
    const uint32_t v = l >> 1;
    uint32_t w;
    if (RESIDENT_ALL) {
        w = g_smem[v >> 5];
    } else {
        const uint32_t rel = v - vbase;                 // wraps to a huge value when v < vbase
        w = (rel < bucket_vars) ? g_smem[rel >> 5] : __ldg(gbits + (v >> 5));
    }
    return ((w >> (v & 31u)) ^ l) & 1u;
}

} // namespace

// ---- per-tile bookkeeping shared by both plane kernels ---------------------------------------------
struct TileCursor {
    uint32_t b, bucket_tile_end, slot_end, loaded;

    __device__ __forceinline__ void init(const SweepParams &p, uint32_t t0)
    {
        b = 0;
        while (b + 1 < p.n_buckets && p.segs[b + 1].tile_begin <= t0) ++b;
        bucket_tile_end = (b + 1 < p.n_buckets) ? p.segs[b + 1].tile_begin : p.n_tiles;
        slot_end = p.segs[b].slot_end;
        loaded = 0xFFFFFFFFu;
    }
    // Moves to `tile`; returns true when its bucket differs from the staged one (caller must then stage()).
    __device__ __forceinline__ bool advance(const SweepParams &p, uint32_t tile)
    {
        while (tile >= bucket_tile_end) {
            ++b;
            bucket_tile_end = (b + 1 < p.n_buckets) ? p.segs[b + 1].tile_begin : p.n_tiles;
            slot_end = p.segs[b].slot_end;
        }
        return b != loaded;
    }
    // Stages bucket b's slice of the assignment into shared memory (whole CTA).
    __device__ __forceinline__ void stage(const SweepParams &p)
    {
        __syncthreads();                      // everyone is done with the previous bucket's bits
        const uint4 *src = reinterpret_cast<const uint4 *>(p.bits + (uint64_t)b * p.bucket_words);
        for (uint32_t i = threadIdx.x; i < p.bucket_words / 4; i += SWEEP_THREADS)
            reinterpret_cast<uint4 *>(g_smem)[i] = __ldg(src + i);
        __syncthreads();
        loaded = b;
    }
    __device__ __forceinline__ void enter(const SweepParams &p, uint32_t tile)
    {
        if (advance(p, tile)) stage(p);
    }
};

// ---- literal evaluation, branch-free ------------------------------------------------------------------
// The upload pass orders every clause's literals bucket-resident first (at most RC of them), so the planes
// fall into three static classes and each class gets the cheapest code:
//   planes [0, RB)  : resident for EVERY clause           -> shared-memory lookup, no range test
//   planes [RB, RC) : resident for some clauses            -> range test, shared memory or L2 gather
//   planes [RC, K)  : never treated as resident           -> L2 gather only
// (RB = K means the whole assignment is staged and nothing is ever gathered.)
// All lookups are predicated on the clause still being alive: a dead lane issues no request, so it costs
// neither a bank conflict nor an L2 sector.

// Shared-memory word load from a 32-bit shared-window byte address.  The hot lookups use this instead of
// g_smem[...]: with the address base held in an (opaque) register the lookup is SHF + LEA + LDS, whereas
// nvcc rebuilds the window base (S2R SR_CgaCtaId, MOV, LEA) for every predicated g_smem[] access.
// Not volatile on purpose (the scheduler may interleave lookups freely); ordering against the staging
// barrier comes from the address base, which is re-materialised through an opaque asm after each barrier.
__device__ __forceinline__ uint32_t lds32(uint32_t byte_addr)
{
    uint32_t w;
    asm("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(byte_addr));
    return w;
}

// sadj = shared byte address of staged word 0 minus 4 * (vbase >> 5): sadj + 4 * (v >> 5) addresses the word
// of a resident variable v.
__device__ __forceinline__ void resident_only_step(uint32_t l, uint32_t &alive, uint32_t sadj)
{
    const bool go = alive != 0;
    const uint32_t w = go ? lds32(sadj + ((l >> 6) << 2)) : 0u;
    const uint32_t lit_true = (__funnelshift_r(w, 0u, l >> 1) ^ l) & 1u;   // bit (v & 31) of w, xor the negation flag
    alive = (go && lit_true) ? 0u : alive;
}

__device__ __forceinline__ void resident_mixed_step(uint32_t l, uint32_t &alive, uint32_t sadj, uint32_t vbase,
                                                    uint32_t bucket_vars)
{
    const uint32_t v = l >> 1;
    const bool go = alive != 0 && (v - vbase) < bucket_vars;               // v - vbase wraps when v < vbase
    const uint32_t w = go ? lds32(sadj + ((v >> 5) << 2)) : 0u;
    const uint32_t lit_true = (__funnelshift_r(w, 0u, v) ^ l) & 1u;
    alive = (go && lit_true) ? 0u : alive;
}

template <bool TEST_RANGE>
__device__ __forceinline__ void gather_issue(uint32_t l, uint32_t alive, const uint32_t *gbits, uint32_t vbase,
                                             uint32_t bucket_vars, uint32_t &w, bool &go)
{
    const uint32_t v = l >> 1;
    go = TEST_RANGE ? (alive != 0 && (v - vbase) >= bucket_vars) : (alive != 0);
    w = go ? __ldg(gbits + (v >> 5)) : 0u;
}
__device__ __forceinline__ void gather_apply(uint32_t l, uint32_t &alive, uint32_t w, bool go)
{
    const uint32_t lit_true = (__funnelshift_r(w, 0u, l >> 1) ^ l) & 1u;
    alive = (go && lit_true) ? 0u : alive;
}

__device__ __forceinline__ uint32_t comp(const uint4 &v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }

// Gathers for planes [J0, J1): all issued back to back (one L2 round trip), then applied.
template <int K, int RB, int RC, int J0, int J1>
__device__ __forceinline__ void gather_round(const uint4 (&L)[K], uint32_t (&a)[4], const uint32_t *gbits,
                                             uint32_t vbase, uint32_t bucket_vars)
{
    if constexpr (J1 > J0) {
        uint32_t w[J1 - J0][4];
        bool go[J1 - J0][4];
#pragma unroll
        for (int j = J0; j < J1; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (j < RC) gather_issue<true>(comp(L[j], q), a[q], gbits, vbase, bucket_vars, w[j - J0][q], go[j - J0][q]);
                else gather_issue<false>(comp(L[j], q), a[q], gbits, vbase, bucket_vars, w[j - J0][q], go[j - J0][q]);
            }
#pragma unroll
        for (int j = J0; j < J1; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) gather_apply(comp(L[j], q), a[q], w[j - J0][q], go[j - J0][q]);
    }
}

// Evaluates the first E literals (the planes held in registers) of 4 clauses; component q of every plane is
// clause slot0+q.  Returns the still-unsatisfied mask (bit q).  Phase R: shared memory.  Phase G: the
// non-resident literals of planes [RB, E), all issued at once -- one L2 round trip.
template <int E, int RB, int RC>
__device__ __forceinline__ uint32_t eval4(const uint4 (&L)[E], uint32_t valid_mask, uint32_t sadj,
                                          const uint32_t *gbits, uint32_t vbase, uint32_t bucket_vars)
{
    constexpr int R_END = RC < E ? RC : E;
    uint32_t a[4] = {valid_mask & 1u, valid_mask & 2u, valid_mask & 4u, valid_mask & 8u};
#pragma unroll
    for (int j = 0; j < R_END; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (j < RB) resident_only_step(comp(L[j], q), a[q], sadj);
            else resident_mixed_step(comp(L[j], q), a[q], sadj, vbase, bucket_vars);
        }
    gather_round<E, RB, RC, (RB < E ? RB : E), E>(L, a, gbits, vbase, bucket_vars);
    return (a[0] ? 1u : 0u) | (a[1] ? 2u : 0u) | (a[2] ? 4u : 0u) | (a[3] ? 8u : 0u);
}

// Clauses that survive their first E literals (a 2^-E fraction) are parked per warp and finished densely, 32
// at a time, one clause per lane: only then are their remaining K-E literals fetched (scalar loads) and
// looked up.  The planes [E, K) are therefore never streamed: like the reference's early exit
// (Clause.h:42-44), most clauses are decided without reading their tail literals.
template <int K, int E, bool RESIDENT_ALL>
struct SurvivorQueue {
    uint32_t qbuf;       // index in g_smem of this warp's queue (QBUF entries)
    uint32_t count;      // warp-uniform
    uint32_t lane;

    __device__ __forceinline__ void push4(uint32_t mask, uint32_t slot0)
    {
        if (!__any_sync(0xffffffffu, mask != 0)) return;
        const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool mine = (mask >> q) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            if (mine) g_smem[qbuf + count + __popc(bal & lt)] = slot0 + q;
            count += __popc(bal);
        }
    }

    // Finishes parked clauses while at least `keep` + 1 are queued (keep = 31: full batches only; 0: everything).
    __device__ __forceinline__ void drain(uint32_t keep, WarpCompactor &out, const SweepParams &p, uint32_t vbase,
                                          uint32_t bucket_vars)
    {
        while (count > keep) {
            __syncwarp();
            const uint32_t n = count < 32u ? count : 32u;
            const bool act = lane < n;
            const uint32_t slot = act ? g_smem[qbuf + count - n + lane] : 0u;
            constexpr int T = K > E ? K - E : 1;     // tail planes (T = 1 only keeps the arrays legal when E == K)
            uint32_t l[T];
#pragma unroll
            for (int j = 0; j < K - E; j++) l[j] = act ? __ldg(p.planes + (uint64_t)(E + j) * p.m_pad + slot) : 0u;
            uint32_t w[T];
#pragma unroll
            for (int j = 0; j < K - E; j++) {                 // all lookups at once: this path is rare and dense
                const uint32_t v = l[j] >> 1;
                if (RESIDENT_ALL) w[j] = act ? g_smem[v >> 5] : 0u;
                else {
                    const uint32_t rel = v - vbase;
                    w[j] = !act ? 0u : (rel < bucket_vars) ? g_smem[rel >> 5] : __ldg(p.bits + (v >> 5));
                }
            }
            bool violated = act;
#pragma unroll
            for (int j = 0; j < K - E; j++) violated = violated && !(((w[j] >> ((l[j] >> 1) & 31u)) ^ l[j]) & 1u);
            count -= n;
            __syncwarp();
            out.push1(violated, slot);
        }
    }
};

// Compile-time clause width K, of which the first E planes are streamed.  One CTA per SM; each thread owns 4
// consecutive clause slots of a tile and keeps TWO tiles of literals in registers: the next tile's E x 128-bit
// loads are in flight while the current tile is evaluated (register double buffering).
// TICKET: (sharded P2P mode) the CTA that finishes last publishes this rank's round to the peers; the persistent solve
// kernel publishes after its grid barrier instead.
template <int K, int RB, int RC, int E, bool TICKET>
__device__ __forceinline__ void sweep_planes_body(const SweepParams &p, unsigned int *n_viol_ctr, uint32_t p2p_parity, bool rec_on)
{
    constexpr bool RESIDENT_ALL = RB >= K;
    constexpr int RBE = RB < E ? RB : E;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    WarpCompactor out{p.bucket_words + warp * WBUF, p.viol, p.ctr, n_viol_ctr, p2p_parity, rec_on, 0u, lane, &p};
    SurvivorQueue<K, E, RESIDENT_ALL> parked{p.bucket_words + (SWEEP_THREADS / 32) * WBUF + warp * QBUF, 0u, lane};

    const uint32_t t0 = (uint32_t)(((uint64_t)blockIdx.x * p.n_tiles) / gridDim.x);
    const uint32_t t1 = (uint32_t)(((uint64_t)(blockIdx.x + 1) * p.n_tiles) / gridDim.x);
    if (t0 >= t1) {
        if (TICKET) p2p_publish(p);
        return;
    }

    TileCursor cur;
    cur.init(p, t0);
    const uint32_t bucket_vars = p.bucket_words * 32u;
    const uint32_t *base = p.planes + threadIdx.x * CLAUSES_PER_THREAD;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(g_smem);

    auto load = [&](uint4 (&L)[E], uint32_t tile) {
        const uint32_t *src = base + (uint64_t)tile * TILE;
#pragma unroll
        for (int j = 0; j < E; j++) L[j] = ld_stream_v4(src + (uint64_t)j * p.m_pad);
    };
    auto process = [&](const uint4 (&L)[E], uint32_t tile) {
        const uint32_t prev_vbase = cur.b * bucket_vars;
        if (cur.advance(p, tile)) {
            if constexpr (E < K) parked.drain(0u, out, p, prev_vbase, bucket_vars);   // parked clauses belong to the old bucket
            cur.stage(p);
        }
        const uint32_t slot0 = tile * TILE + threadIdx.x * CLAUSES_PER_THREAD;
        uint32_t valid = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) valid |= (slot0 + q < cur.slot_end) ? (1u << q) : 0u;
        const uint32_t vbase = cur.b * bucket_vars;
        uint32_t sb = smem_base;
        asm volatile("" : "+r"(sb));          // opaque: lookups below cannot be hoisted above the staging barrier
        const uint32_t alive = eval4<E, RBE, RC>(L, valid, sb - ((vbase >> 5) << 2), p.bits, vbase, bucket_vars);
        if constexpr (E < K) {
            parked.push4(alive, slot0);
            parked.drain(31u, out, p, vbase, bucket_vars);
        } else {
            out.push4(alive, slot0);
        }
    };

    // HBM -> L2: one thread per CTA bulk-prefetches the E plane segments of the tile `dist` ahead of the register
    // double buffer, so enough bytes are in flight to cover the loaded DRAM latency without spending registers.
    const uint32_t dist = p.prefetch_tiles;
    auto prefetch = [&](uint32_t tile) {
        if (threadIdx.x == 0 && dist != 0 && tile < t1) {
#pragma unroll
            for (int j = 0; j < E; j++) tma_prefetch_l2(p.planes + (uint64_t)j * p.m_pad + (uint64_t)tile * TILE, TILE * 4);
        }
    };
    for (uint32_t d = 2; d < 2 + dist; d++) prefetch(t0 + d);

    uint4 A[E], B[E];
    load(A, t0);
    for (uint32_t tile = t0; tile < t1; tile += 2) {
        if (tile + 1 < t1) load(B, tile + 1);
        prefetch(tile + 2 + dist);
        process(A, tile);
        if (tile + 1 >= t1) break;
        if (tile + 2 < t1) load(A, tile + 2);
        prefetch(tile + 3 + dist);
        process(B, tile + 1);
    }
    if constexpr (E < K) parked.drain(0u, out, p, cur.b * bucket_vars, bucket_vars);
    if (out.count) out.flush();
    if (TICKET) p2p_publish(p);
}

template <int K, int RB, int RC, int E>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) sweep_planes_kernel(const SweepParams p)
{
    if (__ldcg(&p.ctr->done) || __ldcg(&p.ctr->incr_next)) return;   // behind the terminal round / this round is incremental
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.round < DBG_ROUNDS) p.ctr->dbg[p.round][0] = global_ns();
    sweep_planes_body<K, RB, RC, E, true>(p, &p.ctr->n_viol, p.p2p_parity, true);
}

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
template <int K, int RB, int RC, int E>
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
    GridBarrier bar{cg::this_grid()};
    Counters *const c = sp.ctr;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    const bool p2p = sp.p2p != nullptr;      // clause-range sharded solve: every GPU runs this kernel on its range
    const uint32_t first = blockIdx.x * SWEEP_THREADS + threadIdx.x, stride = gridDim.x * SWEEP_THREADS;
    unsigned long long t_sweep = 0, t_mis = 0;
    uint32_t prev_n_u = (uint32_t)min((uint64_t)0xFFFFFFFFu, ((uint64_t)sp.n_tiles * TILE) >> K);
    for (uint32_t round = 0; round < max_rounds; ++round) {
        const uint32_t par = round & 1u, tag = ((epoch & 0xFFFu) << 20) | (round + 1u);
        unsigned long long t0 = 0, t1 = 0;
        if (lead) {
            t0 = global_ns();
            if (round < DBG_ROUNDS) c->dbg[round][0] = t0;
        }
        // records next to the violated list only while the violated set is expected to fit them (the previous round's
        // |U|, or m / 2^K before the first round): writing the first urec_cap records of a larger set is wasted work
        const bool rec_on = (uint64_t)prev_n_u <= 2ull * sp.urec_cap;
        if (threadIdx.x == 0) { s_mp.p2p_parity = par; s_mp.p2p_tag = tag; s_mp.urec_cap = rec_on ? sp.urec_cap : 0u; }
        // incremental mode (ip.rows != NULL): the round that just ended decided whether this round's violated set comes
        // from the occurrence lists of the variables it resampled (same set as the sweep's, incremental.cu) or from a sweep
        const bool incremental = ip.rows != nullptr && round > 0 && gm::ld_cg(&c->incr_next) != 0;
        if (incremental) incr_eval_body(ip, gm::ld_cg(&c->last_n_s), &c->n_viol_pp[par]);
        else sweep_planes_body<K, RB, RC, E, false>(sp, &c->n_viol_pp[par], par, rec_on);
        if (p2p) {                                       // this CTA's record stores (NVLink) are ordered before the barrier
            __syncthreads();
            if (threadIdx.x == 0) __threadfence_system();
        }
        bar.sync();
        uint32_t n_u;
        if (p2p) {
            // fused exchange: the violated records went straight into every GPU's region during the sweep; publish our
            // count + arrival flag everywhere, then wait for every peer's flag of this round
            if (lead) {
                const P2PLink &L = *sp.p2p;
                const unsigned int total = gm::ld_cg(&c->n_viol_pp[par]);
                for (uint32_t q = 0; q < L.world; q++) L.hdr[q]->count[par][L.rank] = total;
                __threadfence_system();
                for (uint32_t q = 0; q < L.world; q++) *(volatile unsigned int *)&L.hdr[q]->flag[par][L.rank] = tag;
            }
            n_u = p2p_wait(mp, s_prefix);
        } else {
            n_u = gm::ld_cg(&c->n_viol_pp[par]);
        }
        if (lead) {
            t1 = global_ns();
            t_sweep += t1 - t0;
            if (round < DBG_ROUNDS) { c->dbg[round][1] = t1; c->dbg[round][2] = t1; }
            c->n_viol_pp[par ^ 1u] = 0;
        }
        if (incremental)                                 // the first-visit bits of this round: nobody reads them before the next one
            for (uint32_t i = first; i < visited_words; i += stride) ip.visited[i] = 0u;
        if (n_u == 0xFFFFFFFFu) {                        // a peer overflowed its exchange area or never arrived: stop
            if (lead) {
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
        bar.sync();                                      // new assignment visible to every SM before it is staged again
        prev_n_u = n_u;
        if (lead) t_mis += global_ns() - t1;
    }
    if (lead) {
        c->t_sweep_ns = t_sweep;
        c->t_mis_ns = t_mis;
    }
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
        const uint32_t vbase = cur.b * bucket_vars;
        const uint32_t slot0 = tile * TILE + threadIdx.x * CLAUSES_PER_THREAD;
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

// Variable-width fallback (general DIMACS input): one clause per thread over CSR, assignment words
// gathered through L1/L2.  Not the roofline-graded path.
__global__ void __launch_bounds__(256) sweep_csr_kernel(const uint64_t *__restrict__ off, const uint32_t *__restrict__ lit,
                                                         uint64_t m, const uint32_t *__restrict__ bits,
                                                         uint32_t *viol, Counters *ctr)
{
    if (__ldcg(&ctr->done)) return;
    const uint32_t lane = threadIdx.x & 31u;                       // launched with 8 * WBUF words of dynamic smem
    WarpCompactor comp{(threadIdx.x >> 5) * WBUF, viol, ctr, &ctr->n_viol, 0u, false, 0u, lane, nullptr};
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
            if (violated) g_smem[comp.wbuf + comp.count + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)c;
            comp.count += __popc(bal);
            if (comp.count >= 32) comp.flush();
        }
    }
    if (comp.count) comp.flush();
}

// ---- launchers ------------------------------------------------------------------------

namespace {
enum Op { OP_LAUNCH, OP_CONFIGURE, OP_PERSIST_LAUNCH, OP_PERSIST_CONFIGURE };
struct PersistArgs {
    const MisParams *mp;
    uint32_t max_rounds, epoch;
    const IncrParams *ip;
    uint32_t visited_words;
    int *max_ctas_per_sm;     // OP_PERSIST_CONFIGURE: occupancy of the persistent kernel with the requested shared memory
};
} // namespace

template <int K, int RB, int RC, int E>
static cudaError_t launch_planes_e(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, Op op, const PersistArgs *pa)
{
    switch (op) {
    case OP_CONFIGURE:   // function attributes are per device: the handle configures its kernel once at upload
        return cudaFuncSetAttribute(sweep_planes_kernel<K, RB, RC, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    case OP_LAUNCH:
        sweep_planes_kernel<K, RB, RC, E><<<grid, SWEEP_THREADS, smem, s>>>(p);
        return cudaGetLastError();
    case OP_PERSIST_CONFIGURE: {
        cudaError_t e = cudaFuncSetAttribute(solve_persistent_kernel<K, RB, RC, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(pa->max_ctas_per_sm, solve_persistent_kernel<K, RB, RC, E>,
                                                             SWEEP_THREADS, smem);
    }
    case OP_PERSIST_LAUNCH: {
        uint32_t max_rounds = pa->max_rounds, epoch = pa->epoch, visited_words = pa->visited_words;
        void *args[] = {(void *)&p, (void *)pa->mp, (void *)&max_rounds, (void *)&epoch, (void *)pa->ip, (void *)&visited_words};
        return cudaLaunchCooperativeKernel((const void *)solve_persistent_kernel<K, RB, RC, E>, dim3(grid), dim3(SWEEP_THREADS),
                                           args, smem, s);
    }
    }
    return cudaErrorInvalidValue;
}

// E = min(K, EAGER_PLANES) planes are streamed (4 / 5 / 6 / 8 were measured at k = 8: 5 is fastest, profiles/).
template <int K, int RB, int RC>
static cudaError_t launch_planes(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, Op op, const PersistArgs *pa)
{
    constexpr int E = K < (int)EAGER_PLANES ? K : (int)EAGER_PLANES;
    return launch_planes_e<K, RB, RC, E>(p, grid, smem, s, op, pa);
}

template <bool R>
static cudaError_t launch_generic(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, Op op)
{
    if (op == OP_CONFIGURE)
        return cudaFuncSetAttribute(sweep_planes_generic_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (op != OP_LAUNCH) return cudaErrorNotSupported;      // no persistent kernel for k > 8
    sweep_planes_generic_kernel<R><<<grid, SWEEP_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

// resident_all: every plane is resident-only (RB = RC = K).  Otherwise RC = min(K, RESIDENT_CAP) and
// RB = min(p.min_resident, 2, RC) as measured by the upload pass.
template <int K, int RC>
static cudaError_t dispatch_rb(const SweepParams &p, uint32_t grid, size_t smem, cudaStream_t s, Op op, const PersistArgs *pa)
{
    const uint32_t rb = p.min_resident < 2u ? p.min_resident : 2u;
    if (rb >= 2 && RC >= 2) return launch_planes<K, (RC < 2 ? RC : 2), RC>(p, grid, smem, s, op, pa);
    if (rb >= 1 && RC >= 1) return launch_planes<K, (RC < 1 ? RC : 1), RC>(p, grid, smem, s, op, pa);
    return launch_planes<K, 0, RC>(p, grid, smem, s, op, pa);
}

template <int K>
static cudaError_t dispatch_class(const SweepParams &p, bool resident_all, uint32_t grid, size_t smem, cudaStream_t s, Op op,
                                  const PersistArgs *pa)
{
    if (resident_all) return launch_planes<K, K, K>(p, grid, smem, s, op, pa);
    // RC = min(K, RESIDENT_CAP): 2 / 3 / 4 resident-placed literals were measured at k = 8, 3 is fastest (profiles/)
    constexpr int RC_DEFAULT = K < (int)RESIDENT_CAP ? K : (int)RESIDENT_CAP;
    return dispatch_rb<K, RC_DEFAULT>(p, grid, smem, s, op, pa);
}

static cudaError_t dispatch_k(const SweepParams &p, bool resident_all, uint32_t grid, size_t smem, cudaStream_t s, Op op,
                              const PersistArgs *pa = nullptr)
{
    switch (p.k) {
    case 1: return dispatch_class<1>(p, resident_all, grid, smem, s, op, pa);
    case 2: return dispatch_class<2>(p, resident_all, grid, smem, s, op, pa);
    case 3: return dispatch_class<3>(p, resident_all, grid, smem, s, op, pa);
    case 4: return dispatch_class<4>(p, resident_all, grid, smem, s, op, pa);
    case 5: return dispatch_class<5>(p, resident_all, grid, smem, s, op, pa);
    case 6: return dispatch_class<6>(p, resident_all, grid, smem, s, op, pa);
    case 7: return dispatch_class<7>(p, resident_all, grid, smem, s, op, pa);
    case 8: return dispatch_class<8>(p, resident_all, grid, smem, s, op, pa);
    default: return resident_all ? launch_generic<true>(p, grid, smem, s, op) : launch_generic<false>(p, grid, smem, s, op);
    }
}

size_t sweep_planes_smem_bytes(uint32_t bucket_words)
{
    return (size_t)bucket_words * 4 + (SWEEP_THREADS / 32) * (WBUF + QBUF) * 4;   // bits | violated staging | parked queues
}

cudaError_t configure_sweep_planes(const SweepParams &p, bool resident_all)
{
    const size_t smem = sweep_planes_smem_bytes(p.bucket_words);
    return dispatch_k(p, resident_all, 0, smem, 0, OP_CONFIGURE);
}

cudaError_t launch_sweep_planes(const SweepParams &p, bool resident_all, uint32_t grid, cudaStream_t s)
{
    const size_t smem = sweep_planes_smem_bytes(p.bucket_words);
    return dispatch_k(p, resident_all, grid, smem, s, OP_LAUNCH);
}

// ---- persistent solve kernel: shared memory = the sweep's, or what the independent-set phases need if that is more
static size_t persistent_smem_bytes(uint32_t bucket_words, uint32_t kmax)
{
    const size_t small_words = mis_small_words(SWEEP_THREADS, kmax);
    const size_t one_item = (size_t)SWEEP_THREADS * mis_cache_words(kmax);
    size_t b = sweep_planes_smem_bytes(bucket_words);
    if (small_words * 4 <= 200u * 1024u) b = b > small_words * 4 ? b : small_words * 4;
    else if (one_item * 4 <= 200u * 1024u) b = b > one_item * 4 ? b : one_item * 4;
    return b;
}

static void persistent_fill(const SweepParams &p, MisParams &mp, size_t smem)
{
    mp.cache_items = (uint32_t)((smem / 4 / SWEEP_THREADS) / mis_cache_words(mp.kmax));
    mp.small_ok = mis_small_words(SWEEP_THREADS, mp.kmax) * 4 <= smem ? 1u : 0u;
    (void)p;
}

// ok_out: 1 when the instance can be solved by the persistent kernel on this device (k <= 8, one CTA per SM fits)
cudaError_t configure_solve_persistent(const SweepParams &p, bool resident_all, uint32_t kmax, int *ok_out)
{
    *ok_out = 0;
    if (p.k == 0 || p.k > 8) return cudaSuccess;
    int per_sm = 0;
    PersistArgs pa{nullptr, 0u, 0u, nullptr, 0u, &per_sm};
    const cudaError_t e = dispatch_k(p, resident_all, 0, persistent_smem_bytes(p.bucket_words, kmax), 0, OP_PERSIST_CONFIGURE, &pa);
    if (e != cudaSuccess) return e;
    *ok_out = per_sm >= 1;
    return cudaSuccess;
}

cudaError_t launch_solve_persistent(const SweepParams &p, bool resident_all, uint32_t grid, const ClauseView &cv, uint32_t kmax,
                                    uint8_t *state, uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint64_t seed,
                                    uint32_t max_rounds, uint32_t epoch, const IncrParams *incr, uint32_t visited_words,
                                    uint32_t incr_max_vars, cudaStream_t s)
{
    const size_t smem = persistent_smem_bytes(p.bucket_words, kmax);
    MisParams mp{};
    mp.cv = cv; mp.viol = p.p2p ? nullptr : p.viol; mp.state = state; mp.s_slots = s_slots;
    mp.p2p = p.p2p;                                    // sharded: U = the record blocks in our exchange region
    mp.claim = sc.claim;
    mp.n_vars = n_vars; mp.bits = const_cast<uint32_t *>(p.bits); mp.ctr = p.ctr; mp.seed = seed; mp.kmax = kmax;
    mp.urec = sc.urec; mp.urec_cap = sc.urec_cap;
    persistent_fill(p, mp, smem);
    mp.incr_max_vars = incr ? incr_max_vars : 0u;
    const IncrParams no_incr{};
    PersistArgs pa{&mp, max_rounds, epoch, incr ? incr : &no_incr, incr ? visited_words : 0u, nullptr};
    return dispatch_k(p, resident_all, grid, smem, s, OP_PERSIST_LAUNCH, &pa);
}

cudaError_t launch_sweep_csr(const uint64_t *off, const uint32_t *lit, uint64_t m, const uint32_t *bits,
                             uint32_t *viol, Counters *ctr, uint32_t grid, cudaStream_t s)
{
    sweep_csr_kernel<<<grid, 256, 8 * WBUF * 4, s>>>(off, lit, m, bits, viol, ctr);
    return cudaGetLastError();
}

} // namespace alll
