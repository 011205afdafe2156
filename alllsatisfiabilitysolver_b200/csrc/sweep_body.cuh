// sweep_body.cuh -- device code of the clause-evaluation sweep (K1 + K2; see sweep.cu for the design notes) and the
// dispatch over its compile-time variants, shared by the stand-alone sweep kernels (sweep.cu) and the persistent solve
// kernel (persist.cu).  Two translation units so that the ~60 kernel instantiations compile in parallel.
#pragma once

#include "alll_device.cuh"

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
// link != NULL (sharded P2P mode): the same records go into the receive slot of this rank and round in EVERY GPU's
// exchange region (urec is then the offset of that slot in words): consecutive lanes store consecutive words, so each
// peer gets whole 128-byte NVLink writes instead of one 4-byte packet per word.
__device__ __noinline__ void write_records(uint32_t *__restrict__ urec, const uint32_t *__restrict__ planes, uint64_t m_pad,
                                           const uint32_t *__restrict__ orig_id, uint32_t id_base, uint32_t k, uint32_t wbuf,
                                           uint32_t g, uint32_t count, uint32_t lane, const P2PLink *link = nullptr, uint64_t slot_base = 0)
{
    // One (clause, word) pair per lane and pass; k <= 8, count <= 63 => at most 18 passes.  All scattered reads are issued
    // before the first store (one DRAM round trip per flush instead of one per pass), the stores of a pass are consecutive.
    const uint32_t w = k + 1, total = count * w;
    uint32_t *out = link ? nullptr : urec + (uint64_t)g * w;
    const uint32_t n_dst = link ? link->world : 1u;
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
        for (uint32_t d = 0; d < n_dst; d++) {
            if (link) out = link->rec[d] + slot_base + (uint64_t)g * w;
#pragma unroll
            for (int q = 0; q < 9; q++) {
                const uint32_t t = t0 + q * 32 + lane;
                if (t < total) out[t] = val[q];
            }
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
            if (lane == 0) g_remote_dirty = 1u;
            if ((uint64_t)g + count > L.cap) {
                if (lane == 0) { ctr->p2p_error = 1; for (uint32_t q = 0; q < L.world; q++) L.hdr[q]->abort = sp->p2p_epoch + 1u; }
            } else {
                const uint64_t slot_base = ((uint64_t)p2p_parity * L.world + L.rank) * L.cap * (L.k + 1);
                write_records(nullptr, sp->planes, sp->m_pad, sp->orig_id, sp->id_base, sp->k, wbuf, g, count, lane, sp->p2p, slot_base);
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
    const unsigned long long word = ((unsigned long long)p.p2p_tag << 32) | __ldcg(&p.ctr->n_viol);
    for (uint32_t q = 0; q < L.world; q++) *(volatile unsigned long long *)&L.hdr[q]->cf[p.p2p_parity][L.rank] = word;
}

// Assignment words are read with ld.global.cg (L2, coherent), never through the non-coherent path (__ldg / ld.global.nc):
// inside solve_persistent_kernel the same launch rewrites `bits` every round (resample_var: red.or / red.and), and .nc
// is only defined for data that is read-only for the whole kernel -- a stale line would silently drop a violated clause.
// The immutable literal planes, orig_id and segs keep the .nc path.  (L1 would not help these lookups anyway: the
// staged assignment leaves ~30 KB of L1 against a 1.25 MB array.)
__device__ __forceinline__ uint32_t ld_bits(const uint32_t *p) { return __ldcg(p); }

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
        w = (rel < bucket_vars) ? g_smem[rel >> 5] : ld_bits(gbits + (v >> 5));
    }
    return ((w >> (v & 31u)) ^ l) & 1u;
}

} // namespace

// ---- per-tile bookkeeping shared by both plane kernels ---------------------------------------------
// Stages a bucket's slice of the assignment into shared memory (whole CTA).
__device__ __forceinline__ void stage_bucket(const SweepParams &p, uint32_t bucket)
{
    __syncthreads();                      // everyone is done with the previous bucket's bits
    const uint4 *src = reinterpret_cast<const uint4 *>(p.bits + (uint64_t)bucket * p.bucket_words);
    for (uint32_t i = threadIdx.x; i < p.bucket_words / 4; i += SWEEP_THREADS)
        reinterpret_cast<uint4 *>(g_smem)[i] = __ldcg(src + i);      // coherent (L2) read: see ld_bits
    __syncthreads();
}

// Tile-by-tile walk over the sweep order (the run-time-width kernel; the k <= 8 kernels work on SweepRun lists instead).
struct TileCursor {
    uint32_t b, bucket, bucket_tile_end, slot_end, loaded;    // b: segment index (sweep order); bucket / loaded: variable-range bucket ids
    uint32_t phys, delta;                                     // the current tile's number in slot space; slot space - sweep order within the segment

    __device__ __forceinline__ void read(const SweepParams &p)
    {
        bucket_tile_end = (b + 1 < p.n_segs) ? p.segs[b + 1].tile_begin : p.n_tiles;
        slot_end = p.segs[b].slot_end;
        bucket = p.segs[b].bucket;
        delta = p.segs[b].phys_tile - p.segs[b].tile_begin;
    }
    __device__ __forceinline__ void init(const SweepParams &p, uint32_t t0)
    {
        loaded = 0xFFFFFFFFu;
        phys = t0;
        b = find_segment(p.segs, p.n_segs, t0);
        read(p);
    }
    // Moves to `tile` of the sweep order; returns true when its bucket differs from the staged one (caller must then stage()).
    __device__ __forceinline__ bool advance(const SweepParams &p, uint32_t tile)
    {
        while (tile >= bucket_tile_end) {
            ++b;
            read(p);
        }
        phys = tile + delta;
        return bucket != loaded;
    }
    __device__ __forceinline__ void stage(const SweepParams &p)
    {
        stage_bucket(p, bucket);
        loaded = bucket;
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
    w = go ? ld_bits(gbits + (v >> 5)) : 0u;
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
// sadj_rb: base for planes [0, RB) -- the same as sadj, or (packed eager planes: those literals are stored relative to
// the bucket) the shared byte address of staged word 0 itself.
template <int E, int RB, int RC>
__device__ __forceinline__ uint32_t eval4(const uint4 (&L)[E], uint32_t valid_mask, uint32_t sadj_rb, uint32_t sadj,
                                          const uint32_t *gbits, uint32_t vbase, uint32_t bucket_vars)
{
    constexpr int R_END = RC < E ? RC : E;
    uint32_t a[4] = {valid_mask & 1u, valid_mask & 2u, valid_mask & 4u, valid_mask & 8u};
#pragma unroll
    for (int j = 0; j < R_END; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (j < RB) resident_only_step(comp(L[j], q), a[q], sadj_rb);
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
            if (E >= 4 && K <= 8 && p.rows8 != nullptr) {     // the second half of the clause's row (literals 4..7) instead of K - E scattered plane words
                const uint4 t = act ? __ldg(p.rows8 + 2 * (uint64_t)slot + 1) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < K - E; j++) l[j] = comp(t, E - 4 + j);
            } else {
#pragma unroll
                for (int j = 0; j < K - E; j++) l[j] = act ? __ldg(p.planes + (uint64_t)(E + j) * p.m_pad + slot) : 0u;
            }
            uint32_t w[T];
#pragma unroll
            for (int j = 0; j < K - E; j++) {                 // all lookups at once: this path is rare and dense
                const uint32_t v = l[j] >> 1;
                if (RESIDENT_ALL) w[j] = act ? g_smem[v >> 5] : 0u;
                else {
                    const uint32_t rel = v - vbase;
                    w[j] = !act ? 0u : (rel < bucket_vars) ? g_smem[rel >> 5] : ld_bits(p.bits + (v >> 5));
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
// PK: the E = 5 eager literals come from the four packed planes (EagerPack<RB>) instead of planes 0..4.
template <int K, int RB, int RC, int E, bool TICKET, bool PK>
__device__ __forceinline__ void sweep_planes_body(const SweepParams &p, unsigned int *n_viol_ctr, uint32_t p2p_parity, bool rec_on)
{
    constexpr bool RESIDENT_ALL = RB >= K;
    constexpr int RBE = RB < E ? RB : E;
    constexpr int NS = PK ? 4 : E;               // planes streamed per tile
    static_assert(!PK || (E == 5 && K >= 5 && (RB == 1 || RB == 2)), "packed eager planes: five literals, one or two of them bucket-relative");
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    WarpCompactor out{p.bucket_words + warp * WBUF, p.viol, p.ctr, n_viol_ctr, p2p_parity, rec_on, 0u, lane, &p};
    SurvivorQueue<K, E, RESIDENT_ALL> parked{p.bucket_words + (SWEEP_THREADS / 32) * WBUF + warp * QBUF, 0u, lane};

    // this CTA's runs (SweepRun): consecutive tiles of one bucket segment each
    const uint32_t r0 = __ldg(p.run_begin + blockIdx.x), r1 = __ldg(p.run_begin + blockIdx.x + 1);
    const uint32_t bucket_vars = p.bucket_words * 32u;
    const uint32_t *const stream = PK ? p.packed : p.planes;
    const uint32_t *base = stream + threadIdx.x * CLAUSES_PER_THREAD;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(g_smem);
    const uint32_t dist = p.prefetch_tiles;
    uint32_t loaded = 0xFFFFFFFFu;               // bucket whose slice of the assignment is staged
    uint32_t vbase = 0;

    for (uint32_t r = r0; r < r1; r++) {
        const uint4 run = __ldg(reinterpret_cast<const uint4 *>(p.runs) + r);     // {tile_begin, tile_end, slot_end, bucket}
        const uint32_t t0 = run.x, t1 = run.y, slot_end = run.z;
        if (run.w != loaded) {
            if constexpr (E < K) parked.drain(0u, out, p, vbase, bucket_vars);    // parked clauses belong to the old bucket
            stage_bucket(p, run.w);
            loaded = run.w;
            vbase = run.w * bucket_vars;
        }

        auto load = [&](uint4 (&L)[NS], uint32_t tile) {
            const uint32_t *src = base + (uint64_t)tile * TILE;
#pragma unroll
            for (int j = 0; j < NS; j++) L[j] = ld_stream_v4(src + (uint64_t)j * p.m_pad);
        };
        auto process = [&](const uint4 (&S)[NS], uint32_t tile) {
            const uint32_t slot0 = tile * TILE + threadIdx.x * CLAUSES_PER_THREAD;
            uint32_t valid = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) valid |= (slot0 + q < slot_end) ? (1u << q) : 0u;
            uint32_t sb = smem_base;
            asm volatile("" : "+r"(sb));          // opaque: lookups below cannot be hoisted above the staging barrier
            const uint32_t sadj = sb - ((vbase >> 5) << 2);
            uint32_t alive;
            if constexpr (PK) {
                using P = EagerPack<(RB < 2 ? 1 : 2)>;
                uint4 L[E];
#define ALLL_UNPACK(J) L[J] = make_uint4(P::template field<J>(S[0].x, S[1].x, S[2].x, S[3].x), P::template field<J>(S[0].y, S[1].y, S[2].y, S[3].y), \
                                         P::template field<J>(S[0].z, S[1].z, S[2].z, S[3].z), P::template field<J>(S[0].w, S[1].w, S[2].w, S[3].w))
                ALLL_UNPACK(0); ALLL_UNPACK(1); ALLL_UNPACK(2); ALLL_UNPACK(3); ALLL_UNPACK(4);
#undef ALLL_UNPACK
                alive = eval4<E, RBE, RC>(L, valid, sb, sadj, p.bits, vbase, bucket_vars);
            } else {
                alive = eval4<E, RBE, RC>(S, valid, sadj, sadj, p.bits, vbase, bucket_vars);
            }
            if constexpr (E < K) {
                parked.push4(alive, slot0);
                parked.drain(31u, out, p, vbase, bucket_vars);
            } else {
                out.push4(alive, slot0);
            }
        };
        // HBM -> L2: one thread per CTA bulk-prefetches the plane segments of the tile `dist` ahead of the register
        // double buffer, so enough bytes are in flight to cover the loaded DRAM latency without spending registers.
        auto prefetch = [&](uint32_t tile) {
            if (threadIdx.x == 0 && dist != 0 && tile < t1) {
#pragma unroll
                for (int j = 0; j < NS; j++) tma_prefetch_l2(stream + (uint64_t)j * p.m_pad + (uint64_t)tile * TILE, TILE * 4);
            }
        };
        // (Measured and rejected for the packed planes, profiles/r02_packed_planes.md: a third register buffer -- loads
        // issued two tiles ahead -- is slower: the wait at a tile's first use is arrival rate, not latency.)
        for (uint32_t d = 2; d < 2 + dist; d++) prefetch(t0 + d);
        uint4 A[NS], B[NS];
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
    }
    if constexpr (E < K) parked.drain(0u, out, p, vbase, bucket_vars);
    if (out.count) out.flush();
    if (TICKET) p2p_publish(p);
}

// ---- dispatch over the compile-time variants ---------------------------------------------------------------------
// f.template run<K, RB, RC, E, PK>() for the variant that fits p.k / the bucket classes measured at upload; returns
// cudaErrorNotSupported for k > 8 (only sweep.cu has a run-time-width kernel).
// E = min(K, EAGER_PLANES) planes are streamed (4 / 5 / 6 / 8 were measured at k = 8: 5 is fastest, profiles/).
// resident_all: every plane is resident-only (RB = RC = K).  Otherwise RC = min(K, RESIDENT_CAP) (2 / 3 / 4 measured at
// k = 8, 3 is fastest) and RB = min(p.min_resident, 2, RC) as measured by the upload pass.
template <int K, int RB, int RC, class F>
static cudaError_t dispatch_e(F &f)
{
    constexpr int E = K < (int)EAGER_PLANES ? K : (int)EAGER_PLANES;
    return f.template run<K, RB, RC, E, false>();
}

template <int K, class F>
static cudaError_t dispatch_class(const SweepParams &p, bool resident_all, F &f)
{
    if (resident_all) return dispatch_e<K, K, K>(f);
    constexpr int RC = K < (int)RESIDENT_CAP ? K : (int)RESIDENT_CAP;
    const uint32_t rb = p.min_resident < 2u ? p.min_resident : 2u;
    if (p.packed != nullptr) {                           // packed eager planes (capi.cu decides eligibility at upload)
        if constexpr (K >= (int)EAGER_PLANES && RC == 3) {
            if (rb >= 2) return f.template run<K, 2, RC, (int)EAGER_PLANES, true>();
            if (rb == 1) return f.template run<K, 1, RC, (int)EAGER_PLANES, true>();
        }
        return cudaErrorNotSupported;
    }
    if (rb >= 2 && RC >= 2) return dispatch_e<K, (RC < 2 ? RC : 2), RC>(f);
    if (rb >= 1 && RC >= 1) return dispatch_e<K, (RC < 1 ? RC : 1), RC>(f);
    return dispatch_e<K, 0, RC>(f);
}

template <class F>
static cudaError_t dispatch_variant(const SweepParams &p, bool resident_all, F &f)
{
    switch (p.k) {
    case 1: return dispatch_class<1>(p, resident_all, f);
    case 2: return dispatch_class<2>(p, resident_all, f);
    case 3: return dispatch_class<3>(p, resident_all, f);
    case 4: return dispatch_class<4>(p, resident_all, f);
    case 5: return dispatch_class<5>(p, resident_all, f);
    case 6: return dispatch_class<6>(p, resident_all, f);
    case 7: return dispatch_class<7>(p, resident_all, f);
    case 8: return dispatch_class<8>(p, resident_all, f);
    default: return cudaErrorNotSupported;
    }
}

// shared memory of a sweep: bits | violated staging | parked queues
static inline size_t sweep_smem_bytes_for(uint32_t bucket_words)
{
    return (size_t)bucket_words * 4 + (SWEEP_THREADS / 32) * (WBUF + QBUF) * 4;
}

} // namespace alll
