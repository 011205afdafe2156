// csr_body.cuh -- K1 + K2 over a CSR literal array of arbitrary clause widths: the warp-cooperative sweep.
//
// Replaces Clause::is_not_satisfied over a ClauseArray of mixed widths (Clause.h:20-46; example/main.cpp:157-178 builds
// clauses of whatever width the DIMACS file holds) for input that is NOT padded onto the plane layout
// (ALLL_FLAG_FORCE_CSR, clauses wider than 32 literals, or padding that would more than double the literal count).
//
// Layout (built once at upload, csr.cu):
//   lit[l_pad]          the caller's literal array, padded to a multiple of 128 with at least one padding position
//   start[l_pad / 32]   bit p: position p holds the first literal of a clause (padding positions are all starts);
//                       lane i's four bits are nibble i & 7 of word i >> 3 of the chunk's four words
//   chunk_rank[c]       number of clause starts before position 128 * c  (clause id of a start = its rank)
// A warp streams the literal array in chunks of 128 consecutive literals -- one 128-bit load per lane, fully coalesced
// whatever the clause widths are -- plus the 16 bytes of start bits of the chunk.  Nothing else is read: off[] is not
// touched by the sweep.  Clause boundaries are resolved in registers:
//   * every lane knows, from the start bits, how far each of its 4 literals is from the beginning of its clause;
//   * lookups are predicated and branch-free; shared-memory lookups (STAGED) are done for every literal, L2 lookups are
//     lazy in two phases: the first three literals of every clause, then the later ones only where the first three were
//     all false (lanes exchange their truth nibbles with two shuffles) -- 3.3 instead of 5.5 L2 sectors per clause on
//     widths 3..8;
//   * a clause is violated iff no literal of its segment is true: segments inside one lane are decided there, segments
//     spanning lanes by two ballots (lanes holding a start / lanes with a true literal in the part that belongs to the
//     segment entering them) and a bit-range test; a segment still open at the end of a chunk is carried (warp-uniform
//     state) into the next chunk, across any number of chunks (clauses wider than 128 literals);
//   * a warp owns the clauses that START in its contiguous range of chunks: it reads on past the end of its range until
//     its last clause is closed, and skips the literals at the beginning of its range that belong to its predecessor.
// Violated clause ids are compacted per warp (ballot / popc, WarpCompactor of sweep_body.cuh).
// The assignment is staged in shared memory when it fits (STAGED), otherwise looked up through L2 (coherent loads).
#pragma once

#include "sweep_body.cuh"

namespace alll {

constexpr uint32_t CSR_CHUNK = 128;          // literals per warp step: 32 lanes x one 128-bit load
constexpr uint32_t CSR_PREFETCH = 8;         // chunks of L2 prefetch distance per warp (4 KB ahead of the loads)

struct CsrSweepParams {
    const uint32_t *lit;         // [l_pad]
    const uint32_t *start;       // [l_pad / 32]
    const uint32_t *chunk_rank;  // [l_pad / 128 + 1]
    uint32_t n_chunks;
    uint32_t m;
    const uint32_t *bits;
    uint32_t n_words;            // assignment words (padded to a multiple of 4)
    uint32_t staged_words;       // == n_words when the assignment is staged in shared memory, 0 = lookups through L2
    uint32_t *viol;              // out: violated clause ids
    Counters *ctr;
};

// One assignment lookup, predicated and branch-free (a dead lane issues no request): 1 iff literal l is TRUE.
template <bool STAGED>
__device__ __forceinline__ uint32_t csr_lookup(uint32_t l, bool go, const uint32_t *gbits, uint32_t smem_base)
{
    const uint32_t w = go ? (STAGED ? lds32(smem_base + ((l >> 6) << 2)) : ld_bits(gbits + (l >> 6))) : 0u;
    return go ? ((__funnelshift_r(w, 0u, l >> 1) ^ l) & 1u) : 0u;     // bit (var & 31) of w, xor the negation flag
}

// Out of line: runs once per >= 32 violated clauses and must not bloat the streaming loop.
static __device__ __noinline__ void csr_flush(uint32_t *viol, unsigned int *n_viol, uint32_t wbuf, uint32_t count, uint32_t lane)
{
    __syncwarp();
    unsigned int g = 0;
    if (lane == 0) g = atomicAdd(n_viol, count);
    g = __shfl_sync(0xffffffffu, g, 0);
    for (uint32_t i = lane; i < count; i += 32) viol[g + i] = g_smem[wbuf + i];
    __syncwarp();
}

template <bool STAGED>
__device__ __forceinline__ void sweep_csr_body(const CsrSweepParams &p, unsigned int *n_viol_ctr)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    if (STAGED) {
        __syncthreads();                                              // nobody still uses the shared memory of the previous phase
        const uint4 *src = reinterpret_cast<const uint4 *>(p.bits);
        for (uint32_t i = threadIdx.x; i < p.staged_words / 4; i += blockDim.x) reinterpret_cast<uint4 *>(g_smem)[i] = __ldcg(src + i);
        __syncthreads();
    }
    uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(g_smem);
    asm volatile("" : "+r"(smem_base));                               // lookups below stay behind the staging barrier
    const uint32_t wbuf = p.staged_words + warp * WBUF;               // this warp's violated-id staging (WBUF entries)
    uint32_t n_out = 0;                                               // warp-uniform
    auto push = [&](bool mine, uint32_t id) {                         // whole warp; compaction by ballot / popc
        const uint32_t bal = __ballot_sync(0xffffffffu, mine);
        if (mine) g_smem[wbuf + n_out + __popc(bal & ((1u << lane) - 1u))] = id;
        n_out += __popc(bal);
        if (n_out >= 32) { csr_flush(p.viol, n_viol_ctr, wbuf, n_out, lane); n_out = 0; }
    };

    const uint32_t n_warps = gridDim.x * wpc, w = blockIdx.x * wpc + warp;
    const uint32_t c0 = (uint32_t)(((uint64_t)p.n_chunks * w) / n_warps);
    const uint32_t c1 = (uint32_t)(((uint64_t)p.n_chunks * (w + 1)) / n_warps);
    // the clause entering the current chunk from the left (warp-uniform): valid only if it started in OUR range
    bool carry_valid = false, carry_sat = false;
    uint32_t carry_id = 0, carry_len = 0;
    const uint32_t pos = lane * 4;                                    // my first position inside a chunk
    const uint32_t lt = (1u << lane) - 1u;                            // lanes below me
    const uint32_t gt = ~((2u << lane) - 1u);                         // lanes above me
    // Bytes in flight: the next chunk's literals and start bits are loaded before the current chunk is evaluated
    // (register double buffer), and one lane bulk-prefetches the chunk CSR_PREFETCH steps ahead into L2 (TMA prefetch,
    // no register cost) -- one chunk per warp in flight would cap the stream at a fraction of the HBM rate.
    uint4 Ln = make_uint4(0u, 0u, 0u, 0u);
    uint32_t fn = 0;                                                  // the 32-bit start word holding my nibble
    const uint32_t *my_start = p.start + (lane >> 3);
    if (c0 < c1) {
        Ln = ld_stream_v4(p.lit + (uint64_t)c0 * CSR_CHUNK + pos);
        fn = __ldg(my_start + (uint64_t)c0 * 4);
        if (lane == 0)
            for (uint32_t a = 1; a < CSR_PREFETCH && c0 + a < c1; a++) tma_prefetch_l2(p.lit + (uint64_t)(c0 + a) * CSR_CHUNK, CSR_CHUNK * 4);
    }
    for (uint32_t c = c0; c < p.n_chunks; ++c) {
        const bool finishing = c >= c1;                               // beyond our range: only close the clause we still hold
        if (finishing && !carry_valid) break;
        const uint4 L = Ln;
        const uint32_t fword = fn;
        if (c + 1 < p.n_chunks) {
            Ln = ld_stream_v4(p.lit + (uint64_t)(c + 1) * CSR_CHUNK + pos);
            fn = __ldg(my_start + (uint64_t)(c + 1) * 4);
        }
        if (lane == 0 && c + CSR_PREFETCH < c1) tma_prefetch_l2(p.lit + (uint64_t)(c + CSR_PREFETCH) * CSR_CHUNK, CSR_CHUNK * 4);
        const uint32_t lits[4] = {L.x, L.y, L.z, L.w};
        const uint32_t f = (fword >> ((lane & 7u) * 4u)) & 0xFu;      // start bits of my 4 positions
        const bool has_start = f != 0u;
        const uint32_t H = __ballot_sync(0xffffffffu, has_start);     // lanes holding a clause start
        const uint32_t first_q = has_start ? (uint32_t)__ffs((int)f) - 1u : 4u;
        const uint32_t last_q = has_start ? 31u - (uint32_t)__clz((int)f) : 0u;
        const bool none_below = (H & lt) == 0u;                       // no start before my lane: my first literals belong to the entering clause
        const uint32_t head_m = none_below ? (1u << first_q) - 1u : 0u;   // my positions that belong to the entering clause
        // which literals take part at all: not the ones of a predecessor's clause, and in finishing mode only the head
        const uint32_t act = finishing ? head_m : (carry_valid ? 0xFu : 0xFu & ~head_m);
        uint32_t t = 0;                                               // truth nibble of my literals
        if (STAGED) {
            // shared-memory lookups are cheap: every literal that takes part is looked up
#pragma unroll
            for (int q = 0; q < 4; q++) t |= csr_lookup<true>(lits[q], (act >> q) & 1u, p.bits, smem_base) << q;
        } else {
            // L2 lookups cost a sector each: lazy in two phases.  d = distance of a literal from the first literal of its clause
            uint32_t d[4];
            {
                const uint32_t j = none_below ? 0u : 31u - (uint32_t)__clz((int)(H & lt));  // nearest lane below me holding a start
                const uint32_t lq_j = __shfl_sync(0xffffffffu, last_q, j);                   // (whole warp: no shuffle under divergence)
                uint32_t dprev = none_below ? pos + carry_len - 1u : pos - (4u * j + lq_j) - 1u;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    d[q] = ((f >> q) & 1u) ? 0u : dprev + 1u;
                    dprev = d[q];
                }
            }
            // phase A: the first three literals of every clause
#pragma unroll
            for (int q = 0; q < 4; q++) t |= csr_lookup<false>(lits[q], ((act >> q) & 1u) && d[q] < 3u, p.bits, smem_base) << q;
            // phase B: later literals, only where the clause's first three were all false (truth nibbles of the two lanes below)
            const uint32_t p1 = __shfl_up_sync(0xffffffffu, t, 1), p2 = __shfl_up_sync(0xffffffffu, t, 2);
            const uint32_t win = (t << 8) | ((lane >= 1 ? p1 : 0u) << 4) | (lane >= 2 ? p2 : 0u);   // bit b <-> position pos - 8 + b
            uint32_t tb = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                bool go = ((act >> q) & 1u) && d[q] >= 3u;
                // first three literals of my clause inside the window (clause began <= 7 positions ago, in this chunk)?
                const bool in_win = d[q] <= 7u && pos + q >= d[q];
                go = go && (!in_win || ((win >> ((8 + q - d[q]) & 31u)) & 7u) == 0u);
                tb |= csr_lookup<false>(lits[q], go, p.bits, smem_base) << q;
            }
            t |= tb;
        }
        // ---- segments: inside my lane, then across lanes
        const bool head_sat = (t & ((1u << first_q) - 1u)) != 0u;                       // my part of the clause entering my lane (all 4 without a start)
        const bool tail_sat = has_start && (t >> last_q) != 0u;                          // from my last start to the end of my lane
        const uint32_t B = __ballot_sync(0xffffffffu, head_sat);
        // the clause entering the chunk closes at the first start of the chunk
        bool emit_carry = false;
        if (H != 0u) {
            const uint32_t j0 = (uint32_t)__ffs((int)H) - 1u;
            emit_carry = carry_valid && !(carry_sat || (B & ((2u << j0) - 1u)) != 0u) && carry_id < p.m;
        } else {
            carry_sat = carry_sat || B != 0u;
            carry_len += CSR_CHUNK;
        }
        if (finishing) {
            if (emit_carry) push(lane == 0, carry_id);
            if (H != 0u) break;                                                          // our last clause is closed
            continue;
        }
        if (H != 0u) {
            // my starts: all but the last close inside my lane; the last one closes at the next lane holding a start
            const uint32_t above = H & gt;
            uint32_t vmask = 0;                                                          // bit q: the clause starting at my position q is violated
            bool open_sat = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t later = (f >> (q + 1)) & 0x7u;
                const uint32_t qn = (uint32_t)q + (uint32_t)__ffs((int)later);           // next start in my lane (if any)
                const bool v_in = (t & (((1u << qn) - 1u) & ~((1u << q) - 1u))) == 0u;
                const bool mine = ((f >> q) & 1u) && later != 0u && v_in;
                vmask |= mine ? (1u << q) : 0u;
            }
            if (has_start) {
                if (above) {
                    const uint32_t j = (uint32_t)__ffs((int)above) - 1u;                 // next lane holding a start
                    const uint32_t range = ((2u << j) - 1u) & gt;                        // lanes (lane, j]
                    if (!(tail_sat || (B & range) != 0u)) vmask |= 1u << last_q;
                } else {
                    open_sat = tail_sat || (B & gt) != 0u;                               // still open at the end of the chunk
                }
            }
            // emission (rare: a 2^-width fraction of the clauses): clause id of a start = its rank
            if (__any_sync(0xffffffffu, vmask != 0u) || emit_carry) {
                // starts before my first position: whole start words below mine + the bits of my word below my nibble
                const uint32_t wi = lane >> 3;
                uint32_t sb = (uint32_t)__popc(fword & ((1u << ((lane & 7u) * 4u)) - 1u));
                const uint32_t pw = (uint32_t)__popc(fword);
                const uint32_t w0 = __shfl_sync(0xffffffffu, pw, 0), w1 = __shfl_sync(0xffffffffu, pw, 8), w2 = __shfl_sync(0xffffffffu, pw, 16);
                sb += (wi > 0 ? w0 : 0u) + (wi > 1 ? w1 : 0u) + (wi > 2 ? w2 : 0u);
                const uint32_t rank0 = __ldg(p.chunk_rank + c) + sb;
                if (emit_carry) push(lane == 0, carry_id);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t id = rank0 + (uint32_t)__popc(f & ((1u << q) - 1u));
                    const bool mine = ((vmask >> q) & 1u) && id < p.m;
                    if (__any_sync(0xffffffffu, mine)) push(mine, id);
                }
                // new carry below needs the id of the last start of the highest lane with a start
                const uint32_t hi = 31u - (uint32_t)__clz((int)H);
                carry_id = __shfl_sync(0xffffffffu, rank0 + (uint32_t)__popc(f) - 1u, hi);
            } else {
                // nobody emits: the carried id is still needed later -- computed by the highest lane with a start
                const uint32_t hi = 31u - (uint32_t)__clz((int)H);
                const uint32_t wi = hi >> 3;
                const uint32_t pw = (uint32_t)__popc(fword);
                const uint32_t w0 = __shfl_sync(0xffffffffu, pw, 0), w1 = __shfl_sync(0xffffffffu, pw, 8), w2 = __shfl_sync(0xffffffffu, pw, 16);
                const uint32_t fw_hi = __shfl_sync(0xffffffffu, fword, hi);
                // starts at or below the last start of lane hi = starts in the words below + bits of its word up to its nibble's top start
                const uint32_t upto = (hi & 7u) * 4u + __shfl_sync(0xffffffffu, last_q, hi);      // bit index of that start in its word
                const uint32_t in_word = (uint32_t)__popc(fw_hi & ((2u << upto) - 1u));
                carry_id = __ldg(p.chunk_rank + c) + (wi > 0 ? w0 : 0u) + (wi > 1 ? w1 : 0u) + (wi > 2 ? w2 : 0u) + in_word - 1u;
            }
            const uint32_t hi = 31u - (uint32_t)__clz((int)H);
            carry_valid = true;
            carry_sat = __shfl_sync(0xffffffffu, (uint32_t)open_sat, hi) != 0u;
            carry_len = CSR_CHUNK - (4u * hi + __shfl_sync(0xffffffffu, last_q, hi));
        }
    }
    if (n_out) csr_flush(p.viol, n_viol_ctr, wbuf, n_out, lane);
}

// shared memory of a CSR sweep: staged assignment | violated staging
static inline size_t sweep_csr_smem_bytes(uint32_t staged_words, uint32_t threads)
{
    return (size_t)staged_words * 4 + (threads / 32) * WBUF * 4;
}

} // namespace alll
