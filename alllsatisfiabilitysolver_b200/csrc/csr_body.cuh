// csr_body.cuh -- K1 + K2 over a CSR literal array of arbitrary clause widths: the warp-cooperative sweep.
//
// Replaces Clause::is_not_satisfied over a ClauseArray of mixed widths (Clause.h:20-46; example/main.cpp:157-178 builds
// clauses of whatever width the DIMACS file holds) for input that is NOT padded onto the plane layout
// (ALLL_FLAG_FORCE_CSR, clauses wider than 32 literals, or padding that would more than double the literal count).
//
// Layout (built once at upload, csr.cu):
//   lit[l_pad]          the caller's literal array, padded to a multiple of 128 with at least one padding position
//   start[l_pad / 32]   bit p: position p holds the first literal of a clause (padding positions are all starts)
//   chunk_rank[c]       number of clause starts before position 128 * c  (clause id of a start = its rank)
// A warp streams the literal array in chunks of 128 consecutive literals -- one 128-bit load per lane, fully coalesced
// whatever the clause widths are -- plus the 16 bytes of start bits of the chunk.  Nothing else is read: off[] is not
// touched by the sweep.  Clause boundaries are resolved in registers:
//   * every lane knows, from the start bits, how far each of its 4 literals is from the beginning of its clause;
//   * lookups are lazy in two phases: the first three literals of every clause, then the later ones only where the first
//     three were all false (lanes exchange their truth nibbles with two shuffles) -- 3.3 instead of 5.5 lookups per clause
//     on widths 3..8; a predicated-off lookup issues no request;
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

template <bool STAGED>
__device__ __forceinline__ uint32_t csr_lookup(uint32_t l, bool go, const uint32_t *gbits, uint32_t smem_base)
{
    const uint32_t v = l >> 1;
    uint32_t w = 0;
    if (go) w = STAGED ? lds32(smem_base + ((v >> 5) << 2)) : ld_bits(gbits + (v >> 5));
    return go ? (((w >> (v & 31u)) ^ l) & 1u) : 0u;
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
    WarpCompactor out{p.staged_words + warp * WBUF, p.viol, p.ctr, n_viol_ctr, 0u, false, 0u, lane, nullptr};

    const uint32_t n_warps = gridDim.x * wpc, w = blockIdx.x * wpc + warp;
    const uint32_t c0 = (uint32_t)(((uint64_t)p.n_chunks * w) / n_warps);
    const uint32_t c1 = (uint32_t)(((uint64_t)p.n_chunks * (w + 1)) / n_warps);
    // the clause entering the current chunk from the left (warp-uniform): valid only if it started in OUR range
    bool carry_valid = false, carry_sat = false;
    uint32_t carry_id = 0, carry_len = 0;
    const uint32_t pos = lane * 4;                                    // my first position inside a chunk
    // Bytes in flight: the next chunk's literals and start bits are loaded before the current chunk is evaluated
    // (register double buffer), and one lane bulk-prefetches the chunk CSR_PREFETCH steps ahead into L2 (TMA prefetch,
    // no register cost) -- one chunk per warp in flight would cap the stream at a fraction of the HBM rate.
    uint4 Ln = make_uint4(0u, 0u, 0u, 0u), Fn = Ln;
    if (c0 < c1) {
        Ln = ld_stream_v4(p.lit + (uint64_t)c0 * CSR_CHUNK + pos);
        Fn = __ldg(reinterpret_cast<const uint4 *>(p.start) + c0);
        if (lane == 0)
            for (uint32_t a = 1; a < CSR_PREFETCH && c0 + a < c1; a++) tma_prefetch_l2(p.lit + (uint64_t)(c0 + a) * CSR_CHUNK, CSR_CHUNK * 4);
    }
    for (uint32_t c = c0; c < p.n_chunks; ++c) {
        const bool finishing = c >= c1;                               // beyond our range: only close the clause we still hold
        if (finishing && !carry_valid) break;
        const uint4 L = Ln, F = Fn;
        if (c + 1 < p.n_chunks) {
            Ln = ld_stream_v4(p.lit + (uint64_t)(c + 1) * CSR_CHUNK + pos);
            Fn = __ldg(reinterpret_cast<const uint4 *>(p.start) + c + 1);
        }
        if (lane == 0 && c + CSR_PREFETCH < c1) tma_prefetch_l2(p.lit + (uint64_t)(c + CSR_PREFETCH) * CSR_CHUNK, CSR_CHUNK * 4);
        const uint32_t lits[4] = {L.x, L.y, L.z, L.w};
        const unsigned long long flo = (unsigned long long)F.x | ((unsigned long long)F.y << 32);
        const unsigned long long fhi = (unsigned long long)F.z | ((unsigned long long)F.w << 32);
        const unsigned long long mine64 = pos < 64 ? flo : fhi;
        const uint32_t f = (uint32_t)(mine64 >> (pos & 63u)) & 0xFu;  // start bits of my 4 positions
        // latest start strictly before my first position (in this chunk), and the rank of my first position
        int last = -1;
        uint32_t starts_below;
        if (pos < 64) {
            const unsigned long long mk = flo & ((1ull << pos) - 1ull);
            if (mk) last = 63 - __clzll((long long)mk);
            starts_below = (uint32_t)__popcll(mk);
        } else {
            const unsigned long long mk = fhi & ((1ull << (pos - 64)) - 1ull);
            if (mk) last = 127 - __clzll((long long)mk);
            else if (flo) last = 63 - __clzll((long long)flo);
            starts_below = (uint32_t)__popcll(mk) + (uint32_t)__popcll(flo);
        }
        const uint32_t rank0 = __ldg(p.chunk_rank + c) + starts_below;
        // distance of each of my literals from the first literal of its clause; "head" = belongs to the entering clause
        uint32_t d[4];
        bool head[4];
        {
            uint32_t dprev = last >= 0 ? pos - (uint32_t)last - 1u : pos + carry_len - 1u;   // distance of position pos - 1 (wraps to ~0 at a fresh start: fixed by +1)
            bool hprev = last < 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool st = (f >> q) & 1u;
                d[q] = st ? 0u : dprev + 1u;
                head[q] = hprev && !st;
                dprev = d[q];
                hprev = head[q];
            }
        }
        // which literals take part at all: not the ones of a predecessor's clause, and in finishing mode only the head
        bool act[4];
#pragma unroll
        for (int q = 0; q < 4; q++) act[q] = finishing ? head[q] : (!head[q] || carry_valid);
        // ---- phase A: the first three literals of every clause
        uint32_t t = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) t |= csr_lookup<STAGED>(lits[q], act[q] && d[q] < 3u, p.bits, smem_base) << q;
        // ---- phase B: later literals, only where the clause's first three were all false
        {
            const uint32_t p1 = __shfl_up_sync(0xffffffffu, t, 1), p2 = __shfl_up_sync(0xffffffffu, t, 2);
            const uint32_t win = (t << 8) | ((lane >= 1 ? p1 : 0u) << 4) | (lane >= 2 ? p2 : 0u);   // bit b <-> position pos - 8 + b
            uint32_t tb = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                bool go = act[q] && d[q] >= 3u;
                // first three literals of my clause inside the window (clause began <= 7 positions ago, in this chunk)?
                if (go && d[q] <= 7u && pos + q >= d[q]) go = ((win >> (8 + q - d[q])) & 7u) == 0u;
                tb |= csr_lookup<STAGED>(lits[q], go, p.bits, smem_base) << q;
            }
            t |= tb;
        }
        // ---- segments: inside my lane, then across lanes
        const bool has_start = f != 0u;
        const uint32_t first_q = has_start ? (uint32_t)__ffs((int)f) - 1u : 4u;
        const uint32_t last_q = has_start ? 31u - (uint32_t)__clz((int)f) : 0u;
        const bool head_sat = (t & ((1u << first_q) - 1u)) != 0u;                       // my part of the entering clause (all 4 without a start)
        const bool tail_sat = has_start && (t & (0xFu & ~((1u << last_q) - 1u))) != 0u; // from my last start to the end of my lane
        const uint32_t H = __ballot_sync(0xffffffffu, has_start);
        const uint32_t B = __ballot_sync(0xffffffffu, head_sat);
        // the clause entering the chunk closes at the first start of the chunk
        bool emit_carry = false;
        if (H != 0u) {
            const uint32_t j0 = (uint32_t)__ffs((int)H) - 1u;
            emit_carry = carry_valid && !(carry_sat || (B & ((2u << j0) - 1u)) != 0u);
        } else {
            carry_sat = carry_sat || B != 0u;
            carry_len += CSR_CHUNK;
        }
        out.push1(lane == 0 && emit_carry && carry_id < p.m, carry_id);
        if (finishing) {
            if (H != 0u) break;                                                          // our last clause is closed
            continue;
        }
        if (H != 0u) {
            // my starts: all but the last close inside my lane; the last one closes at the next lane holding a start
            const uint32_t above = lane == 31u ? 0u : H >> (lane + 1u);
            bool open_sat = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                bool violated = false;
                const bool st = (f >> q) & 1u;
                if (st) {
                    const uint32_t later = (f >> (q + 1)) & 0x7u;
                    if (later) {
                        const uint32_t qn = (uint32_t)q + (uint32_t)__ffs((int)later);   // next start in my lane
                        violated = (t & (((1u << qn) - 1u) & ~((1u << q) - 1u))) == 0u;
                    } else if (above) {
                        const uint32_t j = lane + (uint32_t)__ffs((int)above);           // next lane holding a start
                        const uint32_t range = ((2u << j) - 1u) & ~((2u << lane) - 1u);  // lanes (lane, j]
                        violated = !(tail_sat || (B & range) != 0u);
                    } else {
                        open_sat = tail_sat || (B & ~((2u << lane) - 1u)) != 0u;          // still open at the end of the chunk
                    }
                }
                const uint32_t id = rank0 + (uint32_t)__popc(f & ((1u << q) - 1u));
                out.push1(violated && id < p.m, id);
            }
            // new carry: the clause that started last in this chunk (held by the highest lane with a start)
            const uint32_t hi = 31u - (uint32_t)__clz((int)H);
            carry_valid = true;
            carry_sat = __shfl_sync(0xffffffffu, (uint32_t)open_sat, hi) != 0u;
            carry_id = __shfl_sync(0xffffffffu, rank0 + (uint32_t)__popc(f) - 1u, hi);
            carry_len = CSR_CHUNK - (4u * hi + __shfl_sync(0xffffffffu, last_q, hi));
        }
    }
    if (out.count) out.flush();
}

// shared memory of a CSR sweep: staged assignment | violated staging
static inline size_t sweep_csr_smem_bytes(uint32_t staged_words, uint32_t threads)
{
    return (size_t)staged_words * 4 + (threads / 32) * WBUF * 4;
}

} // namespace alll
