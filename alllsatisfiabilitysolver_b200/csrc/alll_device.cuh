// alll_device.cuh -- shared device helpers and kernel parameter blocks (sm_100a).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace alll {

// ---- deterministic round specification (mirrors oracle/alll_oracle.c, which is the checker) ----
constexpr uint32_t STREAM_INIT = 0u;
constexpr uint32_t STREAM_RESAMPLE = 1u;
constexpr uint32_t STREAM_PRIORITY = 2u;

#ifndef ALLL_SWEEP_THREADS
#define ALLL_SWEEP_THREADS 512
#endif
constexpr uint32_t SWEEP_THREADS = ALLL_SWEEP_THREADS;           // one CTA per SM (the staged assignment owns the shared memory), <=128 regs/thread
constexpr uint32_t CLAUSES_PER_THREAD = 4;        // one 128-bit load per literal plane
constexpr uint32_t TILE = SWEEP_THREADS * CLAUSES_PER_THREAD;   // clause slots per sweep tile
constexpr uint32_t WBUF = 64;                     // per-warp violated-id staging entries (flushed at >= 32, +32 max per push)
constexpr uint32_t QBUF = 160;                    // per-warp parked-clause queue entries (<= 31 kept + 128 max per push)
constexpr uint32_t MAX_BUCKETS = 256;
constexpr uint32_t EAGER_PLANES = 5;              // planes streamed by the sweep; the tail planes are fetched only for surviving clauses
constexpr uint32_t RESIDENT_CAP = 3;              // literals of a clause placed as bucket-resident by default (measured best of 2/3/4 at k=8)
constexpr uint32_t RESIDENT_CAP_MAX = 4;
constexpr uint32_t MAX_K = 32;
constexpr uint32_t MIS_CLUSTER_MAX_U = 8192;       // violated sets up to this size are handled by one 8 x 1024-thread cluster
constexpr uint32_t MAX_SHARDS = 64;                // clause-range shards (GPUs) of one instance
constexpr uint32_t INVALID_ID = 0xFFFFFFFFu;

constexpr unsigned long long CLAIM_FREE = ~0ull;  // nobody claims this variable
constexpr unsigned long long CLAIM_TAKEN = 0ull;  // variable belongs to a clause already in the independent set
constexpr uint32_t TAGS = 62;                     // claim tag = TAGS - (step % TAGS), in bits 63..58

struct Philox {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox{c0, c1, c2, c3};
}

// Fair bit for variable v in (stream, round): word (v>>5)&3, bit v&31 of Philox(ctr={v>>7, round, stream, 0}).
__device__ __forceinline__ uint32_t random_bit(uint64_t seed, uint32_t stream, uint32_t round, uint32_t v)
{
    const Philox o = philox4x32_10(v >> 7, round, stream, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t sel = (v >> 5) & 3u;
    const uint32_t word = sel == 0 ? o.x : sel == 1 ? o.y : sel == 2 ? o.z : o.w;
    return (word >> (v & 31u)) & 1u;
}

// 26-bit priority of clause id c in a round; MIS order is (priority, id) ascending.
__device__ __forceinline__ uint32_t clause_priority(uint64_t seed, uint32_t round, uint32_t c)
{
    return philox4x32_10(c, round, STREAM_PRIORITY, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)).x >> 6;
}

__device__ __forceinline__ unsigned long long claim_key(uint32_t step, uint32_t prio26, uint32_t id)
{
    const unsigned long long tag = TAGS - (step % TAGS);          // 62 .. 1: later steps always undercut stale claims
    return (tag << 58) | ((unsigned long long)prio26 << 32) | id;
}

// ---- streaming 128-bit load that does not pollute L1 (the literal planes are read exactly once per sweep)
__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// TMA bulk prefetch of a contiguous global range into L2 (no destination, no register cost): one thread pulls
// a whole 8 KB plane segment of a future tile from HBM so that the later 128-bit loads hit L2.
__device__ __forceinline__ void tma_prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(bytes) : "memory");
}

// (Measured and rejected, profiles/r02_l2_policy.md: an explicit L2 evict_first policy on the stream -- loads and bulk
// prefetches through createpolicy / .L2::cache_hint -- runs exactly like these hint-less forms, an evict_normal policy is
// 10 % slower, and prefetch.global.L2 (SASS CCTL.E.PML2) of the claim sectors of violated clauses from inside the sweep
// costs the stream far more than it saves the independent-set phase.)

// ---- packed eager planes ---------------------------------------------------------------------------------------
// The sweep streams the first five literals of every clause (planes 0..4) and is bound by those bytes.  On a bucketed
// layout the leading RB literals of a clause lie in its bucket's variable range, so they are stored relative to the
// bucket (2 * bucket_vars <= 2^22), and the others need 1 + log2(n_vars) bits: five literals fit ONE 128-bit word per
// clause -- four packed planes instead of five, 16 streamed bytes per clause instead of 20.
//   RB = 2: [l0' : 22][l1' : 22][l2 : 28][l3 : 28][l4 : 28]   (n_vars <= 2^27)
//   RB = 1: [l0' : 22][l1 : 26][l2 : 26][l3 : 26][l4 : 26]    (n_vars <= 2^25), two spare bits
// l' = l - 2 * vbase(bucket).  Word w of clause slot p is packed[w * m_pad + p].  The unpacked planes stay in HBM next to
// them for everything that reads single clauses (records, independent set, tail literals, incremental rows).
template <int RB>
struct EagerPack {
    static_assert(RB == 1 || RB == 2, "packed eager planes need one or two bucket-resident leading literals");
    static constexpr int N = 5;                                   // == EAGER_PLANES
    static constexpr int WORDS = 4;
    static constexpr int REL_BITS = 22, GLOB_BITS = RB == 2 ? 28 : 26;
    __host__ __device__ static constexpr int width(int j) { return j < RB ? REL_BITS : GLOB_BITS; }
    __host__ __device__ static constexpr int offset(int j) { return j <= RB ? j * REL_BITS : RB * REL_BITS + (j - RB) * GLOB_BITS; }
    static_assert(offset(4) + width(4) <= 128, "five literals must fit 128 bits");

    // l[0 .. RB) already relative to the bucket
    __host__ __device__ static void encode(const uint32_t (&l)[N], uint32_t (&w)[WORDS])
    {
        for (int i = 0; i < WORDS; i++) w[i] = 0u;
        for (int j = 0; j < N; j++) {
            const int o = offset(j), i = o >> 5, sh = o & 31;
            w[i] |= l[j] << sh;
            if (sh + width(j) > 32) w[i + 1] |= l[j] >> (32 - sh);
        }
    }
    template <int J>
    __device__ __forceinline__ static uint32_t field(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
    {
        constexpr int o = offset(J), i = o >> 5, sh = o & 31, wd = width(J);
        constexpr uint32_t mask = (1u << wd) - 1u;
        const uint32_t lo = i == 0 ? w0 : i == 1 ? w1 : i == 2 ? w2 : w3;
        if constexpr (sh + wd == 32) return lo >> sh;
        else if constexpr (sh + wd < 32) return (lo >> sh) & mask;
        else {
            const uint32_t hi = i == 0 ? w1 : i == 1 ? w2 : w3;
            return __funnelshift_r(lo, hi, sh) & mask;
        }
    }
};

// ---- clause access for the sparse kernels (MIS / resample / id mapping) ----
struct ClauseView {
    // fixed-k literal planes: lit j of slot p is planes[j * m_pad + p]
    const uint32_t *planes;
    uint64_t m_pad;
    uint32_t k;                 // 0 => CSR
    // CSR
    const uint64_t *off;
    const uint32_t *csr_lit;
    // slot -> caller clause id (NULL = identity), plus the first id of this clause range (sharded mode)
    const uint32_t *orig_id;
    uint32_t id_base;
    // padded planes (ragged input laid out with k = widest clause): true width per slot, NULL when all equal k
    const uint8_t *width_arr;
    // enumerated clauses (rec != NULL; k = width): only the violated ones exist, as records {index, k literals} that the
    // generator sweep of this round wrote; "slot" p is the record number
    const uint32_t *rec;
    // incremental mode: row-major copy of the literals, rows[slot][row_stride] (NULL otherwise) -- the sparse kernels then
    // fetch a clause with one or two sectors instead of one per literal plane
    const uint32_t *rows;
    uint32_t row_stride;

    __device__ __forceinline__ uint32_t width(uint32_t p) const
    {
        return width_arr ? width_arr[p] : (k ? k : (uint32_t)(off[p + 1] - off[p]));
    }
    __device__ __forceinline__ uint32_t literal(uint32_t p, uint32_t j) const
    {
        if (rec) return rec[(uint64_t)p * (k + 1) + 1 + j];
        if (rows) return rows[(uint64_t)p * row_stride + j];
        return k ? planes[(uint64_t)j * m_pad + p] : csr_lit[off[p] + j];
    }
    __device__ __forceinline__ uint32_t id(uint32_t p) const
    {
        if (rec) return rec[(uint64_t)p * (k + 1)];
        return (orig_id ? orig_id[p] : p) + id_base;
    }
};

// Device-side counters of one handle.
struct Counters {
    unsigned int n_viol;        // |U| of the current sweep
    unsigned int n_s;           // |S| of the current round
    unsigned int step_live[64]; // clauses still undecided entering Luby step (index = step % 64)
    unsigned long long n_resampled_round;
    // running totals of a solve
    unsigned long long n_iterations;
    unsigned long long sum_mis;
    unsigned long long n_resamples;
    unsigned long long n_luby_steps;
    // snapshot left by the most recent MIS kernel (what the host reads); n_viol/n_s/n_resampled_round are
    // zeroed by that kernel so the next sweep starts from a clean slate without an extra launch
    unsigned int last_n_viol;
    unsigned int last_n_s;
    unsigned long long last_resampled;
    // set by the MIS kernel that saw an empty violated set: kernels of rounds the host enqueued speculatively
    // behind it return immediately, so the round loop never has to wait for the host between rounds
    unsigned int done;
    unsigned int cta_done;      // sweep CTAs that have finished (sharded P2P mode: the last one publishes the round)
    unsigned int p2p_error;     // 1: capacity overflow (P2P exchange region / enumerated-clause record buffer), 2: a peer did not arrive in time, 3: a peer aborted the solve
    unsigned int handled_tag;   // sharded P2P mode: tag of the last round an MIS kernel has completed
    // incremental re-evaluation: decided by the MIS kernel at the end of a round for the NEXT round
    unsigned int incr_next;     // 1: the next violated set comes from incr_eval_kernel, the sweep kernel returns at entry
    unsigned int n_incr_rounds; // rounds evaluated incrementally so far
    unsigned long long n_evals_incr;   // clauses actually evaluated by incremental rounds
    // persistent solve kernel: |U| by round parity; time spent in sweeps / between sweeps as block 0 saw it
    unsigned int n_viol_pp[2];
    unsigned long long t_sweep_ns, t_mis_ns;
    // Luby step barrier of the grid-wide independent-set phase (mis_body.cuh: luby_barrier_fused), by step parity:
    // CTAs arrived (low half) | clauses that claimed for the next step (high half); zero between rounds
    unsigned long long luby_bar[2];
    // ALLL_TRACE: %globaltimer stamps of the first DBG_ROUNDS rounds (one thread writes them; a few stores per round)
    // [0] sweep entry  [1] MIS kernel entry  [2] |U| known  [3] gather + first claims done  [4] Luby steps done
    // [5] resample done  [6] round finished  [7] (steps << 8) | path (0 small, 1 cluster, 2 grid)
    unsigned long long dbg[32][8];
    // ALLL_TRACE: inside the Luby steps of round 0 (grid-wide path), as block 0 / thread 0 sees them:
    // [s][0] step entry  [1] own clauses decided (claim reads + stores issued)  [2] CTA reduced + live count published  [3] grid barrier passed
    unsigned long long dbg_step[16][4];
    // ALLL_TRACE, sharded persistent solve, block 0 / thread 0: [r][0] own sweep body done  [1] record stores fenced (system
    // scope) + ticket drawn  [2] (the CTA with the last ticket) count + flag stored into every GPU  [3] every rank's flag seen
    unsigned long long dbg_x[32][4];
    // ALLL_TRACE: %globaltimer at the end of every CTA's sweep body in round 1 of a persistent solve (tail / imbalance study)
    unsigned long long dbg_cta[256];
};
constexpr uint32_t DBG_ROUNDS = 32;

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// What the round loop on the host needs to know about a finished round.  Lives in pinned host memory; the MIS
// kernel's last thread stores it directly over PCIe (no copy-engine hop between the kernels of consecutive rounds)
// and publishes it by writing `seq` last.
struct RoundNote {
    unsigned int n_viol;
    unsigned int n_s;
    unsigned long long seq;
};

// Per-job result of the batched small-instance kernel (mirrors alll_batch_stats of the C ABI).
struct BatchJobStats {
    unsigned long long n_iterations, n_resamples, sum_mis_size;
    int status, reserved;
};
constexpr int BATCH_PREEMPTED = 8;      // == ALLL_PREEMPTED: portfolio job stopped because another seed finished first

// ---- peer-to-peer exchange of the clause-range sharded mode (one process per GPU, CUDA IPC mappings) ----
// Every GPU owns one exchange region: a header, then records[2 parities][world][cap][k+1].  In round r the sweep
// kernel of rank q stores its violated records into slot [r & 1][q] of EVERY GPU's region over NVLink, then
// publishes count and arrival flag; the MIS kernel of each GPU waits for all flags of the round and works on its
// own region.  Two parities suffice: a rank cannot start sweep r+2 before every peer has finished MIS r.
struct P2PHeader {
    // written by peers, one 8-byte store per (parity, source rank): (tag of the round whose records are complete) << 32 |
    // number of records of that round.  Count and arrival flag travel in ONE store, so the publisher needs no
    // system-scope fence between them (a fence costs an NVLink round trip on the critical path of every round).
    unsigned long long cf[2][MAX_SHARDS];
    unsigned int abort;                  // == (epoch & 0xFFF) + 1 of the solve in which a peer overflowed its capacity or a wait timed out
                                         // (tagged by solve, so that a failed solve does not poison the next one on the same link)
};
constexpr uint32_t P2P_HEADER_BYTES = 4096;
static_assert(sizeof(P2PHeader) <= P2P_HEADER_BYTES, "header must fit its page");

struct P2PLink {                         // device-resident, one per handle
    uint32_t world, rank, k;
    uint64_t cap;                        // records per (parity, source rank)
    long long timeout_cycles;            // bounded wait for a peer's round (clock64 ticks; ALLL_P2P_TIMEOUT_MS, default 20 s)
    P2PHeader *hdr[MAX_SHARDS];          // every GPU's header (own one included), peer-mapped
    uint32_t *rec[MAX_SHARDS];           // every GPU's record area
};

// Claim storage of a handle (+ the records the sweep of this round wrote next to the violated list, if any).
// Sharded persistent solve: set by any thread of the CTA that stored into a peer's exchange region this round.  A CTA that
// stored nothing draws its ticket without the system-scope fence (~3 us on the critical path of every round; in the late
// rounds of a solve most CTAs find no violated clause at all).
__shared__ uint32_t g_remote_dirty;

struct MisScratch {
    unsigned long long *claim;  // [n_vars][2] claim words (even / odd Luby steps), CLAIM_FREE between rounds
    const uint32_t *urec;       // [urec_cap][k+1] records {id, literals} parallel to viol[] (NULL: none)
    uint32_t urec_cap;
};

// One run of clause slots that all belong to the same variable-range bucket.  The bucketing pass lays the clauses out
// upload chunk by upload chunk (so that it runs behind the H2D copy of each chunk), bucket by bucket inside a chunk, every
// (chunk, bucket) segment starting on a sweep-tile boundary.  The sweep walks the segments bucket by bucket instead (all
// chunks' segments of bucket 0, then bucket 1, ...): a CTA's contiguous range of that order stays inside one bucket almost
// always and stages its slice of the assignment once.  Two tables: in slot order (layout passes) and in sweep order
// (SweepParams::segs; empty segments dropped), where tile_begin counts tiles of the sweep order.
struct BucketSeg {
    uint32_t tile_begin;        // first tile of this segment in the table's own order
    uint32_t slot_end;          // one past its last valid clause slot
    uint32_t bucket;            // variable-range bucket its clauses are resident in
    uint32_t phys_tile;         // first tile in slot space (== tile_begin in the slot-order table)
};

// What one CTA of the sweep streams: runs of consecutive tiles (slot space), each inside one bucket segment.  The host cuts
// the sweep order (bucket-major over the segments) into sweep_grid equal ranges and every range into its runs at upload, so
// the streaming loop works on plain tile numbers -- the loop is latency-bound and has neither registers nor dependent loads
// to spare for a per-tile mapping (profiles/r02_pipelined_layout.md).
struct SweepRun {
    uint32_t tile_begin, tile_end;   // tiles [begin, end) in slot space
    uint32_t slot_end;               // one past the last valid clause slot of the run's segment
    uint32_t bucket;                 // variable-range bucket the run's clauses are resident in
};

// The segment a sweep tile belongs to: the last one that starts at or before it (empty segments share their successor's
// first tile; segs[0].tile_begin == 0).
__device__ __forceinline__ uint32_t find_segment(const BucketSeg *__restrict__ segs, uint32_t n_segs, uint32_t tile)
{
    uint32_t lo = 0, hi = n_segs;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (segs[mid].tile_begin <= tile) lo = mid;
        else hi = mid;
    }
    return lo;
}

struct SweepParams {
    const uint32_t *planes;     // [k][m_pad]
    uint64_t m_pad;
    const uint32_t *bits;       // bit-packed assignment, n_words (padded to a multiple of 4) words
    uint32_t n_words;
    uint32_t bucket_words;      // words of assignment staged per bucket (multiple of 4)
    uint32_t n_segs;            // bucket segments (BucketSeg)
    uint32_t n_tiles;
    const BucketSeg *segs;      // [n_segs], sweep order
    const SweepRun *runs;       // the runs of CTA c are runs[run_begin[c] .. run_begin[c + 1]) (sweep_planes kernels, k <= 8)
    const uint32_t *run_begin;  // [run_grid + 1]
    uint32_t run_grid;          // grid size the run lists were cut for
    uint32_t *viol;             // out: violated slots
    Counters *ctr;
    uint32_t k;
    uint32_t eager;             // tuning: planes streamed eagerly (0 = default EAGER_PLANES)
    uint32_t prefetch_tiles;    // tiles of L2 prefetch distance ahead of the register double buffer (0 = off)
    uint32_t resident_cap;      // resident cap the layout was built with (0 = RESIDENT_CAP)
    uint32_t min_resident;      // min over clauses of the number of resident-placed literals (0 when unknown)
    // sharded P2P mode (p2p == NULL otherwise): records go to every peer instead of the local violated list
    const P2PLink *p2p;
    uint32_t p2p_parity, p2p_tag;
    uint32_t p2p_epoch;         // epoch & 0xFFF of this solve (the abort word of the exchange headers is tagged with it)
    const uint32_t *orig_id;
    uint32_t id_base;
    uint32_t round;             // solve round this sweep belongs to (trace stamps only)
    // records {caller id, k literals} of the first urec_cap violated clauses, parallel to viol[] (NULL = not written)
    uint32_t *urec;
    uint32_t urec_cap;
    // measurement knobs (environment ALLL_TUNE, read at alll_create): TUNE_* bits
    uint32_t tune;
    // packed eager planes [4][m_pad] (EagerPack<min(min_resident, 2)>; NULL = the sweep streams planes 0..4)
    const uint32_t *packed;
    // row-major copy [m_pad][8] (5 < k <= 8; NULL = none), 32 bytes per clause: a clause that survives its five eager
    // literals gets literals 5 .. k-1 with ONE 16-byte fetch of the row's second half instead of one sector per tail plane
    const uint4 *rows8;
};
constexpr uint32_t TUNE_NO_TAIL_ROWS = 2u;         // survivors fetch their tail literals from the planes (one sector per literal)
constexpr uint32_t TUNE_NO_PACKED_PLANES = 4u;     // the sweep streams planes 0..4 even where the packed eager planes apply
constexpr uint32_t TUNE_NO_NEXT_SWEEP_PREFETCH = 8u; // no L2 prefetch of the next sweep's first tiles at the end of a round
constexpr uint32_t TUNE_CG_LUBY_BARRIER = 1u;      // Luby steps end with grid.sync() + a counter of their own instead of luby_barrier_fused

} // namespace alll
