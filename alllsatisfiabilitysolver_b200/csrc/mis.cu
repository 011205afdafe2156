// mis.cu -- K3 (maximal independent set of the violated clauses) + K4 (resample).
//
// Replaces populate_mis_parallel (SATInstance.h:391-451; greedy, O(|S|*|U|*k^2), one OpenMP fork/join per
// picked clause) and resample_clauses (SATInstance.h:340-365).
//
// K3: fixed-priority Luby.  Every violated clause c gets the key (Philox priority(seed, round, id), id).
//     Every undecided clause atomicMin-s its key into the claim word of every variable it touches; a clause
//     that owns all its claims joins S and marks its variables TAKEN; a clause that sees a TAKEN variable
//     drops out; the rest claim again for the next step (in the other of two claim arrays).
//     Iterated to a fixed point this is exactly the greedy independent set in ascending key order
//     (what oracle/alll_oracle.c:alll_oracle_priority_mis computes sequentially) -- independent and
//     maximal like the reference's set (SATInstance.h:415-447), and a pure function of (seed, round, U):
//     no dependence on thread schedule, compaction order or shard count.
//     Keys carry a step tag in the top bits that decreases every step, so claims of clauses that lost
//     or dropped out in earlier steps are undercut without a reset pass.
// K4: every variable of every clause in S is redrawn from Philox(seed, RESAMPLE, round, var); bits are
//     written with atomicOr/atomicAnd on the packed word (idempotent, so a variable occurring twice in a
//     clause is harmless).  The same pass restores the claim words (and table keys) of everything U touched.
//
// Where the claim words live.  A violated set touches |U|*k variables out of n -- at 10 M variables a claim
// array indexed by variable is 80 MB of which a round touches scattered 32-byte sectors, all of them cold in
// DRAM after the sweep has streamed the literals through L2.  So claims are indexed through a COMPACT TABLE
// whenever that is smaller: an open-addressing table of T = 2*|U|*k entries (keys hvar[], linear probing,
// atomicCAS inserts) maps each touched variable to an entry once per round; the entry index is cached next
// to the literal and all later steps address claim words by entry.  The table occupies a dense prefix of
// its buffers, so its sectors are shared by many variables and stay L2-resident for the whole round.  When
// the table would not be smaller than the per-variable array (small n or huge U) the entry index is the
// variable itself.  Either way every touched word is restored at the end of the round (clean-after-use).
//
// Latency structure: literals, entry indices, priority, id and state of every violated clause are held in
// shared memory; all claim reads of a step are issued together, and deciding step s is fused with claiming
// for step s+1: one barrier and one memory round trip per Luby step however wide the clauses are.
// Code size matters here: these kernels run for a few microseconds right after the sweep has flushed L2,
// so every instruction line is fetched cold from DRAM.  Philox is a single out-of-line function and the
// literal loops are only unrolled where the loads need to overlap (the first version of this file compiled
// to 160 KB per kernel and spent most of each phase waiting for instructions).
//
// Two kernels are enqueued per round and decide on the device which one acts (the host does not know
// |U|: rounds are enqueued speculatively, see capi.cu):
//   mis_cluster_kernel : |U| <= 8192 -- one thread-block cluster of 8 x 1024 threads, one clause per
//                        thread, phases separated by the hardware cluster barrier; |U| <= 512 is done by
//                        CTA 0 alone with the claims in a shared-memory hash table (mis_small_body);
//   mis_grid_kernel    : larger |U| -- cooperative launch, phases separated by grid-wide barriers.
#include <cooperative_groups.h>

#include "alll_device.cuh"

namespace cg = cooperative_groups;

namespace alll {

constexpr uint32_t GRID_THREADS = 256;
constexpr uint32_t CL_THREADS = 1024;
constexpr uint32_t CL_SIZE = 8;
constexpr uint32_t CLUSTER_U = MIS_CLUSTER_MAX_U;        // violated sets up to this size go to the cluster kernel
static_assert(CLUSTER_U == CL_THREADS * CL_SIZE, "one clause per cluster thread");
constexpr uint32_t GRID_SMEM_WORDS_PER_THREAD = 48;     // 48 KB per 256-thread CTA: 4 CTAs per SM
constexpr uint32_t EXTRA = 3;                           // cached per clause besides literals and entries: priority, id, width | state << 8
constexpr uint32_t SMALL_U = 512;                       // violated sets up to this size: one CTA, claims in a shared-memory hash table
constexpr uint32_t HSLOTS = 8192;                       // hash slots (power of two); load factor <= 0.5 => SMALL path needs |U| * k <= 4096
constexpr uint32_t H_EMPTY = 0xFFFFFFFFu, C_FREE = 0xFFFFFFFFu, C_TAKEN = 0u;
constexpr uint32_t PENDING = 0x80000000u;               // entry index whose first probe hit another variable
constexpr uint32_t MIN_TABLE = 256;

enum : uint32_t { UNDECIDED = 0, IN_SET = 1, DROPPED = 2 };

struct MisParams {
    ClauseView cv;
    const uint32_t *viol;       // U as clause slots
    uint8_t *state;             // per U entry (only for entries that do not fit the shared-memory cache)
    uint32_t *s_slots;          // out: S as clause slots
    unsigned long long *claim;  // [2][claim_stride] (even / odd Luby steps), FREE between rounds
    uint64_t claim_stride;      // >= max(n_vars, tcap)
    uint32_t *hvar;             // [tcap] keys of the compact table (variable ids), H_EMPTY between rounds; NULL = no table
    uint32_t tcap;
    uint64_t n_vars;
    uint32_t *bits;
    Counters *ctr;
    uint64_t seed;
    uint32_t round;
    uint32_t kmax;              // widest clause
    uint32_t cache_items;       // clauses per thread that fit the shared-memory cache
    uint32_t grid_follows;      // cluster kernel only: a grid kernel is enqueued behind it and takes large sets
    RoundNote *note;            // pinned host memory (may be NULL): where the finished round is announced
    unsigned long long seq;     // value to publish in note->seq
    // records {id, k literals} parallel to viol[], written by the sweep of this round for the first urec_cap violated
    // clauses (NULL = none): one contiguous read per clause instead of k scattered literal planes
    const uint32_t *urec;
    uint32_t urec_cap;
    // sharded P2P mode (NULL otherwise): U is the union of the record blocks all ranks stored into OUR exchange region
    const P2PLink *p2p;
    uint32_t p2p_parity, p2p_tag;
    uint32_t incr_max_vars;     // incremental mode: next round is incremental iff this round resampled <= this many variables (0 = off)
    uint32_t u_cap;             // enumerated clauses: records the sweep could store (0 = no limit); a larger |U| aborts the solve
};

extern __shared__ uint32_t mis_smem[];

// claim[] and the counters are written by other SMs between barriers: read them at L2 (L1 is not coherent).
__device__ __forceinline__ unsigned long long ld_claim(const unsigned long long *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned int ld_u32(const unsigned int *p) { return __ldcg(p); }

// ALLL_TRACE stamps (alll_device.cuh: Counters::dbg); called by one thread
__device__ __forceinline__ void stamp(const MisParams &p, uint32_t what)
{
    if (p.round < DBG_ROUNDS) p.ctr->dbg[p.round][what] = global_ns();
}

// The one copy of Philox4x32-10 in these kernels (see the note on code size at the top of the file).
__device__ __noinline__ Philox philox_call(uint32_t c0, uint32_t c1, uint32_t c2, uint64_t seed)
{
    return philox4x32_10(c0, c1, c2, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint32_t mis_priority(const MisParams &p, uint32_t id)     // == clause_priority()
{
    return philox_call(id, p.round, STREAM_PRIORITY, p.seed).x >> 6;
}
// K4 for one variable: fresh fair bit (== random_bit(seed, STREAM_RESAMPLE, round, v)) written into the packed word
__device__ __forceinline__ void resample_var(const MisParams &p, uint32_t v)
{
    const Philox o = philox_call(v >> 7, p.round, STREAM_RESAMPLE, p.seed);
    const uint32_t sel = (v >> 5) & 3u;
    const uint32_t word = sel == 0 ? o.x : sel == 1 ? o.y : sel == 2 ? o.z : o.w;
    const uint32_t mask = 1u << (v & 31u);
    if ((word >> (v & 31u)) & 1u) atomicOr(&p.bits[v >> 5], mask);
    else atomicAnd(&p.bits[v >> 5], ~mask);
}

struct GridBarrier {
    cg::grid_group g;
    __device__ __forceinline__ void sync() { g.sync(); }
};
struct ClusterBarrier {
    __device__ __forceinline__ void sync() { cg::this_cluster().sync(); }
};

// Where entry i of U comes from: a record {id, literals} (P2P exchange region, enumerated-clause records, or the
// records the sweep wrote next to viol[]), or the stored clause in slot viol[i].
struct Src {
    const uint32_t *rec;
    uint32_t slot, k;
};

// prefix: exclusive prefix sums of the per-rank record counts (sharded P2P mode; unused otherwise)
__device__ __forceinline__ Src locate(const MisParams &p, const uint32_t *prefix, uint32_t i, bool use_urec)
{
    Src s;
    if (p.p2p) {
        const P2PLink &L = *p.p2p;
        uint32_t q = 0;
        while (prefix[q + 1] <= i) ++q;
        s.rec = L.rec[L.rank] + (((uint64_t)p.p2p_parity * L.world + q) * L.cap + (i - prefix[q])) * (L.k + 1);
        s.slot = i;
        s.k = L.k;
    } else if (p.viol == nullptr) {                // enumerated clauses: U is the record buffer itself
        s.rec = p.cv.rec + (uint64_t)i * (p.cv.k + 1);
        s.slot = i;
        s.k = p.cv.k;
    } else {
        s.slot = p.viol[i];
        s.rec = use_urec ? p.urec + (uint64_t)i * (p.cv.k + 1) : nullptr;
        s.k = use_urec ? p.cv.k : p.cv.width(s.slot);
    }
    return s;
}
__device__ __forceinline__ uint32_t src_id(const MisParams &p, const Src &s) { return s.rec ? s.rec[0] : p.cv.id(s.slot); }
__device__ __forceinline__ uint32_t src_lit(const MisParams &p, const Src &s, uint32_t j)
{
    return s.rec ? s.rec[1 + j] : p.cv.literal(s.slot, j);
}
__device__ __forceinline__ uint32_t slot_of(const MisParams &p, uint32_t i) { return (p.p2p || !p.viol) ? i : p.viol[i]; }

// S gets one more member: one atomic per converged group of winners instead of one per winner
__device__ __forceinline__ void append_s(const MisParams &p, uint32_t slot)
{
    const unsigned int grp = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(grp) - 1;
    unsigned int at = 0;
    if ((int)lane == leader) at = atomicAdd(&p.ctr->n_s, (unsigned int)__popc(grp));
    at = __shfl_sync(grp, at, leader);
    p.s_slots[at + __popc(grp & ((1u << lane) - 1u))] = slot;
}

// Sharded P2P mode: wait until every rank's sweep of this round has published its records in OUR region, then
// build the prefix sums of the counts.  Returns the total; 0xFFFFFFFF on abort / timeout.  Whole CTA calls it.
__device__ __forceinline__ uint32_t p2p_wait(const MisParams &p, uint32_t *s_prefix)
{
    const P2PLink &L = *p.p2p;
    if (threadIdx.x == 0) {
        P2PHeader *me = L.hdr[L.rank];
        const long long t_start = clock64();
        bool bad = false;
        for (uint32_t q = 0; q < L.world && !bad; q++) {
            while (*(volatile unsigned int *)&me->flag[p.p2p_parity][q] != p.p2p_tag) {
                if (*(volatile unsigned int *)&me->abort || clock64() - t_start > 6000000000ll) { bad = true; break; }
            }
        }
        __threadfence_system();                       // acquire: the records behind the flags are now visible
        if (*(volatile unsigned int *)&me->abort) bad = true;     // a peer overflowed even though every flag arrived
        uint32_t run = 0;
        for (uint32_t q = 0; q < L.world; q++) {
            s_prefix[q] = run;
            const unsigned int cnt = *(volatile unsigned int *)&me->count[p.p2p_parity][q];
            if (cnt > L.cap) bad = true;
            run += cnt;
        }
        s_prefix[L.world] = bad ? 0xFFFFFFFFu : run;
    }
    __syncthreads();
    return s_prefix[L.world];
}

__device__ __forceinline__ uint32_t table_home(uint32_t v, uint32_t T) { return __umulhi(v * 2654435761u, T); }

// Threads `first`, `first + stride`, ... of the participating group own the same U entries in every phase.
//
// One barrier and one memory round trip per Luby step: step s decides on the claims standing in array s&1 and,
// in the same pass, clauses that neither won nor dropped claim for step s+1 in the OTHER array, so readers of
// step s are never disturbed by claims of step s+1.  A clause that misses a TAKEN mark written concurrently by a
// winner merely claims once more in vain (its TAKEN variable can never read back its key) and drops out one step
// later; the set of winners is unchanged: a clause wins only when every neighbour with a smaller key has dropped.
//
// ALL_CACHED: every clause of U fits the shared-memory cache (the case the kernels are built for).  The other
// instantiation re-reads the clauses beyond the cache from their source at every step -- slow, and there for
// exactness on extreme inputs (|U| in the millions); its code is never fetched otherwise.
template <class Barrier, bool ALL_CACHED>
__device__ void mis_resample_body(const MisParams &p, Barrier &bar, const uint32_t *prefix, uint32_t first, uint32_t stride,
                                  uint32_t n_u)
{
    __shared__ unsigned int s_live;
    const uint32_t bd = blockDim.x, km = p.kmax, slotw = 2 * km + EXTRA;
    const bool use_urec = p.urec != nullptr && n_u <= p.urec_cap;
    // claims through the compact table iff it is smaller than the per-variable array and every clause is cached
    // (an uncached clause would have to probe the table again at every step)
    const uint64_t touches = (uint64_t)n_u * km;
    const bool compact = ALL_CACHED && p.hvar != nullptr && 2 * touches <= p.tcap && 4 * touches <= p.n_vars;
    const uint32_t T = compact ? (uint32_t)(2 * touches < MIN_TABLE ? MIN_TABLE : 2 * touches) : 0u;
    unsigned long long *const claim0 = p.claim;
    unsigned long long *const claim1 = p.claim + p.claim_stride;
    const uint32_t META = 2 * km + 2;      // cache word holding width | state << 8

    {   // ---- gather: literals -> shared memory, table entries, the claims of step 0
        uint32_t it = 0;
        for (uint32_t i = first; i < n_u; i += stride, ++it) {
            const Src s = locate(p, prefix, i, use_urec);
            const uint32_t id = src_id(p, s);
            if (ALL_CACHED || it < p.cache_items) {
                const uint32_t base = it * slotw * bd + threadIdx.x;
                // literal loads and (compact) first probes of all literals overlap; collisions are resolved below
#pragma unroll 4
                for (uint32_t j = 0; j < s.k; j++) {
                    const uint32_t l = src_lit(p, s, j), v = l >> 1;
                    uint32_t e = v;
                    if (compact) {
                        e = table_home(v, T);
                        const uint32_t old = atomicCAS(&p.hvar[e], H_EMPTY, v);
                        if (old != H_EMPTY && old != v) e |= PENDING;
                    }
                    mis_smem[base + j * bd] = l;
                    mis_smem[base + (km + j) * bd] = e;
                }
                const uint32_t prio = mis_priority(p, id);
                const unsigned long long key = claim_key(0, prio, id);
                mis_smem[base + (2 * km) * bd] = prio;
                mis_smem[base + (2 * km + 1) * bd] = id;
                mis_smem[base + META * bd] = s.k | (UNDECIDED << 8);
                for (uint32_t j = 0; j < s.k; j++) {
                    uint32_t e = mis_smem[base + (km + j) * bd];
                    if (e & PENDING) {                         // linear probing from the home entry
                        const uint32_t v = mis_smem[base + j * bd] >> 1;
                        e &= ~PENDING;
                        uint32_t old;
                        do {
                            e = e + 1 == T ? 0u : e + 1;
                            old = atomicCAS(&p.hvar[e], H_EMPTY, v);
                        } while (old != H_EMPTY && old != v);
                        mis_smem[base + (km + j) * bd] = e;
                    }
                    atomicMin(&claim0[e], key);
                }
            } else {
                p.state[i] = UNDECIDED;
                const unsigned long long key = claim_key(0, mis_priority(p, id), id);
                for (uint32_t j = 0; j < s.k; j++) atomicMin(&claim0[src_lit(p, s, j) >> 1], key);
            }
        }
    }
    if (first < 64) p.ctr->step_live[first] = 0;     // only the acting MIS kernel touches step_live
    bar.sync();
    if (first == 0) stamp(p, 3);

    uint32_t step = 0;
    for (;;) {
        unsigned long long *const cur = (step & 1u) ? claim1 : claim0;
        unsigned long long *const nxt = (step & 1u) ? claim0 : claim1;
        const bool wrap = (step + 1) % TAGS == 0;
        // wrap: the tag of step+1 wraps to the largest value, stale claims in `nxt` would undercut fresh ones, so
        // in a pass of its own the still-undecided clauses clear what they touch there (nobody reads `nxt` now)
        for (uint32_t pass = wrap ? 0u : 1u; pass < 2u; pass++) {
            if (pass == 1u) {
                if (threadIdx.x == 0) s_live = 0;
                __syncthreads();
            }
            uint32_t live = 0, it = 0;
            for (uint32_t i = first; i < n_u; i += stride, ++it) {
                const bool cached = ALL_CACHED || it < p.cache_items;
                const uint32_t base = it * slotw * bd + threadIdx.x;
                uint32_t k, id, prio;
                Src s{};
                if (cached) {
                    const uint32_t w = mis_smem[base + META * bd];
                    if ((w >> 8) != UNDECIDED) continue;
                    k = w & 0xFFu;
                    prio = mis_smem[base + (2 * km) * bd];
                    id = mis_smem[base + (2 * km + 1) * bd];
                } else {
                    if (p.state[i] != UNDECIDED) continue;
                    s = locate(p, prefix, i, use_urec);
                    id = src_id(p, s);
                    k = s.k;
                    prio = mis_priority(p, id);
                }
                auto entry = [&](uint32_t j) { return cached ? mis_smem[base + (km + j) * bd] : src_lit(p, s, j) >> 1; };
                if (pass == 0u) {
                    for (uint32_t j = 0; j < k; j++) {
                        const uint32_t e = entry(j);
                        if (ld_claim(&nxt[e]) != CLAIM_TAKEN) nxt[e] = CLAIM_FREE;
                    }
                    continue;
                }
                const unsigned long long key = claim_key(step, prio, id);
                bool win = true, taken = false;
#pragma unroll 4
                for (uint32_t j = 0; j < k; j++) {               // no early exit: the loads overlap
                    const unsigned long long c = ld_claim(&cur[entry(j)]);
                    win &= c == key;
                    taken |= c == CLAIM_TAKEN;
                }
                if (win) {
                    for (uint32_t j = 0; j < k; j++) {
                        const uint32_t e = entry(j);
                        claim0[e] = CLAIM_TAKEN;
                        claim1[e] = CLAIM_TAKEN;
                    }
                    append_s(p, slot_of(p, i));
                } else if (!taken) {
                    const unsigned long long next_key = claim_key(step + 1, prio, id);
                    for (uint32_t j = 0; j < k; j++) atomicMin(&nxt[entry(j)], next_key);
                    live++;
                    continue;
                }
                const uint32_t st = win ? IN_SET : DROPPED;
                if (cached) mis_smem[base + META * bd] = k | (st << 8);
                else p.state[i] = (uint8_t)st;
            }
            if (pass == 0u) { bar.sync(); continue; }
            if (live) atomicAdd(&s_live, live);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_live) atomicAdd(&p.ctr->step_live[(step + 1) & 63u], s_live);
        // the slot of step+3 (mod 64) is next written two steps from now: clear it while nobody touches it
        if (first == 0) p.ctr->step_live[(step + 3) & 63u] = 0;
        bar.sync();
        step++;
        if (ld_u32(&p.ctr->step_live[step & 63u]) == 0) break;      // nobody claimed for this step: all decided
    }

    if (first == 0) stamp(p, 4);
    // ---- K4 + clean-after-use.  All claim reads of this round are behind the last barrier.
    unsigned long long resampled = 0;
    uint32_t it = 0;
    for (uint32_t i = first; i < n_u; i += stride, ++it) {
        const bool cached = ALL_CACHED || it < p.cache_items;
        const uint32_t base = it * slotw * bd + threadIdx.x;
        uint32_t k, st;
        Src s{};
        if (cached) {
            const uint32_t w = mis_smem[base + META * bd];
            k = w & 0xFFu;
            st = w >> 8;
        } else {
            s = locate(p, prefix, i, use_urec);
            k = s.k;
            st = p.state[i];
        }
        for (uint32_t j = 0; j < k; j++) {
            const uint32_t e = cached ? mis_smem[base + (km + j) * bd] : src_lit(p, s, j) >> 1;
            claim0[e] = CLAIM_FREE;
            claim1[e] = CLAIM_FREE;
            if (compact) p.hvar[e] = H_EMPTY;
            if (st == IN_SET) resample_var(p, cached ? mis_smem[base + j * bd] >> 1 : e);     // (uncached => e is the variable)
        }
        if (st == IN_SET) resampled += k;                          // SATInstance.h:363 counts literals->size()
    }
    // warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) resampled += __shfl_down_sync(0xffffffffu, resampled, o);
    if ((threadIdx.x & 31u) == 0 && resampled) atomicAdd(&p.ctr->n_resampled_round, resampled);
    if (first == 0) {
        atomicAdd(&p.ctr->n_luby_steps, (unsigned long long)step);
        if (p.round < DBG_ROUNDS) p.ctr->dbg[p.round][7] = ((unsigned long long)step << 8) | (compact ? 0x80u : 0u);
    }
}

// ---- small violated sets: the whole independent-set computation in ONE CTA's shared memory ---------------------
// (|U| <= SMALL_U and |U| * kmax <= HSLOTS / 2).  No global claim traffic, no memory fences between steps: a Luby
// step is two __syncthreads().  Exactly the same set as the large paths: the 64-bit (priority, id) keys are replaced
// by their ranks within U (ids are unique, so ranks are a strict order), claims are 32-bit (step tag | rank) words in
// an open-addressing table keyed by variable; each literal's table slot is found once and cached.
__device__ void mis_small_body(const MisParams &p, const uint32_t *prefix, uint32_t n_u)
{
    const uint32_t km = p.kmax, bd = CL_THREADS;
    uint32_t *hvar = mis_smem + (size_t)CL_THREADS * (2 * km + EXTRA);
    uint32_t *hclaim = hvar + HSLOTS;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(hclaim + HSLOTS);     // [SMALL_U]
    __shared__ unsigned int s_cnt, s_sum;
    const uint32_t t = threadIdx.x;
    const bool mine = t < n_u;
    const bool use_urec = p.urec != nullptr && n_u <= p.urec_cap;

    for (uint32_t i = t; i < HSLOTS; i += CL_THREADS) { hvar[i] = H_EMPTY; hclaim[i] = C_FREE; }
    if (t == 0) { s_cnt = 0; s_sum = 0; }
    uint32_t k = 0;
    if (mine) {
        const Src s = locate(p, prefix, t, use_urec);
        const uint32_t id = src_id(p, s);
        k = s.k;
#pragma unroll 4
        for (uint32_t j = 0; j < k; j++) mis_smem[t + j * bd] = src_lit(p, s, j);
        keys[t] = ((unsigned long long)mis_priority(p, id) << 32) | id;
    }
    __syncthreads();
    uint32_t rank = 0;
    if (mine) {
        const unsigned long long k0 = keys[t];
        for (uint32_t j = 0; j < n_u; j++) rank += keys[j] < k0;
        for (uint32_t j = 0; j < k; j++) {                     // register this clause's variables in the table
            const uint32_t v = mis_smem[t + j * bd] >> 1;
            uint32_t s = (v * 2654435761u) & (HSLOTS - 1);
            for (;;) {
                const uint32_t old = atomicCAS(&hvar[s], H_EMPTY, v);
                if (old == H_EMPTY || old == v) break;
                s = (s + 1) & (HSLOTS - 1);
            }
            mis_smem[t + (km + j) * bd] = s;
        }
    }
    __syncthreads();
    if (t == 0) stamp(p, 3);

    uint32_t state = mine ? UNDECIDED : DROPPED;
    uint32_t step = 0;
    for (;;) {
        const uint32_t key = ((TAGS - (step % TAGS)) << 16) | rank;
        bool live = false;
        if (state == UNDECIDED) {
            bool taken = false;
            for (uint32_t j = 0; j < k; j++) taken |= hclaim[mis_smem[t + (km + j) * bd]] == C_TAKEN;
            if (taken) state = DROPPED;
            else {
                for (uint32_t j = 0; j < k; j++) atomicMin(&hclaim[mis_smem[t + (km + j) * bd]], key);
                live = true;
            }
        }
        if (__syncthreads_count(live) == 0) break;
        bool win = false;
        if (state == UNDECIDED) {
            win = true;
            for (uint32_t j = 0; j < k; j++) win &= hclaim[mis_smem[t + (km + j) * bd]] == key;
        }
        __syncthreads();                                       // every win test has read before TAKEN marks land
        if (win) {
            state = IN_SET;
            for (uint32_t j = 0; j < k; j++) hclaim[mis_smem[t + (km + j) * bd]] = C_TAKEN;
        }
        step++;
        if (step % TAGS == 0) {                                // tag wrap: survivors clear their stale claims
            __syncthreads();
            if (state == UNDECIDED)
                for (uint32_t j = 0; j < k; j++) {
                    const uint32_t s = mis_smem[t + (km + j) * bd];
                    if (hclaim[s] != C_TAKEN) hclaim[s] = C_FREE;
                }
        }
        __syncthreads();
    }

    if (t == 0) stamp(p, 4);
    // ---- K4: winners redraw their variables (global bit-packed assignment) and report themselves
    const bool in_s = state == IN_SET;
    if (in_s) {
        for (uint32_t j = 0; j < k; j++) resample_var(p, mis_smem[t + j * bd] >> 1);
        p.s_slots[atomicAdd(&s_cnt, 1u)] = slot_of(p, t);
    }
    uint32_t resampled = in_s ? k : 0u;                        // SATInstance.h:363 counts literals->size()
    for (int o = 16; o > 0; o >>= 1) resampled += __shfl_down_sync(0xffffffffu, resampled, o);
    if ((t & 31u) == 0 && resampled) atomicAdd(&s_sum, resampled);
    __syncthreads();
    if (t == 0) {
        p.ctr->n_s = s_cnt;
        p.ctr->n_resampled_round = s_sum;
        atomicAdd(&p.ctr->n_luby_steps, (unsigned long long)step);
        if (p.round < DBG_ROUNDS) p.ctr->dbg[p.round][7] = (unsigned long long)step << 8;
        __threadfence();
    }
}

// Round bookkeeping by one thread after the last barrier.  n_iterations counts every sweep (SATInstance.h:261).
__device__ __forceinline__ void announce(const MisParams &p, unsigned int n_viol, unsigned int n_s)
{
    if (!p.note) return;
    p.note->n_viol = n_viol;
    p.note->n_s = n_s;
    __threadfence_system();
    *(volatile unsigned long long *)&p.note->seq = p.seq;
}

__device__ __forceinline__ void finish_round(const MisParams &p, uint32_t n_u, uint32_t path)
{
    Counters *c = p.ctr;
    stamp(p, 5);
    if (ld_u32(&c->incr_next)) c->n_incr_rounds += 1;      // the round that just ended was evaluated incrementally
    const unsigned int n_s = ld_u32(&c->n_s);              // (both loads in flight together; the totals below are
    const unsigned long long n_r = __ldcg(&c->n_resampled_round);   //  fire-and-forget atomics: no read-modify-write chain)
    atomicAdd(&c->n_iterations, 1ull);
    atomicAdd(&c->sum_mis, (unsigned long long)n_s);       // SATInstance.h:291
    atomicAdd(&c->n_resamples, n_r);                       // SATInstance.h:313-315
    c->last_n_viol = n_u;
    c->last_n_s = n_s;
    c->last_resampled = n_r;
    c->n_viol = 0;                                         // clean slate for the next sweep
    c->n_s = 0;
    c->n_resampled_round = 0;
    c->handled_tag = p.p2p_tag;
    c->incr_next = (p.incr_max_vars != 0 && n_r <= p.incr_max_vars) ? 1u : 0u;
    announce(p, n_u, n_s);
    if (p.round < DBG_ROUNDS) { c->dbg[p.round][6] = global_ns(); c->dbg[p.round][7] |= path; }
}

// First MIS kernel of a round: owns the terminal case (|U| == 0) and violated sets that fit one cluster.
__global__ void __cluster_dims__(CL_SIZE, 1, 1) __launch_bounds__(CL_THREADS) mis_cluster_kernel(const MisParams p)
{
    __shared__ uint32_t s_prefix[MAX_SHARDS + 1];
    const unsigned long long t_entry = global_ns();
    if (ld_u32(&p.ctr->done)) return;             // speculative round behind the terminal one
    // |U|: left by the sweep kernel that ran before us, or (sharded P2P mode) the sum over all ranks' record blocks
    const uint32_t n_u = p.p2p ? p2p_wait(p, s_prefix) : ld_u32(&p.ctr->n_viol);
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.round < DBG_ROUNDS) {
        p.ctr->dbg[p.round][1] = t_entry;
        p.ctr->dbg[p.round][2] = global_ns();
    }
    if (n_u == 0xFFFFFFFFu || (p.u_cap && n_u > p.u_cap)) {   // a peer overflowed or never arrived / records did not fit: stop the solve
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.ctr->p2p_error = p.ctr->p2p_error ? p.ctr->p2p_error : (n_u == 0xFFFFFFFFu ? 2 : 1);
            p.ctr->n_viol = 0;
            p.ctr->done = 1;
            announce(p, 0xFFFFFFFFu, 0u);
        }
        return;
    }
    if (n_u == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&p.ctr->n_iterations, 1ull);  // the terminal all-satisfied sweep counts (SATInstance.h:261,285-287)
            if (ld_u32(&p.ctr->incr_next)) p.ctr->n_incr_rounds += 1;
            p.ctr->last_n_viol = 0;
            p.ctr->last_n_s = 0;
            p.ctr->last_resampled = 0;
            p.ctr->done = 1;
            p.ctr->n_viol = 0;
            p.ctr->handled_tag = p.p2p_tag;
            announce(p, 0u, 0u);
        }
        return;
    }
    if (n_u > CLUSTER_U && p.grid_follows) return;  // the grid kernel behind us takes it (else: strided, slower, still exact)
    if (n_u <= SMALL_U && (uint64_t)n_u * p.kmax <= HSLOTS / 2 && p.cache_items) {
        if (blockIdx.x != 0) return;              // uniform over the cluster: nobody waits on a cluster barrier below
        mis_small_body(p, s_prefix, n_u);
        if (threadIdx.x == 0) finish_round(p, n_u, 0u);
        return;
    }
    ClusterBarrier bar;
    if (n_u <= CLUSTER_U * p.cache_items) mis_resample_body<ClusterBarrier, true>(p, bar, s_prefix, blockIdx.x * CL_THREADS + threadIdx.x, CLUSTER_U, n_u);
    else mis_resample_body<ClusterBarrier, false>(p, bar, s_prefix, blockIdx.x * CL_THREADS + threadIdx.x, CLUSTER_U, n_u);
    bar.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) finish_round(p, n_u, 1u);
}

// Second MIS kernel of a round (cooperative launch): violated sets too large for one cluster.
__global__ void __launch_bounds__(GRID_THREADS) mis_grid_kernel(const MisParams p)
{
    __shared__ uint32_t s_prefix[MAX_SHARDS + 1];
    const unsigned long long t_entry = global_ns();
    if (ld_u32(&p.ctr->done)) return;
    if (p.p2p && ld_u32(&p.ctr->handled_tag) == p.p2p_tag) return;     // the cluster kernel already did this round
    const uint32_t n_u = p.p2p ? p2p_wait(p, s_prefix) : ld_u32(&p.ctr->n_viol);  // 0 when the cluster kernel handled it
    if (n_u <= CLUSTER_U || n_u == 0xFFFFFFFFu) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.round < DBG_ROUNDS) {
        p.ctr->dbg[p.round][1] = t_entry;
        p.ctr->dbg[p.round][2] = global_ns();
    }
    GridBarrier bar{cg::this_grid()};
    const uint32_t stride = gridDim.x * GRID_THREADS;
    if ((uint64_t)n_u <= (uint64_t)stride * p.cache_items) mis_resample_body<GridBarrier, true>(p, bar, s_prefix, blockIdx.x * GRID_THREADS + threadIdx.x, stride, n_u);
    else mis_resample_body<GridBarrier, false>(p, bar, s_prefix, blockIdx.x * GRID_THREADS + threadIdx.x, stride, n_u);
    bar.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) finish_round(p, n_u, 2u);
}

// Per-round scratch reset + (optionally) totals reset.
__global__ void reset_counters_kernel(Counters *c, int reset_totals)
{
    c->n_viol = 0;
    c->n_s = 0;
    c->n_resampled_round = 0;
    c->last_n_viol = 0;
    c->last_n_s = 0;
    c->last_resampled = 0;
    c->done = 0;
    c->cta_done = 0;
    c->p2p_error = 0;
    c->incr_next = 0;
    if (reset_totals) {
        c->n_incr_rounds = 0;
        c->n_evals_incr = 0;
        c->n_iterations = 0;
        c->sum_mis = 0;
        c->n_resamples = 0;
        c->n_luby_steps = 0;
    }
}

// slots -> caller clause ids (for alll_eval / alll_round outputs)
__global__ void map_ids_kernel(const ClauseView cv, const uint32_t *slots, uint32_t n, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cv.id(slots ? slots[i] : i);
}

// ---- host side --------------------------------------------------------------------------------------

// per cached clause: kmax literals, kmax claim entries, priority, id, width | state
static uint32_t grid_cache_items(uint32_t kmax) { return GRID_SMEM_WORDS_PER_THREAD / (2 * kmax + EXTRA); }
static size_t grid_smem_bytes(uint32_t kmax) { return (size_t)grid_cache_items(kmax) * (2 * kmax + EXTRA) * GRID_THREADS * 4; }
// the cluster kernel caches its one clause per thread whenever that fits the opt-in shared memory
static uint32_t cluster_cache_items(uint32_t kmax) { return (size_t)(2 * kmax + EXTRA) * CL_THREADS * 4 <= 144u * 1024u ? 1u : 0u; }
static size_t cluster_smem_bytes(uint32_t kmax)
{
    if (!cluster_cache_items(kmax)) return 0;
    return (size_t)(2 * kmax + EXTRA) * CL_THREADS * 4 + (size_t)2 * HSLOTS * 4 + (size_t)SMALL_U * 8;   // clause cache | hash table | keys
}

// Called once per upload on the handle's device: shared-memory opt-in + cooperative grid size.
cudaError_t mis_configure(int device, uint32_t kmax, uint32_t *grid_out)
{
    cudaError_t e = cudaFuncSetAttribute(mis_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cluster_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mis_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mis_grid_kernel, GRID_THREADS, grid_smem_bytes(kmax));
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm > 4) per_sm = 4;          // 4 x 256 threads per SM is plenty for an atomics-bound kernel
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid_out = (uint32_t)(per_sm * sms);
    return cudaSuccess;
}

// Enqueues the MIS kernels of one round: the cluster kernel, and (with_grid) the cooperative grid kernel behind it.
// Callers that know the violated set is small skip the grid kernel: a cooperative launch is not free even as a no-op.
cudaError_t launch_mis_resample_args(const ClauseView &cv, uint32_t kmax, const uint32_t *viol, uint8_t *state,
                                     uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint32_t *bits,
                                     Counters *ctr, uint64_t seed, uint32_t round, uint32_t grid, bool with_grid,
                                     RoundNote *note, unsigned long long seq, const P2PLink *p2p, uint32_t p2p_parity,
                                     uint32_t p2p_tag, uint32_t incr_max_vars, uint32_t u_cap, cudaStream_t s)
{
    MisParams p{cv, viol, state, s_slots, sc.claim, sc.claim_stride, sc.hvar, sc.tcap, n_vars, bits, ctr, seed, round, kmax,
                cluster_cache_items(kmax), with_grid ? 1u : 0u, note, seq, sc.urec, sc.urec_cap, p2p, p2p_parity, p2p_tag,
                incr_max_vars, u_cap};
    mis_cluster_kernel<<<CL_SIZE, CL_THREADS, cluster_smem_bytes(kmax), s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !with_grid) return e;
    p.cache_items = grid_cache_items(kmax);
    void *args[] = {(void *)&p};
    return cudaLaunchCooperativeKernel((const void *)mis_grid_kernel, dim3(grid), dim3(GRID_THREADS), args,
                                       grid_smem_bytes(kmax), s);
}

cudaError_t launch_reset_counters(Counters *c, int reset_totals, cudaStream_t s)
{
    reset_counters_kernel<<<1, 1, 0, s>>>(c, reset_totals);
    return cudaGetLastError();
}

cudaError_t launch_map_ids(const ClauseView &cv, const uint32_t *slots, uint32_t n, uint32_t *out, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    map_ids_kernel<<<(n + 255) / 256, 256, 0, s>>>(cv, slots, n, out);
    return cudaGetLastError();
}

} // namespace alll
