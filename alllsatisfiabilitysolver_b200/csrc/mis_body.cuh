// mis_body.cuh -- device code of K3 + K4 (see mis.cu for the algorithm), shared by the per-round kernels of mis.cu and
// the persistent solve kernel of persist.cu.
#pragma once

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
constexpr uint32_t EXTRA = 3;                           // cached per clause besides its literals: priority, id, width | state << 8
constexpr uint32_t SMALL_U = 512;                       // violated sets up to this size: one CTA, claims in a shared-memory hash table
constexpr uint32_t HSLOTS = 8192;                       // hash slots (power of two); load factor <= 0.5 => SMALL path needs |U| * k <= 4096
constexpr uint32_t H_EMPTY = 0xFFFFFFFFu, C_FREE = 0xFFFFFFFFu, C_TAKEN = 0u;

enum : uint32_t { UNDECIDED = 0, IN_SET = 1, DROPPED = 2 };

// shared-memory words: per cached clause of mis_resample_body / all of mis_small_body in a CTA of `threads` threads
__host__ __device__ constexpr uint32_t mis_cache_words(uint32_t kmax) { return kmax + EXTRA; }
__host__ __device__ constexpr size_t mis_small_words(uint32_t threads, uint32_t kmax)
{
    return (size_t)threads * (2 * kmax + EXTRA) + 2 * (size_t)HSLOTS + 2 * (size_t)SMALL_U;   // literals + table slots | table | keys
}

struct MisParams {
    ClauseView cv;
    const uint32_t *viol;       // U as clause slots
    uint8_t *state;             // per U entry (only for entries that do not fit the shared-memory cache)
    uint32_t *s_slots;          // out: S as clause slots
    unsigned long long *claim;  // [n_vars][2]: claim words of a variable for even / odd Luby steps, FREE between rounds
    uint64_t n_vars;
    uint32_t *bits;
    Counters *ctr;
    uint64_t seed;
    uint32_t kmax;              // widest clause
    uint32_t cache_items;       // clauses per thread that fit the shared-memory cache
    uint32_t small_ok;          // the shared memory of the launch holds mis_small_body's tables (persistent solve kernel)
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
    uint32_t tune;              // TUNE_* measurement knobs
};

extern __shared__ uint32_t mis_smem[];

// Explicit global-space memory operations.  The bodies below are out-of-line functions that receive the parameter
// block by reference, so the compiler no longer knows that the pointers in it address global memory; a generic
// atomic compiles to a 25-instruction address-space dispatch, a generic load to a slower LD.  These wrappers pin the
// state space (and say exactly which operations are fire-and-forget reductions and which return a value).
// volatile, but no "memory" clobber: the wrappers keep their order among themselves (and against barriers), while
// the shared-memory traffic around them may be scheduled freely -- with a clobber every result would have to be
// stored before the next atomic may issue, which serialises the probes / claims of a clause into k round trips.
namespace gm {
__device__ __forceinline__ void red_min(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.global.min.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v));
}
__device__ __forceinline__ void red_add(unsigned int *p, unsigned int v)
{
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v));
}
__device__ __forceinline__ void red_add(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v));
}
__device__ __forceinline__ void red_or(uint32_t *p, uint32_t v)
{
    asm volatile("red.global.or.b32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v));
}
__device__ __forceinline__ void red_and(uint32_t *p, uint32_t v)
{
    asm volatile("red.global.and.b32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v));
}
__device__ __forceinline__ unsigned int add_ret(unsigned int *p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(__cvta_generic_to_global(p)), "r"(v));
    return old;
}
// L2 loads (claims and counters are written by other SMs between barriers; L1 is not coherent)
__device__ __forceinline__ unsigned long long ld_cg(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ unsigned int ld_cg(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
// read-only data of this launch (records, literal planes, violated list): may be cached and scheduled freely
__device__ __forceinline__ uint32_t ld(const uint32_t *p)
{
    uint32_t v;
    asm("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(__cvta_generic_to_global(p)));
    return v;
}
__device__ __forceinline__ void st(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v));
}
// both claim words of a variable (even / odd Luby steps, one aligned 16-byte pair) in ONE store: the independent-set
// phases of large violated sets are bound by the number of L2 write / atomic operations, not by bytes
__device__ __forceinline__ void st_pair(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.global.v2.u64 [%0], {%1, %1};" ::"l"(__cvta_generic_to_global(p)), "l"(v));
}
__device__ __forceinline__ void st(uint32_t *p, uint32_t v)
{
    asm volatile("st.global.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v));
}
} // namespace gm

// acquire / release at system scope without the sequential-consistency part of __threadfence_system() (membar.sys)
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_claim(const unsigned long long *p) { return gm::ld_cg(p); }
__device__ __forceinline__ unsigned int ld_u32(const unsigned int *p) { return gm::ld_cg(p); }

// ALLL_TRACE stamps (alll_device.cuh: Counters::dbg); called by one thread
__device__ __forceinline__ void stamp(const MisParams &p, uint32_t round, uint32_t what)
{
    if (round < DBG_ROUNDS) p.ctr->dbg[round][what] = global_ns();
}

// The one copy of Philox4x32-10 in these kernels (see the note on code size at the top of the file).
static __device__ __noinline__ Philox philox_call(uint32_t c0, uint32_t c1, uint32_t c2, uint64_t seed)
{
    return philox4x32_10(c0, c1, c2, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint32_t mis_priority(const MisParams &p, uint32_t round, uint32_t id)     // == clause_priority()
{
    return philox_call(id, round, STREAM_PRIORITY, p.seed).x >> 6;
}
// K4 for one variable: fresh fair bit (== random_bit(seed, STREAM_RESAMPLE, round, v)) written into the packed word
__device__ __forceinline__ void resample_var(const MisParams &p, uint32_t round, uint32_t v)
{
    const Philox o = philox_call(v >> 7, round, STREAM_RESAMPLE, p.seed);
    const uint32_t sel = (v >> 5) & 3u;
    const uint32_t word = sel == 0 ? o.x : sel == 1 ? o.y : sel == 2 ? o.z : o.w;
    const uint32_t mask = 1u << (v & 31u);
    if ((word >> (v & 31u)) & 1u) gm::red_or(&p.bits[v >> 5], mask);
    else gm::red_and(&p.bits[v >> 5], ~mask);
}

struct GridBarrier {
    cg::grid_group g;
    static constexpr bool GRID = true;
    __device__ __forceinline__ void sync() { g.sync(); }
};
struct ClusterBarrier {
    static constexpr bool GRID = false;
    __device__ __forceinline__ void sync() { cg::this_cluster().sync(); }
};

// Grid-wide end of a Luby step with the live count riding on the arrival word: every CTA adds (its clauses that claimed
// for the next step) << 32 | 1 to the word of the step's parity and polls it until all CTAs have arrived; the value that
// ends the poll already carries the grid-wide count.  Against grid.sync() + a counter of its own this saves the separate
// reduction atomic before the barrier and a dependent L2 round trip after it -- a Luby step of a few thousand clauses IS
// these round trips.  Two words by step parity: a CTA that has passed step s may arrive for step s+1 while slower CTAs
// still poll the word of step s, and nobody arrives for step s+2 before everyone has left that poll.  The words count up
// over the steps of a round (`seen` = high half after this CTA's previous barrier of the same parity); the thread that
// finishes the round clears them (finish_round), when no CTA is inside a Luby loop.
// Called by thread 0 of every CTA of the grid, after a __syncthreads() that made the CTA's claims of this step complete.
__device__ __forceinline__ uint32_t luby_barrier_fused(unsigned long long *words, uint32_t step, uint32_t cta_live, uint32_t n_ctas,
                                                       uint32_t (&seen)[2])
{
    unsigned long long *w = words + (step & 1u);
    const uint32_t target = (step / 2u + 1u) * n_ctas;
    __threadfence();                                   // the CTA's claim stores / reductions are ordered before its arrival
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(w)), "l"(((unsigned long long)cta_live << 32) | 1ull) : "memory");
    unsigned long long v;
    do {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(__cvta_generic_to_global(w)) : "memory");
    } while ((uint32_t)v < target);
    const uint32_t hi = (uint32_t)(v >> 32), live = hi - seen[step & 1u];
    seen[step & 1u] = hi;
    return live;
}

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
        s.slot = gm::ld(p.viol + i);
        s.rec = use_urec ? p.urec + (uint64_t)i * (p.cv.k + 1) : nullptr;
        s.k = use_urec ? p.cv.k : p.cv.width(s.slot);
    }
    return s;
}
__device__ __forceinline__ uint32_t src_id(const MisParams &p, const Src &s) { return s.rec ? gm::ld(s.rec) : p.cv.id(s.slot); }
__device__ __forceinline__ uint32_t src_lit(const MisParams &p, const Src &s, uint32_t j)
{
    return s.rec ? gm::ld(s.rec + 1 + j) : p.cv.literal(s.slot, j);
}
__device__ __forceinline__ uint32_t slot_of(const MisParams &p, uint32_t i) { return (p.p2p || !p.viol) ? i : gm::ld(p.viol + i); }

// S gets one more member: one atomic per converged group of winners instead of one per winner
__device__ __forceinline__ void append_s(const MisParams &p, uint32_t slot)
{
    const unsigned int grp = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(grp) - 1;
    unsigned int at = 0;
    if ((int)lane == leader) at = gm::add_ret(&p.ctr->n_s, (unsigned int)__popc(grp));
    at = __shfl_sync(grp, at, leader);
    p.s_slots[at + __popc(grp & ((1u << lane) - 1u))] = slot;
}

// Sharded P2P mode: wait until every rank's sweep of this round has published its records in OUR region, then
// build the prefix sums of the counts.  Returns the total; 0xFFFFFFFF on abort / timeout.  Whole CTA calls it.
__device__ __forceinline__ uint32_t p2p_wait(const MisParams &p, uint32_t *s_prefix)
{
    const P2PLink &L = *p.p2p;
    if (threadIdx.x < 32) {
        // One lane per source rank (two at world > 32): all arrival words are polled side by side -- one L2 round trip
        // for the whole round instead of one per rank on the critical path of every round.  The word carries the count.
        const uint32_t lane = threadIdx.x;
        P2PHeader *me = L.hdr[L.rank];
        const long long t_start = clock64();
        const unsigned int aval = (p.p2p_tag >> 20) + 1u;     // abort words are tagged with the solve's epoch
        unsigned int why = 0;                         // 2: a peer never published this round in time, 3: a peer aborted
        uint32_t cnt[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t q = lane + 32u * h;
            if (q < L.world) {
                for (;;) {
                    const unsigned long long w = *(volatile unsigned long long *)&me->cf[p.p2p_parity][q];
                    if ((uint32_t)(w >> 32) == p.p2p_tag) { cnt[h] = (uint32_t)w; break; }
                    if (*(volatile unsigned int *)&me->abort == aval) { why = 3; break; }
                    if (clock64() - t_start > L.timeout_cycles) { why = 2; break; }
                }
            }
        }
        fence_acq_rel_sys();                          // acquire: the records behind the flags are now visible
        if (why == 0 && lane == 0 && *(volatile unsigned int *)&me->abort == aval) why = 3;   // a peer overflowed even though every flag arrived
        const bool over = cnt[0] > L.cap || cnt[1] > L.cap;
        const unsigned int any_timeout = __ballot_sync(0xffffffffu, why == 2);
        const unsigned int any_abort = __ballot_sync(0xffffffffu, why == 3);
        const bool bad = any_timeout != 0u || any_abort != 0u || __any_sync(0xffffffffu, over);
        if (any_timeout && lane < L.world)            // tell the peers as well: they would otherwise wait for OUR next round until their own time-out
            *(volatile unsigned int *)&L.hdr[lane]->abort = aval;
        if (any_timeout && lane + 32u < L.world) *(volatile unsigned int *)&L.hdr[lane + 32u]->abort = aval;
        if (bad && lane == 0 && blockIdx.x == 0 && p.ctr->p2p_error == 0) p.ctr->p2p_error = any_timeout ? 2u : (any_abort ? 3u : 0u);
        // exclusive prefix sums of the counts in rank order: ranks 0..31 by lane, then ranks 32..63
        uint32_t inc0 = cnt[0], inc1 = cnt[1];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, inc0, o), b = __shfl_up_sync(0xffffffffu, inc1, o);
            if ((int)lane >= o) { inc0 += a; inc1 += b; }
        }
        const uint32_t total0 = __shfl_sync(0xffffffffu, inc0, 31), total1 = __shfl_sync(0xffffffffu, inc1, 31);
        if (lane < L.world) s_prefix[lane] = inc0 - cnt[0];
        if (lane + 32u < L.world) s_prefix[lane + 32u] = total0 + inc1 - cnt[1];
        if (lane == 0) s_prefix[L.world] = bad ? 0xFFFFFFFFu : total0 + total1;
    }
    __syncthreads();
    return s_prefix[L.world];
}

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
__device__ __noinline__ void mis_resample_body(const MisParams &p, uint32_t round, Barrier &bar, const uint32_t *prefix,
                                               uint32_t first, uint32_t stride, uint32_t n_u)
{
    __shared__ unsigned int s_live, s_total;
    const bool fused = Barrier::GRID && (p.tune & TUNE_CG_LUBY_BARRIER) == 0;
    uint32_t seen[2] = {0u, 0u};
    const uint32_t bd = blockDim.x, km = p.kmax, slotw = km + EXTRA;
    const bool use_urec = p.urec != nullptr && n_u <= p.urec_cap;
    // claim words of variable v: p.claim[2v] (even Luby steps) and p.claim[2v + 1] (odd steps) -- one 16-byte pair, so
    // everything a round does to a variable touches a single 32-byte sector (the sectors are cold in DRAM after a sweep)
    unsigned long long *const claim = p.claim;
    const uint32_t META = km + 2;          // cache word holding width | state << 8

    {   // ---- gather: literals -> shared memory, and the claims of step 0
        uint32_t it = 0;
        for (uint32_t i = first; i < n_u; i += stride, ++it) {
            const Src s = locate(p, prefix, i, use_urec);
            const uint32_t id = src_id(p, s);
            const uint32_t prio = mis_priority(p, round, id);
            const unsigned long long key = claim_key(0, prio, id);
            if (ALL_CACHED || it < p.cache_items) {
                const uint32_t base = it * slotw * bd + threadIdx.x;
#pragma unroll 4
                for (uint32_t j = 0; j < s.k; j++) {
                    const uint32_t l = src_lit(p, s, j);
                    mis_smem[base + j * bd] = l;
                    gm::red_min(&claim[2 * (uint64_t)(l >> 1)], key);
                }
                mis_smem[base + km * bd] = prio;
                mis_smem[base + (km + 1) * bd] = id;
                mis_smem[base + META * bd] = s.k | (UNDECIDED << 8);
            } else {
                p.state[i] = UNDECIDED;
                for (uint32_t j = 0; j < s.k; j++) gm::red_min(&claim[2 * (uint64_t)(src_lit(p, s, j) >> 1)], key);
            }
        }
    }
    if (first < 64) p.ctr->step_live[first] = 0;     // only the acting MIS kernel touches step_live
    bar.sync();
    if (first == 0) stamp(p, round, 3);

    uint32_t step = 0;
    const bool trace_steps = first == 0 && round == 0;
    for (;;) {
        const uint32_t cur = step & 1u, nxt = cur ^ 1u;
        const bool wrap = (step + 1) % TAGS == 0;
        if (trace_steps && step < 16) p.ctr->dbg_step[step][0] = global_ns();
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
                    prio = mis_smem[base + km * bd];
                    id = mis_smem[base + (km + 1) * bd];
                } else {
                    if (p.state[i] != UNDECIDED) continue;
                    s = locate(p, prefix, i, use_urec);
                    id = src_id(p, s);
                    k = s.k;
                    prio = mis_priority(p, round, id);
                }
                // claim pair of literal j
                auto pair = [&](uint32_t j) { return claim + 2 * (uint64_t)((cached ? mis_smem[base + j * bd] : src_lit(p, s, j)) >> 1); };
                if (pass == 0u) {
                    for (uint32_t j = 0; j < k; j++) {
                        unsigned long long *c = pair(j) + nxt;
                        if (ld_claim(c) != CLAIM_TAKEN) gm::st(c, CLAIM_FREE);
                    }
                    continue;
                }
                const unsigned long long key = claim_key(step, prio, id);
                bool win = true, taken = false;
#pragma unroll 4
                for (uint32_t j = 0; j < k; j++) {               // no early exit: the loads overlap
                    const unsigned long long c = ld_claim(pair(j) + cur);
                    win &= c == key;
                    taken |= c == CLAIM_TAKEN;
                }
                if (win) {
                    for (uint32_t j = 0; j < k; j++) gm::st_pair(pair(j), CLAIM_TAKEN);
                    append_s(p, slot_of(p, i));
                } else if (!taken) {
                    const unsigned long long next_key = claim_key(step + 1, prio, id);
                    for (uint32_t j = 0; j < k; j++) gm::red_min(pair(j) + nxt, next_key);
                    live++;
                    continue;
                }
                const uint32_t st = win ? IN_SET : DROPPED;
                if (cached) mis_smem[base + META * bd] = k | (st << 8);
                else p.state[i] = (uint8_t)st;
            }
            if (pass == 0u) { bar.sync(); continue; }
            if (trace_steps && step < 16) p.ctr->dbg_step[step][1] = global_ns();
            if (live) atomicAdd(&s_live, live);
        }
        __syncthreads();
        if (fused) {
            if (threadIdx.x == 0) {
                if (trace_steps && step < 16) p.ctr->dbg_step[step][2] = global_ns();
                s_total = luby_barrier_fused(p.ctr->luby_bar, step, s_live, gridDim.x, seen);
                if (trace_steps && step < 16) p.ctr->dbg_step[step][3] = global_ns();
            }
            __syncthreads();
            step++;
            if (s_total == 0) break;                                    // nobody claimed for this step: all decided
            continue;
        }
        if (threadIdx.x == 0 && s_live) gm::red_add(&p.ctr->step_live[(step + 1) & 63u], s_live);
        // the slot of step+3 (mod 64) is next written two steps from now: clear it while nobody touches it
        if (first == 0) p.ctr->step_live[(step + 3) & 63u] = 0;
        if (trace_steps && step < 16) p.ctr->dbg_step[step][2] = global_ns();
        bar.sync();
        if (trace_steps && step < 16) p.ctr->dbg_step[step][3] = global_ns();
        step++;
        if (ld_u32(&p.ctr->step_live[step & 63u]) == 0) break;      // nobody claimed for this step: all decided
    }

    if (first == 0) stamp(p, round, 4);
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
            const uint32_t v = (cached ? mis_smem[base + j * bd] : src_lit(p, s, j)) >> 1;
            gm::st_pair(&claim[2 * (uint64_t)v], CLAIM_FREE);
            if (st == IN_SET) resample_var(p, round, v);
        }
        if (st == IN_SET) resampled += k;                          // SATInstance.h:363 counts literals->size()
    }
    // warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) resampled += __shfl_down_sync(0xffffffffu, resampled, o);
    if ((threadIdx.x & 31u) == 0 && resampled) gm::red_add(&p.ctr->n_resampled_round, resampled);
    if (first == 0) {
        gm::red_add(&p.ctr->n_luby_steps, (unsigned long long)step);
        if (round < DBG_ROUNDS) p.ctr->dbg[round][7] = (unsigned long long)step << 8;
    }
}

// ---- small violated sets: the whole independent-set computation in ONE CTA's shared memory ---------------------
// (|U| <= SMALL_U and |U| * kmax <= HSLOTS / 2).  No global claim traffic, no memory fences between steps: a Luby
// step is two __syncthreads().  Exactly the same set as the large paths: the 64-bit (priority, id) keys are replaced
// by their ranks within U (ids are unique, so ranks are a strict order), claims are 32-bit (step tag | rank) words in
// an open-addressing table keyed by variable; each literal's table slot is found once and cached.
// Runs in one CTA of >= SMALL_U threads (1024 in the cluster kernel, 512 in the persistent solve kernel).
static __device__ __noinline__ void mis_small_body(const MisParams &p, uint32_t round, const uint32_t *prefix, uint32_t n_u)
{
    const uint32_t km = p.kmax, bd = blockDim.x;
    uint32_t *hvar = mis_smem + (size_t)bd * (2 * km + EXTRA);
    uint32_t *hclaim = hvar + HSLOTS;
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(hclaim + HSLOTS);     // [SMALL_U]
    __shared__ unsigned int s_cnt, s_sum;
    const uint32_t t = threadIdx.x;
    const bool mine = t < n_u;
    const bool use_urec = p.urec != nullptr && n_u <= p.urec_cap;

    for (uint32_t i = t; i < HSLOTS; i += bd) { hvar[i] = H_EMPTY; hclaim[i] = C_FREE; }
    if (t == 0) { s_cnt = 0; s_sum = 0; }
    uint32_t k = 0;
    if (mine) {
        const Src s = locate(p, prefix, t, use_urec);
        const uint32_t id = src_id(p, s);
        k = s.k;
#pragma unroll 4
        for (uint32_t j = 0; j < k; j++) mis_smem[t + j * bd] = src_lit(p, s, j);
        keys[t] = ((unsigned long long)mis_priority(p, round, id) << 32) | id;
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
    if (t == 0) stamp(p, round, 3);

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

    if (t == 0) stamp(p, round, 4);
    // ---- K4: winners redraw their variables (global bit-packed assignment) and report themselves
    const bool in_s = state == IN_SET;
    if (in_s) {
        for (uint32_t j = 0; j < k; j++) resample_var(p, round, mis_smem[t + j * bd] >> 1);
        p.s_slots[atomicAdd(&s_cnt, 1u)] = slot_of(p, t);
    }
    uint32_t resampled = in_s ? k : 0u;                        // SATInstance.h:363 counts literals->size()
    for (int o = 16; o > 0; o >>= 1) resampled += __shfl_down_sync(0xffffffffu, resampled, o);
    if ((t & 31u) == 0 && resampled) atomicAdd(&s_sum, resampled);
    __syncthreads();
    if (t == 0) {
        p.ctr->n_s = s_cnt;
        p.ctr->n_resampled_round = s_sum;
        gm::red_add(&p.ctr->n_luby_steps, (unsigned long long)step);
        if (round < DBG_ROUNDS) p.ctr->dbg[round][7] = (unsigned long long)step << 8;
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

__device__ __forceinline__ void finish_round(const MisParams &p, uint32_t round, uint32_t n_u, uint32_t path)
{
    Counters *c = p.ctr;
    stamp(p, round, 5);
    if (ld_u32(&c->incr_next)) c->n_incr_rounds += 1;      // the round that just ended was evaluated incrementally
    const unsigned int n_s = ld_u32(&c->n_s);              // (both loads in flight together; the totals below are
    const unsigned long long n_r = gm::ld_cg(&c->n_resampled_round);   //  fire-and-forget atomics: no read-modify-write chain)
    gm::red_add(&c->n_iterations, 1ull);
    gm::red_add(&c->sum_mis, (unsigned long long)n_s);     // SATInstance.h:291
    gm::red_add(&c->n_resamples, n_r);                     // SATInstance.h:313-315
    c->last_n_viol = n_u;
    c->last_n_s = n_s;
    c->last_resampled = n_r;
    c->n_viol = 0;                                         // clean slate for the next sweep
    c->n_s = 0;
    c->n_resampled_round = 0;
    c->handled_tag = p.p2p_tag;
    c->luby_bar[0] = 0;                                    // (luby_barrier_fused: no CTA is inside a Luby loop here)
    c->luby_bar[1] = 0;
    c->incr_next = (p.incr_max_vars != 0 && n_r <= p.incr_max_vars) ? 1u : 0u;
    announce(p, n_u, n_s);
    if (round < DBG_ROUNDS) { c->dbg[round][6] = global_ns(); c->dbg[round][7] |= path; }
}


} // namespace alll
