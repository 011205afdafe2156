// mis.cu -- K3 (maximal independent set of the violated clauses) + K4 (resample), one cooperative kernel.
//
// Replaces populate_mis_parallel (SATInstance.h:391-451; greedy, O(|S|*|U|*k^2), one OpenMP fork/join per
// picked clause) and resample_clauses (SATInstance.h:340-365).
//
// K3: fixed-priority Luby.  Every violated clause c gets the key (Philox priority(seed, round, id), id).
//     step:  A) undecided clauses first drop out if one of their variables is TAKEN, otherwise
//               atomicMin their key into claim[var] for every variable they touch;
//            B) a clause that still owns all its claims joins S and marks its variables TAKEN.
//     Iterated to a fixed point this is exactly the greedy independent set in ascending key order
//     (what oracle/alll_oracle.c:alll_oracle_priority_mis computes sequentially) -- independent and
//     maximal like the reference's set (SATInstance.h:415-447), and a pure function of (seed, round, U):
//     no dependence on thread schedule, compaction order or shard count.
//     Keys carry a step tag in the top bits that decreases every step, so claims of clauses that lost
//     or dropped out in earlier steps are undercut without a reset pass.
// K4: every variable of every clause in S is redrawn from Philox(seed, RESAMPLE, round, var); bits are
//     written with atomicOr/atomicAnd on the packed word (idempotent, so a variable occurring twice in a
//     clause is harmless).  The same pass restores claim[var] = FREE for everything U touched.
//
// |U| <= SMALL_U runs in CTA 0 alone with __syncthreads(); larger sets use grid-wide barriers.
#include <cooperative_groups.h>

#include "alll_device.cuh"

namespace cg = cooperative_groups;

namespace alll {

constexpr uint32_t MIS_THREADS = 256;
constexpr uint32_t SMALL_U = 4096;

enum : uint8_t { UNDECIDED = 0, IN_SET = 1, DROPPED = 2 };

struct MisParams {
    ClauseView cv;
    const uint32_t *viol;       // U as clause slots
    uint8_t *state;             // per U entry
    uint32_t *s_slots;          // out: S as clause slots
    unsigned long long *claim;  // [n_vars], FREE between rounds
    uint32_t *bits;
    Counters *ctr;
    uint64_t seed;
    uint32_t round;
    uint32_t do_resample;       // 0: MIS only (alll_round with inspection still resamples; kept for tests)
};

// claim[] and the counters are written by other SMs between barriers: read them at L2 (L1 is not coherent).
__device__ __forceinline__ unsigned long long ld_claim(const unsigned long long *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned int ld_u32(const unsigned int *p) { return __ldcg(p); }

template <bool SMALL>
__device__ __forceinline__ void phase_barrier(cg::grid_group &grid)
{
    if (SMALL) __syncthreads(); else grid.sync();
}

template <bool SMALL>
__device__ void mis_resample_body(const MisParams &p, cg::grid_group &grid, uint32_t n_u)
{
    __shared__ unsigned int s_live;
    const uint32_t stride = SMALL ? blockDim.x : gridDim.x * blockDim.x;
    const uint32_t first = SMALL ? threadIdx.x : blockIdx.x * blockDim.x + threadIdx.x;

    for (uint32_t i = first; i < n_u; i += stride) p.state[i] = UNDECIDED;
    if (first < 64) p.ctr->step_live[first] = 0;     // only this kernel touches step_live; visible after the first barrier
    phase_barrier<SMALL>(grid);

    uint32_t step = 0;
    for (;;) {
        if (step != 0 && step % TAGS == 0) {
            // the step tag wraps: stale claims would now undercut fresh ones, so clear what the survivors touch
            for (uint32_t i = first; i < n_u; i += stride) {
                if (p.state[i] != UNDECIDED) continue;
                const uint32_t slot = p.viol[i];
                const uint32_t k = p.cv.width(slot);
                for (uint32_t j = 0; j < k; j++) {
                    const uint32_t v = p.cv.literal(slot, j) >> 1;
                    if (ld_claim(&p.claim[v]) != CLAIM_TAKEN) p.claim[v] = CLAIM_FREE;
                }
            }
            phase_barrier<SMALL>(grid);
        }
        // ---- phase A: drop out next to winners, otherwise claim
        if (threadIdx.x == 0) s_live = 0;
        __syncthreads();
        uint32_t live = 0;
        for (uint32_t i = first; i < n_u; i += stride) {
            if (p.state[i] != UNDECIDED) continue;
            const uint32_t slot = p.viol[i];
            const uint32_t k = p.cv.width(slot);
            bool taken = false;
            for (uint32_t j = 0; j < k; j++) {
                const uint32_t v = p.cv.literal(slot, j) >> 1;
                if (ld_claim(&p.claim[v]) == CLAIM_TAKEN) { taken = true; break; }
            }
            if (taken) { p.state[i] = DROPPED; continue; }
            const uint32_t id = p.cv.id(slot);
            const unsigned long long key = claim_key(step, clause_priority(p.seed, p.round, id), id);
            for (uint32_t j = 0; j < k; j++) atomicMin(&p.claim[p.cv.literal(slot, j) >> 1], key);
            live++;
        }
        if (live) atomicAdd(&s_live, live);
        __syncthreads();
        if (threadIdx.x == 0 && s_live) atomicAdd(&p.ctr->step_live[step & 63u], s_live);
        phase_barrier<SMALL>(grid);
        const unsigned int total_live = ld_u32(&p.ctr->step_live[step & 63u]);
        if (total_live == 0) break;

        // ---- phase B: owners of all their claims win
        for (uint32_t i = first; i < n_u; i += stride) {
            if (p.state[i] != UNDECIDED) continue;
            const uint32_t slot = p.viol[i];
            const uint32_t k = p.cv.width(slot);
            const uint32_t id = p.cv.id(slot);
            const unsigned long long key = claim_key(step, clause_priority(p.seed, p.round, id), id);
            bool win = true;
            for (uint32_t j = 0; j < k; j++) {
                // another winner may be storing TAKEN to ITS variables concurrently; ours still read == key
                if (ld_claim(&p.claim[p.cv.literal(slot, j) >> 1]) != key) { win = false; break; }
            }
            if (!win) continue;
            p.state[i] = IN_SET;
            for (uint32_t j = 0; j < k; j++) p.claim[p.cv.literal(slot, j) >> 1] = CLAIM_TAKEN;
            p.s_slots[atomicAdd(&p.ctr->n_s, 1u)] = slot;
        }
        // the slot of step+2 (mod 64) is reused two steps from now: clear it while nobody reads it
        if (first == 0) p.ctr->step_live[(step + 2) & 63u] = 0;
        phase_barrier<SMALL>(grid);
        step++;
    }

    // ---- K4 + claim reset.  All claim reads of this round are behind the last barrier.
    unsigned long long resampled = 0;
    for (uint32_t i = first; i < n_u; i += stride) {
        const uint32_t slot = p.viol[i];
        const uint32_t k = p.cv.width(slot);
        const bool in_s = p.state[i] == IN_SET && p.do_resample;
        for (uint32_t j = 0; j < k; j++) {
            const uint32_t v = p.cv.literal(slot, j) >> 1;
            p.claim[v] = CLAIM_FREE;
            if (in_s) {
                const uint32_t mask = 1u << (v & 31u);
                if (random_bit(p.seed, STREAM_RESAMPLE, p.round, v)) atomicOr(&p.bits[v >> 5], mask);
                else atomicAnd(&p.bits[v >> 5], ~mask);
            }
        }
        if (p.state[i] == IN_SET) resampled += k;          // SATInstance.h:363 counts literals->size()
    }
    // warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) resampled += __shfl_down_sync(0xffffffffu, resampled, o);
    if ((threadIdx.x & 31u) == 0 && resampled) atomicAdd(&p.ctr->n_resampled_round, resampled);
    if (first == 0) atomicAdd(&p.ctr->n_luby_steps, (unsigned long long)step);
}

// One launch per round.  n_iterations counts every sweep, including the terminal one (SATInstance.h:261).
__global__ void __launch_bounds__(MIS_THREADS) mis_resample_kernel(const MisParams p)
{
    cg::grid_group grid = cg::this_grid();
    const uint32_t n_u = ld_u32(&p.ctr->n_viol);  // written by the sweep kernel that ran before us
    if (n_u == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            p.ctr->n_iterations += 1;
            p.ctr->last_n_viol = 0;
            p.ctr->last_n_s = 0;
            p.ctr->last_resampled = 0;
        }
        return;
    }
    if (n_u <= SMALL_U) {
        if (blockIdx.x != 0) return;
        mis_resample_body<true>(p, grid, n_u);
        __syncthreads();
    } else {
        mis_resample_body<false>(p, grid, n_u);
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        Counters *c = p.ctr;
        c->n_iterations += 1;
        const unsigned int n_s = ld_u32(&c->n_s);
        const unsigned long long n_r = __ldcg(&c->n_resampled_round);
        c->sum_mis += n_s;                                     // SATInstance.h:291
        c->n_resamples += n_r;                                 // SATInstance.h:313-315
        c->last_n_viol = n_u;
        c->last_n_s = n_s;
        c->last_resampled = n_r;
        c->n_viol = 0;                                         // clean slate for the next sweep
        c->n_s = 0;
        c->n_resampled_round = 0;
    }
}

// Per-round scratch reset + (optionally) totals reset, fused into one tiny launch ahead of the sweep.
__global__ void reset_counters_kernel(Counters *c, int reset_totals)
{
    c->n_viol = 0;
    c->n_s = 0;
    c->n_resampled_round = 0;
    c->last_n_viol = 0;
    c->last_n_s = 0;
    c->last_resampled = 0;
    if (reset_totals) {
        c->n_iterations = 0;
        c->sum_mis = 0;
        c->n_resamples = 0;
        c->n_luby_steps = 0;
    }
}

// slots -> caller clause ids (for alll_eval / alll_round outputs)
__global__ void map_ids_kernel(const uint32_t *slots, const uint32_t *orig_id, uint32_t n, uint32_t *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = orig_id ? orig_id[slots[i]] : slots[i];
}

cudaError_t mis_max_grid(int device, uint32_t *grid_out)
{
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mis_resample_kernel, MIS_THREADS, 0);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm > 4) per_sm = 4;          // 4 x 256 threads per SM is plenty for an atomics-bound kernel
    *grid_out = (uint32_t)(per_sm * sms);
    return cudaSuccess;
}

cudaError_t launch_mis_resample_args(const ClauseView &cv, const uint32_t *viol, uint8_t *state, uint32_t *s_slots,
                                     unsigned long long *claim, uint32_t *bits, Counters *ctr, uint64_t seed,
                                     uint32_t round, uint32_t grid, cudaStream_t s)
{
    MisParams p{cv, viol, state, s_slots, claim, bits, ctr, seed, round, 1u};
    void *args[] = {(void *)&p};
    return cudaLaunchCooperativeKernel((const void *)mis_resample_kernel, dim3(grid), dim3(MIS_THREADS), args, 0, s);
}

cudaError_t launch_reset_counters(Counters *c, int reset_totals, cudaStream_t s)
{
    reset_counters_kernel<<<1, 1, 0, s>>>(c, reset_totals);
    return cudaGetLastError();
}

cudaError_t launch_map_ids(const uint32_t *slots, const uint32_t *orig_id, uint32_t n, uint32_t *out, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    map_ids_kernel<<<(n + 255) / 256, 256, 0, s>>>(slots, orig_id, n, out);
    return cudaGetLastError();
}

} // namespace alll
