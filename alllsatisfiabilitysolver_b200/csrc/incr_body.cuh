// incr_body.cuh -- per-round device code of the incremental re-evaluation (see incremental.cu), shared by
// incr_eval_kernel and the persistent solve kernel of persist.cu.
#pragma once

#include "alll_device.cuh"

namespace alll {

struct IncrParams {
    const uint32_t *s_slots;       // S of the round that just finished (clause slots)
    const uint32_t *rows;          // [m_pad][stride]
    uint32_t stride, k;
    const uint32_t *occ_off, *occ;
    uint32_t *visited;
    const uint32_t *bits;
    uint32_t *viol;
    Counters *ctr;
};

// One warp per (clause of S, literal): walks the occurrence list of that variable, 32 clauses at a time.
// n_s: |S| of the round that just ended; n_viol: where |U| of the new round is accumulated.
__device__ __forceinline__ void incr_eval_body(const IncrParams &p, uint32_t n_s, unsigned int *n_viol)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t n_items = n_s * p.k;
    unsigned long long evals = 0;
    for (uint32_t item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < n_items; item += warps_total) {
        const uint32_t slot_s = p.s_slots[item / p.k];
        const uint32_t v = p.rows[(uint64_t)slot_s * p.stride + item % p.k] >> 1;
        const uint32_t lo = p.occ_off[v], hi = p.occ_off[v + 1];
        for (uint32_t base = lo; base < hi; base += 32) {
            const uint32_t e = base + lane;
            bool violated = false;
            uint32_t c = 0;
            if (e < hi) {
                c = p.occ[e];
                const uint32_t bit = 1u << (c & 31u);
                if (!(atomicOr(&p.visited[c >> 5], bit) & bit)) {       // first visit this round: evaluate
                    evals++;
                    const uint32_t *row = p.rows + (uint64_t)c * p.stride;
                    violated = true;
                    for (uint32_t j0 = 0; j0 < p.k && violated; j0 += 4) {
                        const uint4 L = *reinterpret_cast<const uint4 *>(row + j0);
                        const uint32_t l[4] = {L.x, L.y, L.z, L.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if (j0 + q < p.k && violated) {
                                const uint32_t var = l[q] >> 1;
                                if (((__ldcg(p.bits + (var >> 5)) >> (var & 31u)) ^ l[q]) & 1u) violated = false;
                            }
                        }
                    }
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, violated);
            if (bal) {
                unsigned int g = 0;
                if (lane == 0) g = atomicAdd(n_viol, (unsigned int)__popc(bal));
                g = __shfl_sync(0xffffffffu, g, 0);
                if (violated) p.viol[g + __popc(bal & ((1u << lane) - 1u))] = c;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_down_sync(0xffffffffu, evals, o);
    if (lane == 0 && evals) atomicAdd(&p.ctr->n_evals_incr, evals);
}

} // namespace alll
