// incr_body.cuh -- per-round device code of the incremental re-evaluation (see incremental.cu), shared by
// incr_eval_kernel and the persistent solve kernel of persist.cu.
#pragma once

#include "alll_device.cuh"

namespace alll {

struct IncrParams {
    const uint32_t *s_slots;       // S of the round that just finished (clause slots; sharded P2P mode: indices into that round's gathered U)
    const uint32_t *rows;          // [m_pad][stride]
    uint32_t stride, k;
    const uint32_t *occ_off, *occ;
    uint32_t *visited;
    const uint32_t *bits;
    uint32_t *viol;
    Counters *ctr;
};

// Sharded P2P mode of the persistent solve kernel: where S's literals come from (the records of the PREVIOUS round in
// our own exchange region) and where the new violated clauses go (records in every GPU's region, like the sweep's).
struct IncrP2P {
    const P2PLink *link;           // NULL: single-GPU form (S as slots, violated slots into IncrParams::viol)
    const uint32_t *prev_prefix;   // exclusive prefix sums of the previous round's per-rank record counts
    uint32_t par;                  // parity of THIS round (previous round's records: par ^ 1)
    const uint32_t *orig_id;       // slot -> caller id (NULL = identity), + id_base = global clause id
    uint32_t id_base;
    Counters *ctr;
    uint32_t abort_val;            // what to write into the peers' abort words on a capacity overflow ((epoch & 0xFFF) + 1)
};

// One warp per (clause of S, literal): walks the occurrence list of that variable, 32 clauses at a time.
// n_s: |S| of the round that just ended; n_viol: where |U| of the new round is accumulated.
__device__ __forceinline__ void incr_eval_body(const IncrParams &p, uint32_t n_s, unsigned int *n_viol, const IncrP2P &x)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t n_items = n_s * p.k;
    const uint32_t w = p.k + 1;
    unsigned long long evals = 0;
    for (uint32_t item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < n_items; item += warps_total) {
        const uint32_t s_entry = p.s_slots[item / p.k];
        uint32_t v;
        if (x.link) {                                   // entry i of the previous round's gathered U: a record in our region
            const P2PLink &L = *x.link;
            uint32_t q = 0;
            while (x.prev_prefix[q + 1] <= s_entry) ++q;
            const uint32_t *rec = L.rec[L.rank] + (((uint64_t)(x.par ^ 1u) * L.world + q) * L.cap + (s_entry - x.prev_prefix[q])) * w;
            v = __ldcg(rec + 1 + item % p.k) >> 1;
        } else {
            v = p.rows[(uint64_t)s_entry * p.stride + item % p.k] >> 1;
        }
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
            uint32_t bal = __ballot_sync(0xffffffffu, violated);
            if (bal) {
                unsigned int g = 0;
                if (lane == 0) g = atomicAdd(n_viol, (unsigned int)__popc(bal));
                g = __shfl_sync(0xffffffffu, g, 0);
                if (!x.link) {
                    if (violated) p.viol[g + __popc(bal & ((1u << lane) - 1u))] = c;
                } else {
                    // fused exchange, as in the sweep's flush: the record {global id, k literals} of every violated clause
                    // goes into this rank's receive slot in every GPU's region; lanes 0..k write one record (36 bytes,
                    // contiguous) per peer
                    const P2PLink &L = *x.link;
                    const uint32_t n_new = __popc(bal);
                    if (lane == 0) g_remote_dirty = 1u;
                    if ((uint64_t)g + n_new > L.cap) {
                        if (lane == 0) { x.ctr->p2p_error = 1; for (uint32_t q = 0; q < L.world; q++) L.hdr[q]->abort = x.abort_val; }
                    } else {
                        uint32_t at = g;
                        while (bal) {
                            const int src = __ffs(bal) - 1;
                            bal &= bal - 1;
                            const uint32_t cs = __shfl_sync(0xffffffffu, c, src);
                            if (lane < w) {
                                const uint32_t word = lane == 0 ? (x.orig_id ? __ldg(x.orig_id + cs) : cs) + x.id_base
                                                                : p.rows[(uint64_t)cs * p.stride + lane - 1];
                                const uint64_t off = (((uint64_t)x.par * L.world + L.rank) * L.cap + at) * w + lane;
                                for (uint32_t q = 0; q < L.world; q++) L.rec[q][off] = word;
                            }
                            ++at;
                        }
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_down_sync(0xffffffffu, evals, o);
    if (lane == 0 && evals) atomicAdd(&p.ctr->n_evals_incr, evals);
}

} // namespace alll
