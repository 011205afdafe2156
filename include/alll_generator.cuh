/*
 * alll_generator.cuh -- device side of the enumerated-clause solve (alll_upload_generator, alll_b200.h).
 *
 * Reference counterpart: the clause callback `Clause<T>* (*)(T index, unsigned short t_id)` consumed by
 * SATInstance::solve (SATInstance.h:70-153) through ClauseGenerator::yieldRandomUNSATClauseBatch
 * (ClauseGenerator.h:37-70): clause `index` is produced on demand, tested against the assignment, and only
 * violated clauses are kept.  Here that callback is a device functor
 *
 *     struct MyClauses {
 *         // fill lits[0..K) of clause `index` with 2*var+neg (Clause.h:40 encoding); must be a pure function
 *         __device__ void operator()(uint64_t index, uint32_t (&lits)[K]) const;
 *     };
 *
 * and the whole "yield, test, keep the violated" pass is one kernel, alll_gen::sweep_kernel<K, MyClauses>: no clause is
 * ever stored, the sweep reads only the bit-packed assignment (L2-resident) and is bound by integer throughput instead
 * of HBM.  The user's translation unit (nvcc, sm_100a) instantiates the kernel and gives the library its launcher:
 *
 *     static int launch(void* user, const alll_gen_sweep_args* a, void* stream) {
 *         return alll_gen::launch_sweep<K>(*static_cast<MyClauses*>(user), *a, stream);
 *     }
 *     alll_upload_generator(h, n_vars, m, K, launch, &my_clauses, 0);
 *     alll_randomize(h, seed); alll_solve(h, seed, max_rounds, &stats);
 *
 * Visiting order: the reference walks the index range with an additive stride (ClauseGenerator.h:45,109) to
 * decorrelate consecutive batches; a full sweep has no batches, and the independent set chosen from the violated clauses
 * does not depend on the order they were found in, so indices are simply visited grid-stride.
 */
#ifndef ALLL_GENERATOR_CUH
#define ALLL_GENERATOR_CUH

#include <cstdint>
#include <cuda_runtime.h>

#include "alll_b200.h"

namespace alll_gen {

constexpr int SWEEP_THREADS = 256;

struct Philox4 {
    uint32_t x, y, z, w;
};

// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the whole path uses; handy for generators whose
// clauses are pseudo-random functions of the index.
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                           uint32_t k1)
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
    return Philox4{c0, c1, c2, c3};
}

// One pass over all clause indices: generate, evaluate (Clause.h:34-46: violated <=> every literal is false), compact the
// violated ones warp-wide (ballot + popc, one atomic per warp that found any) and store their records.
template <int K, class Gen>
__global__ void __launch_bounds__(SWEEP_THREADS) sweep_kernel(const Gen gen, const alll_gen_sweep_args a)
{
    static_assert(K >= 1 && K <= 32, "1 <= K <= 32");
    if (*reinterpret_cast<const volatile unsigned int *>(a.skip)) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * SWEEP_THREADS;
    const uint64_t first = (uint64_t)blockIdx.x * SWEEP_THREADS + threadIdx.x;
    // whole warps stay in the loop together: the ballot below needs all 32 lanes
    for (uint64_t base = first - lane; base < a.m; base += stride) {
        const uint64_t index = base + lane;
        uint32_t lits[K];
        bool violated = false;
        if (index < a.m) {
            gen(index, lits);
            violated = true;
            // Early exit like Clause.h:36-43: a lane stops looking things up at its first true literal (2 lookups per
            // clause on average instead of K).  The lookups are scattered 32-byte sectors of the L2-resident
            // assignment, and an SM's load pipe takes about one such sector per clock: they, not the arithmetic, bound
            // this kernel.
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (violated) {
                    const uint32_t var = lits[j] >> 1;
                    const uint32_t value = (__ldg(a.bits + (var >> 5)) >> (var & 31u)) & 1u;
                    violated = value == (lits[j] & 1u);     // literal 2*var+neg is true iff value != neg
                }
            }
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, violated);
        if (mask == 0) continue;
        unsigned int at = 0;
        if (lane == 0) at = atomicAdd(a.n_violated, (unsigned int)__popc(mask));
        at = __shfl_sync(0xFFFFFFFFu, at, 0) + __popc(mask & ((1u << lane) - 1u));
        if (violated && at < a.cap) {
            uint32_t *rec = a.records + (uint64_t)at * (K + 1);
            rec[0] = (uint32_t)index;
#pragma unroll
            for (int j = 0; j < K; j++) rec[1 + j] = lits[j];
        }
    }
}

// Launches sweep_kernel<K, Gen> on `stream`; returns the cudaError_t as int (what alll_gen_launch_fn must return).
template <int K, class Gen>
inline int launch_sweep(const Gen &gen, const alll_gen_sweep_args &a, void *stream)
{
    if (a.k != (uint32_t)K) return (int)cudaErrorInvalidValue;
    if (a.m == 0) return 0;
    const uint64_t want = (a.m + SWEEP_THREADS - 1) / SWEEP_THREADS;
    const uint32_t grid = (uint32_t)(want < a.grid_hint ? want : a.grid_hint);
    sweep_kernel<K, Gen><<<grid ? grid : 1u, SWEEP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(gen, a);
    return (int)cudaGetLastError();
}

} // namespace alll_gen

#endif
