// shard.cu -- kernels of the clause-range sharded mode ("largest instance").
//
// Every GPU holds a contiguous clause range [id_base, id_base + m) and a full replica of the bit-packed
// assignment.  Per round each GPU sweeps its range (sweep.cu), exports its violated clauses as records
// {global id, k literals}, the records are all-gathered (NCCL, driven by the host side), and every GPU runs
// the identical MIS + resample (mis.cu) on the full violated set -- Philox priorities keyed on global ids make
// the replicas stay bit-identical with no second exchange.
#include "alll_device.cuh"

namespace alll {

// viol[0..n_viol) (local slots) -> records[i] = {global id, lit_0 .. lit_{k-1}}, row-major, (k+1) words each.
__global__ void __launch_bounds__(256) export_records_kernel(ClauseView cv, const uint32_t *viol, const Counters *ctr,
                                                              uint32_t *records, uint64_t cap)
{
    const uint32_t n = min((uint64_t)__ldcg(&ctr->n_viol), cap);
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t slot = viol[i];
        uint32_t *rec = records + (uint64_t)i * (cv.k + 1);
        rec[0] = cv.id(slot);
        for (uint32_t j = 0; j < cv.k; j++) rec[1 + j] = cv.literal(slot, j);
    }
}

struct RepackArgs {
    uint32_t n_blocks;
    uint32_t prefix[MAX_SHARDS + 1];   // exclusive prefix sums of the per-block record counts
};

// gathered row-major records [n_blocks][block_cap][k+1] -> dense literal planes [k][dense_cap] + ids + identity list;
// also publishes the total as ctr->n_viol so the MIS kernels find it where the sweep would have left it.
__global__ void __launch_bounds__(256) repack_records_kernel(const uint32_t *records, uint64_t block_cap, uint32_t k,
                                                              RepackArgs a, uint32_t *planes, uint64_t dense_cap,
                                                              uint32_t *ids, uint32_t *iota, Counters *ctr)
{
    const uint32_t total = a.prefix[a.n_blocks];
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->n_viol = total;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < total; d += stride) {
        uint32_t b = 0;
        while (a.prefix[b + 1] <= d) ++b;
        const uint32_t *rec = records + ((uint64_t)b * block_cap + (d - a.prefix[b])) * (k + 1);
        ids[d] = rec[0];
        iota[d] = d;
        for (uint32_t j = 0; j < k; j++) planes[(uint64_t)j * dense_cap + d] = rec[1 + j];
    }
}

cudaError_t launch_export_records(const ClauseView &cv, const uint32_t *viol, const Counters *ctr, uint32_t *records,
                                  uint64_t cap, uint32_t grid, cudaStream_t s)
{
    export_records_kernel<<<grid, 256, 0, s>>>(cv, viol, ctr, records, cap);
    return cudaGetLastError();
}

cudaError_t launch_repack_records(const uint32_t *records, uint64_t block_cap, uint32_t k, uint32_t n_blocks,
                                  const uint32_t *prefix, uint32_t *planes, uint64_t dense_cap, uint32_t *ids,
                                  uint32_t *iota, Counters *ctr, uint32_t grid, cudaStream_t s)
{
    RepackArgs a;
    a.n_blocks = n_blocks;
    for (uint32_t b = 0; b <= n_blocks; b++) a.prefix[b] = prefix[b];
    repack_records_kernel<<<grid, 256, 0, s>>>(records, block_cap, k, a, planes, dense_cap, ids, iota, ctr);
    return cudaGetLastError();
}

} // namespace alll
