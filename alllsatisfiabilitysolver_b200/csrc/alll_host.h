// alll_host.h -- launcher prototypes shared between the kernel translation units and the C ABI.
#pragma once

#include "../../include/alll_b200.h"
#include "alll_device.cuh"
#include "incr_body.cuh"

namespace alll {

struct MisParams;   // mis.cu

// sweep.cu
size_t sweep_planes_smem_bytes(uint32_t bucket_words);
cudaError_t configure_sweep_planes(const SweepParams &p, bool resident_all);
cudaError_t launch_sweep_planes(const SweepParams &p, bool resident_all, uint32_t grid, cudaStream_t s);
// the whole solve in one cooperative launch (plane layout, k <= 8)
cudaError_t configure_solve_persistent(const SweepParams &p, bool resident_all, uint32_t kmax, int *ok_out);
cudaError_t launch_solve_persistent(const SweepParams &p, bool resident_all, uint32_t grid, const ClauseView &cv, uint32_t kmax,
                                    uint8_t *state, uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint64_t seed,
                                    uint32_t max_rounds, uint32_t epoch, const IncrParams *incr, uint32_t visited_words,
                                    uint32_t incr_max_vars, cudaStream_t s);   // p.p2p != NULL: sharded solve; incr != NULL: incremental mode
// csr.cu (variable-width clauses: layout build, warp-cooperative sweep, whole solve in one cooperative launch)
struct CsrSweepParams;
cudaError_t launch_csr_build(const uint64_t *off, uint64_t m, uint64_t n_lit, uint64_t l_pad, uint32_t *start, uint32_t *chunk_rank,
                             cudaStream_t s);
cudaError_t configure_sweep_csr(const CsrSweepParams &p, int *ctas_per_sm);
cudaError_t launch_sweep_csr(const CsrSweepParams &p, uint32_t grid, cudaStream_t s);
cudaError_t configure_solve_persistent_csr(const CsrSweepParams &p, uint32_t kmax, int *ok_out);
cudaError_t launch_solve_persistent_csr(const CsrSweepParams &p, uint32_t grid, const ClauseView &cv, uint32_t kmax, uint8_t *state,
                                        uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint64_t seed, uint32_t max_rounds,
                                        cudaStream_t s);

// mis.cu
cudaError_t mis_configure(int device, uint32_t kmax, uint32_t *grid_out);
cudaError_t launch_mis_resample_args(const ClauseView &cv, uint32_t kmax, const uint32_t *viol, uint8_t *state,
                                     uint32_t *s_slots, const MisScratch &sc, uint64_t n_vars, uint32_t *bits,
                                     Counters *ctr, uint64_t seed, uint32_t round, uint32_t grid, bool with_grid,
                                     RoundNote *note, unsigned long long seq, const P2PLink *p2p, uint32_t p2p_parity,
                                     uint32_t p2p_tag, uint32_t incr_max_vars, uint32_t u_cap, cudaStream_t s);
cudaError_t launch_reset_counters(Counters *c, int reset_totals, cudaStream_t s);
cudaError_t launch_map_ids(const ClauseView &cv, const uint32_t *slots, uint32_t n, uint32_t *out, cudaStream_t s);   // slots == NULL: identity

// generator.cu (built-in clause generators of the enumerated-clause mode)
struct BuiltinGenerator;
const char *builtin_generator_create(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                     BuiltinGenerator **out);       // NULL on success, else the error text
void builtin_generator_destroy(BuiltinGenerator *g);
int builtin_generator_launch(void *user, const alll_gen_sweep_args *a, void *stream);
const char *builtin_generator_clause(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                     uint64_t index, uint32_t *lits);

// shard.cu
cudaError_t launch_export_records(const ClauseView &cv, const uint32_t *viol, const Counters *ctr, uint32_t *records,
                                  uint64_t cap, uint32_t grid, cudaStream_t s);
cudaError_t launch_repack_records(const uint32_t *records, uint64_t block_cap, uint32_t k, uint32_t n_blocks,
                                  const uint32_t *prefix, uint32_t *planes, uint64_t dense_cap, uint32_t *ids,
                                  uint32_t *iota, Counters *ctr, uint32_t grid, cudaStream_t s);

// incremental.cu
cudaError_t launch_incr_build(const uint32_t *planes, uint64_t m_pad, uint32_t k, uint32_t stride, const BucketSeg *segs,
                              uint32_t n_buckets, uint64_t n_vars, uint32_t *occ_off, uint32_t *cursor, uint32_t *block_sums,
                              uint32_t *rows, uint32_t *occ, uint32_t *d_total, const uint8_t *width, cudaStream_t s);
cudaError_t launch_incr_eval(const uint32_t *s_slots, const uint32_t *rows, uint32_t stride, uint32_t k,
                             const uint32_t *occ_off, const uint32_t *occ, uint32_t *visited, uint64_t visited_words,
                             const uint32_t *bits, uint32_t *viol, Counters *ctr, uint32_t grid, cudaStream_t s);

// batch.cu
size_t batch_smem_bytes(uint32_t n_vars, uint32_t n_words, uint32_t m_max);
cudaError_t launch_batch_solve(const uint32_t *planes, uint64_t m_pad, const uint32_t *inst_off, const uint32_t *inst_m,
                               uint32_t n_instances, uint32_t n_vars, uint32_t n_words, uint32_t k, uint32_t m_max,
                               const uint64_t *seeds, uint64_t max_rounds, uint32_t *bits_out, BatchJobStats *stats,
                               int portfolio, int *winner, int job_base, int shared, uint32_t n_jobs, uint32_t *retry,
                               int *n_launches, cudaStream_t s);
cudaError_t launch_batch_transpose(const uint32_t *lit, const uint64_t *src_off, const uint32_t *inst_off, uint32_t n_instances,
                                   uint32_t k, uint32_t n_vars, uint32_t *planes, uint64_t m_pad, uint32_t *err, cudaStream_t s);
cudaError_t launch_batch_unpack(const uint32_t *bits, uint32_t n_words, uint32_t n_vars, uint64_t total, uint8_t *out, cudaStream_t s);

// layout.cu
cudaError_t launch_transpose(const uint32_t *lit, uint64_t c0, uint64_t c1, uint32_t k, uint64_t n_vars, uint32_t *planes,
                             uint64_t m_pad, uint32_t *err, cudaStream_t s);               // clauses [c0, c1)
cudaError_t launch_validate_csr(const uint32_t *lit, uint64_t n_lit, uint64_t n_vars, uint32_t *err, cudaStream_t s);
uint32_t bucket_pass_ctas(uint64_t m);
uint32_t bucket_pass_clauses_per_cta();
cudaError_t launch_bucket_count(const uint32_t *lit, uint64_t m, uint64_t c0, uint64_t c1, uint32_t k, uint64_t n_vars,
                                uint32_t bucket_vars, uint32_t n_buckets, uint8_t *bkt, uint32_t *cta_counts, uint32_t *err,
                                cudaStream_t s);                                            // clauses [c0, c1) of m
cudaError_t launch_rows8(const uint32_t *planes, uint64_t m_pad, uint32_t k, uint4 *rows, cudaStream_t s);
cudaError_t launch_pack_eager(const uint32_t *planes, uint64_t m_pad, const BucketSeg *segs, uint32_t n_buckets,
                              uint32_t bucket_vars, uint32_t rb, uint32_t *packed, cudaStream_t s);
cudaError_t launch_bucket_scan(uint32_t *cta_counts, uint64_t m, uint64_t c0, uint64_t c1, uint32_t n_buckets, BucketSeg *segs_out,
                               uint32_t *tile_cursor, cudaStream_t s);
cudaError_t launch_bucket_scatter(const uint32_t *lit, uint64_t m, uint64_t c0, uint64_t c1, uint32_t k, uint32_t bucket_vars,
                                  uint32_t n_buckets, const uint8_t *bkt, const uint32_t *cta_base, uint32_t *planes, uint64_t m_pad,
                                  uint32_t *orig_id, uint32_t *min_resident, uint32_t resident_cap, const uint8_t *width_in,
                                  uint8_t *width_out, uint32_t *packed, uint4 *rows, cudaStream_t s);
bool bucket_scatter_fuses(uint32_t k, bool with_widths);
cudaError_t launch_pack_bits(const uint8_t *bools, uint64_t n_vars, uint32_t *bits, uint32_t n_words_alloc, cudaStream_t s);
cudaError_t launch_unpack_bits(const uint32_t *bits, uint64_t n_vars, uint8_t *bools, cudaStream_t s);
cudaError_t launch_randomize(uint32_t *bits, uint64_t n_vars, uint32_t n_words_alloc, uint64_t seed, cudaStream_t s);
cudaError_t launch_fill_u64(unsigned long long *p, uint64_t n, unsigned long long value, cudaStream_t s);
cudaError_t launch_unpack25(const uint8_t *lo3, const uint8_t *hi, uint32_t *out, uint64_t n, cudaStream_t s);    // packed H2D transport

// hostpack.cpp: n literals -> lo3[3 n] (low three bytes of each) + hi[ceil(n / 8)] (bit 24 of eight literals per byte); returns their OR
uint32_t host_pack25(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi);

// capi.cu: what the single-process multi-GPU layer (multi.cu) needs beyond the C ABI
int internal_upload_fixedk_streamed(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit, alll_filled_fn filled,
                                    void *user, uint64_t filled_base);
void *internal_p2p_region(alll_handle h);
int internal_p2p_create_local(alll_handle h, uint32_t world, uint32_t rank, uint64_t cap_records);   // exchange region, no IPC export
bool internal_csr_is_uniform(const uint64_t *off, uint64_t m);
int internal_p2p_connect_ptrs(alll_handle h, void *const *regions);         // regions[rank]: exchange-region base of every rank
bool internal_p2p_persistent_possible(alll_handle h);
int internal_solve_p2p_begin(alll_handle h, uint64_t seed, uint64_t max_rounds, uint32_t epoch, uint64_t *launches0);   // enqueue only
int internal_solve_p2p_end(alll_handle h, uint64_t m_global, uint64_t launches0, alll_stats *stats);                    // wait + statistics
int *internal_flag_ptr(alll_handle h);
void internal_h2d_pack_default_off(alll_handle h);
int internal_flag_attach(alll_handle h, int *word);
int internal_device(alll_handle h);

} // namespace alll
