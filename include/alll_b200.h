/*
 * alll_b200.h -- C ABI of the B200 (sm_100a) parallel Moser-Tardos resampling path.
 *
 * This is the drop-in boundary underneath the reference's public solver surface
 * (/root/reference/library/include/SATInstance.h).  The reference has no FFI of its
 * own -- it is a header-only C++ template library -- so the boundary is inserted
 * directly beneath SATInstance<T>::solve / verify_validity; the same-named C++
 * headers in alllsatisfiabilitysolver_b200/include/ call only the functions below.
 * Each entry point cites the reference interface it replaces (file:line relative
 * to /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a HOST pointer unless the
 *     parameter name starts with d_ (device pointer on the handle's device);
 *   - literal encoding lit = 2*var + neg, var 0-based (example/main.cpp:168, Clause.h:40);
 *   - clause ids are positions in the concatenation of the caller's batches
 *     (SATInstance.h:60-66 receives vector<ClauseArray*>; batches are contiguous);
 *   - every function returns an alll_status; ALLL_OK == 0;
 *   - blocking calls from one host thread per handle (like SATInstance::solve);
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     ALLL_CUDA_ERROR;
 *   - no C++ exception leaves the library: a host-side failure inside a call (out of
 *     memory, no thread to start) comes back as ALLL_CUDA_ERROR with the text
 *     "host-side failure: ..." in alll_last_error / alll_multi_last_error.
 */
#ifndef ALLL_B200_H
#define ALLL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALLL_ABI_VERSION 8

#if defined(__GNUC__)
#define ALLL_API __attribute__((visibility("default")))
#else
#define ALLL_API
#endif

typedef enum {
    ALLL_OK = 0,
    ALLL_MAX_ROUNDS = 1,    /* round cap hit before all clauses were satisfied (reference: loops forever, SATInstance.h:260) */
    ALLL_EMPTY_CLAUSE = 2,  /* an empty clause can never be satisfied (Clause.h:35-45) -- refused at upload              */
    ALLL_BAD_ARG = 3,
    ALLL_CUDA_ERROR = 4,    /* a CUDA call failed, or the host side of a call did (see alll_last_error) */
    ALLL_NCCL_ERROR = 5,
    ALLL_NO_INSTANCE = 6,   /* call needs an uploaded instance */
    ALLL_CAPACITY = 7,      /* caller buffer too small */
    ALLL_PREEMPTED = 8      /* portfolio job stopped: another seed reached a satisfying assignment first */
} alll_status;

typedef struct alll_solver *alll_handle;

typedef struct {
    int32_t  device;            /* CUDA device ordinal; -1 = current device                                    */
    uint32_t sweep_smem_bytes;  /* shared-memory budget for the staged assignment; 0 = default (192 KiB)      */
    uint32_t flags;             /* ALLL_FLAG_* */
    uint32_t reserved;
} alll_config;

#define ALLL_FLAG_NO_BUCKETING 1u  /* keep clause order; gather non-resident assignment words from L2 (debug / comparison) */
/* tuning knob: bits 16..23 = L2 bulk-prefetch distance in tiles (0 = measured default 2; 0xFF = off).  The number of
 * planes streamed eagerly (5) and of resident-placed literals per clause (3) are compile-time choices; the measured
 * alternatives are recorded in profiles/. */
#define ALLL_FLAG_PREFETCH_TILES(d) ((uint32_t)(d) << 16)
/* incremental re-evaluation: builds variable->clause occurrence lists and a row-major literal
 * copy at upload (about 2x the literal bytes of extra HBM); once a round resamples few variables, the next violated set
 * is computed from the clauses containing them instead of a full sweep.  Results are bit-identical to the default
 * mode.  Bits 24..27: log2 of the switch-over divisor (default 3: incremental when <= m/8 clauses would be touched). */
#define ALLL_FLAG_INCREMENTAL 4u
#define ALLL_FLAG_FORCE_CSR 8u     /* alll_upload_csr: keep the input on the CSR kernels (warp-cooperative sweep) instead of the plane layout */
#define ALLL_FLAG_INCR_DIVISOR_LOG2(x) ((uint32_t)(x) << 24)
/* alll_solve runs the whole round loop of a plane-layout instance (k <= 8) as ONE cooperative kernel (sweep -> grid
 * barrier -> independent set + resample -> grid barrier, all rounds on the device).  This flag keeps the round loop on
 * the host instead, one kernel per phase with per-round CUDA events -- same trajectory, slower between sweeps; it is
 * what incremental, enumerated, CSR and sharded solves use anyway. */
#define ALLL_FLAG_HOST_ROUND_LOOP 16u
/* alll_solve_p2p as one persistent kernel per GPU as well (the exchange stays fused: records stored into the peers
 * during the sweep, count + flag published after the grid barrier, every GPU waits for all flags of the round).
 * Opt-in because the kernels of all ranks must be resident at the same time: set it only when every rank of the
 * sharded solve has a GPU of its own (several ranks on ONE device would wait for each other until the 3 s time-out). */
#define ALLL_FLAG_P2P_PERSISTENT 32u
/* alll_multi_*: shard every stored instance of uniform width k <= 8 with at least one clause per device (default: only
 * instances with >= 4096 clauses per device; smaller ones are solved on the first device alone).  For tests. */
#define ALLL_FLAG_FORCE_SHARDING 64u
/* The sweep of a bucketed instance with 5 <= k <= 8 streams its five eager literals per clause from four packed planes
 * (one 128-bit word per clause: leading literals relative to the clause's bucket, the others in 26 / 28 bits; 16 bytes
 * per clause instead of 20, built at upload next to the plain planes).  This flag keeps the sweep on the plain planes
 * (comparison / tests); results are identical either way. */
#define ALLL_FLAG_NO_PACKING 128u

/* Statistics{} of SATInstance.h:25-32 plus device-side counters.
 * n_iterations = resample rounds + 1 (the terminal all-satisfied sweep counts, :261,:285-287);
 * n_resamples  = sum over rounds of sum_{c in S} k_c (variables, not clauses, :363);
 * avg_mis_size = floor(sum|S| / n_iterations) (:291,:317). */
typedef struct {
    uint64_t n_iterations;
    uint64_t n_resamples;
    uint64_t avg_mis_size;
    uint64_t sum_mis_size;      /* sum |S| before the division                         */
    uint64_t n_clause_evals;    /* m * n_iterations (clauses actually evaluated in incremental mode) */
    uint64_t n_luby_steps;      /* claim/win iterations summed over rounds             */
    uint64_t n_kernel_launches; /* kernels launched by this call                       */
    double   solve_ms;          /* device-timed: first sweep launched -> last kernel done (upload excluded) */
    double   sweep_ms;          /* device time inside the clause-evaluation sweeps     */
    int32_t  status;            /* alll_status of the solve                            */
    int32_t  reserved;
    double   between_sweeps_ms; /* device time from the end of a sweep to the start of the next: independent-set
                                   and resample kernels plus launch gaps                                       */
    uint64_t n_incremental_rounds; /* rounds whose violated set came from incremental re-evaluation (0 by default);
                                      n_clause_evals then counts the clauses actually evaluated                 */
} alll_stats;

/* ---- lifetime ---------------------------------------------------------------------- */

/* Replaces: SATInstance(VariablesArray<T>*, int n_threads), SATInstance.h:51-56. */
ALLL_API int alll_create(const alll_config *cfg, alll_handle *out);
ALLL_API int alll_destroy(alll_handle h);
/* Last error text of this handle (or of the failed alll_create when h == NULL). Never NULL. */
ALLL_API const char *alll_last_error(alll_handle h);
ALLL_API int alll_abi_version(void);
/* Number of CUDA devices this process sees (0 without a usable device -- every other call then fails: no CPU fallback). */
ALLL_API int alll_device_count(int32_t *n);
/* Page-locked host memory for upload buffers (cudaMallocHost / cudaFreeHost): literals copied from it reach the full
 * PCIe rate (about 5x pageable memory).  For callers that do not link the CUDA runtime themselves -- the drop-in
 * SATInstance.h flattens the caller's Clause objects straight into such a buffer. */
ALLL_API int alll_host_alloc(uint64_t bytes, void **out);
ALLL_API int alll_host_free(void *p);

/* ---- instance upload (the flattening of vector<ClauseArray*> that solve() receives) - */

/* Fixed clause width k (1..32): lit is row-major [m][k].  Device layout: k literal-major
 * planes, clauses bucketed by variable range when the bit-packed assignment exceeds the
 * shared-memory budget.  Replaces the Clause object graph of Clause.h:17-28. */
ALLL_API int alll_upload_fixedk(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit);
/* (The host buffer is copied in 64 MB chunks on a stream of its own while ALL layout passes -- bucket count, device-side
 * scan, scatter into the planes, packed eager planes, row-major copy -- already work on the chunks that have arrived, so
 * the call ends ~0.2 ms after the last byte; page-locked `lit` gets the full PCIe rate -- 1.28 GB in 23.3 ms on B200 --
 * pageable memory about a fifth of it.  `lit` is the caller's again when the call returns.) */
/* Same with the host buffer still being PRODUCED while the call runs (the drop-in SATInstance flattens the caller's Clause
 * objects into page-locked memory with all host threads): the copy and the layout passes of a 64 MB chunk run while the
 * producer fills the next one, so upload time hides behind the flatten instead of following it.  `filled(user, c)` must block
 * until rows [0, c) of `lit` are complete and return 0 (it is asked for increasing c, up to m; from several host threads at
 * once in the multi-GPU form); a non-zero return means the producer gave up (e.g. found a clause of another width): the
 * upload is abandoned with ALLL_BAD_ARG and no instance is loaded.  filled == NULL: the buffer is complete (== alll_upload_fixedk). */
typedef int (*alll_filled_fn)(void *user, uint64_t rows_needed);
ALLL_API int alll_upload_fixedk_streamed(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit,
                                         alll_filled_fn filled, void *user);
/* Same, literals already in device memory (row-major [m][k]); the buffer is only read during the call. */
ALLL_API int alll_upload_fixedk_device(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *d_lit);
/* Variable width: off[m+1] into lit[].  Uniform-width input is routed to the fixed-k layout; ragged input whose widest
 * clause has <= 32 literals and whose padding at most doubles the literal count is padded onto the plane layout (a
 * repeated literal never changes a clause's value; true widths are kept for the statistics); anything else uses CSR. */
ALLL_API int alll_upload_csr(alll_handle h, uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit);

/* ---- enumerated clauses -------------------------------------------------------------------
 * Replaces SATInstance::solve(Clause<T>* (*)(T, unsigned short), ull n_clauses, T batch_size), SATInstance.h:70-153,
 * and the ClauseGenerator that feeds it (ClauseGenerator.h:16-114): the clauses are a pure function of their index and
 * are never stored.  A host callback cannot run on the device, so the device-side form of that callback is a functor
 * the user compiles into a sweep kernel with include/alll_generator.cuh (templates in a header: no device linking
 * across shared libraries); the library is handed the LAUNCHER of that kernel.  Every round the launcher is asked to
 * evaluate all m clauses against the bit-packed assignment and to append {index, k literals} of the violated ones to
 * `records`; the independent-set and resample kernels then work on those records.  Everything else (alll_randomize,
 * alll_set/get_assignment, alll_eval, alll_verify, alll_round, alll_solve, statistics) behaves as for stored clauses. */
typedef struct alll_gen_sweep_args {
    const uint32_t *bits;        /* device: bit-packed assignment, bit v&31 of word v>>5                             */
    uint64_t m;                  /* clause indices are [0, m)                                                        */
    uint32_t k;                  /* literals per clause                                                              */
    uint32_t grid_hint;          /* CTAs the library suggests (a multiple of the SM count)                           */
    uint32_t *records;           /* device out: [cap][k+1] = {clause index, literals} of the violated clauses        */
    uint64_t cap;                /* records that fit; count beyond it, store nothing there (the solve then fails      */
                                 /* with ALLL_CAPACITY)                                                              */
    unsigned int *n_violated;    /* device counter, 0 on entry                                                       */
    const unsigned int *skip;    /* device flag: non-zero => return at once (round enqueued behind the terminal one) */
} alll_gen_sweep_args;
/* Enqueues the sweep kernel on `cuda_stream` (a cudaStream_t) and returns the cudaError_t of the launch (0 = ok). */
typedef int (*alll_gen_launch_fn)(void *user, const alll_gen_sweep_args *args, void *cuda_stream);

/* cap_records: capacity of the violated-record buffer; 0 = m (always enough).  A random assignment violates about
 * m / 2^k clauses.  `user` must stay valid while the instance is loaded. */
ALLL_API int alll_upload_generator(alll_handle h, uint64_t n_vars, uint64_t m, uint32_t k, alll_gen_launch_fn launch,
                                   void *user, uint64_t cap_records);

/* Generators that ship with the library (tests, bench; csrc/generator.cu is also the worked example of a user TU):
 *   ALLL_GEN_UNIFORM : clause i = k variables drawn uniformly: with w = word (j & 3) of
 *                      Philox4x32-10(ctr = {lo32(i), hi32(i), 0x47454E31, j >> 2}, key = seed):
 *                      var = (w * n_vars) >> 32, neg = w & 1;
 *   ALLL_GEN_BOUNDED : every variable occurs at most d times: literal j of clause i sits at position p = i*k + j,
 *                      var = ((a*p + b) mod (n_vars*d)) mod n_vars, neg = bit j of
 *                      Philox4x32-10(ctr = {lo32(i), hi32(i), 0x47454E32, 0}, key = seed).x.  With
 *                      (x, y, z, w) = Philox4x32-10(ctr = {0, 0, 0x47454E33, 0}, key = seed): a = the first value coprime
 *                      to n_vars*d at or after ((((y << 32 | x) mod 2^26) | 2^20 | 1) mod (n_vars*d)) (0 counts as 1),
 *                      b = (w << 32 | z) mod (n_vars*d).  Needs m*k <= n_vars*d < 2^36.
 * Both take 1 <= k <= 16. */
#define ALLL_GEN_UNIFORM 0u
#define ALLL_GEN_BOUNDED 1u
ALLL_API int alll_upload_builtin_generator(alll_handle h, uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k,
                                           uint64_t seed, uint32_t d, uint64_t cap_records);
/* The same generators on the host (used to cross-check a device generator; writes k literals of clause `index`). */
ALLL_API int alll_builtin_generator_clause(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                           uint64_t index, uint32_t *lits);

/* ---- assignment (VariablesArray<T>::vars, VariablesArray.h:18-35; 1 byte per variable on the host) */

/* `bools` has n_vars bytes.  Any host memory works; a page-locked buffer (cudaMallocHost / cudaHostRegister) is copied
 * from / to directly, anything else goes through a pinned staging buffer of the handle. */
ALLL_API int alll_set_assignment(alll_handle h, const uint8_t *bools);
ALLL_API int alll_get_assignment(alll_handle h, uint8_t *bools);
/* Uniform random assignment from Philox4x32-10 keyed by `seed` (replaces VariablesArray.h:24-33). */
ALLL_API int alll_randomize(alll_handle h, uint64_t seed);

/* ---- the hot path --------------------------------------------------------------------- */

/* Violated-clause sweep + compaction (K1+K2).  Replaces SATInstance.h:273-280 /
 * Clause::is_not_satisfied (Clause.h:34-46).  ids (may be NULL) receives up to `cap`
 * violated clause ids in unspecified order; *n_violated is always the full count. */
ALLL_API int alll_eval(alll_handle h, uint32_t *ids, uint64_t cap, uint64_t *n_violated);

/* verify_validity, SATInstance.h:156-173: *valid = 1 iff no clause is violated. */
ALLL_API int alll_verify(alll_handle h, int *valid);

/* One full Moser-Tardos round on the device: sweep -> maximal independent set of the
 * violated clauses (fixed-priority Luby with atomicMin claims == greedy in ascending
 * (Philox priority, clause id) order; replaces populate_mis_parallel, SATInstance.h:391-451)
 * -> resample (replaces resample_clauses, SATInstance.h:340-365).
 * u_ids / s_ids (may be NULL) receive the violated set and the independent set, unspecified order. */
ALLL_API int alll_round(alll_handle h, uint64_t seed, uint32_t round,
               uint32_t *u_ids, uint64_t u_cap, uint64_t *n_u,
               uint32_t *s_ids, uint64_t s_cap, uint64_t *n_s,
               uint64_t *n_resampled);

/* Round loop until no clause is violated or max_rounds resample rounds were done.
 * Replaces SATInstance::solve(vector<ClauseArray*>*), SATInstance.h:60-66 -> parallel_solve :217-320.
 * The assignment on the device is the in/out state; fetch it with alll_get_assignment. */
ALLL_API int alll_solve(alll_handle h, uint64_t seed, uint64_t max_rounds, alll_stats *stats);

/* ---- clause-range sharded mode: one large instance over several GPUs -----------------
 * Each GPU (one handle, one process) uploads the contiguous clause range [id_base, id_base+m) and holds a full
 * replica of the assignment.  Per round: alll_shard_sweep on every GPU -> all-gather of the record buffers
 * (NCCL, done by the caller) -> alll_shard_round on every GPU with identical data.  Replicas stay bit-identical
 * because priorities and resample bits are keyed on global clause ids / variable ids.  No reference counterpart
 * (the reference is single-process, shared-memory OpenMP). */

/* First global clause id of this handle's clause range; ids reported and used for priorities are global. */
ALLL_API int alll_set_id_base(alll_handle h, uint64_t id_base);
/* Sweep the local range and write its violated clauses as row-major records {global id, k literals}
 * ((k+1) uint32 each) to the DEVICE buffer d_records (cap_records records).  *n_local = local |U|. */
ALLL_API int alll_shard_sweep(alll_handle h, uint32_t *d_records, uint64_t cap_records, uint64_t *n_local);
/* MIS + resample over the gathered records: n_blocks blocks of block_cap records each (DEVICE buffer), counts[b]
 * valid records in block b (HOST array).  Sum of counts == 0 is the terminal round.  Updates the statistics
 * (alll_get_stats) exactly like one iteration of alll_solve. */
ALLL_API int alll_shard_round(alll_handle h, const uint32_t *d_records, const uint64_t *counts, uint32_t n_blocks,
                     uint64_t block_cap, uint64_t seed, uint32_t round, uint64_t *n_total, uint64_t *n_s,
                     uint64_t *n_resampled);
/* Running Statistics totals of the handle (since upload / alll_reset_stats / the start of the last alll_solve). */
ALLL_API int alll_get_stats(alll_handle h, alll_stats *stats);
ALLL_API int alll_reset_stats(alll_handle h);

/* Sharded mode with the exchange fused into the kernels: the sweep stores its violated records straight into every
 * GPU's exchange region over NVLink (CUDA IPC mappings, one process per GPU) and publishes a per-round arrival
 * flag; the MIS kernel waits on the flags.  No NCCL call and no host round trip inside the round loop.
 *   1. every rank: alll_upload_fixedk* (its clause range), alll_set_id_base, alll_p2p_create -> 64-byte handle;
 *   2. exchange the handles (any host mechanism), every rank: alll_p2p_connect(all handles, rank order);
 *   3. every rank: same initial assignment (alll_randomize with one seed), then alll_solve_p2p with the same
 *      seed / max_rounds / epoch.  Ranks must not start a solve before all ranks have returned from the previous
 *      one (a host barrier between solves); epoch distinguishes the solves (same value on all ranks).
 * cap_records: records one rank may publish per round (ALLL_CAPACITY if exceeded). */
ALLL_API int alll_p2p_create(alll_handle h, uint32_t world, uint32_t rank, uint64_t cap_records, uint8_t handle_out[64]);
ALLL_API int alll_p2p_connect(alll_handle h, const uint8_t *handles);
ALLL_API int alll_solve_p2p(alll_handle h, uint64_t seed, uint64_t max_rounds, uint64_t m_global, uint32_t epoch,
                   alll_stats *stats);

/* ---- batched small instances and seed portfolio (BASELINE config 5) --------------------
 * Many independent small instances (same n_vars and k, e.g. 8,192 x 5-SAT n=10k): one CTA per instance, the whole
 * solver state in shared memory, ALL rounds inside one kernel launch.  Each instance follows exactly the round
 * specification of alll_solve (same result for the same seed).  Portfolio: one instance, many seeds, the first
 * CTA to satisfy everything claims a device-wide winner word and the others stop.  The reference has no batch
 * API; this replaces running SATInstance::solve (SATInstance.h:60-66) once per instance / per seed. */
typedef struct {
    uint64_t n_iterations;   /* Statistics semantics of SATInstance.h:25-32 */
    uint64_t n_resamples;
    uint64_t sum_mis_size;   /* avg_mis_size = sum_mis_size / n_iterations */
    int32_t  status;         /* ALLL_OK, ALLL_MAX_ROUNDS or ALLL_PREEMPTED */
    int32_t  reserved;
} alll_batch_stats;

/* clause_off[n_instances+1] indexes the rows of lit (row-major [total_clauses][k], host memory). */
ALLL_API int alll_batch_upload(alll_handle h, uint32_t n_instances, uint64_t n_vars, uint32_t k,
                      const uint64_t *clause_off, const uint32_t *lit);
/* n_jobs == n_instances (portfolio == 0: job i solves instance i with seeds[i], starting from the Philox
 * assignment of that seed), or any n_jobs with portfolio != 0 (every job solves instance 0).
 * assignments (host, [n_jobs][n_vars] bytes, may be NULL): rows of finished jobs; in portfolio mode only the
 * winner's row is written.  *winner: portfolio winner id (job_base + job) or -1.  portfolio == 2: the winner word is
 * the one shared between GPUs (alll_flag_*).  *device_ms: kernel time from CUDA events. */
ALLL_API int alll_batch_solve(alll_handle h, uint32_t n_jobs, const uint64_t *seeds, uint64_t max_rounds, int portfolio,
                     uint8_t *assignments, alll_batch_stats *stats, int32_t *winner, double *device_ms);

/* Multi-GPU portfolio (BASELINE config 5: seeds spread over the GPUs of a box, "the first GPU to find SAT wins via a
 * device flag"): ONE winner word for all ranks.  One rank creates it (device memory, -1 = open) and passes the 64-byte
 * CUDA IPC handle to the other processes, which map it over NVLink; alll_batch_solve(portfolio = 2) then claims and
 * polls that word with system-scope atomics instead of the per-handle one, writing job_base + job (give every rank a
 * distinct job_base).  A job that finishes after another rank's claim reports ALLL_PREEMPTED like a local loser.  The
 * owner resets the word between portfolios (host barrier before and after, the caller's). */
ALLL_API int alll_flag_create(alll_handle h, uint8_t handle_out[64]);
ALLL_API int alll_flag_open(alll_handle h, const uint8_t *handle);
ALLL_API int alll_flag_reset(alll_handle h);
ALLL_API int alll_flag_read(alll_handle h, int64_t *value);
ALLL_API int alll_batch_set_job_base(alll_handle h, uint32_t job_base);

/* ---- several GPUs behind ONE call from ONE process (alll_solve_sharded / alll_solve_batch) -----
 * The reference's parallel-resource knob is the constructor argument / CLI flag -p (SATInstance.h:51-56,259;
 * example/main.cpp:56-61,76-84): a single process, a single blocking call.  This group gives a C or C++ caller of the
 * drop-in headers the same shape over a list of GPUs -- no torch.distributed, no process per GPU: peer access
 * (cudaDeviceEnablePeerAccess) instead of CUDA IPC, one host thread starts every GPU's persistent solve kernel.
 *   - one large instance: contiguous clause ranges, one per device, replicated bit-packed assignment; every device
 *     uploads ONLY its own 1/N of the caller's host literals (N PCIe links in parallel); the round loop is the fused
 *     P2P exchange of alll_solve_p2p.  Needs stored clauses of uniform width k <= 8; anything else (and n_devices == 1)
 *     is solved on the first device alone -- same results, alll_multi_info tells which.
 *   - batched small instances / seed portfolio: instance (or seed) blocks per device, one first-SAT word for all devices.
 * The same device may be listed several times (simulated ranks sharing one GPU: one kernel per phase instead of
 * persistent kernels, host threads instead of one launcher) -- that is how single-GPU machines test the exchange.
 * Results are bit-identical to the single-GPU calls for the same seed, whatever the device list. */
typedef struct alll_multi *alll_multi_handle;

/* cfg (may be NULL): sweep_smem_bytes and flags apply to every device; cfg->device is ignored.
 * ALLL_FLAG_INCREMENTAL is honoured by the sharded solve as well (each device walks the occurrence lists of its own range). */
ALLL_API int alll_multi_create(const int32_t *devices, uint32_t n_devices, const alll_config *cfg, alll_multi_handle *out);
ALLL_API int alll_multi_destroy(alll_multi_handle mh);
ALLL_API const char *alll_multi_last_error(alll_multi_handle mh);      /* mh == NULL: error of the failed alll_multi_create */
/* Replaces the flattening + ownership of SATInstance::solve's vector<ClauseArray*> (SATInstance.h:60-66) for N GPUs. */
ALLL_API int alll_multi_upload_fixedk(alll_multi_handle mh, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit);
ALLL_API int alll_multi_upload_csr(alll_multi_handle mh, uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit);
/* alll_upload_fixedk_streamed over the device list: every device copies its own clause range as soon as the producer's fill
 * position has passed it. */
ALLL_API int alll_multi_upload_fixedk_streamed(alll_multi_handle mh, uint64_t n_vars, uint64_t m, uint32_t k, const uint32_t *lit,
                                               alll_filled_fn filled, void *user);
ALLL_API int alll_multi_set_assignment(alll_multi_handle mh, const uint8_t *bools);
ALLL_API int alll_multi_get_assignment(alll_multi_handle mh, uint8_t *bools);
ALLL_API int alll_multi_randomize(alll_multi_handle mh, uint64_t seed);
/* verify_validity (SATInstance.h:156-173) over all clause ranges. */
ALLL_API int alll_multi_verify(alll_multi_handle mh, int *valid);
/* SATInstance::solve -> parallel_solve (SATInstance.h:60-66, :217-320) over all devices.  stats->solve_ms is the
 * device-timed span (CUDA events), max over the devices; n_kernel_launches sums over them. */
ALLL_API int alll_multi_solve(alll_multi_handle mh, uint64_t seed, uint64_t max_rounds, alll_stats *stats);
/* {devices in use, 1 if the instance is clause-range sharded (0: first device alone), clauses of the widest range,
 *  records one rank may publish per round} */
ALLL_API int alll_multi_info(alll_multi_handle mh, uint64_t info[4]);
/* Handle of device slot i (for alll_layout_info, alll_eval, ... on one range); owned by mh. */
ALLL_API int alll_multi_device_handle(alll_multi_handle mh, uint32_t i, alll_handle *out);
/* Batched small instances / seed portfolio over the device list: instance blocks [i*n/N, (i+1)*n/N) per device (seed
 * blocks in portfolio mode, every device then holds instance 0); arguments as for alll_batch_upload / alll_batch_solve,
 * portfolio != 0 uses ONE winner word for all devices.  *device_ms: max over the devices. */
ALLL_API int alll_multi_batch_upload(alll_multi_handle mh, uint32_t n_instances, uint64_t n_vars, uint32_t k,
                                     const uint64_t *clause_off, const uint32_t *lit);
ALLL_API int alll_multi_batch_solve(alll_multi_handle mh, uint32_t n_jobs, const uint64_t *seeds, uint64_t max_rounds, int portfolio,
                                    uint8_t *assignments, alll_batch_stats *stats, int32_t *winner, double *device_ms);

/* ---- measurement hooks -------------------------------------------------------------- */

/* `reps` back-to-back sweeps of the current assignment; *ms_per_sweep is the mean kernel
 * duration from CUDA events recorded around each launch on the launching stream. */
ALLL_API int alll_time_sweep(alll_handle h, uint32_t reps, double *ms_per_sweep, uint64_t *n_violated);
/* Kernels launched through this handle so far. */
ALLL_API int alll_launch_count(alll_handle h, uint64_t *n);
/* Layout facts: {m, k (0 = CSR), n_buckets, m_padded, bytes of literal planes, smem bytes of the sweep}. */
ALLL_API int alll_layout_info(alll_handle h, uint64_t info[6]);

/* What the sweep of the uploaded instance streams: {1 if it reads the packed eager planes (see ALLL_FLAG_NO_PACKING),
 *  leading literals stored relative to their bucket (0 when unpacked), bytes streamed per clause (the remaining literals
 *  are fetched only for clauses that survive the streamed ones), minimum number of bucket-resident leading literals}. */
ALLL_API int alll_sweep_info(alll_handle h, uint64_t info[4]);

/* How the most recent host-buffer upload (alll_upload_fixedk / alll_upload_csr on a uniform-width instance) crossed the link:
 * {bytes sent host -> device, chunks sent packed, chunks sent as they are, host threads that packed}.  Packed transport:
 * with n_vars <= 2^24 a literal (2*var+neg, example/main.cpp:168) has 25 significant bits; for uploads of >= 128 MB the
 * library's host threads re-pack every ~64 MB chunk to 25 bits per literal in page-locked memory while earlier chunks are
 * on the link and a kernel expands it on the device -- 0.78 of the bytes cross PCIe, and a pageable caller buffer is read
 * by all host threads instead of by the driver's staging copy.  A chunk whose packing is not finished when the link runs
 * dry goes as it is (page-locked callers).  Environment ALLL_H2D_PACK=0 turns it off, =1 forces it at any size; handles
 * behind alll_multi_create with several devices default to off.  The uploaded instance is bit-identical either way. */
ALLL_API int alll_upload_info(alll_handle h, uint64_t info[4]);

#ifdef __cplusplus
}
#endif
#endif /* ALLL_B200_H */
