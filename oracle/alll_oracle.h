/*
 * alll_oracle.h -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * CPU restatement (plain C, single thread) of the reference's parallel
 * Moser-Tardos path (/root/reference/library/include/{Clause,SATInstance}.h)
 * plus the deterministic Philox round specification that the CUDA path
 * implements.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product
 * (alllsatisfiabilitysolver_b200/) never links, imports or calls it.
 *
 * Parity pinning: the reference ships no tests, fixtures or golden vectors
 * (SURVEY.md section 4 / 8c), so this restatement is pinned by executing the
 * unmodified reference headers (oracle/_ref/liballl_ref.so, built by
 * oracle/Makefile from the sources where they lie under /root/reference) on
 * the same inputs -- see tests/test_oracle_vs_ref.py and tests/golden/.
 *
 * Literal encoding everywhere: lit = 2*var + neg, var 0-based
 * (reference example/main.cpp:168, Clause.h:40).
 * Clause storage: CSR, off[m+1] (uint64) into lit[] (uint32).
 */
#ifndef ALLL_ORACLE_H
#define ALLL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Philox stream tags of the deterministic round specification. */
#define ALLL_ORACLE_STREAM_INIT     0u
#define ALLL_ORACLE_STREAM_RESAMPLE 1u
#define ALLL_ORACLE_STREAM_PRIORITY 2u

#define ALLL_ORACLE_OK          0
#define ALLL_ORACLE_MAX_ROUNDS  1
#define ALLL_ORACLE_EMPTY_CLAUSE 2

typedef struct {
    uint64_t n_iterations;   /* rounds + 1 (terminal sweep counted): SATInstance.h:261,285-287 */
    uint64_t n_resamples;    /* sum over rounds of sum_{c in S} k_c: SATInstance.h:363,313-315  */
    uint64_t avg_mis_size;   /* floor(sum|S| / n_iterations): SATInstance.h:291,317            */
    uint64_t sum_mis_size;   /* sum|S| before the division (extra, for tests)                  */
    uint64_t n_clause_evals; /* m * n_iterations (extra, for throughput)                       */
} alll_oracle_stats;

/* ---- reference restatements ------------------------------------------- */

/* Clause::is_not_satisfied, Clause.h:34-46.  vars is 1 byte per variable. */
int alll_oracle_clause_is_not_satisfied(const uint32_t *lits, uint64_t k, const uint8_t *vars);

/* Violated-clause sweep over the concatenation of all batches,
 * SATInstance.h:273-280.  Writes ascending clause ids; returns |U|. */
uint64_t alll_oracle_sweep(uint64_t m, const uint64_t *off, const uint32_t *lit,
                           const uint8_t *vars, uint32_t *out_ids);

/* verify_validity, SATInstance.h:156-173: 1 iff no clause is violated. */
int alll_oracle_verify(uint64_t m, const uint64_t *off, const uint32_t *lit, const uint8_t *vars);

/* dependent_clauses, SATInstance.h:369-389: share a variable, sign ignored. */
int alll_oracle_dependent(const uint32_t *l1, uint64_t k1, const uint32_t *l2, uint64_t k2);

/* Batch index of every clause as example/main.cpp:149-178 assigns it
 * (chunk = ceil(m/n_threads); the split test is `c > (t+1)*chunk`). */
void alll_oracle_batches(uint64_t m, int n_threads, uint16_t *batch_of);

/* populate_mis_parallel on the non-stream path (mis empty on entry),
 * SATInstance.h:414-447: round-robin over per-batch violated lists starting
 * at (0+1)%n_lists, pop front, erase all dependents everywhere.
 * u_ids: violated ids in ascending order (each batch list is the ascending
 * subsequence with that batch index).  Writes the picked ids in pick order;
 * returns |S|. */
uint64_t alll_oracle_greedy_mis(uint64_t m, const uint64_t *off, const uint32_t *lit,
                                const uint32_t *u_ids, uint64_t n_u,
                                int n_threads, uint32_t *out_s);

/* cnf_evaluate semantics, example/cnf_io/cnf_io.cpp:392-484, on raw signed
 * DIMACS literals (1-based, sign = polarity), no early exit inside a clause.
 * This is the independently coded checker. */
int alll_oracle_check_signed(int64_t c_num, const int32_t *l_c_num, const int32_t *l_val,
                             const uint8_t *v_val);

/* ---- deterministic round specification (what the CUDA path computes) ---- */

/* Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 constants). */
void alll_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Fair bit for variable v: Philox(ctr={v>>7, round, stream, 0}, key=seed),
 * word (v>>5)&3, bit v&31. */
uint32_t alll_oracle_random_bit(uint64_t seed, uint32_t stream, uint32_t round, uint32_t v);

/* 26-bit priority of clause c in a round: Philox(ctr={c, round, PRIORITY, 0})[0] >> 6.
 * MIS order is lexicographic in (priority, clause id). */
uint32_t alll_oracle_priority(uint64_t seed, uint32_t round, uint32_t c);

/* vars[v] = random_bit(seed, INIT, 0, v) for all v (replaces VariablesArray.h:24-33). */
void alll_oracle_randomize(uint64_t n_vars, uint64_t seed, uint8_t *vars);

/* Enumerated-clause instances (SATInstance.h:70-153 takes the clauses as a callback of the index): the two index ->
 * clause functions the CUDA library ships (include/alll_b200.h, ALLL_GEN_UNIFORM = 0 / ALLL_GEN_BOUNDED = 1), restated
 * from that specification with plain 64-bit % arithmetic.  Writes all m clauses, row-major [m][k].
 * Returns 0, or -1 on parameters outside the specification. */
int alll_oracle_gen_materialize(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                uint32_t *lits);

/* Maximal independent set of U = greedy in ascending (priority, id) order;
 * identical to iterating fixed-priority Luby steps to a fixed point.
 * Writes S ascending by (priority,id); returns |S|.  scratch: n_vars bytes, zeroed on entry and exit. */
uint64_t alll_oracle_priority_mis(uint64_t n_vars, const uint64_t *off, const uint32_t *lit,
                                  const uint32_t *u_ids, uint64_t n_u,
                                  uint64_t seed, uint32_t round,
                                  uint8_t *scratch, uint32_t *out_s);

/* Resample, SATInstance.h:353-364 with the RNG replaced by random_bit(seed, RESAMPLE, round, var).
 * Returns sum of k_c (the n_resamples increment). */
uint64_t alll_oracle_resample(const uint64_t *off, const uint32_t *lit,
                              const uint32_t *s_ids, uint64_t n_s,
                              uint64_t seed, uint32_t round, uint8_t *vars);

/* One full round: sweep -> (stop if none) -> priority MIS -> resample.
 * u_out/s_out may be NULL.  Returns |U| (0 means satisfied, nothing changed). */
uint64_t alll_oracle_round(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                           uint8_t *vars, uint64_t seed, uint32_t round,
                           uint32_t *u_out, uint32_t *s_out, uint64_t *n_s_out,
                           uint64_t *n_resampled_out);

/* parallel_solve round loop, SATInstance.h:260-320, with the deterministic
 * MIS/resample above.  vars in/out.  trace_u / trace_s (length >= max_rounds+1,
 * may be NULL) receive |U| and |S| per round. */
int alll_oracle_solve(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                      uint8_t *vars, uint64_t seed, uint64_t max_rounds,
                      alll_oracle_stats *stats, uint64_t *trace_u, uint64_t *trace_s);

/* Same loop but with the reference's greedy MIS (alll_oracle_greedy_mis) and
 * a caller-seeded splitmix64 bit source for resampling: the closest
 * restatement of the reference's own trajectory law (used for the
 * round-count distribution checks). */
int alll_oracle_solve_greedy(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                             uint8_t *vars, uint64_t seed, int n_threads, uint64_t max_rounds,
                             alll_oracle_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
