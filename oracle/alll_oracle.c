/*
 * alll_oracle.c -- TEST INFRASTRUCTURE ONLY (see alll_oracle.h).
 *
 * Plain-C, single-thread restatement of the reference hot path.  Each function
 * cites the reference file:line it follows (paths relative to /root/reference).
 * Written from the behaviour of those lines, not copied: the reference works on
 * heap objects (vector<Clause*>), this works on CSR arrays.
 */
#include "alll_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* Reference restatements                                                    */
/* ------------------------------------------------------------------------ */

/* Clause.h:34-46 -- a clause is "not satisfied" iff no literal is true;
 * literal l is true when (l&1) ? !vars[l>>1] : vars[l>>1].  Early exit on the
 * first true literal.  An empty clause is never satisfied. */
int alll_oracle_clause_is_not_satisfied(const uint32_t *lits, uint64_t k, const uint8_t *vars)
{
    for (uint64_t j = 0; j < k; j++) {
        uint32_t l = lits[j];
        int value = vars[l >> 1] != 0;
        int lit_true = (l & 1u) ? !value : value;
        if (lit_true) return 0;
    }
    return 1;
}

/* SATInstance.h:273-280 -- every batch scans its clauses in order and keeps the
 * violated ones.  Batches are contiguous in the global order, so the union in
 * batch order is the ascending id list. */
uint64_t alll_oracle_sweep(uint64_t m, const uint64_t *off, const uint32_t *lit,
                           const uint8_t *vars, uint32_t *out_ids)
{
    uint64_t n = 0;
    for (uint64_t c = 0; c < m; c++) {
        if (alll_oracle_clause_is_not_satisfied(lit + off[c], off[c + 1] - off[c], vars)) {
            if (out_ids) out_ids[n] = (uint32_t)c;
            n++;
        }
    }
    return n;
}

/* SATInstance.h:156-173 */
int alll_oracle_verify(uint64_t m, const uint64_t *off, const uint32_t *lit, const uint8_t *vars)
{
    for (uint64_t c = 0; c < m; c++)
        if (alll_oracle_clause_is_not_satisfied(lit + off[c], off[c + 1] - off[c], vars)) return 0;
    return 1;
}

/* SATInstance.h:369-389 -- exists (l1,l2) with l1>>1 == l2>>1. */
int alll_oracle_dependent(const uint32_t *l1, uint64_t k1, const uint32_t *l2, uint64_t k2)
{
    for (uint64_t a = 0; a < k1; a++)
        for (uint64_t b = 0; b < k2; b++)
            if ((l1[a] >> 1) == (l2[b] >> 1)) return 1;
    return 0;
}

/* example/main.cpp:149-178 -- chunk = ceil(m / n_threads); the running batch
 * index t is bumped (once) before clause c is pushed when c > (t+1)*chunk. */
void alll_oracle_batches(uint64_t m, int n_threads, uint16_t *batch_of)
{
    if (n_threads < 1) n_threads = 1;
    uint64_t chunk = (m + (uint64_t)n_threads - 1) / (uint64_t)n_threads;
    uint64_t t = 0;
    for (uint64_t c = 0; c < m; c++) {
        if (c > (t + 1) * chunk) t++;
        batch_of[c] = (uint16_t)t;
    }
}

/* SATInstance.h:414-447 (mis empty on entry, so :392-412 is skipped).
 * Lists are arrays of ids; "erase" compacts in place, which keeps the order
 * exactly as vector::erase does. */
uint64_t alll_oracle_greedy_mis(uint64_t m, const uint64_t *off, const uint32_t *lit,
                                const uint32_t *u_ids, uint64_t n_u,
                                int n_threads, uint32_t *out_s)
{
    if (n_threads < 1) n_threads = 1;
    uint16_t *batch_of = (uint16_t *)malloc((m ? m : 1) * sizeof(uint16_t));
    alll_oracle_batches(m, n_threads, batch_of);

    /* one list per batch, as unsat_clauses is built at SATInstance.h:265-280 */
    uint64_t n_lists = (uint64_t)n_threads;
    uint32_t **list = (uint32_t **)calloc(n_lists, sizeof(uint32_t *));
    uint64_t *len = (uint64_t *)calloc(n_lists, sizeof(uint64_t));
    for (uint64_t i = 0; i < n_u; i++) len[batch_of[u_ids[i]]]++;
    for (uint64_t t = 0; t < n_lists; t++) {
        list[t] = (uint32_t *)malloc((len[t] ? len[t] : 1) * sizeof(uint32_t));
        len[t] = 0;
    }
    for (uint64_t i = 0; i < n_u; i++) {
        uint64_t t = batch_of[u_ids[i]];
        list[t][len[t]++] = u_ids[i];
    }

    uint64_t n_s = 0;
    uint64_t t = 0;
    while (n_lists > 0) {                       /* :415 */
        t = (t + 1) % n_lists;                  /* :416 */
        if (len[t] == 0) {                      /* :417-423 drop the exhausted list */
            free(list[t]);
            for (uint64_t j = t; j + 1 < n_lists; j++) { list[j] = list[j + 1]; len[j] = len[j + 1]; }
            n_lists--;
            continue;
        }
        uint32_t pick = list[t][0];             /* :425 */
        memmove(list[t], list[t] + 1, (len[t] - 1) * sizeof(uint32_t)); /* :426 */
        len[t]--;
        out_s[n_s++] = pick;                    /* :428 */
        const uint32_t *pl = lit + off[pick];
        uint64_t pk = off[pick + 1] - off[pick];
        for (uint64_t k = 0; k < n_lists; k++) { /* :430-446 */
            uint64_t w = 0;
            for (uint64_t r = 0; r < len[k]; r++) {
                uint32_t y = list[k][r];
                if (!alll_oracle_dependent(pl, pk, lit + off[y], off[y + 1] - off[y]))
                    list[k][w++] = y;
            }
            len[k] = w;
        }
    }
    free(list);
    free(len);
    free(batch_of);
    return n_s;
}

/* example/cnf_io/cnf_io.cpp:392-484 -- signed literal is true when
 * v_val[|l|-1] == (0 < l); a clause is true if any literal is; no early exit
 * inside a clause; formula false at the first false clause. */
int alll_oracle_check_signed(int64_t c_num, const int32_t *l_c_num, const int32_t *l_val,
                             const uint8_t *v_val)
{
    int64_t l = 0;
    for (int64_t c = 0; c < c_num; c++) {
        int c_val = 0;
        for (int32_t j = 0; j < l_c_num[c]; j++) {
            int32_t x = l_val[l++];
            int s_val = (0 < x);
            int64_t v_index = (0 <= x) ? x : -(int64_t)x;
            if ((v_val[v_index - 1] != 0) == s_val) c_val = 1;
        }
        if (!c_val) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------ */
/* Deterministic round specification                                          */
/* ------------------------------------------------------------------------ */

void alll_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

uint32_t alll_oracle_random_bit(uint64_t seed, uint32_t stream, uint32_t round, uint32_t v)
{
    uint32_t ctr[4] = { v >> 7, round, stream, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    alll_oracle_philox4x32_10(ctr, key, out);
    return (out[(v >> 5) & 3u] >> (v & 31u)) & 1u;
}

uint32_t alll_oracle_priority(uint64_t seed, uint32_t round, uint32_t c)
{
    uint32_t ctr[4] = { c, round, ALLL_ORACLE_STREAM_PRIORITY, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    alll_oracle_philox4x32_10(ctr, key, out);
    return out[0] >> 6;
}

void alll_oracle_randomize(uint64_t n_vars, uint64_t seed, uint8_t *vars)
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (uint64_t base = 0; base < n_vars; base += 128) {
        uint32_t ctr[4] = { (uint32_t)(base >> 7), 0u, ALLL_ORACLE_STREAM_INIT, 0u };
        uint32_t out[4];
        alll_oracle_philox4x32_10(ctr, key, out);
        for (uint64_t i = 0; i < 128 && base + i < n_vars; i++)
            vars[base + i] = (uint8_t)((out[i >> 5] >> (i & 31u)) & 1u);
    }
}

static uint64_t gcd_u64(uint64_t a, uint64_t b)
{
    while (b) { uint64_t t = a % b; a = b; b = t; }
    return a;
}

int alll_oracle_gen_materialize(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                uint32_t *lits)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    if (k < 1 || k > 16 || n_vars == 0 || n_vars > (1ull << 31)) return -1;
    if (kind == 0) {                                   /* uniform: var = (w * n) >> 32, neg = w & 1 */
        for (uint64_t i = 0; i < m; i++)
            for (uint32_t j = 0; j < k; j++) {
                const uint32_t ctr[4] = { (uint32_t)i, (uint32_t)(i >> 32), 0x47454E31u, j >> 2 };
                alll_oracle_philox4x32_10(ctr, key, out);
                const uint32_t w = out[j & 3u];
                lits[i * k + j] = 2u * (uint32_t)(((uint64_t)w * n_vars) >> 32) + (w & 1u);
            }
        return 0;
    }
    if (kind != 1 || d == 0) return -1;
    const uint64_t span = n_vars * (uint64_t)d;        /* bounded: every variable at most d times */
    if (span >= (1ull << 36) || m > span / k) return -1;
    const uint32_t setup[4] = { 0u, 0u, 0x47454E33u, 0u };
    alll_oracle_philox4x32_10(setup, key, out);
    uint64_t a = ((((((uint64_t)out[1] << 32) | out[0]) % (1ull << 26)) | (1ull << 20) | 1ull)) % span;
    if (a == 0) a = 1;
    while (gcd_u64(a, span) != 1) a = (a + 1 < span) ? a + 1 : 1;
    const uint64_t b = (((uint64_t)out[3] << 32) | out[2]) % span;
    for (uint64_t i = 0; i < m; i++) {
        const uint32_t ctr[4] = { (uint32_t)i, (uint32_t)(i >> 32), 0x47454E32u, 0u };
        alll_oracle_philox4x32_10(ctr, key, out);
        for (uint32_t j = 0; j < k; j++) {
            const uint64_t p = i * k + j;
            const uint64_t var = ((a * p + b) % span) % n_vars;
            lits[i * k + j] = 2u * (uint32_t)var + ((out[0] >> j) & 1u);
        }
    }
    return 0;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

uint64_t alll_oracle_priority_mis(uint64_t n_vars, const uint64_t *off, const uint32_t *lit,
                                  const uint32_t *u_ids, uint64_t n_u,
                                  uint64_t seed, uint32_t round,
                                  uint8_t *scratch, uint32_t *out_s)
{
    (void)n_vars;
    uint64_t *keys = (uint64_t *)malloc((n_u ? n_u : 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n_u; i++)
        keys[i] = ((uint64_t)alll_oracle_priority(seed, round, u_ids[i]) << 32) | u_ids[i];
    qsort(keys, n_u, sizeof(uint64_t), cmp_u64);
    uint64_t n_s = 0;
    for (uint64_t i = 0; i < n_u; i++) {
        uint32_t c = (uint32_t)keys[i];
        int free_ = 1;
        for (uint64_t j = off[c]; j < off[c + 1]; j++)
            if (scratch[lit[j] >> 1]) { free_ = 0; break; }
        if (!free_) continue;
        for (uint64_t j = off[c]; j < off[c + 1]; j++) scratch[lit[j] >> 1] = 1;
        out_s[n_s++] = c;
    }
    for (uint64_t i = 0; i < n_s; i++) {
        uint32_t c = out_s[i];
        for (uint64_t j = off[c]; j < off[c + 1]; j++) scratch[lit[j] >> 1] = 0;
    }
    free(keys);
    return n_s;
}

uint64_t alll_oracle_resample(const uint64_t *off, const uint32_t *lit,
                              const uint32_t *s_ids, uint64_t n_s,
                              uint64_t seed, uint32_t round, uint8_t *vars)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_s; i++) {
        uint32_t c = s_ids[i];
        for (uint64_t j = off[c]; j < off[c + 1]; j++) {
            uint32_t v = lit[j] >> 1;
            vars[v] = (uint8_t)alll_oracle_random_bit(seed, ALLL_ORACLE_STREAM_RESAMPLE, round, v);
        }
        n += off[c + 1] - off[c];   /* SATInstance.h:363 counts literals->size() */
    }
    return n;
}

uint64_t alll_oracle_round(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                           uint8_t *vars, uint64_t seed, uint32_t round,
                           uint32_t *u_out, uint32_t *s_out, uint64_t *n_s_out,
                           uint64_t *n_resampled_out)
{
    uint64_t n_u = alll_oracle_sweep(m, off, lit, vars, NULL);
    if (n_s_out) *n_s_out = 0;
    if (n_resampled_out) *n_resampled_out = 0;
    if (n_u == 0) return 0;
    uint32_t *u = u_out ? u_out : (uint32_t *)malloc(n_u * sizeof(uint32_t));
    uint32_t *s = s_out ? s_out : (uint32_t *)malloc(n_u * sizeof(uint32_t));
    alll_oracle_sweep(m, off, lit, vars, u);
    uint8_t *scratch = (uint8_t *)calloc(n_vars ? n_vars : 1, 1);
    uint64_t n_s = alll_oracle_priority_mis(n_vars, off, lit, u, n_u, seed, round, scratch, s);
    uint64_t n_r = alll_oracle_resample(off, lit, s, n_s, seed, round, vars);
    free(scratch);
    if (n_s_out) *n_s_out = n_s;
    if (n_resampled_out) *n_resampled_out = n_r;
    if (!u_out) free(u);
    if (!s_out) free(s);
    return n_u;
}

/* SATInstance.h:260-320 */
int alll_oracle_solve(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                      uint8_t *vars, uint64_t seed, uint64_t max_rounds,
                      alll_oracle_stats *stats, uint64_t *trace_u, uint64_t *trace_s)
{
    memset(stats, 0, sizeof(*stats));
    for (uint64_t c = 0; c < m; c++)
        if (off[c + 1] == off[c]) return ALLL_ORACLE_EMPTY_CLAUSE;
    int status = ALLL_ORACLE_OK;
    uint64_t round = 0;
    for (;;) {
        stats->n_iterations += 1;                       /* :261 */
        uint64_t n_s = 0, n_r = 0;
        uint64_t n_u = alll_oracle_round(n_vars, m, off, lit, vars, seed, (uint32_t)round,
                                         NULL, NULL, &n_s, &n_r);
        if (trace_u) trace_u[round] = n_u;
        if (trace_s) trace_s[round] = n_s;
        if (n_u == 0) break;                            /* :285-287 */
        stats->sum_mis_size += n_s;                     /* :291 */
        stats->n_resamples += n_r;                      /* :363, :313-315 */
        round++;
        if (round >= max_rounds) { status = ALLL_ORACLE_MAX_ROUNDS; break; }
    }
    stats->avg_mis_size = stats->sum_mis_size / stats->n_iterations; /* :317 */
    stats->n_clause_evals = m * stats->n_iterations;
    return status;
}

static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int alll_oracle_solve_greedy(uint64_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit,
                             uint8_t *vars, uint64_t seed, int n_threads, uint64_t max_rounds,
                             alll_oracle_stats *stats)
{
    (void)n_vars;
    memset(stats, 0, sizeof(*stats));
    uint32_t *u = (uint32_t *)malloc((m ? m : 1) * sizeof(uint32_t));
    uint32_t *s = (uint32_t *)malloc((m ? m : 1) * sizeof(uint32_t));
    uint64_t rng = seed, pool = 0;
    int pool_bits = 0;
    int status = ALLL_ORACLE_OK;
    uint64_t round = 0;
    for (;;) {
        stats->n_iterations += 1;
        uint64_t n_u = alll_oracle_sweep(m, off, lit, vars, u);
        if (n_u == 0) break;
        uint64_t n_s = alll_oracle_greedy_mis(m, off, lit, u, n_u, n_threads, s);
        stats->sum_mis_size += n_s;
        for (uint64_t i = 0; i < n_s; i++) {
            uint32_t c = s[i];
            for (uint64_t j = off[c]; j < off[c + 1]; j++) {
                if (pool_bits == 0) { pool = splitmix64(&rng); pool_bits = 63; } /* RBG hands out 63 bits per draw, RandomBoolGenerator.h:37-43 */
                vars[lit[j] >> 1] = (uint8_t)(pool & 1u);
                pool >>= 1; pool_bits--;
            }
            stats->n_resamples += off[c + 1] - off[c];
        }
        round++;
        if (round >= max_rounds) { status = ALLL_ORACLE_MAX_ROUNDS; break; }
    }
    stats->avg_mis_size = stats->sum_mis_size / stats->n_iterations;
    stats->n_clause_evals = m * stats->n_iterations;
    free(u); free(s);
    return status;
}
