/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * A thin extern "C" shim around the UNMODIFIED reference headers
 * (/root/reference/library/include/*.h) and the vendored DIMACS reader
 * (/root/reference/example/cnf_io/cnf_io.cpp), compiled where those sources
 * lie by oracle/Makefile into oracle/_ref/liballl_ref.so.  No reference source
 * is copied into this repository; this file only *calls* the reference API.
 *
 * It exists to (a) pin oracle/alll_oracle.c against the real reference,
 * (b) generate tests/golden/ fixtures, and (c) serve as the CPU baseline
 * (`bench.py --impl reference`, cpu_baseline.kind == "reference").
 *
 * Built with -fno-access-control so the private helpers
 * (dependent_clauses, populate_mis_parallel; SATInstance.h:369-451) can be
 * driven directly.
 */
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "SATInstance.h"
#include "cnf_io/cnf_io.h"

typedef uint32_t UINT_T;
typedef SATInstance<UINT_T>::ClauseArray ClauseArray;

struct RefInstance {
    SATInstance<UINT_T> *inst = nullptr;
    std::vector<ClauseArray *> *clauses = nullptr;       // n_threads batches
    std::vector<Clause<UINT_T> *> flat;                  // global order
    std::unordered_map<const Clause<UINT_T> *, uint32_t> id_of;
    int n_threads = 1;
};

extern "C" {

/* Builds the object graph the way the reference CLI does (example/main.cpp:149-181):
 * n_threads batches, chunk = ceil(m/n_threads), split test `c > (t+1)*chunk`. */
void *ref_create(uint32_t n_vars, uint64_t m, const uint64_t *off, const uint32_t *lit, int n_threads)
{
    auto *r = new RefInstance;
    if (n_threads < 1) n_threads = 1;
    r->n_threads = n_threads;
    r->clauses = new std::vector<ClauseArray *>();
    for (int t = 0; t < n_threads; t++) r->clauses->push_back(new ClauseArray());
    int chunk_size = (int)std::ceil((double)m / (double)n_threads);
    unsigned short t = 0;
    r->flat.reserve(m);
    for (uint64_t c = 0; c < m; c++) {
        auto *literals = new std::vector<UINT_T>(lit + off[c], lit + off[c + 1]);
        if ((long long)c > (long long)(t + 1) * chunk_size) t += 1;
        auto *cl = new Clause<UINT_T>(literals, t);
        r->clauses->at(t)->push_back(cl);
        r->flat.push_back(cl);
        r->id_of.emplace(cl, (uint32_t)c);
    }
    r->inst = new SATInstance<UINT_T>(new VariablesArray<UINT_T>(n_vars), n_threads);
    return r;
}

void ref_destroy(void *h)
{
    auto *r = (RefInstance *)h;
    for (auto *cl : r->flat) { delete cl->literals; delete cl; }
    for (auto *b : *r->clauses) delete b;
    delete r->clauses;
    delete[] r->inst->var_arr->vars;
    delete r->inst->var_arr;
    delete r->inst;
    delete r;
}

void ref_set_assignment(void *h, const uint8_t *bools)
{
    auto *r = (RefInstance *)h;
    for (UINT_T i = 0; i < r->inst->n_vars; i++) r->inst->var_arr->vars[i] = bools[i] != 0;
}

void ref_get_assignment(void *h, uint8_t *bools)
{
    auto *r = (RefInstance *)h;
    for (UINT_T i = 0; i < r->inst->n_vars; i++) bools[i] = r->inst->var_arr->vars[i] ? 1 : 0;
}

/* Clause::is_not_satisfied (Clause.h:34) over all clauses in global order. */
uint64_t ref_sweep(void *h, uint32_t *out_ids)
{
    auto *r = (RefInstance *)h;
    uint64_t n = 0;
    const bool *vars = r->inst->var_arr->vars;
    for (size_t c = 0; c < r->flat.size(); c++) {
        if (r->flat[c]->is_not_satisfied(vars)) {
            if (out_ids) out_ids[n] = (uint32_t)c;
            n++;
        }
    }
    return n;
}

/* SATInstance::verify_validity (SATInstance.h:156) */
int ref_verify(void *h)
{
    auto *r = (RefInstance *)h;
    omp_set_num_threads(r->n_threads);
    return r->inst->verify_validity(r->clauses) ? 1 : 0;
}

/* SATInstance::dependent_clauses (private, SATInstance.h:369) */
int ref_dependent(void *h, uint32_t c1, uint32_t c2)
{
    auto *r = (RefInstance *)h;
    return r->inst->dependent_clauses(r->flat[c1], r->flat[c2]) ? 1 : 0;
}

/* Builds the per-batch violated lists exactly as parallel_solve does
 * (SATInstance.h:265-280) and hands them to the reference's own
 * populate_mis_parallel (SATInstance.h:391).  Returns picks in pick order. */
uint64_t ref_greedy_mis(void *h, uint32_t *out_ids)
{
    auto *r = (RefInstance *)h;
    omp_set_num_threads(r->n_threads);
    auto *unsat = new std::vector<ClauseArray *>;
    const bool *vars = r->inst->var_arr->vars;
    for (int t = 0; t < r->n_threads; t++) {
        auto *lst = new ClauseArray();
        for (auto *cl : *r->clauses->at(t))
            if (cl->is_not_satisfied(vars)) lst->push_back(cl);
        unsat->push_back(lst);
    }
    ClauseArray mis;
    r->inst->populate_mis_parallel(unsat, &mis, false);   // frees `unsat`
    uint64_t n = 0;
    for (auto *cl : mis) out_ids[n++] = r->id_of.at(cl);
    return n;
}

/* SATInstance::solve (SATInstance.h:60), timed as example/main.cpp:216-222 does.
 * stats_out = {n_iterations, n_resamples, avg_mis_size, n_clauses_field}. */
double ref_solve(void *h, uint64_t *stats_out)
{
    auto *r = (RefInstance *)h;
    r->inst->n_clauses = 0;   // solve() accumulates; reset so repeated calls stay meaningful
    auto start = std::chrono::high_resolution_clock::now();
    Statistics *st = r->inst->solve(r->clauses);
    auto stop = std::chrono::high_resolution_clock::now();
    stats_out[0] = st->n_iterations;
    stats_out[1] = st->n_resamples;
    stats_out[2] = st->avg_mis_size;
    stats_out[3] = r->inst->n_clauses;
    delete st;
    return std::chrono::duration<double>(stop - start).count();
}

/* Re-randomise the assignment the way VariablesArray's constructor does
 * (VariablesArray.h:24-33) so repeated solves start from fresh random points. */
void ref_rerandomize(void *h)
{
    auto *r = (RefInstance *)h;
    auto *fresh = new VariablesArray<UINT_T>(r->inst->n_vars);
    std::memcpy(r->inst->var_arr->vars, fresh->vars, r->inst->n_vars * sizeof(bool));
    delete[] fresh->vars;
    delete fresh;
}

/* verify_validity timed `reps` times: a pure clause-eval sweep rate. Returns seconds total. */
double ref_time_verify(void *h, int reps, int *all_valid)
{
    auto *r = (RefInstance *)h;
    omp_set_num_threads(r->n_threads);
    int ok = 1;
    auto start = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < reps; i++) ok &= r->inst->verify_validity(r->clauses) ? 1 : 0;
    auto stop = std::chrono::high_resolution_clock::now();
    if (all_valid) *all_valid = ok;
    return std::chrono::duration<double>(stop - start).count();
}

int ref_num_procs(void) { return omp_get_num_procs(); }

/* cnf_io pass 1 (cnf_io.cpp:487).  Returns 1 on error, like the reference. */
int ref_cnf_header_read(const char *path, int *v_num, int *c_num, int *l_num)
{
    return cnf_header_read(std::string(path), v_num, c_num, l_num) ? 1 : 0;
}

/* cnf_io pass 2 (cnf_io.cpp:126).  Caller sizes l_c_num[c_num], l_val[l_num]
 * (+ slack; see SURVEY section 5 on the trailing-`0` out-of-bounds hazard). */
int ref_cnf_data_read(const char *path, int v_num, int c_num, int l_num, int *l_c_num, int *l_val)
{
    return cnf_data_read(std::string(path), v_num, c_num, l_num, l_c_num, l_val) ? 1 : 0;
}

/* cnf_evaluate (cnf_io.cpp:392): the differently coded checker. */
int ref_cnf_evaluate(int v_num, int c_num, int l_num, const int *l_c_num, int *l_val, const uint8_t *v_val)
{
    std::vector<char> tmp(v_val, v_val + v_num);
    bool *b = new bool[v_num > 0 ? v_num : 1];
    for (int i = 0; i < v_num; i++) b[i] = v_val[i] != 0;
    bool res = cnf_evaluate(v_num, c_num, l_num, l_c_num, l_val, b);
    delete[] b;
    return res ? 1 : 0;
}

} // extern "C"
