// Test program written against the reference's public API only (SATInstance.h:51,60,70,156,175; Clause.h:25;
// VariablesArray.h:23): builds clauses the way example/main.cpp:149-178 does, solves, verifies, and also
// exercises the enumerated-clause overload and writeDIMACS.  Prints one JSON line.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <random>
#include "SATInstance.h"

typedef uint32_t UINT_T;
typedef SATInstance<UINT_T>::ClauseArray ClauseArray;

static std::vector<std::vector<UINT_T>> g_clauses;
static Clause<UINT_T> *enumerate(UINT_T idx, unsigned short t_id)
{
    if (idx >= g_clauses.size()) return nullptr;
    return new Clause<UINT_T>(new std::vector<UINT_T>(g_clauses[idx]), t_id);
}

int main(int argc, char **argv)
{
    const UINT_T n = 3000; const int k = 5, d = 3, n_threads = 3;
    // bounded-occurrence instance: d slots per variable, shuffled, cut into k-tuples
    std::mt19937 rng(7);
    std::vector<UINT_T> slots;
    for (UINT_T v = 0; v < n; v++) for (int i = 0; i < d; i++) slots.push_back(v);
    std::shuffle(slots.begin(), slots.end(), rng);
    for (size_t i = 0; i + k <= slots.size(); i += k) {
        std::vector<UINT_T> c;
        bool dup = false;
        for (int j = 0; j < k; j++) { for (auto l : c) dup |= (l >> 1) == slots[i + j]; c.push_back(2 * slots[i + j] + (rng() & 1)); }
        if (!dup) g_clauses.push_back(c);
    }
    const int c_num = (int)g_clauses.size();
    int chunk = (c_num + n_threads - 1) / n_threads;
    auto clauses = new std::vector<ClauseArray *>();
    for (int t = 0; t < n_threads; t++) clauses->push_back(new ClauseArray());
    unsigned short t = 0;
    for (int c = 0; c < c_num; c++) {
        if (c > (t + 1) * chunk) t++;
        clauses->at(t)->push_back(new Clause<UINT_T>(new std::vector<UINT_T>(g_clauses[c]), t));
    }
    auto inst = new SATInstance<UINT_T>(new VariablesArray<UINT_T>(n), n_threads);
    inst->set_seed(11);
    Statistics *st = inst->solve(clauses);
    const bool valid = inst->verify_validity(clauses);
    // host-side check through the public Clause API
    bool host_ok = true;
    for (auto b : *clauses) for (auto cl : *b) host_ok &= !cl->is_not_satisfied(inst->var_arr->vars);

    // enumerated-clause overload on a fresh instance
    auto inst2 = new SATInstance<UINT_T>(new VariablesArray<UINT_T>(n), 1);
    Statistics *st2 = inst2->solve(enumerate, (ull)c_num, (UINT_T)64);
    bool host_ok2 = true;
    for (auto &c : g_clauses) { Clause<UINT_T> cl(&c, 0); host_ok2 &= !cl.is_not_satisfied(inst2->var_arr->vars); }
    if (argc > 1) { std::ofstream f(argv[1]); inst2->writeDIMACS(enumerate, (ull)c_num, &f); }

    printf("{\"n_clauses\": %llu, \"iterations\": %llu, \"resamples\": %llu, \"avg_mis\": %llu, \"thread_entries\": %zu, "
           "\"thread0\": %llu, \"valid\": %s, \"host_ok\": %s, \"status\": %d, \"iterations2\": %llu, \"host_ok2\": %s}\n",
           inst->n_clauses, st->n_iterations, st->n_resamples, st->avg_mis_size, st->n_thread_resamples.size(),
           st->n_thread_resamples[0], valid ? "true" : "false", host_ok ? "true" : "false", inst->last_status(),
           st2->n_iterations, host_ok2 ? "true" : "false");
    return (valid && host_ok && host_ok2) ? 0 : 1;
}
