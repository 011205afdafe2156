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

    // mixed clause widths (general DIMACS input, Clause.h:20-28): the flatten leaves its uniform-width fast path
    bool host_ok3 = true;
    {
        auto ragged = new std::vector<ClauseArray *>();
        for (int t = 0; t < 2; t++) ragged->push_back(new ClauseArray());
        std::mt19937 r2(21);
        for (int c = 0; c < 900; c++) {
            const int w = 2 + (int)(r2() % 5);
            auto *ls = new std::vector<UINT_T>();
            while ((int)ls->size() < w) {
                const UINT_T v = r2() % n;
                bool dup = false;
                for (auto l : *ls) dup |= (l >> 1) == v;
                if (!dup) ls->push_back(2 * v + (r2() & 1));
            }
            ragged->at(c < 500 ? 0 : 1)->push_back(new Clause<UINT_T>(ls, (unsigned short)(c < 500 ? 0 : 1)));
        }
        auto inst3 = new SATInstance<UINT_T>(new VariablesArray<UINT_T>(n), 2);
        inst3->set_seed(5);
        Statistics *st3 = inst3->solve(ragged);
        host_ok3 = inst3->last_status() == 0 && inst3->verify_validity(ragged) && inst3->verify_last() && st3->n_iterations >= 1;
        for (auto b : *ragged) for (auto cl : *b) host_ok3 &= !cl->is_not_satisfied(inst3->var_arr->vars);
    }
    if (!host_ok3) { printf("{\"valid\": false, \"ragged\": false}\n"); return 1; }

    printf("{\"n_clauses\": %llu, \"iterations\": %llu, \"resamples\": %llu, \"avg_mis\": %llu, \"thread_entries\": %zu, "
           "\"thread0\": %llu, \"valid\": %s, \"host_ok\": %s, \"status\": %d, \"iterations2\": %llu, \"host_ok2\": %s}\n",
           inst->n_clauses, st->n_iterations, st->n_resamples, st->avg_mis_size, st->n_thread_resamples.size(),
           st->n_thread_resamples[0], valid ? "true" : "false", host_ok ? "true" : "false", inst->last_status(),
           st2->n_iterations, host_ok2 ? "true" : "false");
    return (valid && host_ok && host_ok2) ? 0 : 1;
}
