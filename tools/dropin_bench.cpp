// dropin_bench.cpp -- end-to-end time of the reference-facing C++ entry point, SATInstance::solve(vector<ClauseArray*>*)
// (SATInstance.h:60-66), on BASELINE-shaped instances: the caller owns heap Clause objects built exactly the way
// example/main.cpp:149-178 builds them (n_threads batches, one `new Clause(new vector)` per clause); the timed call
// flattens them, uploads, solves on the GPU(s) and writes var_arr->vars.
//
//   dropin_bench [--n 1000000] [--k 7] [--d 28] [--threads 16] [--gpus 1[,2,...]] [--steps 3] [--seed 1] [--ragged 0|1]
//
// Prints one JSON line per entry of --gpus (the object graph is built once).  Build: see tools/build_dropin_bench.sh (g++ against include/ + liballl_b200.so; no nvcc needed).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "SATInstance.h"

typedef uint32_t UINT_T;
typedef SATInstance<UINT_T>::ClauseArray ClauseArray;

static double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    uint64_t n = 1000000, seed = 1;
    int k = 7, d = 28, n_threads = (int)std::max(1u, std::thread::hardware_concurrency()), steps = 3;
    int ragged = 0;          // 1: the LAST clause loses a literal -- a streamed flatten finds the other width only at its very end
    std::vector<int> gpu_list{1};
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string a = argv[i];
        if (a == "--n") n = strtoull(argv[i + 1], nullptr, 0);
        else if (a == "--k") k = atoi(argv[i + 1]);
        else if (a == "--d") d = atoi(argv[i + 1]);
        else if (a == "--threads") n_threads = atoi(argv[i + 1]);
        else if (a == "--gpus") {
            gpu_list.clear();
            for (char *tok = strtok(argv[i + 1], ","); tok; tok = strtok(nullptr, ",")) gpu_list.push_back(atoi(tok));
        }
        else if (a == "--steps") steps = atoi(argv[i + 1]);
        else if (a == "--seed") seed = strtoull(argv[i + 1], nullptr, 0);
        else if (a == "--ragged") ragged = atoi(argv[i + 1]);
    }
    // configuration model: d slots per variable, shuffle, cut into k-tuples, drop repeats
    double t0 = now_ms();
    std::vector<UINT_T> slots(n * (uint64_t)d);
    for (uint64_t v = 0; v < n; v++)
        for (int i = 0; i < d; i++) slots[v * d + i] = (UINT_T)v;
    std::mt19937_64 rng(0xA111 + seed);
    for (uint64_t i = slots.size() - 1; i > 0; i--) {                  // Fisher-Yates with a 64-bit engine
        const uint64_t j = rng() % (i + 1);
        std::swap(slots[i], slots[j]);
    }
    const uint64_t m_all = slots.size() / k;
    // clause objects, batched like example/main.cpp:149-178
    auto clauses = new std::vector<ClauseArray *>();
    for (int t = 0; t < n_threads; t++) clauses->push_back(new ClauseArray());
    const uint64_t chunk = (m_all + n_threads - 1) / n_threads;
    uint64_t m = 0;
    for (uint64_t c = 0; c < m_all; c++) {
        const UINT_T *s = &slots[c * k];
        bool dup = false;
        for (int a = 0; a < k && !dup; a++)
            for (int b = a + 1; b < k; b++) dup |= s[a] == s[b];
        if (dup) continue;
        auto *lits = new std::vector<UINT_T>(k);
        const uint64_t bits = rng();
        for (int j = 0; j < k; j++) (*lits)[j] = 2 * s[j] + (UINT_T)((bits >> j) & 1);
        const unsigned short t = (unsigned short)std::min<uint64_t>(c / chunk, n_threads - 1);
        clauses->at(t)->push_back(new Clause<UINT_T>(lits, t));
        m++;
    }
    std::vector<UINT_T>().swap(slots);
    if (ragged && m > 0) {
        for (int t = n_threads - 1; t >= 0; t--)
            if (!clauses->at(t)->empty()) { clauses->at(t)->back()->literals->pop_back(); break; }
    }
    const double build_ms = now_ms() - t0;

    bool all = true;
    for (int gpus : gpu_list) {
    auto *vars = new VariablesArray<UINT_T>((UINT_T)n, 99);
    auto *inst = new SATInstance<UINT_T>(vars, n_threads);
    inst->set_gpus(gpus);
    std::vector<bool> start(vars->vars, vars->vars + n);
    double first_ms = 0, sum_ms = 0, sum_dev = 0, sum_flat = 0;
    unsigned long long iters = 0;
    bool ok = true;
    for (int i = -1; i < steps; i++) {                                 // step -1: warm-up (context, buffer growth)
        for (uint64_t v = 0; v < n; v++) vars->vars[v] = start[v];
        inst->set_seed(1000 + i);
        const double a = now_ms();
        Statistics *st = inst->solve(clauses);
        const double b = now_ms();
        if (i < 0) first_ms = b - a;
        else { sum_ms += b - a; sum_dev += inst->last_device_stats().solve_ms; sum_flat += inst->last_flatten_ms(); iters += st->n_iterations; }
        ok = ok && inst->last_status() == ALLL_OK;
        delete st;
    }
    const double v0 = now_ms();
    const bool valid_dev = inst->verify_last();
    const double v1 = now_ms();
    const bool valid_full = inst->verify_validity(clauses);           // flatten + upload + sweep, like the reference's call
    const double v2 = now_ms();
    bool host_ok = true;                                               // the caller's own check through the public Clause API
    for (auto b : *clauses)
        for (auto cl : *b) host_ok = host_ok && !cl->is_not_satisfied(vars->vars);
    const double v3 = now_ms();
    printf("{\"n\": %llu, \"m\": %llu, \"k\": %d, \"d\": %d, \"host_threads\": %d, \"gpus_requested\": %d, \"gpus_in_use\": %d, "
           "\"steps\": %d, \"build_objects_ms\": %.1f, \"first_call_ms\": %.2f, \"solve_call_ms\": %.3f, \"flatten_ms\": %.3f, "
           "\"device_solve_ms\": %.3f, \"sweeps_per_solve\": %.1f, \"clause_evals_per_sec_e2e\": %.4g, \"verify_last_ms\": %.3f, "
           "\"verify_validity_ms\": %.3f, \"host_check_ms\": %.1f, \"all_ok\": %s}\n",
           (unsigned long long)n, (unsigned long long)m, k, d, n_threads, gpus, inst->gpus_in_use(), steps, build_ms, first_ms,
           sum_ms / steps, sum_flat / steps, sum_dev / steps, (double)iters / steps, (double)m * iters / (sum_ms * 1e-3),
           v1 - v0, v2 - v1, v3 - v2, (ok && valid_dev && valid_full && host_ok) ? "true" : "false");
    fflush(stdout);
    all = all && ok && valid_dev && valid_full && host_ok;
    delete inst;
    delete[] vars->vars;
    delete vars;
    }
    return all ? 0 : 1;
}
