// SATInstance.h -- drop-in for library/include/SATInstance.h of the reference.
//
// Same public surface (SATInstance.h:25-32, 45-56, 60-66, 70-153, 156-173, 175-203 of the reference):
//   struct Statistics { n_iterations, n_resamples, avg_mis_size, n_thread_resamples }
//   SATInstance<T>(VariablesArray<T>*, int n_threads); T n_vars; ull n_clauses; VariablesArray<T>* var_arr;
//   Statistics* solve(vector<ClauseArray*>*);
//   Statistics* solve(Clause<T>* (*)(T, unsigned short), ull n_clauses, T batch_size);
//   bool verify_validity(vector<ClauseArray*>*) const;
//   void writeDIMACS(Clause<T>* (*)(T, unsigned short), ull n_clauses, ofstream*);
// so example/main.cpp of the reference compiles unchanged against this directory.
//
// What is different underneath: there is no OpenMP loop here.  solve() flattens the caller's batches once
// (global clause id = position in the concatenation), uploads them through the C ABI (include/alll_b200.h)
// and the whole round loop -- sweep, independent set, resample -- runs on the B200.  There is no CPU
// fallback: a missing device or library error throws std::runtime_error.
//
// Deliberate deviations from the reference: n_clauses is assigned, not accumulated,
// by solve() (Q6); an optional seed makes runs reproducible (Q8); a round cap turns the reference's
// non-termination on unsatisfiable input into a status (Q11); n_thread_resamples has n_threads entries with
// the device total in entry 0 (there are no host worker threads to attribute resamples to).
#ifndef ALLL_B200_SATINSTANCE_H
#define ALLL_B200_SATINSTANCE_H

#include <cstdint>
#include <fstream>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#else
#include <thread>
// The reference header pulls in <omp.h> (SATInstance.h:16) and its CLI calls omp_get_num_procs()
// (example/main.cpp:77-78); keep that call compiling without OpenMP.
static inline int omp_get_num_procs() { const unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }
#endif
#include <cmath>

#include "../../include/alll_b200.h"
#include "Clause.h"
#include "ClauseGenerator.h"
#include "RandomBoolGenerator.h"
#include "VariablesArray.h"

using namespace std;

typedef struct Statistics {
    ull n_iterations = 0;               // resample rounds + 1: the terminal all-satisfied sweep counts
    ull n_resamples = 0;                // variables resampled (sum of clause widths over all independent sets)
    ull avg_mis_size = 0;               // floor(sum |S| / n_iterations)
    vector<ull> n_thread_resamples;     // [0] = device total; remaining n_threads-1 entries are 0
} Statistics;

template <typename T>
class SATInstance {
public:
    using ClauseArray = typename Clause<T>::ClauseArray;

    T n_vars;
    ull n_clauses = 0;

    VariablesArray<T> *var_arr;

    SATInstance(VariablesArray<T> *var_arr, int n_threads) : n_vars(var_arr->n_vars), var_arr(var_arr), n_threads(n_threads < 1 ? 1 : n_threads) {}

    SATInstance(const SATInstance &) = delete;
    SATInstance &operator=(const SATInstance &) = delete;

    ~SATInstance()
    {
        if (handle) alll_destroy(handle);
    }

    // Parallel Moser-Tardos solve of the clauses in `clauses` (n_threads batches, any split); on return
    // var_arr->vars holds the assignment.  Replaces SATInstance.h:60-66 -> parallel_solve :217-320.
    Statistics *solve(vector<ClauseArray *> *clauses)
    {
        n_clauses = 0;
        for (auto c : *clauses) n_clauses += c->size();
        upload(clauses);
        return run_solve();
    }

    // Enumerated-clause variant (SATInstance.h:70-153): the enumeration is materialised once through the
    // caller's callback and solved by the same device loop; batch_size is accepted for compatibility.
    Statistics *solve(Clause<T> *(*getEnumeratedClause)(T, unsigned short int), ull n_clauses, T batch_size)
    {
        (void)batch_size;
        this->n_clauses = n_clauses;
        ClauseGenerator<T> generator(getEnumeratedClause, 0, (T)n_clauses, 0, (T)n_clauses);
        vector<uint64_t> off(1, 0);
        vector<uint32_t> lit;
        for (ull i = 0; i < n_clauses; i++) {
            Clause<T> *cl = generator.yieldNextClause();
            if (cl == nullptr) throw std::runtime_error("SATInstance::solve: clause enumeration ended early");
            for (auto &l : *cl->literals) lit.push_back((uint32_t)l);
            off.push_back(lit.size());
            delete cl->literals;
            delete cl;
        }
        upload_flat(off, lit);
        return run_solve();
    }

    // true iff the assignment in var_arr satisfies every clause (SATInstance.h:156-173); evaluated on the device.
    bool verify_validity(vector<ClauseArray *> *clauses) const
    {
        auto *self = const_cast<SATInstance *>(this);
        self->upload(clauses);
        self->check(alll_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        int valid = 0;
        self->check(alll_verify(handle, &valid), "alll_verify");
        return valid != 0;
    }

    // DIMACS dump of an enumerated instance (SATInstance.h:175-203): "p cnf V C", then " l1 l2 ... 0" per clause.
    void writeDIMACS(Clause<T> *(*getEnumeratedClause)(T, unsigned short int), ull n_clauses, ofstream *out_f)
    {
        this->n_clauses = n_clauses;
        ClauseGenerator<T> generator(getEnumeratedClause, 0, (T)n_clauses, 0, (T)n_clauses);
        *out_f << "p cnf " << n_vars << " " << n_clauses << endl;
        for (ull i = 0; i < n_clauses; i++) {
            Clause<T> *cl = generator.yieldNextClause();
            for (auto &l : *cl->literals) {
                const intmax_t v = (intmax_t)(l >> 1) + 1;
                *out_f << " " << to_string((l & 1) ? -v : v);
            }
            *out_f << " 0" << endl;
            delete cl->literals;
            delete cl;
            if (i % 1000 == 0) out_f->flush();
        }
        out_f->flush();
    }

    // ---- extensions (not in the reference) --------------------------------------------------------------
    // The same solve / verify on clauses that already are a CSR (off[m+1], lit[L], literal = 2*var+neg): no
    // Clause object graph (Clause.h:17-28 costs ~120 B of heap per clause, main.cpp:157-178 builds it one `new`
    // at a time).  cnf_read_csr (cli/cnf_io/cnf_io.h) produces this form straight from a DIMACS file.
    Statistics *solve_csr(const vector<uint64_t> &off, const vector<uint32_t> &lit)
    {
        if (off.empty()) throw std::runtime_error("SATInstance::solve_csr: off must hold m+1 entries");
        n_clauses = off.size() - 1;
        upload_flat(off, lit);
        return run_solve();
    }
    bool verify_validity_csr(const vector<uint64_t> &off, const vector<uint32_t> &lit) const
    {
        if (off.empty()) throw std::runtime_error("SATInstance::verify_validity_csr: off must hold m+1 entries");
        auto *self = const_cast<SATInstance *>(this);
        self->upload_flat(off, lit);
        self->check(alll_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        int valid = 0;
        self->check(alll_verify(handle, &valid), "alll_verify");
        return valid != 0;
    }
    // Enumerated clauses produced ON THE DEVICE.  The reference's callback form above hands out heap Clause objects
    // from host code, which cannot run inside a kernel; its device-side equivalent is a functor compiled into the
    // sweep kernel of include/alll_generator.cuh, passed here by its launcher (see tests/cpp/user_generator.cu).
    // Nothing is stored or materialised: n_clauses may exceed what fits in memory as a literal array.
    Statistics *solve_generator(alll_gen_launch_fn launch, void *user, ull n_clauses, unsigned k, ull cap_records = 0)
    {
        this->n_clauses = n_clauses;
        ensure_handle();
        check(alll_upload_generator(handle, (uint64_t)n_vars, n_clauses, k, launch, user, cap_records), "alll_upload_generator");
        have_upload = false;
        return run_solve();
    }
    bool verify_validity_generator(alll_gen_launch_fn launch, void *user, ull n_clauses, unsigned k, ull cap_records = 0) const
    {
        auto *self = const_cast<SATInstance *>(this);
        self->ensure_handle();
        self->check(alll_upload_generator(handle, (uint64_t)n_vars, n_clauses, k, launch, user, cap_records), "alll_upload_generator");
        self->have_upload = false;
        self->check(alll_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        int valid = 0;
        self->check(alll_verify(handle, &valid), "alll_verify");
        return valid != 0;
    }

    void set_seed(uint64_t s) { seed = s; have_seed = true; }          // reproducible rounds (Philox key)
    void set_max_rounds(uint64_t r) { max_rounds = r; }                // default: effectively unbounded
    void set_device(int ordinal) { device = ordinal; }                 // before the first solve/verify
    int last_status() const { return status; }                         // alll_status of the last solve
    const alll_stats &last_device_stats() const { return dev_stats; }  // device-timed ms, launches, ...

private:
    int n_threads{};
    alll_handle handle = nullptr;
    int device = -1;
    uint64_t seed = 0;
    bool have_seed = false;
    uint64_t max_rounds = ~0ull;
    int status = ALLL_OK;
    alll_stats dev_stats{};
    uint64_t uploaded_fingerprint = 0;
    bool have_upload = false;

    void check(int rc, const char *what) const
    {
        if (rc != ALLL_OK) throw std::runtime_error(string(what) + ": " + alll_last_error(handle));
    }

    void ensure_handle()
    {
        if (handle) return;
        alll_config cfg{};
        cfg.device = device;
        if (alll_create(&cfg, &handle) != ALLL_OK)
            throw std::runtime_error(string("alll_create: ") + alll_last_error(nullptr));
    }

    // Concatenate the batches in order -> CSR; identical content is not uploaded twice (solve then verify).
    void upload(vector<ClauseArray *> *clauses)
    {
        vector<uint64_t> off(1, 0);
        vector<uint32_t> lit;
        size_t m = 0;
        for (auto batch : *clauses) m += batch->size();
        off.reserve(m + 1);
        for (auto batch : *clauses) {
            for (auto cl : *batch) {
                for (auto &l : *cl->literals) lit.push_back((uint32_t)l);
                off.push_back(lit.size());
            }
        }
        upload_flat(off, lit);
    }

    void upload_flat(const vector<uint64_t> &off, const vector<uint32_t> &lit)
    {
        ensure_handle();
        uint64_t fp = 1469598103934665603ull ^ off.size();                    // FNV-1a over widths and literals
        for (size_t c = 1; c < off.size(); c++) fp = (fp ^ (off[c] - off[c - 1])) * 1099511628211ull;
        for (uint32_t l : lit) fp = (fp ^ l) * 1099511628211ull;
        if (have_upload && fp == uploaded_fingerprint) return;
        check(alll_upload_csr(handle, (uint64_t)n_vars, off.size() - 1, off.data(), lit.data()), "alll_upload_csr");
        uploaded_fingerprint = fp;
        have_upload = true;
    }

    Statistics *run_solve()
    {
        check(alll_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        const uint64_t s = have_seed ? seed : (((uint64_t)std::random_device{}() << 32) | std::random_device{}());
        status = alll_solve(handle, s, max_rounds, &dev_stats);
        if (status != ALLL_OK && status != ALLL_MAX_ROUNDS) check(status, "alll_solve");
        check(alll_get_assignment(handle, reinterpret_cast<uint8_t *>(var_arr->vars)), "alll_get_assignment");
        auto *st = new Statistics;
        st->n_iterations = dev_stats.n_iterations;
        st->n_resamples = dev_stats.n_resamples;
        st->avg_mis_size = dev_stats.avg_mis_size;
        st->n_thread_resamples.assign((size_t)n_threads, 0);
        st->n_thread_resamples[0] = dev_stats.n_resamples;
        return st;
    }
};

#endif
