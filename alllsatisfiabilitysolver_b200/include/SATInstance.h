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
// (global clause id = position in the concatenation; all host threads, straight into a page-locked staging buffer),
// uploads them through the C ABI (include/alll_b200.h) and the whole round loop -- sweep, independent set, resample --
// runs on the B200s.  There is no CPU fallback: a missing device or library error throws std::runtime_error.
//
// The reference's parallel-resource knob is this constructor's n_threads (SATInstance.h:51-56,259; -p of the CLI).  Here
// the resource is GPUs: set_gpus(n) -- or the environment variable ALLL_GPUS=n|all for programs compiled unchanged --
// solves one large instance over min(n, visible) devices behind this same blocking call (alll_multi_*: contiguous
// clause ranges, every device uploads its own 1/N of the staging buffer, fused NVLink exchange per round).  n_threads
// keeps its second role (number of input batches, size of n_thread_resamples) and is used for the host-side flatten.
//
// Deliberate deviations from the reference: n_clauses is assigned, not accumulated,
// by solve() (Q6); an optional seed makes runs reproducible (Q8); a round cap turns the reference's
// non-termination on unsatisfiable input into a status (Q11); n_thread_resamples has n_threads entries with
// the device total in entry 0 (there are no host worker threads to attribute resamples to).
#ifndef ALLL_B200_SATINSTANCE_H
#define ALLL_B200_SATINSTANCE_H

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#else
// The reference header pulls in <omp.h> (SATInstance.h:16) and its CLI calls omp_get_num_procs()
// (example/main.cpp:77-78); keep that call compiling without OpenMP.
static inline int omp_get_num_procs() { const unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }
#endif
#include <cmath>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/alll_b200.h"
#include "Clause.h"
#include "ClauseGenerator.h"
#include "RandomBoolGenerator.h"
#include "VariablesArray.h"

using namespace std;

// One literal into the page-locked staging buffer.  The buffer is written once, front to back, and next read by the GPU's
// copy engine: a non-temporal store keeps the line out of the cache and spares the memory system the read-for-ownership
// of a line that is overwritten completely (the flatten runs at the host's memory bandwidth, so a third less traffic on
// the store side is time).  alll_stage_fence() orders those stores before the unit is announced to the uploader.
static inline void alll_stage_store(uint32_t *dst, uint32_t v)
{
#if defined(__SSE2__)
    _mm_stream_si32(reinterpret_cast<int *>(dst), (int)v);
#else
    *dst = v;
#endif
}
static inline void alll_stage_fence()
{
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

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
        if (handle) alll_multi_destroy(handle);
        if (stage_lit) alll_host_free(stage_lit);
        if (stage_off) alll_host_free(stage_off);
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
        // (the caller's callback is host code handing out heap objects one at a time: inherently serial)
        vector<uint64_t> off(1, 0);
        vector<uint32_t> lit;
        off.reserve(n_clauses + 1);
        for (ull i = 0; i < n_clauses; i++) {
            Clause<T> *cl = generator.yieldNextClause();
            if (cl == nullptr) throw std::runtime_error("SATInstance::solve: clause enumeration ended early");
            if (i == 0) lit.reserve(cl->literals->size() * n_clauses);
            for (auto &l : *cl->literals) lit.push_back((uint32_t)l);
            off.push_back(lit.size());
            delete cl->literals;
            delete cl;
        }
        upload_flat(off.data(), off.size() - 1, lit.data());
        return run_solve();
    }

    // true iff the assignment in var_arr satisfies every clause (SATInstance.h:156-173); evaluated on the device.
    bool verify_validity(vector<ClauseArray *> *clauses) const
    {
        // (const in the reference's signature; the device handle is this object's cache, not its value)
        auto *self = const_cast<SATInstance *>(this);
        self->upload(clauses);
        return self->verify_uploaded();
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
        upload_flat(off.data(), off.size() - 1, lit.data());
        return run_solve();
    }
    bool verify_validity_csr(const vector<uint64_t> &off, const vector<uint32_t> &lit) const
    {
        if (off.empty()) throw std::runtime_error("SATInstance::verify_validity_csr: off must hold m+1 entries");
        auto *self = const_cast<SATInstance *>(this);
        self->upload_flat(off.data(), off.size() - 1, lit.data());
        return self->verify_uploaded();
    }
    // var_arr against the clauses of the most recent solve / verify call, which are still on the device: no flatten, no
    // upload.  (verify_validity(clauses) itself always uploads what it is given -- it cannot know whether the caller
    // changed a clause since; round 1 guessed with a fingerprint, which a collision would have turned into a wrong answer.)
    bool verify_last() const
    {
        if (!have_upload) throw std::runtime_error("SATInstance::verify_last: nothing has been solved or verified yet");
        return const_cast<SATInstance *>(this)->verify_uploaded();
    }
    // Enumerated clauses produced ON THE DEVICE.  The reference's callback form above hands out heap Clause objects
    // from host code, which cannot run inside a kernel; its device-side equivalent is a functor compiled into the
    // sweep kernel of include/alll_generator.cuh, passed here by its launcher (see tests/cpp/user_generator.cu).
    // Nothing is stored or materialised: n_clauses may exceed what fits in memory as a literal array.
    Statistics *solve_generator(alll_gen_launch_fn launch, void *user, ull n_clauses, unsigned k, ull cap_records = 0)
    {
        this->n_clauses = n_clauses;
        ensure_handle();
        check1(alll_upload_generator(first(), (uint64_t)n_vars, n_clauses, k, launch, user, cap_records), "alll_upload_generator");
        have_upload = false;
        return run_solve_single();
    }
    bool verify_validity_generator(alll_gen_launch_fn launch, void *user, ull n_clauses, unsigned k, ull cap_records = 0) const
    {
        auto *self = const_cast<SATInstance *>(this);
        self->ensure_handle();
        alll_handle h = self->first();
        self->check1(alll_upload_generator(h, (uint64_t)n_vars, n_clauses, k, launch, user, cap_records), "alll_upload_generator");
        self->have_upload = false;
        self->check1(alll_set_assignment(h, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        int valid = 0;
        self->check1(alll_verify(h, &valid), "alll_verify");
        return valid != 0;
    }

    void set_seed(uint64_t s) { seed = s; have_seed = true; }          // reproducible rounds (Philox key)
    void set_max_rounds(uint64_t r) { max_rounds = r; }                // default: effectively unbounded
    void set_device(int ordinal) { device = ordinal; }                 // first (or only) GPU; before the first solve/verify
    // GPUs for one large instance (0 = every visible device); before the first solve/verify.  Default: ALLL_GPUS, else 1.
    void set_gpus(int n) { n_gpus = n; gpus_set = true; }
    int gpus_in_use() const { uint64_t info[4] = {1, 0, 0, 0}; if (handle && have_upload) alll_multi_info(handle, info); return (int)info[0]; }
    double last_flatten_ms() const { return flatten_ms; }              // host time of the last flatten of Clause objects
    int last_status() const { return status; }                         // alll_status of the last solve
    const alll_stats &last_device_stats() const { return dev_stats; }  // device-timed ms, launches, ...

private:
    int n_threads{};
    alll_multi_handle handle = nullptr;
    int device = -1, n_gpus = 1;
    bool gpus_set = false;
    uint64_t seed = 0;
    bool have_seed = false;
    uint64_t max_rounds = ~0ull;
    int status = ALLL_OK;
    alll_stats dev_stats{};
    bool have_upload = false;
    double flatten_ms = 0.0;
    // page-locked staging of the flattened clauses (grow-only): the H2D copy runs at the PCIe rate from here
    uint32_t *stage_lit = nullptr;
    uint64_t *stage_off = nullptr;
    size_t stage_lit_cap = 0, stage_off_cap = 0;

    void check(int rc, const char *what) const
    {
        if (rc != ALLL_OK) throw std::runtime_error(string(what) + ": " + alll_multi_last_error(handle));
    }
    void check1(int rc, const char *what) const          // calls on the first device's own handle
    {
        if (rc != ALLL_OK) throw std::runtime_error(string(what) + ": " + alll_last_error(first()));
    }
    alll_handle first() const
    {
        alll_handle h = nullptr;
        if (!handle || alll_multi_device_handle(handle, 0, &h) != ALLL_OK) throw std::runtime_error("SATInstance: no device handle");
        return h;
    }

    void ensure_handle()
    {
        if (handle) return;
        int32_t visible = 0;
        alll_device_count(&visible);
        if (visible <= 0) throw std::runtime_error("alll_multi_create: no CUDA device: the solver has no CPU fallback");
        int want = n_gpus;
        if (!gpus_set) {
            if (const char *e = std::getenv("ALLL_GPUS")) want = (std::strcmp(e, "all") == 0) ? 0 : std::atoi(e);
        }
        if (want <= 0 || want > visible) want = visible;
        const int base = device < 0 ? 0 : device;
        vector<int32_t> devs;
        for (int i = 0; i < want; i++) devs.push_back((int32_t)((base + i) % visible));
        alll_config cfg{};
        if (alll_multi_create(devs.data(), (uint32_t)devs.size(), &cfg, &handle) != ALLL_OK)
            throw std::runtime_error(string("alll_multi_create: ") + alll_multi_last_error(nullptr));
    }

    template <typename U> void ensure_stage(U *&buf, size_t &cap, size_t count)
    {
        if (cap >= count) return;
        if (buf) { alll_host_free(buf); buf = nullptr; cap = 0; }
        void *p = nullptr;
        const size_t want = count + count / 8 + 64;
        if (alll_host_alloc(want * sizeof(U), &p) != ALLL_OK) {
            int32_t visible = 0;
            alll_device_count(&visible);
            throw std::runtime_error(visible > 0 ? "alll_host_alloc failed (page-locked staging memory)"
                                                 : "alll_host_alloc: no CUDA device: the solver has no CPU fallback");
        }
        buf = static_cast<U *>(p);
        cap = want;
    }

    // progress of a streamed flatten: units of clauses are filled by the flatten threads (any order), `filled_rows` is the
    // contiguous prefix -- what alll_multi_upload_fixedk_streamed waits on
    struct Streamed {
        size_t unit = 0, n_units = 0, m = 0;
        std::atomic<size_t> next{0}, filled_rows{0};
        std::atomic<int> abort{0};
        vector<char> done;
        size_t prefix = 0;
        std::mutex mu;
        std::chrono::steady_clock::time_point t_last;
        void finish(size_t u)
        {
            std::lock_guard<std::mutex> lk(mu);
            done[u] = 1;
            while (prefix < n_units && done[prefix]) ++prefix;
            filled_rows.store(std::min(m, prefix * unit), std::memory_order_release);
            t_last = std::chrono::steady_clock::now();
        }
        double flatten_ms(std::chrono::steady_clock::time_point t0) const { return std::chrono::duration<double, std::milli>(t_last - t0).count(); }
        static int filled(void *user, uint64_t rows_needed)       // alll_filled_fn
        {
            Streamed *s = static_cast<Streamed *>(user);
            while (s->filled_rows.load(std::memory_order_acquire) < rows_needed) {
                if (s->abort.load()) return 1;
                std::this_thread::yield();
            }
            return s->abort.load();
        }
    };

    static unsigned flatten_threads(size_t m)
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return (unsigned)std::max<size_t>(1, std::min<size_t>(std::min(hw, 64u), m / 16384));
    }
    template <class F> static void run_threads(unsigned nt, F &&fn)      // fn(t)
    {
        vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back([&, t] { fn(t); });
        fn(0u);
        for (auto &x : th) x.join();
    }

    // Concatenate the batches in order -> CSR in page-locked memory, split over the host's threads by clause position.
    // The pointer chase Clause -> vector -> data is what the reference pays in EVERY sweep (Clause.h:34-46); here it
    // is paid once per solve, with the next clauses' three levels software-prefetched (the chase is latency-bound).
    // Fast path: one pass that assumes every clause has the width of the first one and writes literals to their final
    // place; the first clause of another width sends the whole flatten down the general two-pass path
    // (widths -> offsets by a prefix sum -> literals).
    void upload(vector<ClauseArray *> *clauses)
    {
        const auto t0 = std::chrono::steady_clock::now();
        size_t m = 0;
        vector<size_t> batch_first;
        for (auto batch : *clauses) { batch_first.push_back(m); m += batch->size(); }
        batch_first.push_back(m);
        ensure_stage(stage_off, stage_off_cap, m + 1);
        const unsigned nt = flatten_threads(m);
        // clause position -> (batch, index): each thread walks its own contiguous position range
        auto for_range = [&](size_t p0, size_t p1, auto &&fn) {
            size_t b = 0;
            while (b + 1 < batch_first.size() - 1 && batch_first[b + 1] <= p0) ++b;
            for (size_t p = p0; p < p1;) {
                while (batch_first[b + 1] <= p) ++b;
                const ClauseArray &arr = *(*clauses)[b];
                const size_t base = batch_first[b], n_in = arr.size();
                const size_t end = std::min(p1, batch_first[b + 1]);
                for (; p < end; ++p) {
                    const size_t i = p - base;
                    // three-stage prefetch: Clause object, its vector header, the vector's storage
                    if (i + 24 < n_in) __builtin_prefetch(arr[i + 24]);
                    if (i + 16 < n_in) __builtin_prefetch(arr[i + 16]->literals);
                    if (i + 8 < n_in) __builtin_prefetch(arr[i + 8]->literals->data());
                    if (!fn(p, arr[i])) return false;
                }
            }
            return true;
        };
        bool done = false;
        if (m > 0) {
            size_t b0 = 0;
            while ((*clauses)[b0]->empty()) ++b0;
            const uint64_t w0 = (*(*clauses)[b0])[0]->literals->size();
            if (w0 > 0 && w0 <= 32 && m >= (size_t)1 << 18 && !std::getenv("ALLL_NO_STREAMED_UPLOAD")) {
                // Streamed: the flatten threads fill the staging buffer unit by unit IN ORDER while this thread sits in
                // alll_multi_upload_fixedk_streamed, which copies and lays out every 64 MB chunk as soon as the fill position
                // has passed it -- the upload hides behind the flatten instead of following it.
                ensure_stage(stage_lit, stage_lit_cap, m * w0);
                ensure_handle();
                have_upload = false;
                Streamed st;
                st.unit = (size_t)1 << 16;
                st.n_units = (m + st.unit - 1) / st.unit;
                st.m = m;
                st.done.assign(st.n_units, 0);
                vector<std::thread> th;
                for (unsigned t = 0; t < nt; t++)
                    th.emplace_back([&] {
                        for (;;) {
                            const size_t u = st.next.fetch_add(1);
                            if (u >= st.n_units || st.abort.load()) return;
                            const bool ok = for_range(u * st.unit, std::min(m, (u + 1) * st.unit), [&](size_t p, const Clause<T> *cl) {
                                const vector<T> &ls = *cl->literals;
                                if (ls.size() != w0) return false;
                                uint32_t *dst = stage_lit + p * w0;
                                for (size_t j = 0; j < w0; j++) alll_stage_store(dst + j, (uint32_t)ls[j]);
                                return true;
                            });
                            if (!ok) { st.abort.store(1); return; }
                            alll_stage_fence();
                            st.finish(u);
                        }
                    });
                const int rc = alll_multi_upload_fixedk_streamed(handle, (uint64_t)n_vars, m, (uint32_t)w0, stage_lit, &Streamed::filled, &st);
                for (auto &x : th) x.join();
                flatten_ms = st.flatten_ms(t0);
                if (rc == ALLL_OK) { have_upload = true; return; }
                if (!st.abort.load()) check(rc, "alll_multi_upload_fixedk_streamed");
                // a clause of another width: the general two-pass path below
            } else if (w0 > 0 && w0 <= 32) {
                ensure_stage(stage_lit, stage_lit_cap, m * w0);
                vector<char> uniform(nt, 1);
                run_threads(nt, [&](unsigned t) {
                    uniform[t] = for_range(m * t / nt, m * (t + 1) / nt, [&](size_t p, const Clause<T> *cl) {
                        const vector<T> &ls = *cl->literals;
                        if (ls.size() != w0) return false;
                        uint32_t *dst = stage_lit + p * w0;
                        for (size_t j = 0; j < w0; j++) dst[j] = (uint32_t)ls[j];
                        stage_off[p + 1] = (p + 1) * w0;
                        return true;
                    });
                });
                done = true;
                for (char u : uniform) done = done && u;
                stage_off[0] = 0;
            }
        }
        if (!done) {
            vector<uint64_t> part(nt + 1, 0);
            run_threads(nt, [&](unsigned t) {
                uint64_t sum = 0;
                for_range(m * t / nt, m * (t + 1) / nt, [&](size_t p, const Clause<T> *cl) {
                    const uint64_t w = cl->literals->size();
                    stage_off[p + 1] = w;                         // widths for now
                    sum += w;
                    return true;
                });
                part[t + 1] = sum;
            });
            for (unsigned t = 0; t < nt; t++) part[t + 1] += part[t];
            const uint64_t n_lit = part[nt];
            ensure_stage(stage_lit, stage_lit_cap, (size_t)std::max<uint64_t>(n_lit, 1));
            stage_off[0] = 0;
            run_threads(nt, [&](unsigned t) {
                uint64_t at = part[t];
                for_range(m * t / nt, m * (t + 1) / nt, [&](size_t p, const Clause<T> *cl) {
                    const vector<T> &ls = *cl->literals;
                    uint32_t *dst = stage_lit + at;
                    for (size_t j = 0; j < ls.size(); j++) dst[j] = (uint32_t)ls[j];
                    at += ls.size();
                    stage_off[p + 1] = at;
                    return true;
                });
            });
        }
        flatten_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        upload_flat(stage_off, m, stage_lit);
    }

    void upload_flat(const uint64_t *off, size_t m, const uint32_t *lit)
    {
        ensure_handle();
        have_upload = false;
        check(alll_multi_upload_csr(handle, (uint64_t)n_vars, m, off, lit), "alll_multi_upload_csr");
        have_upload = true;
    }

    bool verify_uploaded()
    {
        check(alll_multi_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_multi_set_assignment");
        int valid = 0;
        check(alll_multi_verify(handle, &valid), "alll_multi_verify");
        return valid != 0;
    }

    Statistics *make_statistics()
    {
        auto *st = new Statistics;
        st->n_iterations = dev_stats.n_iterations;
        st->n_resamples = dev_stats.n_resamples;
        st->avg_mis_size = dev_stats.avg_mis_size;
        st->n_thread_resamples.assign((size_t)n_threads, 0);
        st->n_thread_resamples[0] = dev_stats.n_resamples;
        return st;
    }
    uint64_t next_seed() { return have_seed ? seed : (((uint64_t)std::random_device{}() << 32) | std::random_device{}()); }

    Statistics *run_solve()
    {
        check(alll_multi_set_assignment(handle, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_multi_set_assignment");
        status = alll_multi_solve(handle, next_seed(), max_rounds, &dev_stats);
        if (status != ALLL_OK && status != ALLL_MAX_ROUNDS) check(status, "alll_multi_solve");
        check(alll_multi_get_assignment(handle, reinterpret_cast<uint8_t *>(var_arr->vars)), "alll_multi_get_assignment");
        return make_statistics();
    }

    // enumerated clauses produced on the device: first device only (the generator launcher is bound to one device)
    Statistics *run_solve_single()
    {
        alll_handle h = first();
        check1(alll_set_assignment(h, reinterpret_cast<const uint8_t *>(var_arr->vars)), "alll_set_assignment");
        status = alll_solve(h, next_seed(), max_rounds, &dev_stats);
        if (status != ALLL_OK && status != ALLL_MAX_ROUNDS) check1(status, "alll_solve");
        check1(alll_get_assignment(h, reinterpret_cast<uint8_t *>(var_arr->vars)), "alll_get_assignment");
        return make_statistics();
    }
};

#endif
