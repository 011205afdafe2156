// ClauseGenerator.h -- API-compatible stand-in for library/include/ClauseGenerator.h of the reference
// (ClauseGenerator.h:16-114): clauses of an "enumerated" instance are produced on demand by a user callback
// `Clause<T>* (*)(T index, unsigned short t_id)` instead of being stored.
//
// The device path does not stream clauses: SATInstance::solve(callback, ...) materialises the enumeration once
// (yieldNextClause over the whole index range) and runs the ordinary GPU round loop.  This class keeps the
// reference's public members so user code and writeDIMACS keep compiling:
//   n_clauses, yieldRandomUNSATClauseBatch(), yieldNextClause(), has_finished_yielding(), reset().
#ifndef ALLL_B200_CLAUSEGENERATOR_H
#define ALLL_B200_CLAUSEGENERATOR_H

#include <cstdint>
#include <type_traits>

#include "Clause.h"

using namespace std;

template <class T, class Enable = void>
class ClauseGenerator {};   // only integral index types are supported, as in the reference

template <class T>
class ClauseGenerator<T, typename enable_if<is_integral<T>::value>::type> {
public:
    using ClauseArray = typename Clause<T>::ClauseArray;
    typedef unsigned short int t_id_T;
    typedef Clause<T> *(*Producer)(T, t_id_T);

    T n_clauses;   // size of this generator's index range [base_offset, base_offset + n_clauses)

    ClauseGenerator(Producer getEnumeratedClause, t_id_T t_id, T n_clauses, T base_offset, T batch_size)
        : n_clauses(n_clauses), producer_(getEnumeratedClause), owner_(t_id), first_(base_offset), batch_(batch_size),
          hop_(n_clauses ? (T)(kStride % (uint64_t)n_clauses) : 0)
    {
    }

    // One pass over the range is cut into batches of `batch_size` indices.  Indices are visited in the
    // additive-stride order c <- (c + P) mod n_clauses with the prime P = 2^63 - 25 (ClauseGenerator.h:45,109);
    // only clauses violated under `var_arr` are handed to the caller, the others are released here.
    ClauseArray *yieldRandomUNSATClauseBatch(const bool *var_arr)
    {
        restart_if_done();
        ClauseArray *violated = new ClauseArray();
        const T remaining = n_clauses - served_;
        const T quota = batch_ < remaining ? batch_ : remaining;
        for (T step = 0; step < quota && !done_; ++step) {
            stride_pos_ = advance(stride_pos_);
            Clause<T> *c = fetch(first_ + stride_pos_);
            if (!c) break;
            ++served_;
            if (c->is_not_satisfied(var_arr)) {
                violated->push_back(c);
            } else {
                release(c);
            }
        }
        if (served_ == n_clauses) done_ = true;
        return violated;
    }

    // Plain sequential enumeration (ClauseGenerator.h:73-93); nullptr once the callback runs out of clauses.
    Clause<T> *yieldNextClause()
    {
        restart_if_done();
        Clause<T> *c = fetch(first_ + served_);
        if (c && ++served_ == n_clauses) done_ = true;
        return c;
    }

    bool has_finished_yielding() { return done_; }

    void reset()
    {
        served_ = 0;
        done_ = false;
    }

private:
    static constexpr uint64_t kStride = 9223372036854775783ull;   // 2^63 - 25

    Producer producer_;
    t_id_T owner_;
    T first_, batch_;
    T hop_;                 // kStride mod n_clauses: (c + P) mod n == (c + hop_) mod n
    T stride_pos_ = 0;
    T served_ = 0;
    bool done_ = false;

    void restart_if_done() { if (done_) reset(); }

    T advance(T pos) const { return (T)(((uint64_t)pos + (uint64_t)hop_) % (uint64_t)n_clauses); }

    Clause<T> *fetch(T index)
    {
        Clause<T> *c = producer_(index, owner_);
        if (!c) {
            cerr << "WARNING: Clause generator went out of range and yielded nullptr." << endl;
            done_ = true;
        }
        return c;
    }

    static void release(Clause<T> *c)
    {
        delete c->literals;
        delete c;
    }
};

#endif
