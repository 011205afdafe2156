// ClauseGenerator.h -- drop-in for library/include/ClauseGenerator.h of the reference (ClauseGenerator.h:16-114).
//
// Enumerated-clause access: clauses are produced on demand by a user callback
// `Clause<T>* (*)(T idx, unsigned short t_id)`.  Kept for the streaming solve() overload and writeDIMACS();
// the device path materialises the enumeration once (SATInstance.h of this directory).
#ifndef ALLL_B200_CLAUSEGENERATOR_H
#define ALLL_B200_CLAUSEGENERATOR_H

#include <cstdint>
#include <type_traits>

#include "Clause.h"

using namespace std;

template <class T, class Enable = void>
class ClauseGenerator {};

template <class T>
class ClauseGenerator<T, typename enable_if<is_integral<T>::value>::type> {
public:
    using ClauseArray = typename Clause<T>::ClauseArray;
    typedef unsigned short int t_id_T;

    T n_clauses;

    ClauseGenerator(Clause<T> *(*getEnumeratedClause)(T, t_id_T), t_id_T t_id, T n_clauses, T base_offset, T batch_size)
        : n_clauses(n_clauses), get_(getEnumeratedClause), batch_size_(batch_size), base_offset_(base_offset), t_id_(t_id)
    {
    }

    // Next batch of violated clauses, visiting this generator's index range in the additive-stride order
    // c <- (c + P) mod n_clauses, P = 2^63 - 25 (ClauseGenerator.h:45,109).  Satisfied clauses are freed.
    ClauseArray *yieldRandomUNSATClauseBatch(const bool *var_arr)
    {
        if (finished_) reset();
        auto *out = new ClauseArray();
        T todo = batch_size_;
        if (yielded_ + batch_size_ >= n_clauses) todo = n_clauses - yielded_;
        for (T i = 0; i < todo; i++) {
            cursor_ = (T)(((uint64_t)cursor_ + STRIDE) % (uint64_t)n_clauses);
            Clause<T> *cl = get_(base_offset_ + cursor_, t_id_);
            if (cl == nullptr) {
                cerr << "WARNING: Clause generator went out of range and yielded nullptr." << endl;
                finished_ = true;
                break;
            }
            if (cl->is_not_satisfied(var_arr)) out->push_back(cl);
            else { delete cl->literals; delete cl; }
            yielded_++;
        }
        if (yielded_ == n_clauses) finished_ = true;
        return out;
    }

    // Sequential enumeration (ClauseGenerator.h:73-93).
    Clause<T> *yieldNextClause()
    {
        if (finished_) reset();
        Clause<T> *cl = get_(base_offset_ + yielded_, t_id_);
        if (cl == nullptr) {
            cerr << "WARNING: Clause generator went out of range and yielded nullptr." << endl;
            finished_ = true;
            return nullptr;
        }
        if (++yielded_ == n_clauses) finished_ = true;
        return cl;
    }

    bool has_finished_yielding() { return finished_; }

    void reset()
    {
        yielded_ = 0;
        finished_ = false;
    }

private:
    static constexpr uint64_t STRIDE = 9223372036854775783ull;   // 2^63 - 25, prime
    Clause<T> *(*get_)(T, unsigned short int);
    T batch_size_, base_offset_;
    t_id_T t_id_{};
    T cursor_ = 0;
    bool finished_ = false;
    T yielded_ = 0;
};

#endif
