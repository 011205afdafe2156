// Clause.h -- drop-in for library/include/Clause.h of the reference (Clause.h:17-47).
//
// Same public surface: `vector<tV>* literals` (encoding lit = 2*var + neg, var 0-based), `t_id`, the
// ClauseArray typedef and is_not_satisfied().  The caller still builds these objects (example/main.cpp:157-178);
// SATInstance::solve flattens them once into the device layout.  is_not_satisfied() is kept for source
// compatibility of user code that checks a single clause on the host -- SATInstance never calls it: every
// clause evaluation of solve()/verify_validity() runs in csrc/sweep.cu.
#ifndef ALLL_B200_CLAUSE_H
#define ALLL_B200_CLAUSE_H

#include <vector>

#include "VariablesArray.h"

template <typename tV>
class Clause {
public:
    typedef vector<Clause<tV> *> ClauseArray;

    vector<tV> *literals;
    unsigned short int t_id{};

    explicit Clause(vector<tV> *literals, unsigned short int t_id) : literals(literals), t_id(t_id) {}

    // true iff no literal of the clause is true under var_arr (Clause.h:34-46); an empty clause is never satisfied.
    bool is_not_satisfied(const bool *var_arr) const
    {
        bool any_true = false;
        for (const tV l : *literals) any_true = any_true || (var_arr[l >> 1] != ((l & 1) != 0));
        return !any_true;
    }
};

#endif
