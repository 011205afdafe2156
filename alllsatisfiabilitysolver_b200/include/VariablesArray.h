// VariablesArray.h -- drop-in for library/include/VariablesArray.h of the reference (VariablesArray.h:18-35).
//
// Same public surface: `tV n_vars; bool* vars;` (1 byte per variable, the solver's in/out state) and a
// constructor that fills it with uniform random booleans.  On the device the assignment lives bit-packed
// (csrc/layout.cu pack_bits_kernel / unpack_bits_kernel convert at the boundary).
#ifndef ALLL_B200_VARIABLESARRAY_H
#define ALLL_B200_VARIABLESARRAY_H

#include <iostream>
#include <random>

#include "RandomBoolGenerator.h"

using namespace std;   // the reference's headers export this; example/main.cpp relies on it

template <typename tV>
class VariablesArray {
public:
    tV n_vars;
    bool *vars;

    // Random initial assignment from a random_device-seeded engine (VariablesArray.h:23-34).
    explicit VariablesArray(tV n_vars) : VariablesArray(n_vars, std::random_device{}()) {}

    // Extension: reproducible initial assignment.
    VariablesArray(tV n_vars, unsigned long seed) : n_vars(n_vars), vars(new bool[n_vars > 0 ? n_vars : 1])
    {
        default_random_engine engine(seed);
        RBG<default_random_engine> rbg(engine);
        for (tV i = 0; i < n_vars; i++) vars[i] = rbg.sample();
    }
};

#endif
