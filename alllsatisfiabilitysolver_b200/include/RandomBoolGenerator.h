// RandomBoolGenerator.h -- drop-in for library/include/RandomBoolGenerator.h of the reference
// (RBG<E>, RandomBoolGenerator.h:28-50; `typedef unsigned long long ull` lives here, :14).
//
// Host-side helper only: the solver's own randomness is Philox4x32-10 on the device
// (csrc/alll_device.cuh).  The public contract carried over is "i.i.d. fair bits from engine E".
#ifndef ALLL_B200_RANDOMBOOLGENERATOR_H
#define ALLL_B200_RANDOMBOOLGENERATOR_H

#include <cstdint>
#include <random>

typedef unsigned long long ull;

template <typename E>
class RBG {
public:
    explicit RBG(E &engine) : engine_(engine) {}

    // One fair bit; a 64-bit engine draw is consumed bit by bit before the next draw.
    bool sample()
    {
        if (bits_left_ == 0) {
            pool_ = std::uniform_int_distribution<ull>{}(engine_);
            bits_left_ = 64;
        }
        const bool bit = (pool_ & 1ull) != 0;
        pool_ >>= 1;
        --bits_left_;
        return bit;
    }

private:
    E engine_;              // held by value, as the reference does
    ull pool_ = 0;
    unsigned bits_left_ = 0;
};

#endif
