// A user's translation unit for the enumerated-clause solve: the device-side form of the reference's clause callback
// `Clause<T>* (*)(T index, unsigned short t_id)` (SATInstance.h:70, ClauseGenerator.h:27).  Built by
// tests/test_generator.py with
//   nvcc -gencode arch=compute_100a,code=sm_100a -I include user_generator.cu -L... -lalll_b200
// Prints one JSON line; writes the final assignment (1 byte per variable) to argv[4].
//
// The family: "ring" 5-SAT.  Clause i holds the variables (i + j * stride) mod n, j = 0..4, with signs taken from a
// multiplicative hash of i -- a pure function of the index, as the reference requires of its callback.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "alll_generator.cuh"

constexpr int K = 5;

struct RingClauses {
    uint32_t n_vars, stride;
    __host__ __device__ void operator()(uint64_t index, uint32_t (&lits)[K]) const
    {
        const uint32_t h = (uint32_t)index * 2654435761u;
        for (int j = 0; j < K; j++) {
            const uint32_t var = (uint32_t)((index + (uint64_t)j * stride) % n_vars);
            lits[j] = 2u * var + ((h >> (j + 7)) & 1u);
        }
    }
};

static int launch(void *user, const alll_gen_sweep_args *a, void *stream)
{
    return alll_gen::launch_sweep<K>(*static_cast<const RingClauses *>(user), *a, stream);
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const uint64_t n = strtoull(argv[1], nullptr, 0), m = strtoull(argv[2], nullptr, 0), seed = strtoull(argv[3], nullptr, 0);
    RingClauses gen{(uint32_t)n, 7919u};
    alll_handle h = nullptr;
    if (alll_create(nullptr, &h) != ALLL_OK) { fprintf(stderr, "%s\n", alll_last_error(nullptr)); return 1; }
    int rc = alll_upload_generator(h, n, m, K, launch, &gen, 0);
    if (rc == ALLL_OK) rc = alll_randomize(h, seed);
    alll_stats st{};
    if (rc == ALLL_OK) rc = alll_solve(h, seed, 10000, &st);
    int valid = 0;
    if (rc == ALLL_OK) rc = alll_verify(h, &valid);
    std::vector<uint8_t> vars(n);
    if (rc == ALLL_OK) rc = alll_get_assignment(h, vars.data());
    if (rc != ALLL_OK) { fprintf(stderr, "error %d: %s\n", rc, alll_last_error(h)); return 1; }
    // independent host check with the same functor
    uint64_t host_violated = 0;
    for (uint64_t i = 0; i < m; i++) {
        uint32_t lits[K];
        gen(i, lits);
        bool violated = true;
        for (int j = 0; j < K; j++) violated &= vars[lits[j] >> 1] == (lits[j] & 1u);
        host_violated += violated;
    }
    if (FILE *f = fopen(argv[4], "wb")) { fwrite(vars.data(), 1, n, f); fclose(f); }
    printf("{\"n_iterations\": %llu, \"n_resamples\": %llu, \"avg_mis_size\": %llu, \"valid\": %d, \"host_violated\": %llu, "
           "\"solve_ms\": %.3f}\n",
           (unsigned long long)st.n_iterations, (unsigned long long)st.n_resamples, (unsigned long long)st.avg_mis_size, valid,
           (unsigned long long)host_violated, st.solve_ms);
    alll_destroy(h);
    return 0;
}
