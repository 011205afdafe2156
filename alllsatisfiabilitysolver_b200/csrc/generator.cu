// generator.cu -- the generators that ship with the library (alll_upload_builtin_generator) and the worked example of a
// user translation unit for the enumerated-clause solve: a device functor, include/alll_generator.cuh instantiated with
// it, and a launcher handed to alll_upload_generator.  Specification of the two generators: include/alll_b200.h.
#include <numeric>

#include "../../include/alll_generator.cuh"

namespace alll {

namespace {

constexpr uint32_t TAG_UNIFORM = 0x47454E31u, TAG_BOUNDED = 0x47454E32u, TAG_BOUNDED_SETUP = 0x47454E33u;

// q = floor(x / d) from the precomputed reciprocal r = floor((2^64 - 1) / d): the estimate is exact or one short.
__host__ __device__ __forceinline__ uint64_t mod_by(uint64_t x, uint64_t d, uint64_t recip)
{
#ifdef __CUDA_ARCH__
    const uint64_t q = __umul64hi(x, recip);
#else
    const uint64_t q = (uint64_t)(((unsigned __int128)x * recip) >> 64);
#endif
    uint64_t r = x - q * d;
    if (r >= d) r -= d;
    return r;
}

template <int K>
struct UniformClauses {
    uint32_t n_vars, k0, k1;
    __host__ __device__ void operator()(uint64_t index, uint32_t (&lits)[K]) const
    {
#pragma unroll
        for (int j0 = 0; j0 < K; j0 += 4) {
            const alll_gen::Philox4 o = alll_gen::philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), TAG_UNIFORM, (uint32_t)(j0 >> 2), k0, k1);
            const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (j0 + u < K) {
#ifdef __CUDA_ARCH__
                    const uint32_t var = __umulhi(w[u], n_vars);
#else
                    const uint32_t var = (uint32_t)(((uint64_t)w[u] * n_vars) >> 32);
#endif
                    lits[j0 + u] = 2u * var + (w[u] & 1u);
                }
        }
    }
};

// n_vars divides span = n_vars * d, so ((a*p + b) mod span) mod n_vars == (a*p + b) mod n_vars: one 64-bit reduction per
// clause, then literal j+1 is literal j plus (a mod n_vars), wrapped -- all in 32 bits.
template <int K>
struct BoundedClauses {
    uint64_t a, b, recip_n;
    uint32_t n_vars, a_mod_n, k0, k1;
    __host__ __device__ void operator()(uint64_t index, uint32_t (&lits)[K]) const
    {
        const uint32_t signs = alll_gen::philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), TAG_BOUNDED, 0u, k0, k1).x;
        uint32_t var = (uint32_t)mod_by(a * (index * (uint64_t)K) + b, n_vars, recip_n);   // a < 2^26 + a few, index*K < 2^36: no overflow
#pragma unroll
        for (int j = 0; j < K; j++) {
            lits[j] = 2u * var + ((signs >> j) & 1u);
            var += a_mod_n;                                  // n_vars <= 2^31: no 32-bit overflow
            if (var >= n_vars) var -= n_vars;
        }
    }
};

struct BuiltinSpec {
    uint32_t kind = 0, k = 0;
    UniformClauses<1> uni{};        // parameters only; re-typed per K at launch
    BoundedClauses<1> bnd{};
};

const char *make_spec(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d, BuiltinSpec *out)
{
    if (k < 1 || k > 16) return "built-in generators support 1 <= k <= 16";
    if (n_vars == 0 || n_vars > (1ull << 31)) return "n_vars must be in [1, 2^31]";
    out->kind = kind;
    out->k = k;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    if (kind == ALLL_GEN_UNIFORM) {
        out->uni.n_vars = (uint32_t)n_vars; out->uni.k0 = k0; out->uni.k1 = k1;
        return nullptr;
    }
    if (kind == ALLL_GEN_BOUNDED) {
        if (d == 0) return "d must be positive";
        const uint64_t span = n_vars * (uint64_t)d;
        if (span >= (1ull << 36)) return "n_vars * d must be below 2^36";
        if (m > span / k) return "m * k must not exceed n_vars * d";
        const alll_gen::Philox4 o = alll_gen::philox4x32_10(0u, 0u, TAG_BOUNDED_SETUP, 0u, k0, k1);
        uint64_t a = ((((uint64_t)o.y << 32) | o.x) & ((1ull << 26) - 1)) | (1ull << 20) | 1ull;
        a %= span;
        if (a == 0) a = 1;
        while (std::gcd(a, span) != 1) a = a + 1 < span ? a + 1 : 1;
        out->bnd.n_vars = (uint32_t)n_vars; out->bnd.a = a; out->bnd.a_mod_n = (uint32_t)(a % n_vars);
        out->bnd.b = (((uint64_t)o.w << 32) | o.z) % span;
        out->bnd.recip_n = ~0ull / n_vars;
        out->bnd.k0 = k0; out->bnd.k1 = k1;
        return nullptr;
    }
    return "unknown generator kind";
}

template <int K> UniformClauses<K> retype(const UniformClauses<1> &s) { return UniformClauses<K>{s.n_vars, s.k0, s.k1}; }
template <int K> BoundedClauses<K> retype(const BoundedClauses<1> &s)
{
    return BoundedClauses<K>{s.a, s.b, s.recip_n, s.n_vars, s.a_mod_n, s.k0, s.k1};
}

template <int K>
int launch_k(const BuiltinSpec &sp, const alll_gen_sweep_args &a, void *stream)
{
    if (sp.kind == ALLL_GEN_UNIFORM) return alll_gen::launch_sweep<K>(retype<K>(sp.uni), a, stream);
    return alll_gen::launch_sweep<K>(retype<K>(sp.bnd), a, stream);
}

template <int K>
void clause_k(const BuiltinSpec &sp, uint64_t index, uint32_t *lits)
{
    uint32_t tmp[K];
    if (sp.kind == ALLL_GEN_UNIFORM) retype<K>(sp.uni)(index, tmp);
    else retype<K>(sp.bnd)(index, tmp);
    for (int j = 0; j < K; j++) lits[j] = tmp[j];
}

#define FOR_EACH_K(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

} // namespace

// ---- what capi.cu calls ------------------------------------------------------------------------------------------

struct BuiltinGenerator {
    BuiltinSpec spec;
};

const char *builtin_generator_create(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                     BuiltinGenerator **out)
{
    auto *g = new BuiltinGenerator();
    if (const char *e = make_spec(kind, n_vars, m, k, seed, d, &g->spec)) { delete g; return e; }
    *out = g;
    return nullptr;
}

void builtin_generator_destroy(BuiltinGenerator *g) { delete g; }

// alll_gen_launch_fn
int builtin_generator_launch(void *user, const alll_gen_sweep_args *a, void *stream)
{
    const BuiltinSpec &sp = static_cast<BuiltinGenerator *>(user)->spec;
    switch (sp.k) {
#define CASE(K) case K: return launch_k<K>(sp, *a, stream);
        FOR_EACH_K(CASE)
#undef CASE
    }
    return (int)cudaErrorInvalidValue;
}

const char *builtin_generator_clause(uint32_t kind, uint64_t n_vars, uint64_t m, uint32_t k, uint64_t seed, uint32_t d,
                                     uint64_t index, uint32_t *lits)
{
    BuiltinSpec sp;
    if (const char *e = make_spec(kind, n_vars, m, k, seed, d, &sp)) return e;
    if (index >= m) return "index out of range";
    switch (k) {
#define CASE(K) case K: clause_k<K>(sp, index, lits); return nullptr;
        FOR_EACH_K(CASE)
#undef CASE
    }
    return "unsupported k";
}

} // namespace alll
