// hostpack.cpp -- host side of the packed host-to-device transport (capi.cu: PackPipe; device side: layout.cu unpack25_kernel).
//
// The caller's literals are 32-bit words (2*var+neg, the reference's encoding: example/main.cpp:168) but carry at most 25
// significant bits for n_vars <= 2^24, and the end-to-end time of a large upload IS the PCIe transfer of those words.  The
// host threads of an upload therefore re-pack every chunk into 25 bits per literal in page-locked memory -- the low three
// bytes of every literal back to back, bit 24 of eight literals in one byte of a second array -- and the device expands
// the chunk again (0.78 of the bytes on the link).  Plain g++ translation unit (no CUDA): AVX2 where the CPU has it.
#include <cstddef>
#include <cstdint>

#include <immintrin.h>

namespace alll {

// n literals -> lo3[3 n] and hi[ceil(n / 8)] (bit j of hi[i] = bit 24 of literal 8 i + j).  Returns the OR of all literals:
// a bit above 24 in it means the chunk was not representable (the caller falls back / reports the literal).
static uint32_t pack25_scalar(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi)
{
    uint32_t acc = 0;
    for (size_t i = 0; i < n; i += 8) {
        uint32_t hb = 0;
        const size_t e = n - i < 8 ? n - i : 8;
        for (size_t j = 0; j < e; j++) {
            const uint32_t v = src[i + j];
            acc |= v;
            lo3[3 * (i + j)] = (uint8_t)v;
            lo3[3 * (i + j) + 1] = (uint8_t)(v >> 8);
            lo3[3 * (i + j) + 2] = (uint8_t)(v >> 16);
            hb |= ((v >> 24) & 1u) << j;
        }
        hi[i / 8] = (uint8_t)hb;
    }
    return acc;
}

// 32 literals (128 bytes) per iteration -> three 32-byte non-temporal stores + four bit-24 bytes.  lo3 must be 32-byte
// aligned (the units of an upload start on such boundaries); the remainder goes through the scalar loop.
__attribute__((target("avx2"))) static uint32_t pack25_avx2(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi)
{
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1,
                                          0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
    const __m256i compact = _mm256_setr_epi32(0, 1, 2, 4, 5, 6, 7, 7);        // 2 x 12 bytes -> 24 bytes in dwords 0..5
    const __m256i b_lo = _mm256_setr_epi32(0, 0, 0, 0, 0, 0, 0, 1);           // B0 B1 -> dwords 6, 7
    const __m256i b_hi = _mm256_setr_epi32(2, 3, 4, 5, 0, 0, 0, 0);           // B2..B5 -> dwords 0..3
    const __m256i c_lo = _mm256_setr_epi32(0, 0, 0, 0, 0, 1, 2, 3);           // C0..C3 -> dwords 4..7
    const __m256i c_hi = _mm256_setr_epi32(4, 5, 0, 0, 0, 0, 0, 0);           // C4 C5 -> dwords 0, 1
    const __m256i d_all = _mm256_setr_epi32(0, 0, 0, 1, 2, 3, 4, 5);          // D0..D5 -> dwords 2..7
    __m256i acc = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 8));
        const __m256i v2 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 16));
        const __m256i v3 = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i + 24));
        acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(v0, v1), _mm256_or_si256(v2, v3)));
        const __m256i A = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(v0, shuf), compact);
        const __m256i B = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(v1, shuf), compact);
        const __m256i C = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(v2, shuf), compact);
        const __m256i D = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(v3, shuf), compact);
        const __m256i o0 = _mm256_blend_epi32(A, _mm256_permutevar8x32_epi32(B, b_lo), 0xC0);
        const __m256i o1 = _mm256_blend_epi32(_mm256_permutevar8x32_epi32(B, b_hi), _mm256_permutevar8x32_epi32(C, c_lo), 0xF0);
        const __m256i o2 = _mm256_blend_epi32(_mm256_permutevar8x32_epi32(C, c_hi), _mm256_permutevar8x32_epi32(D, d_all), 0xFC);
        __m256i *dst = reinterpret_cast<__m256i *>(lo3 + 3 * i);
        _mm256_stream_si256(dst, o0);
        _mm256_stream_si256(dst + 1, o1);
        _mm256_stream_si256(dst + 2, o2);
        const uint32_t h0 = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v0, 7)));
        const uint32_t h1 = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v1, 7)));
        const uint32_t h2 = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v2, 7)));
        const uint32_t h3 = (uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(_mm256_slli_epi32(v3, 7)));
        const uint32_t hw = h0 | (h1 << 8) | (h2 << 16) | (h3 << 24);
        __builtin_memcpy(hi + i / 8, &hw, 4);
    }
    _mm_sfence();                                         // the streamed stores are ordered before the unit is announced
    alignas(32) uint32_t lanes[8];
    _mm256_store_si256(reinterpret_cast<__m256i *>(lanes), acc);
    uint32_t o = 0;
    for (int j = 0; j < 8; j++) o |= lanes[j];
    if (i < n) o |= pack25_scalar(src + i, n - i, lo3 + 3 * i, hi + i / 8);
    return o;
}

// n == 0 mod 8 except for the last piece of a chunk; lo3 + 3 * (literal offset) as the caller computed it
uint32_t host_pack25(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(lo3) & 31u) == 0) return pack25_avx2(src, n, lo3, hi);
    return pack25_scalar(src, n, lo3, hi);
}

} // namespace alll
