// hostpack_check.cpp -- CPU check of the host packer of the packed H2D transport (csrc/hostpack.cpp), compiled together
// with it by tests/test_boundary.py: AVX2 path against the scalar path byte for byte, both against an independent
// unpack that follows the device kernel's arithmetic (layout.cu: unpack25_kernel), odd lengths and the OR word included.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace alll { uint32_t host_pack25(const uint32_t *src, size_t n, uint8_t *lo3, uint8_t *hi); }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 20); }

// the device kernel's per-group arithmetic on six 32-bit words + one bit-24 byte
static void unpack_group(const uint8_t *lo3, uint32_t hb, uint32_t out[8])
{
    uint32_t w[6];
    std::memcpy(w, lo3, 24);
    out[0] = (w[0] & 0xFFFFFFu) | ((hb & 1u) << 24);
    out[1] = (w[0] >> 24) | ((w[1] & 0xFFFFu) << 8) | (((hb >> 1) & 1u) << 24);
    out[2] = (w[1] >> 16) | ((w[2] & 0xFFu) << 16) | (((hb >> 2) & 1u) << 24);
    out[3] = (w[2] >> 8) | (((hb >> 3) & 1u) << 24);
    out[4] = (w[3] & 0xFFFFFFu) | (((hb >> 4) & 1u) << 24);
    out[5] = (w[3] >> 24) | ((w[4] & 0xFFFFu) << 8) | (((hb >> 5) & 1u) << 24);
    out[6] = (w[4] >> 16) | ((w[5] & 0xFFu) << 16) | (((hb >> 6) & 1u) << 24);
    out[7] = (w[5] >> 8) | (((hb >> 7) & 1u) << 24);
}

int main()
{
    const size_t sizes[] = {0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 96, 100, 1000, 4096, 131072, 131072 + 24, 1000003};
    for (size_t n : sizes) {
        for (int mode = 0; mode < 2; mode++) {                     // 0: 25-bit literals, 1: one literal above 25 bits somewhere
            std::vector<uint32_t> src(n + 8);
            uint32_t want_or = 0;
            for (size_t i = 0; i < n; i++) { src[i] = rnd() & 0x1FFFFFFu; want_or |= src[i]; }
            if (mode == 1 && n) { src[n / 2] |= 1u << 27; want_or |= src[n / 2]; }
            const size_t lo_bytes = 3 * n + 64, hi_bytes = (n + 7) / 8 + 8;
            // 64-byte aligned buffers (vector path) and a copy shifted by one byte (scalar path)
            uint8_t *raw_a = (uint8_t *)aligned_alloc(64, (lo_bytes + 63) / 64 * 64 + 64), *raw_b = (uint8_t *)aligned_alloc(64, (lo_bytes + 63) / 64 * 64 + 64);
            std::vector<uint8_t> hi_a(hi_bytes, 0xAA), hi_b(hi_bytes, 0x55);
            std::memset(raw_a, 0xCC, lo_bytes); std::memset(raw_b, 0x33, lo_bytes);
            const uint32_t or_a = alll::host_pack25(src.data(), n, raw_a, hi_a.data());
            const uint32_t or_b = alll::host_pack25(src.data(), n, raw_b + 1, hi_b.data());
            if (or_a != want_or || or_b != want_or) { printf("n=%zu mode=%d: OR word %x / %x, want %x\n", n, mode, or_a, or_b, want_or); return 1; }
            if (std::memcmp(raw_a, raw_b + 1, 3 * n) != 0) { printf("n=%zu: vector and scalar low bytes differ\n", n); return 1; }
            if (std::memcmp(hi_a.data(), hi_b.data(), (n + 7) / 8) != 0) { printf("n=%zu: vector and scalar bit-24 bytes differ\n", n); return 1; }
            for (size_t i = 3 * n; i < lo_bytes; i++) if (raw_a[i] != 0xCC) { printf("n=%zu: wrote past the low bytes (%zu)\n", n, i); return 1; }
            for (size_t i = (n + 7) / 8; i < hi_bytes; i++) if (hi_a[i] != 0xAA) { printf("n=%zu: wrote past the bit-24 bytes\n", n); return 1; }
            for (size_t g = 0; g * 8 < n; g++) {
                uint8_t grp[24] = {0};
                const size_t have = (n - g * 8 < 8 ? n - g * 8 : 8);
                std::memcpy(grp, raw_a + 24 * g, 3 * have);
                uint32_t out[8];
                unpack_group(grp, hi_a[g], out);
                for (size_t j = 0; j < have; j++)
                    if (out[j] != (src[g * 8 + j] & 0x1FFFFFFu)) { printf("n=%zu: literal %zu comes back as %x, was %x\n", n, g * 8 + j, out[j], src[g * 8 + j]); return 1; }
            }
            free(raw_a); free(raw_b);
        }
    }
    printf("hostpack ok\n");
    return 0;
}
