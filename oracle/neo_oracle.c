/* TEST INFRASTRUCTURE ONLY.
 *
 * neo_oracle.c -- CPU restatement (plain C11) of the reference's fallback FFT / UPOLS hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
 * The product (libneo_b200.so) never links, loads or calls anything in oracle/.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function below against
 *   (a) the reference's own known-answer tests (rfft_test.cpp:170-186, dct_test.cpp:23-39, fdl_index_test.cpp:13-65,
 *       stft_test.cpp:8-13, uniform_partition_test.cpp:8-38, multiply_add_test.cpp:52-95),
 *   (b) golden vectors in tests/golden/ produced by oracle/_ref/libneo_ref.so = the unmodified reference headers
 *       compiled in place (oracle/ref_wrapper.cpp, oracle/make_golden.py), and
 *   (c) oracle/_ref itself, side by side, whenever the prebuilt library is present.
 *
 * Paths in comments are relative to /root/reference/src/neo/.
 */
#include "neo_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- integers: bit-exact contract ------------------------------------------------------------------ */

/* bit/bit_ceil.hpp:34-41 (std::bit_ceil) */
size_t oracle_bit_ceil(size_t x)
{
    size_t p = 1;
    while (p < x) { p <<= 1; }
    return p;
}

/* fft/order.hpp:33-39: next_order(n) = log2(bit_ceil(n)) */
size_t oracle_next_order(size_t n)
{
    size_t const c = oracle_bit_ceil(n);
    size_t order   = 0;
    while (((size_t)1 << order) < c) { ++order; }
    return order;
}

/* fft/reference/bitrevorder.hpp:65-75: table[i] = bits of i reversed within `order` bits */
void oracle_bitrev_table(size_t order, uint32_t* table)
{
    size_t const size = (size_t)1 << order;
    for (size_t i = 0; i < size; ++i) {
        uint32_t v = 0;
        for (size_t j = 0; j < order; ++j) { v |= (uint32_t)((i >> j) & 1U) << (order - 1 - j); }
        table[i] = v;
    }
}

/* fft/reference/digitrevorder.hpp:27-45: base-`radix` digit-reversal LUT. The loop stops at size-1, so
 * lut[size-1] stays 0 (harmless: apply() tests i < lut[i], :21-25). Reproduced as is. */
void oracle_digitrev_lut(size_t radix, size_t size, uint32_t* lut)
{
    for (size_t i = 0; i < size; ++i) { lut[i] = 0; }
    size_t j = 0;
    for (size_t i = 0; i + 1 < size; ++i) {
        lut[i]   = (uint32_t)j;
        size_t k = (radix - 1U) * size / radix;
        while (k <= j) {
            j -= k;
            k /= radix;
        }
        j += k / (radix - 1U);
    }
}

/* digitrevorder_plan::operator() (:17-25) applied to iota: the permutation a caller observes */
void oracle_digitrev_perm(size_t radix, size_t size, uint32_t* perm)
{
    uint32_t* lut = (uint32_t*)malloc(sizeof(uint32_t) * size);
    oracle_digitrev_lut(radix, size, lut);
    for (size_t i = 0; i < size; ++i) { perm[i] = (uint32_t)i; }
    for (size_t i = 0; i < size; ++i) {
        if (i < lut[i]) {
            uint32_t const t = perm[i];
            perm[i]          = perm[lut[i]];
            perm[lut[i]]     = t;
        }
    }
    free(lut);
}

/* fft/stft.hpp:21-25 with math/idiv.hpp:11-14 (ceil-div): frames = ceil((signal-frame+overlap)/(frame-overlap)) + 1 */
size_t oracle_num_stft_frames(size_t signal, size_t frame, size_t overlap)
{
    size_t const x = signal - frame + overlap;
    size_t const y = frame - overlap;
    return (x + y - 1) / y + 1;
}

/* convolution/fdl_index.hpp:24-36: per call -> write_pos, then P pairs (segment, (write_pos + P - segment) % P),
 * then write_pos = (write_pos + 1) % P */
void oracle_fdl_index_sequence(size_t parts, size_t calls, uint32_t* write_pos, uint32_t* pairs)
{
    size_t w = 0;
    for (size_t c = 0; c < calls; ++c) {
        write_pos[c] = (uint32_t)w;
        for (size_t segment = 0; segment < parts; ++segment) {
            pairs[(c * parts + segment) * 2 + 0] = (uint32_t)segment;
            pairs[(c * parts + segment) * 2 + 1] = (uint32_t)((w + parts - segment) % parts);
        }
        if (++w >= parts) { w = 0; }
    }
}

/* ---- std::mt19937 + libstdc++ generate_canonical (the input distribution, testing/testing.hpp:37-72) ----- */
typedef struct oracle_mt19937
{
    uint32_t mt[624];
    int idx;
} oracle_mt19937;

static void oracle_mt19937_seed(oracle_mt19937* g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) { g->mt[i] = 1812433253U * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i; }
    g->idx = 624;
}

static uint32_t oracle_mt19937_next(oracle_mt19937* g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t const y = (g->mt[i] & 0x80000000U) | (g->mt[(i + 1) % 624] & 0x7fffffffU);
            g->mt[i]         = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= y >> 18;
    return y;
}

/* float: 24 digits <= 32 bits per draw -> one draw, / 2^32 in float; clamp to nextafter(1,0) */
static float oracle_canonical_f32(oracle_mt19937* g)
{
    float const sum = (float)oracle_mt19937_next(g) * 1.0F;
    float ret       = sum / 4294967296.0F;
    if (ret >= 1.0F) { ret = nextafterf(1.0F, 0.0F); }
    return ret;
}

/* double: 53 digits -> two draws: (d0 + d1 * 2^32) / 2^64 */
static double oracle_canonical_f64(oracle_mt19937* g)
{
    double sum = (double)oracle_mt19937_next(g) * 1.0;
    sum += (double)oracle_mt19937_next(g) * 4294967296.0;
    double ret = sum / 18446744073709551616.0;
    if (ret >= 1.0) { ret = nextafter(1.0, 0.0); }
    return ret;
}

#define REAL float
#define SUF(name) name##_f32
#define R_COS cosf
#define R_SIN sinf
#define R_SQRT sqrtf
#define R_FMA fmaf
#define R_LROUND lroundf
#include "neo_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_COS
#undef R_FMA
#undef R_LROUND
#undef R_SIN
#undef R_SQRT

#define REAL double
#define SUF(name) name##_f64
#define R_COS cos
#define R_SIN sin
#define R_SQRT sqrt
#define R_FMA fma
#define R_LROUND lround
#include "neo_oracle_impl.inc"
#undef REAL
#undef SUF
#undef R_COS
#undef R_FMA
#undef R_LROUND
#undef R_SIN
#undef R_SQRT
