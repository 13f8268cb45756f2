/* TEST INFRASTRUCTURE ONLY -- see neo_oracle.c. Complex data is interleaved [re, im]. direction: -1 forward, +1 backward
 * (fft/direction.hpp:8-12). */
#ifndef NEO_ORACLE_H
#define NEO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

size_t oracle_bit_ceil(size_t x);
size_t oracle_next_order(size_t n);
void oracle_bitrev_table(size_t order, uint32_t* table);
void oracle_digitrev_lut(size_t radix, size_t size, uint32_t* lut);
void oracle_digitrev_perm(size_t radix, size_t size, uint32_t* perm);
size_t oracle_num_stft_frames(size_t signal, size_t frame, size_t overlap);
void oracle_fdl_index_sequence(size_t parts, size_t calls, uint32_t* write_pos, uint32_t* pairs);

#define NEO_ORACLE_DECLARE(REAL, S)                                                                                    \
    void oracle_twiddle_lut_##S(size_t size, int direction, REAL* out);                                                \
    int oracle_fft_c2c_##S(size_t order, REAL* inout, int direction);                                                  \
    int oracle_dft_c2c_##S(size_t size, REAL* inout, int direction);                                                   \
    int oracle_dct2_##S(size_t order, REAL* inout);                                                                    \
    void oracle_rfft_##S(size_t order, REAL const* in, REAL* out);                                                     \
    void oracle_irfft_##S(size_t order, REAL const* in, size_t in_len, REAL* out);                                     \
    void oracle_multiply_add_##S(REAL const* x, REAL const* y, REAL const* z, REAL* out, size_t n);                    \
    size_t oracle_uniform_partition_##S(REAL const* ir, size_t channels, size_t len, size_t block, REAL* out);         \
    size_t oracle_stft_##S(REAL const* x, size_t channels, size_t len, size_t frame, size_t transform, size_t overlap, \
                           int window, REAL* out);                                                                     \
    void oracle_normalize_impulse_##S(REAL* ir, size_t channels, size_t len);                                          \
    struct oracle_conv_##S* oracle_conv_create_##S(int kind);                                                          \
    void oracle_conv_destroy_##S(struct oracle_conv_##S* c);                                                           \
    void oracle_conv_filter_##S(struct oracle_conv_##S* c, REAL const* h, size_t parts, size_t bins);                  \
    void oracle_conv_process_##S(struct oracle_conv_##S* c, REAL* inout, size_t num_samples);                          \
    void oracle_compress_row_##S(REAL const* in, size_t n, int bits, int16_t* out);                                    \
    void oracle_decompress_row_##S(int16_t const* in, size_t n, int bits, REAL* out);                                  \
    size_t oracle_csr_build_##S(REAL const* h, size_t rows, size_t cols, REAL threshold, uint64_t* row_ptr, uint64_t* col_idx,      \
                                REAL* values);                                                                         \
    void oracle_conv_filter_sparse_##S(struct oracle_conv_##S* c, REAL const* h, size_t parts, size_t bins, REAL threshold);       \
    void oracle_direct_convolve_##S(REAL const* sig, size_t sig_len, REAL const* ir, size_t ir_len, REAL* out, size_t out_len); \
    void oracle_fft_convolve_##S(REAL const* sig, size_t sig_len, REAL const* patch, size_t patch_len, REAL* out);      \
    void oracle_noise_##S(size_t n, uint32_t seed, REAL* out);

NEO_ORACLE_DECLARE(float, f32)
NEO_ORACLE_DECLARE(double, f64)

#ifdef __cplusplus
}
#endif

#endif
