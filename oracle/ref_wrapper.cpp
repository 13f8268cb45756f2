// TEST INFRASTRUCTURE ONLY. Not part of the product; never linked by libneo_b200.so.
//
// C entry points around the UNMODIFIED reference headers, compiled where they lie
// (/root/reference/src, see oracle/Makefile) into oracle/_ref/libneo_ref.so.
// It is the strongest checker we have: the reference's own `fft_plan`,
// `rfft_plan`, `uniform_partition` and `up{ols,ola}_convolver` running on the CPU.
// It is also the `--impl reference` / `cpu_baseline.kind == "reference"` arm of bench.py
// (threaded over channels, one convolver per channel as extra/cli/src/convolver.cpp:37-40 does).
//
// No xsimd (not installable offline): NEO_HAS_XSIMD is undefined, so the reference
// drops to its scalar loops (algorithm/multiply_add.hpp:298-300) = "neo's own fallback".

#include <neo/algorithm.hpp>
#include <neo/convolution.hpp>
#include <neo/fft.hpp>
#include <neo/fft/dct.hpp>
#include <neo/testing/testing.hpp>

#include <atomic>
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <variant>
#include <vector>

namespace {

template<typename T>
using vec_view = stdex::mdspan<T, stdex::dextents<std::size_t, 1>>;
template<typename T>
using mat_view = stdex::mdspan<T, stdex::dextents<std::size_t, 2>>;

template<typename Float>
auto c2c(std::size_t order, Float* inout, int direction) -> int
{
    using Complex = std::complex<Float>;
    try {
        auto plan = neo::fft::fft_plan<Complex>{neo::fft::from_order, order};
        auto x    = vec_view<Complex>{reinterpret_cast<Complex*>(inout), plan.size()};
        plan(x, direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward);
    } catch (std::exception const&) {
        return 1;
    }
    return 0;
}

template<typename Float>
auto r2c(std::size_t order, Float const* in, Float* out) -> void
{
    using Complex = std::complex<Float>;
    auto plan     = neo::fft::rfft_plan<Float, Complex>{neo::fft::from_order, order};
    auto x        = vec_view<Float const>{in, plan.size()};
    auto y        = vec_view<Complex>{reinterpret_cast<Complex*>(out), plan.size() / 2 + 1};
    neo::fft::rfft(plan, x, y);
}

template<typename Float>
auto c2r(std::size_t order, Float const* in, std::size_t in_len, Float* out) -> void
{
    using Complex = std::complex<Float>;
    auto plan     = neo::fft::rfft_plan<Float, Complex>{neo::fft::from_order, order};
    auto x        = vec_view<Complex const>{reinterpret_cast<Complex const*>(in), in_len};
    auto y        = vec_view<Float>{out, plan.size()};
    neo::fft::irfft(plan, x, y);
}

// stft_plan (fft/stft.hpp:39-109): window 0 rectangular, 1 hann, 2 hamming (math/windowing.hpp); out [C][frames][bins], returns frames
template<typename Float>
auto stft_impl(Float const* x, std::size_t channels, std::size_t len, std::size_t frame, std::size_t transform, std::size_t overlap,
               int window, Float* out) -> std::size_t
{
    auto options = neo::fft::stft_options<Float>{.frame_size = frame, .transform_size = transform, .overlap_size = overlap};
    if (window == 0) { options.window = neo::rectangular_window<Float>{}; }
    else if (window == 1) { options.window = neo::hann_window<Float>{}; }
    else { options.window = neo::hamming_window<Float>{}; }
    auto plan   = neo::fft::stft_plan<Float>{options};
    auto result = plan(mat_view<Float const>{x, channels, len});
    if (out != nullptr) { std::memcpy(out, result.data(), result.size() * sizeof(std::complex<Float>)); }
    return result.extent(1);
}

// fallback_dct2_plan (fft/dct.hpp:24-68)
template<typename Float>
auto dct2(std::size_t order, Float* inout) -> int
{
    if (order > 27) { return 1; }
    auto plan = neo::fft::fallback_dct2_plan<Float>{neo::fft::from_order, order};
    plan(vec_view<Float>{inout, plan.size()});
    return 0;
}

// dft_plan == fallback_dft_plan (fft/dft.hpp:28-30): Bluestein, any size
template<typename Float>
auto bluestein(std::size_t size, Float* inout, int direction) -> int
{
    using Complex = std::complex<Float>;
    if (size == 0) { return 1; }
    auto plan = neo::fft::dft_plan<Complex>{size};
    auto x    = vec_view<Complex>{reinterpret_cast<Complex*>(inout), size};
    plan(x, direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward);
    return 0;
}

template<typename Float>
struct convolver_box
{
    using Complex = std::complex<Float>;
    std::variant<
        neo::convolution::upols_convolver<Complex>,
        neo::convolution::upola_convolver<Complex>,
        neo::convolution::split_upols_convolver<Complex>,
        neo::convolution::split_upola_convolver<Complex>,
        neo::convolution::upola_convolver_v2<Complex>,
        neo::convolution::sparse_upols_convolver<Complex>,
        neo::convolution::sparse_upola_convolver<Complex>>
        impl;

    explicit convolver_box(int kind)
    {
        switch (kind) {
            case 0: impl.template emplace<0>(); break;
            case 1: impl.template emplace<1>(); break;
            case 2: impl.template emplace<2>(); break;
            case 3: impl.template emplace<3>(); break;
            case 5: impl.template emplace<5>(); break;
            case 6: impl.template emplace<6>(); break;
            default: impl.template emplace<4>(); break;
        }
    }

    auto filter(Float const* h, std::size_t parts, std::size_t bins) -> void
    {
        auto view = mat_view<Complex const>{reinterpret_cast<Complex const*>(h), parts, bins};
        std::visit(
            [&](auto& c) {
                using conv_t = std::remove_cvref_t<decltype(c)>;
                if constexpr (!std::is_same_v<conv_t, neo::convolution::sparse_upols_convolver<Complex>>
                              && !std::is_same_v<conv_t, neo::convolution::sparse_upola_convolver<Complex>>) {
                    c.filter(view);
                }
            },
            impl
        );
    }

    // sparse_filter::filter(input, sparsity) (sparse_filter.hpp:25-28); predicate: comparisons only (see neo_oracle_impl.inc)
    auto filter_sparse(Float const* h, std::size_t parts, std::size_t bins, Float threshold) -> void
    {
        auto view       = mat_view<Complex const>{reinterpret_cast<Complex const*>(h), parts, bins};
        auto const keep = [threshold](auto /*row*/, auto /*col*/, Complex v) {
            return std::abs(v.real()) > threshold || std::abs(v.imag()) > threshold;
        };
        if (auto* c = std::get_if<5>(&impl)) { c->filter(view, keep); }
        if (auto* c = std::get_if<6>(&impl)) { c->filter(view, keep); }
    }

    auto process(Float* block, std::size_t n) -> void
    {
        auto view = vec_view<Float>{block, n};
        std::visit([&](auto& c) { c(view); }, impl);
    }
};

}  // namespace

extern "C" {

// ---- plans ---------------------------------------------------------------------------
int ref_fft_c2c_f32(std::size_t order, float* inout, int direction) { return c2c<float>(order, inout, direction); }
int ref_fft_c2c_f64(std::size_t order, double* inout, int direction) { return c2c<double>(order, inout, direction); }
int ref_dct2_f32(std::size_t order, float* inout) { return dct2<float>(order, inout); }
int ref_dct2_f64(std::size_t order, double* inout) { return dct2<double>(order, inout); }
int ref_dft_c2c_f32(std::size_t size, float* inout, int direction) { return bluestein<float>(size, inout, direction); }
int ref_dft_c2c_f64(std::size_t size, double* inout, int direction) { return bluestein<double>(size, inout, direction); }
void ref_rfft_f32(std::size_t order, float const* in, float* out) { r2c<float>(order, in, out); }
void ref_rfft_f64(std::size_t order, double const* in, double* out) { r2c<double>(order, in, out); }
void ref_irfft_f32(std::size_t order, float const* in, std::size_t n, float* out) { c2r<float>(order, in, n, out); }
void ref_irfft_f64(std::size_t order, double const* in, std::size_t n, double* out) { c2r<double>(order, in, n, out); }

// split-complex plan (fft/fallback/fallback_split_fft_plan.hpp:16-137), in place on separate real / imaginary planes
void ref_split_fft_f32(std::size_t order, float* re, float* im, int direction)
{
    auto plan = neo::fft::split_fft_plan<float>{neo::fft::from_order, order};
    auto x    = neo::split_complex{vec_view<float>{re, plan.size()}, vec_view<float>{im, plan.size()}};
    plan(x, direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward);
}

void ref_split_fft_f64(std::size_t order, double* re, double* im, int direction)
{
    auto plan = neo::fft::split_fft_plan<double>{neo::fft::from_order, order};
    auto x    = neo::split_complex{vec_view<double>{re, plan.size()}, vec_view<double>{im, plan.size()}};
    plan(x, direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward);
}

std::size_t ref_fft_max_order() { return neo::fft::fft_plan<std::complex<float>>::max_order(); }
std::size_t ref_next_order(std::size_t n) { return neo::fft::next_order(n); }

// ---- integer tables (bit-exact contract) ------------------------------------------------
// bitrevorder_plan keeps its table private (fft/reference/bitrevorder.hpp:77); applying the
// involution to iota reads it back exactly.
void ref_bitrev_table(std::size_t order, std::uint32_t* out)
{
    auto const n = std::size_t(1) << order;
    auto plan    = neo::fft::bitrevorder_plan{order};
    auto buf     = std::vector<std::complex<double>>(n);
    for (std::size_t i = 0; i < n; ++i) { buf[i] = double(i); }
    plan(vec_view<std::complex<double>>{buf.data(), n});
    for (std::size_t i = 0; i < n; ++i) { out[i] = static_cast<std::uint32_t>(buf[i].real()); }
}

// permutation produced by digitrevorder_plan<Radix> (fft/reference/digitrevorder.hpp:13-48)
void ref_digitrev_perm(std::size_t radix, std::size_t size, std::uint32_t* out)
{
    auto buf = std::vector<std::complex<double>>(size);
    for (std::size_t i = 0; i < size; ++i) { buf[i] = double(i); }
    auto view = vec_view<std::complex<double>>{buf.data(), size};
    switch (radix) {
        case 2: neo::fft::digitrevorder_plan<2>{size}(view); break;
        case 3: neo::fft::digitrevorder_plan<3>{size}(view); break;
        case 4: neo::fft::digitrevorder_plan<4>{size}(view); break;
        case 5: neo::fft::digitrevorder_plan<5>{size}(view); break;
        case 8: neo::fft::digitrevorder_plan<8>{size}(view); break;
        default: break;
    }
    for (std::size_t i = 0; i < size; ++i) { out[i] = static_cast<std::uint32_t>(buf[i].real()); }
}

// twiddle LUT exactly as the plan builds it (fft/twiddle.hpp:47-52), interleaved re/im
void ref_twiddle_lut_f32(std::size_t size, int direction, float* out)
{
    auto lut = neo::fft::make_twiddle_lut_radix2<std::complex<float>>(
        size,
        direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward
    );
    std::memcpy(out, lut.data(), sizeof(float) * size);
}

void ref_twiddle_lut_f64(std::size_t size, int direction, double* out)
{
    auto lut = neo::fft::make_twiddle_lut_radix2<std::complex<double>>(
        size,
        direction < 0 ? neo::fft::direction::forward : neo::fft::direction::backward
    );
    std::memcpy(out, lut.data(), sizeof(double) * size);
}

// fdl_index sequence: for each of `calls` invocations records write_pos and the P (fdl,filter) pairs
void ref_fdl_index_sequence(std::size_t parts, std::size_t calls, std::uint32_t* write_pos, std::uint32_t* pairs)
{
    auto idx = neo::convolution::fdl_index<std::size_t>{parts};
    for (std::size_t c = 0; c < calls; ++c) {
        auto n = std::size_t(0);
        idx(
            [&](std::size_t w) { write_pos[c] = static_cast<std::uint32_t>(w); },
            [&](std::size_t fdl, std::size_t filt) {
                pairs[(c * parts + n) * 2 + 0] = static_cast<std::uint32_t>(fdl);
                pairs[(c * parts + n) * 2 + 1] = static_cast<std::uint32_t>(filt);
                ++n;
            }
        );
    }
}

std::size_t ref_num_stft_frames(std::size_t signal, std::size_t frame, std::size_t overlap)
{
    return neo::fft::detail::num_sftf_frames(signal, frame, overlap);
}

// ---- inputs ------------------------------------------------------------------------------
void ref_noise_f32(std::size_t n, std::uint32_t seed, float* out)
{
    auto sig = neo::generate_noise_signal<float>(n, seed);
    std::memcpy(out, sig.data(), sizeof(float) * n);
}

void ref_noise_f64(std::size_t n, std::uint32_t seed, double* out)
{
    auto sig = neo::generate_noise_signal<double>(n, seed);
    std::memcpy(out, sig.data(), sizeof(double) * n);
}

void ref_noise_c64(std::size_t n, std::uint32_t seed, float* out)
{
    auto sig = neo::generate_noise_signal<std::complex<float>>(n, seed);
    std::memcpy(out, sig.data(), sizeof(float) * 2 * n);
}

void ref_noise_c128(std::size_t n, std::uint32_t seed, double* out)
{
    auto sig = neo::generate_noise_signal<std::complex<double>>(n, seed);
    std::memcpy(out, sig.data(), sizeof(double) * 2 * n);
}

void ref_normalize_impulse_f32(float* ir, std::size_t channels, std::size_t len)
{
    neo::convolution::normalize_impulse(mat_view<float>{ir, channels, len});
}

// ---- elementwise -------------------------------------------------------------------------
void ref_multiply_add_c64(float const* x, float const* y, float const* z, float* out, std::size_t n)
{
    using C = std::complex<float>;
    neo::multiply_add(
        vec_view<C const>{reinterpret_cast<C const*>(x), n},
        vec_view<C const>{reinterpret_cast<C const*>(y), n},
        vec_view<C const>{reinterpret_cast<C const*>(z), n},
        vec_view<C>{reinterpret_cast<C*>(out), n}
    );
}

// ---- filter preparation -------------------------------------------------------------------
// uniform_partition (convolution/uniform_partition.hpp:13-26): out is [C][P][B+1] complex, returns P
// fft_convolve (convolution/fft_convolver.hpp:84-93), mode::full: out is [n + m - 1]
void ref_fft_convolve_f32(float const* signal, std::size_t n, float const* patch, std::size_t m, float* out)
{
    auto result = neo::convolution::fft_convolve(vec_view<float const>{signal, n}, vec_view<float const>{patch, m});
    std::memcpy(out, result.data(), result.size() * sizeof(float));
}
void ref_fft_convolve_f64(double const* signal, std::size_t n, double const* patch, std::size_t m, double* out)
{
    auto result = neo::convolution::fft_convolve(vec_view<double const>{signal, n}, vec_view<double const>{patch, m});
    std::memcpy(out, result.data(), result.size() * sizeof(double));
}
std::size_t ref_stft_f32(float const* x, std::size_t channels, std::size_t len, std::size_t frame, std::size_t transform,
                         std::size_t overlap, int window, float* out)
{
    return stft_impl<float>(x, channels, len, frame, transform, overlap, window, out);
}
std::size_t ref_stft_f64(double const* x, std::size_t channels, std::size_t len, std::size_t frame, std::size_t transform,
                         std::size_t overlap, int window, double* out)
{
    return stft_impl<double>(x, channels, len, frame, transform, overlap, window, out);
}
std::size_t ref_uniform_partition_f32(float const* ir, std::size_t channels, std::size_t len, std::size_t block, float* out)
{
    auto parts = neo::convolution::uniform_partition(mat_view<float const>{ir, channels, len}, block);
    if (out != nullptr) { std::memcpy(out, parts.data(), sizeof(float) * 2 * parts.size()); }
    return parts.extent(1);
}

std::size_t
ref_uniform_partition_f64(double const* ir, std::size_t channels, std::size_t len, std::size_t block, double* out)
{
    auto parts = neo::convolution::uniform_partition(mat_view<double const>{ir, channels, len}, block);
    if (out != nullptr) { std::memcpy(out, parts.data(), sizeof(double) * 2 * parts.size()); }
    return parts.extent(1);
}

// ---- overlap policies with the identity callback (convolution/overlap_test.cpp:21-64) -----------
void ref_overlap_identity_f32(int add, std::size_t block, std::size_t filter_size, float* signal, std::size_t nblocks)
{
    using C = std::complex<float>;
    auto run = [&](auto policy) {
        for (std::size_t b = 0; b < nblocks; ++b) {
            policy(vec_view<float>{signal + b * block, block}, [](auto) {});
        }
    };
    if (add != 0) {
        run(neo::convolution::overlap_add<C>{block, filter_size});
    } else {
        run(neo::convolution::overlap_save<C>{block, filter_size});
    }
}

// ---- convolvers ----------------------------------------------------------------------------
// kind: 0 upols, 1 upola, 2 split_upols, 3 split_upola, 4 upola_v2 (convolution/dense_convolver.hpp:20-41)
void* ref_conv_create_f32(int kind) { return new convolver_box<float>{kind}; }
void* ref_conv_create_f64(int kind) { return new convolver_box<double>{kind}; }
void ref_conv_destroy_f32(void* h) { delete static_cast<convolver_box<float>*>(h); }
void ref_conv_destroy_f64(void* h) { delete static_cast<convolver_box<double>*>(h); }
void ref_conv_filter_f32(void* h, float const* f, std::size_t parts, std::size_t bins)
{
    static_cast<convolver_box<float>*>(h)->filter(f, parts, bins);
}
void ref_conv_filter_f64(void* h, double const* f, std::size_t parts, std::size_t bins)
{
    static_cast<convolver_box<double>*>(h)->filter(f, parts, bins);
}
// compressed_fdl<complex<float>, scalar_complex<int8/int16>> (convolution/compressed_fdl.hpp:17-52): insert `in` [n] as row 0 of a
// 2-row delay line and read it back through operator[] (the compressed_accessor): out [n] complex
void ref_compressed_fdl_roundtrip_f32(float const* in, std::size_t n, int bits, float* out)
{
    using C       = std::complex<float>;
    auto const go = [&](auto fdl) {
        fdl.insert(vec_view<C const>{reinterpret_cast<C const*>(in), n}, 0);
        auto const row = fdl[0];
        for (std::size_t i = 0; i < n; ++i) {
            C const v      = row[i];
            out[2 * i]     = v.real();
            out[2 * i + 1] = v.imag();
        }
    };
    auto const ext = Kokkos::dextents<std::size_t, 2>{2, n};
    if (bits == 8) { go(neo::convolution::compressed_fdl<C, neo::scalar_complex<std::int8_t>>{ext}); }
    else { go(neo::convolution::compressed_fdl<C, neo::scalar_complex<std::int16_t>>{ext}); }
}
void ref_conv_filter_sparse_f32(void* h, float const* f, std::size_t parts, std::size_t bins, float threshold)
{
    static_cast<convolver_box<float>*>(h)->filter_sparse(f, parts, bins, threshold);
}
void ref_conv_filter_sparse_f64(void* h, double const* f, std::size_t parts, std::size_t bins, double threshold)
{
    static_cast<convolver_box<double>*>(h)->filter_sparse(f, parts, bins, threshold);
}
// neo::csr_matrix(matrix, filter) (container/csr_matrix.hpp:64-98) with the same predicate: its three containers, copied out
std::size_t ref_csr_build_f32(float const* h, std::size_t rows, std::size_t cols, float threshold, std::uint64_t* row_ptr,
                              std::uint64_t* col_idx, float* values)
{
    using C   = std::complex<float>;
    auto view = mat_view<C const>{reinterpret_cast<C const*>(h), rows, cols};
    auto csr  = neo::csr_matrix<C>{view, [threshold](auto, auto, C v) {
                                      return std::abs(v.real()) > threshold || std::abs(v.imag()) > threshold;
                                  }};
    auto const& r = csr.row_container();
    auto const& c = csr.column_container();
    auto const& v = csr.value_container();
    if (row_ptr != nullptr) { std::copy(r.begin(), r.end(), row_ptr); }
    if (col_idx != nullptr) { std::copy(c.begin(), c.end(), col_idx); }
    if (values != nullptr) { std::memcpy(values, v.data(), v.size() * sizeof(C)); }
    return v.size();
}
void ref_conv_process_f32(void* h, float* block, std::size_t n) { static_cast<convolver_box<float>*>(h)->process(block, n); }
void ref_conv_process_f64(void* h, double* block, std::size_t n)
{
    static_cast<convolver_box<double>*>(h)->process(block, n);
}

// ---- timed CPU baselines (bench.py cpu_baseline / --impl reference) ------------------------------
// `channels` independent convolvers (own filter each, filter[c] = H + c*filter_stride complex elements; stride 0
// shares one filter so the sample fits host RAM), spread over `threads` std::threads.
// Every channel processes `nblocks` blocks of B = bins-1 samples in place in signal[c][nblocks*B].
// Returns seconds spent in the block loop only (filter setup excluded, as convolution.cpp:26-40 excludes it).
// Worker threads are created first and released together by a start gate, so thread creation is outside the timed region.
namespace {
struct start_gate
{
    std::atomic<bool> go{false};
    void wait() const
    {
        while (!go.load(std::memory_order_acquire)) { std::this_thread::yield(); }
    }
};
}  // namespace

double ref_conv_bench_f32(
    int kind,
    float const* filters,
    std::size_t filter_stride,
    std::size_t parts,
    std::size_t bins,
    float* signal,
    std::size_t channels,
    std::size_t nblocks,
    std::size_t threads
)
{
    auto const block = bins - 1;
    auto boxes       = std::vector<std::unique_ptr<convolver_box<float>>>(channels);

    // returns the seconds between the release of the workers and the last one finishing
    auto for_channels = [&](auto fn) {
        auto next = std::atomic<std::size_t>{0};
        auto gate = start_gate{};
        auto pool = std::vector<std::thread>{};
        for (std::size_t t = 0; t < threads; ++t) {
            pool.emplace_back([&] {
                gate.wait();
                for (auto c = next.fetch_add(1); c < channels; c = next.fetch_add(1)) { fn(c); }
            });
        }
        auto const start = std::chrono::steady_clock::now();
        gate.go.store(true, std::memory_order_release);
        for (auto& t : pool) { t.join(); }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    };

    for_channels([&](std::size_t c) {
        boxes[c] = std::make_unique<convolver_box<float>>(kind);
        boxes[c]->filter(filters + c * filter_stride * 2, parts, bins);
    });

    return for_channels([&](std::size_t c) {
        for (std::size_t b = 0; b < nblocks; ++b) { boxes[c]->process(signal + (c * nblocks + b) * block, block); }
    });
}

// batched rfft+irfft round trips, batch split over threads, one plan per thread (the pair the reference's
// own driver times, extra/benchmark/src/rfft.cpp:22-31). Plans and scratch are built before the start gate opens.
double ref_rfft_bench_f32(std::size_t order, float* data, std::size_t batch, std::size_t threads)
{
    using C      = std::complex<float>;
    auto const n = std::size_t(1) << order;
    auto next    = std::atomic<std::size_t>{0};
    auto ready   = std::atomic<std::size_t>{0};
    auto gate    = start_gate{};
    auto pool    = std::vector<std::thread>{};
    for (std::size_t t = 0; t < threads; ++t) {
        pool.emplace_back([&] {
            auto plan = neo::fft::rfft_plan<float, C>{neo::fft::from_order, order};
            auto spec = std::vector<C>(n);
            ready.fetch_add(1);
            gate.wait();
            for (auto b = next.fetch_add(1); b < batch; b = next.fetch_add(1)) {
                auto x = vec_view<float>{data + b * n, n};
                neo::fft::rfft(plan, x, vec_view<C>{spec.data(), n});
                neo::fft::irfft(plan, vec_view<C>{spec.data(), n}, x);
            }
        });
    }
    while (ready.load() < threads) { std::this_thread::yield(); }
    auto const t0 = std::chrono::steady_clock::now();
    gate.go.store(true, std::memory_order_release);
    for (auto& t : pool) { t.join(); }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// BASELINE config 1 (extra/benchmark/src/fft.cpp:12-38): one complex<float> plan of 2^order points, forward then inverse then 1/N
// scale in place, `reps` round trips on one thread. Returns seconds.
double ref_c2c_bench_f32(std::size_t order, float* inout, std::size_t reps)
{
    using C      = std::complex<float>;
    auto const n = std::size_t(1) << order;
    auto plan    = neo::fft::fft_plan<C>{neo::fft::from_order, order};
    auto x       = vec_view<C>{reinterpret_cast<C*>(inout), n};
    auto const t0 = std::chrono::steady_clock::now();
    for (std::size_t r = 0; r < reps; ++r) {
        neo::fft::fft(plan, x);
        neo::fft::ifft(plan, x);
        neo::scale(1.0F / static_cast<float>(n), x);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
