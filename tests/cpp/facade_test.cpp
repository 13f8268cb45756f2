// facade_test.cpp -- the reference's own plan / convolver tests, restated against the C++ facade (include/neo_b200.hpp).
// Views are cuda::std::mdspan (same interface as the Kokkos::mdspan neo uses). Needs a GPU; run by tests/test_cpp_facade.py.
//
// Mirrors: fft/fft_test.cpp:53-130 (size/order, throw past max_size, round trips in place / out of place / strided view),
//          fft/rfft_test.cpp:39-62,170-186 (round trip, FFT([1,2,3,4]) known answer),
//          convolution/uniform_partitioned_convolver_test.cpp:35-75 (identity filter), plus oracle parity.
#include "../../include/neo_b200.hpp"
#include "../../oracle/neo_oracle.h"

#include <cuda/std/array>
#include <cuda/std/mdspan>

#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

namespace stdex = cuda::std;

template<typename T>
using vec = stdex::mdspan<T, stdex::dextents<std::size_t, 1>>;
template<typename T>
using strided_vec = stdex::mdspan<T, stdex::dextents<std::size_t, 1>, stdex::layout_stride>;
template<typename T>
using mat = stdex::mdspan<T, stdex::dextents<std::size_t, 2>>;

enum struct direction : int
{
    forward  = -1,
    backward = 1,
};

static int failures = 0;
#define REQUIRE(cond)                                                                                                  \
    do {                                                                                                               \
        if (!(cond)) {                                                                                                 \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                              \
            ++failures;                                                                                                \
        }                                                                                                              \
    } while (0)

template<typename A, typename B>
static auto rel_l2(A const& got, B const& want) -> double
{
    double num = 0, den = 0;
    for (std::size_t i = 0; i < want.size(); ++i) {
        num += std::norm(std::complex<double>(got[i]) - std::complex<double>(want[i]));
        den += std::norm(std::complex<double>(want[i]));
    }
    return std::sqrt(num / (den > 0 ? den : 1));
}

template<typename Float>
static auto noise(std::size_t n, unsigned seed) -> std::vector<Float>
{
    auto v = std::vector<Float>(n);
    if constexpr (std::is_same_v<Float, float>) { oracle_noise_f32(n, seed, v.data()); }
    else { oracle_noise_f64(n, seed, v.data()); }
    return v;
}

template<typename Float>
static auto test_fft_plan() -> void
{
    using Complex = std::complex<Float>;
    using Plan    = neo::b200::fft_plan<Complex>;
    auto const tol = std::is_same_v<Float, float> ? 1e-5 : 1e-12;

    // fft_test.cpp:62-67
    bool threw = false;
    try {
        auto p = Plan{neo::b200::from_order, neo_b200_next_order(Plan::max_size() + 1)};
    } catch (std::runtime_error const&) {
        threw = true;
    }
    REQUIRE(threw);

    for (std::size_t order = 2; order <= 14; ++order) {
        auto plan = Plan{neo::b200::from_order, order};
        REQUIRE(plan.order() == order);
        REQUIRE(plan.size() == (std::size_t(1) << order));
        auto const n = plan.size();

        auto raw  = noise<Float>(2 * n, 1);
        auto orig = std::vector<Complex>(n);
        for (std::size_t i = 0; i < n; ++i) { orig[i] = {raw[2 * i], raw[2 * i + 1]}; }

        // oracle parity, in place
        auto want = raw;
        if constexpr (std::is_same_v<Float, float>) { oracle_fft_c2c_f32(order, want.data(), -1); }
        else { oracle_fft_c2c_f64(order, want.data(), -1); }
        auto x = orig;
        plan(vec<Complex>{x.data(), n}, direction::forward);
        auto wantc = std::vector<Complex>(n);
        for (std::size_t i = 0; i < n; ++i) { wantc[i] = {want[2 * i], want[2 * i + 1]}; }
        REQUIRE(rel_l2(x, wantc) <= tol);

        // round trip in place (fft_test.cpp:79-91)
        plan(vec<Complex>{x.data(), n}, direction::backward);
        for (auto& v : x) { v /= Float(n); }
        REQUIRE(rel_l2(x, orig) <= tol);

        // out of place (fft_test.cpp:93-110)
        auto y = std::vector<Complex>(n);
        plan(vec<Complex const>{orig.data(), n}, vec<Complex>{y.data(), n}, direction::forward);
        REQUIRE(rel_l2(y, wantc) <= tol);

        // stride-2 view (fft_test.cpp:114-128)
        auto wide = std::vector<Complex>(2 * n, Complex{});
        for (std::size_t i = 0; i < n; ++i) { wide[2 * i + 1] = orig[i]; }
        auto const map = stdex::layout_stride::mapping<stdex::dextents<std::size_t, 1>>{stdex::dextents<std::size_t, 1>{n}, cuda::std::array<std::size_t, 1>{2}};
        plan(strided_vec<Complex>{wide.data() + 1, map}, direction::forward);
        auto col = std::vector<Complex>(n);
        bool untouched = true;
        for (std::size_t i = 0; i < n; ++i) {
            col[i]    = wide[2 * i + 1];
            untouched = untouched && wide[2 * i] == Complex{};
        }
        REQUIRE(rel_l2(col, wantc) <= tol);
        REQUIRE(untouched);
    }
}

template<typename Float>
static auto test_rfft_plan() -> void
{
    using Complex  = std::complex<Float>;
    auto const tol = std::is_same_v<Float, float> ? 1e-5 : 1e-12;
    for (std::size_t order = 2; order <= 14; ++order) {
        auto plan    = neo::b200::rfft_plan<Float>{neo::b200::from_order, order};
        auto const n = plan.size();
        REQUIRE(plan.order() == order);
        auto sig  = noise<Float>(n, 2);
        auto spec = std::vector<Complex>(n);  // N-long buffer like overlap_save.hpp:57
        plan(vec<Float const>{sig.data(), n}, vec<Complex>{spec.data(), n});
        auto want = std::vector<Float>(2 * n);
        if constexpr (std::is_same_v<Float, float>) { oracle_rfft_f32(order, sig.data(), want.data()); }
        else { oracle_rfft_f64(order, sig.data(), want.data()); }
        auto wantc = std::vector<Complex>(n / 2 + 1), gotc = std::vector<Complex>(n / 2 + 1);
        for (std::size_t i = 0; i <= n / 2; ++i) {
            wantc[i] = {want[2 * i], want[2 * i + 1]};
            gotc[i]  = spec[i];
        }
        REQUIRE(rel_l2(gotc, wantc) <= tol);
        auto back = std::vector<Float>(n);
        plan(vec<Complex const>{spec.data(), n}, vec<Float>{back.data(), n});  // unnormalised
        for (auto& v : back) { v /= Float(n); }
        REQUIRE(rel_l2(back, sig) <= tol);
    }
    // rfft_test.cpp:170-186: FFT([1,2,3,4]) = [10, -2+2i, -2, -2-2i]
    auto plan = neo::b200::fft_plan<Complex>{neo::b200::from_order, 2};
    auto x    = std::vector<Complex>{{1, 0}, {2, 0}, {3, 0}, {4, 0}};
    plan(vec<Complex>{x.data(), 4}, direction::forward);
    auto const kat = std::vector<Complex>{{10, 0}, {-2, 2}, {-2, 0}, {-2, -2}};
    REQUIRE(rel_l2(x, kat) <= 1e-6);
}

template<typename Float, template<typename> class Convolver, int OracleKind>
static auto test_convolver() -> void
{
    using Complex  = std::complex<Float>;
    auto const tol = std::is_same_v<Float, float> ? 1e-5 : 1e-12;
    // identity filter: uniform_partitioned_convolver_test.cpp:35-75
    for (std::size_t block : {128U, 256U, 512U, 1024U}) {
        auto const bins = block + 1;
        auto h          = std::vector<Complex>(3 * bins, Complex{});
        for (std::size_t k = 0; k < bins; ++k) { h[k] = {1, 0}; }
        auto conv = Convolver<Complex>{};
        conv.filter(mat<Complex const>{h.data(), 3, bins});
        auto sig = noise<Float>(block * 20, 7);
        auto out = sig;
        for (std::size_t b = 0; b < 20; ++b) { conv(vec<Float>{out.data() + b * block, block}); }
        REQUIRE(rel_l2(out, sig) <= 1e-5);
    }
    // oracle parity with a real impulse response, filter swapped once (plugin does: DenseConvolution.cpp:78-108)
    std::size_t const block = 128, taps = 1000, nblocks = 16;
    auto conv = Convolver<Complex>{};
    for (unsigned seed : {11U, 31U}) {
        auto ir    = noise<Float>(taps, seed);
        auto parts = neo_b200_num_partitions(taps, block);
        auto h     = std::vector<Complex>(parts * (block + 1));
        neo::b200::uniform_partition(ir.data(), 1, taps, block, h.data());
        conv.filter(mat<Complex const>{h.data(), parts, block + 1});
        auto sig = noise<Float>(block * nblocks, 13);
        auto got = sig, want = sig;
        for (std::size_t b = 0; b < nblocks; ++b) { conv(vec<Float>{got.data() + b * block, block}); }
        if constexpr (std::is_same_v<Float, float>) {
            auto* o = oracle_conv_create_f32(OracleKind);
            oracle_conv_filter_f32(o, reinterpret_cast<float const*>(h.data()), parts, block + 1);
            for (std::size_t b = 0; b < nblocks; ++b) { oracle_conv_process_f32(o, want.data() + b * block, block); }
            oracle_conv_destroy_f32(o);
        } else {
            auto* o = oracle_conv_create_f64(OracleKind);
            oracle_conv_filter_f64(o, reinterpret_cast<double const*>(h.data()), parts, block + 1);
            for (std::size_t b = 0; b < nblocks; ++b) { oracle_conv_process_f64(o, want.data() + b * block, block); }
            oracle_conv_destroy_f64(o);
        }
        REQUIRE(rel_l2(got, want) <= tol);
    }
}

// dft_plan (fft/dft_test.cpp:17-67): Plan{size}, identity and round trip, any size
template<typename Float>
static auto test_dft_plan() -> void
{
    using Complex  = std::complex<Float>;
    auto const tol = std::is_same_v<Float, float> ? 1e-5 : 1e-12;
    for (std::size_t size : {2U, 3U, 21U, 100U, 127U}) {
        auto plan = neo::b200::dft_plan<Complex>{size};
        REQUIRE(plan.size() == size);
        auto x = std::vector<Complex>(size, Complex{});
        x[0]   = Complex{1, 0};
        plan(vec<Complex>{x.data(), size}, direction::forward);
        auto const ones = std::vector<Complex>(size, Complex{1, 0});
        REQUIRE(rel_l2(x, ones) <= tol);
        plan(vec<Complex>{x.data(), size}, direction::backward);
        REQUIRE(std::abs(x[0].real() - Float(size)) <= Float(size) * Float(tol) * 10);
        auto re  = noise<Float>(size, 3);
        auto im  = noise<Float>(size, 4);
        auto sig = std::vector<Complex>(size);
        for (std::size_t i = 0; i < size; ++i) { sig[i] = {re[i], im[i]}; }
        auto y = sig;
        plan(vec<Complex>{y.data(), size}, direction::forward);
        plan(vec<Complex>{y.data(), size}, direction::backward);
        for (auto& v : y) { v /= Float(size); }
        REQUIRE(rel_l2(y, sig) <= 10 * tol);
    }
}

// compressed_fdl_test.cpp:14-76 against the facade: the reference's known-answer row, int8 and int16 parts
template<typename Int>
static auto test_compressed_fdl() -> void
{
    using FloatComplex = std::complex<float>;
    struct IntComplex
    {
        using value_type = Int;
    };
    double const tolerance = sizeof(Int) == 1 ? 0.005 : 0.0001;
    auto input = std::vector<FloatComplex>{{+0.000F, +0.125F}, {+0.250F, +0.333F}, {+0.500F, +0.666F}, {+0.750F, +1.000F},
                                           {-0.000F, -0.125F}, {-0.250F, -0.333F}, {-0.500F, -0.666F}, {-0.750F, -1.000F}};
    auto fdl   = neo::b200::compressed_fdl<FloatComplex, IntComplex>{4, 8};
    fdl.insert(vec<FloatComplex const>{input.data(), input.size()}, 0);
    auto const compressed = fdl[0];
    REQUIRE(compressed.size() == 8);
    auto ints = std::vector<std::int16_t>(16);
    oracle_compress_row_f32(reinterpret_cast<float const*>(input.data()), 8, int(8 * sizeof(Int)), ints.data());
    auto const raw = fdl.raw(0);
    for (std::size_t i = 0; i < 8; ++i) {
        REQUIRE(std::abs(double(compressed[i].real()) - double(input[i].real())) <= tolerance);
        REQUIRE(std::abs(double(compressed[i].imag()) - double(input[i].imag())) <= tolerance);
        REQUIRE(raw[2 * i] == Int(ints[2 * i]));          // the reference's integers, bit for bit
        REQUIRE(raw[2 * i + 1] == Int(ints[2 * i + 1]));
    }
    // neo::add(compressed, output, output) with output = (1, 2) (:66-74)
    REQUIRE(std::abs(double(compressed[1].real()) + 1.0 - 1.250) <= tolerance);
    REQUIRE(std::abs(double(compressed[1].imag()) + 2.0 - 2.333) <= tolerance);
    for (auto v : fdl[3]) { REQUIRE(v == FloatComplex{}); }
}

// sparse_upols_convolver (sparse_convolver.hpp:14-17): filter(H, sparsity) == the dense convolver on H with the rejected bins zeroed
static auto test_sparse_convolver() -> void
{
    using Complex = std::complex<float>;
    std::size_t const block = 128, taps = 1000, nblocks = 12;
    auto ir    = noise<float>(taps, 41);
    auto parts = neo_b200_num_partitions(taps, block);
    auto h     = std::vector<Complex>(parts * (block + 1));
    neo::b200::uniform_partition(ir.data(), 1, taps, block, h.data());
    auto const keep = [](std::size_t row, std::size_t col, Complex v) { return std::abs(v) > 0.5F && (row + col) % 3 != 0; };
    auto masked     = h;
    for (std::size_t p = 0; p < parts; ++p) {
        for (std::size_t k = 0; k <= block; ++k) {
            if (!keep(p, k, h[p * (block + 1) + k])) { masked[p * (block + 1) + k] = Complex{}; }
        }
    }
    auto conv = neo::b200::sparse_upols_convolver<Complex>{};
    conv.filter(mat<Complex const>{h.data(), parts, block + 1}, keep);
    auto sig = noise<float>(block * nblocks, 43);
    auto got = sig, want = sig;
    for (std::size_t b = 0; b < nblocks; ++b) { conv(vec<float>{got.data() + b * block, block}); }
    auto* o = oracle_conv_create_f32(0);
    oracle_conv_filter_f32(o, reinterpret_cast<float const*>(masked.data()), parts, block + 1);
    for (std::size_t b = 0; b < nblocks; ++b) { oracle_conv_process_f32(o, want.data() + b * block, block); }
    oracle_conv_destroy_f32(o);
    REQUIRE(rel_l2(got, want) <= 1e-5);
    // the CSR containers handed to the device are the reference's (csr_matrix.hpp:64-98, restated in the oracle), and the device
    // holds fewer bytes than with the dense layout of the same filter
    {
        float const thr      = 0.3F;
        auto const threshold = [thr](std::size_t, std::size_t, Complex v) { return std::abs(v.real()) > thr || std::abs(v.imag()) > thr; };
        auto rows            = std::vector<std::uint64_t>(parts + 1);
        auto const nnz       = oracle_csr_build_f32(reinterpret_cast<float const*>(h.data()), parts, block + 1, thr, rows.data(), nullptr, nullptr);
        auto cols            = std::vector<std::uint64_t>(nnz + 1);
        auto vals            = std::vector<Complex>(nnz + 1);
        oracle_csr_build_f32(reinterpret_cast<float const*>(h.data()), parts, block + 1, thr, rows.data(), cols.data(), reinterpret_cast<float*>(vals.data()));
        cols.resize(nnz);
        auto sparse = neo::b200::sparse_upola_convolver<Complex>{};
        sparse.filter(mat<Complex const>{h.data(), parts, block + 1}, threshold);
        REQUIRE(sparse.csr_row_container() == rows);
        REQUIRE(sparse.csr_column_container() == cols);
        auto dense = neo::b200::upola_convolver<Complex>{};
        dense.filter(mat<Complex const>{h.data(), parts, block + 1});
        REQUIRE(nnz < parts * (block + 1));
        REQUIRE(sparse.device_bytes() < dense.device_bytes());
        auto y = sig, ref = sig;
        for (std::size_t b = 0; b < nblocks; ++b) { sparse(vec<float>{y.data() + b * block, block}); }
        auto* so = oracle_conv_create_f32(6);
        oracle_conv_filter_sparse_f32(so, reinterpret_cast<float const*>(h.data()), parts, block + 1, thr);
        for (std::size_t b = 0; b < nblocks; ++b) { oracle_conv_process_f32(so, ref.data() + b * block, block); }
        oracle_conv_destroy_f32(so);
        REQUIRE(rel_l2(y, ref) <= 1e-5);
    }
    // identity filter with an all-pass predicate (uniform_partitioned_convolver_test.cpp:57-61)
    auto id = std::vector<Complex>(3 * (block + 1), Complex{});
    for (std::size_t k = 0; k <= block; ++k) { id[k] = {1, 0}; }
    conv.filter(mat<Complex const>{id.data(), 3, block + 1}, [](auto, auto, auto) { return true; });
    auto out = sig;
    for (std::size_t b = 0; b < nblocks; ++b) { conv(vec<float>{out.data() + b * block, block}); }
    REQUIRE(rel_l2(out, sig) <= 1e-5);
}

// stft_plan (fft/stft.hpp:39-109, stft_test.cpp:15-40): frame/bin counts and one frame against the rfft plan
static auto test_stft_plan() -> void
{
    using Complex = std::complex<float>;
    std::size_t const len = 1024, channels = 2;
    auto sig  = noise<float>(len * channels, 21);
    auto plan = neo::b200::stft_plan<float>{256};  // frame 256, half overlap, hann
    REQUIRE(plan.bins() == 129);
    REQUIRE(plan.frames(len) == 8);  // num_sftf_frames: idiv(1024 - 256 + 128, 128) + 1 (fft/stft.hpp:21-25)
    auto out = std::vector<Complex>(channels * plan.frames(len) * plan.bins());
    plan(sig.data(), channels, len, out.data());
    // frame 3 of channel 1 by hand: window, then the r2c plan
    auto frame = std::vector<float>(256);
    for (std::size_t i = 0; i < 256; ++i) {
        auto const w = 0.5 * (1.0 - std::cos(2.0 * 3.14159265358979323846 * double(i) / 255.0));
        frame[i]     = float(double(sig[len + 3 * 128 + i]) * w);
    }
    auto want = std::vector<Complex>(129);
    auto rp   = neo::b200::rfft_plan<float, Complex>{neo::b200::from_order, 8};
    rp(vec<float const>{frame.data(), 256}, vec<Complex>{want.data(), 129});
    auto got = std::vector<Complex>(out.begin() + (8 + 3) * 129, out.begin() + (8 + 4) * 129);
    REQUIRE(rel_l2(got, want) <= 1e-5);
}

// upola_convolver_v2 (overlap_add_convolver.hpp:21-136): calls of several whole blocks, uneven split, against the oracle's upola
static auto test_overlap_add_convolver() -> void
{
    using Complex = std::complex<float>;
    std::size_t const block = 64, taps = 700, nblocks = 40;
    auto ir    = noise<float>(taps, 5);
    auto parts = neo_b200_num_partitions(taps, block);
    auto h     = std::vector<Complex>(parts * (block + 1));
    neo::b200::uniform_partition(ir.data(), 1, taps, block, h.data());
    auto conv = neo::b200::upola_convolver_v2<Complex>{};
    conv.filter(mat<Complex const>{h.data(), parts, block + 1});
    auto sig = noise<float>(block * nblocks, 17);
    auto got = sig, want = sig;
    std::size_t pos = 0;
    for (std::size_t blocks : {1U, 3U, 35U, 1U}) {  // 35 > the 32 blocks one launch takes
        conv(vec<float>{got.data() + pos * block, blocks * block});
        pos += blocks;
    }
    auto* o = oracle_conv_create_f32(1);
    oracle_conv_filter_f32(o, reinterpret_cast<float const*>(h.data()), parts, block + 1);
    for (std::size_t b = 0; b < nblocks; ++b) { oracle_conv_process_f32(o, want.data() + b * block, block); }
    oracle_conv_destroy_f32(o);
    REQUIRE(rel_l2(got, want) <= 1e-5);
    bool threw = false;
    try {
        conv(vec<float>{got.data(), block - 1});  // overlap_add_convolver.hpp:74 asserts block_size <= extent
    } catch (std::invalid_argument const&) {
        threw = true;
    }
    REQUIRE(threw);

    // calls that END INSIDE a block (overlap_add_convolver.hpp:85-132), against the oracle's restatement of the same object (kind 4):
    // whole-block calls first (device-side tail), then ragged lengths, then whole blocks again while a block is half filled
    for (std::size_t blk : {std::size_t(64), std::size_t(256)}) {
        auto const p2 = neo_b200_num_partitions(taps, blk);
        auto h2       = std::vector<Complex>(p2 * (blk + 1));
        neo::b200::uniform_partition(ir.data(), 1, taps, blk, h2.data());
        auto v2 = neo::b200::upola_convolver_v2<Complex>{};
        v2.filter(mat<Complex const>{h2.data(), p2, blk + 1});
        auto* o2 = oracle_conv_create_f32(4);
        oracle_conv_filter_f32(o2, reinterpret_cast<float const*>(h2.data()), p2, blk + 1);
        std::size_t const lengths[] = {2 * blk, blk, blk + 7, 3 * blk - 7, blk + blk / 2, 2 * blk, blk + blk / 2, blk + 1, 4 * blk - 1, blk};
        std::size_t total = 0;
        for (auto len : lengths) { total += len; }
        auto sig2 = noise<float>(total, 23);
        auto g2 = sig2, w2 = sig2;
        std::size_t at = 0;
        for (auto len : lengths) {
            v2(vec<float>{g2.data() + at, len});
            oracle_conv_process_f32(o2, w2.data() + at, len);
            at += len;
        }
        oracle_conv_destroy_f32(o2);
        REQUIRE(rel_l2(g2, w2) <= 1e-5);
    }
}

// multi_gpu_bank: the facade over neo_b200_bank_* with every rank on device 0 (or spread over the box's devices), against the oracle
static auto test_multi_gpu_bank() -> void
{
    std::size_t const channels = 8, block = 64, taps = 64 * 16 - 5, T = 4, steps = 5;
    auto const parts = neo_b200_num_partitions(taps, block);
    auto ir  = noise<float>(channels * taps, 31);
    auto sig = noise<float>(channels * T * block * steps, 37);
    auto want = sig;
    auto h    = std::vector<std::complex<float>>(channels * parts * (block + 1));
    neo::b200::uniform_partition(ir.data(), channels, taps, block, h.data());
    for (std::size_t c = 0; c < channels; ++c) {
        auto* o = oracle_conv_create_f32(0);
        oracle_conv_filter_f32(o, reinterpret_cast<float const*>(h.data() + c * parts * (block + 1)), parts, block + 1);
        // the bank's arrays are [channels][T*block] per call: channel c's samples of call s sit at (s*channels + c)*T*block
        for (std::size_t s = 0; s < steps; ++s) {
            for (std::size_t b = 0; b < T; ++b) { oracle_conv_process_f32(o, want.data() + (s * channels + c) * T * block + b * block, block); }
        }
        oracle_conv_destroy_f32(o);
    }
    int const ndev = neo_b200_device_count();
    for (auto layout : {neo_b200_bank_layout{2, 2}, neo_b200_bank_layout{1, 4}, neo_b200_bank_layout{4, 1}}) {
        for (std::size_t frame : {std::size_t(0), T}) {
            auto devices = std::vector<int>(4);
            for (int r = 0; r < 4; ++r) { devices[r] = r % ndev; }
            auto bank = neo::b200::multi_gpu_bank<float, NEO_B200_UPOLS>{};
            bank.create(devices, layout, NEO_B200_DIAGONAL, channels, channels, block, parts, T, frame);
            bank.impulse(ir.data(), taps);
            auto got = std::vector<float>(sig.size());
            for (std::size_t s = 0; s < steps; ++s) {
                bank.submit(sig.data() + s * channels * T * block, got.data() + s * channels * T * block, T);
                if (s >= 2) { bank.wait(); }
            }
            bank.wait();
            bank.wait();
            bank.wait();
            REQUIRE(rel_l2(got, want) <= 1e-5);
        }
    }
}

int main()
{
    if (neo_b200_device_count() < 1) {
        std::printf("facade_test: no CUDA device (no CPU fallback exists)\n");
        return 2;
    }
    test_fft_plan<float>();
    test_fft_plan<double>();
    test_rfft_plan<float>();
    test_rfft_plan<double>();
    test_stft_plan();
    test_dft_plan<float>();
    test_dft_plan<double>();
    test_convolver<float, neo::b200::upols_convolver, 0>();
    test_convolver<float, neo::b200::upola_convolver, 1>();
    test_convolver<double, neo::b200::upols_convolver, 0>();
    test_convolver<double, neo::b200::upola_convolver, 1>();
    test_convolver<float, neo::b200::split_upols_convolver, 0>();
    test_convolver<float, neo::b200::split_upola_convolver, 1>();
    test_overlap_add_convolver();
    test_sparse_convolver();
    test_compressed_fdl<std::int8_t>();
    test_compressed_fdl<std::int16_t>();
    test_multi_gpu_bank();
    std::printf(failures == 0 ? "facade_test: all passed\n" : "facade_test: %d FAILED\n", failures);
    return failures == 0 ? 0 : 1;
}
