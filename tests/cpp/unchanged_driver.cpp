// unchanged_driver.cpp -- the reference's own driver loops, compiled against the B200 facade with the convolver type as the only change.
//
//   convolve<Convolver>()   = extra/cli/src/convolver.cpp:25-59: for every channel build one Convolver, filter(partitions of that
//                             channel), then one call per block through a block buffer in ordinary (pageable) host memory;
//   bench<Convolver>()      = extra/benchmark/src/convolution.cpp:12-45: one convolver, copy noise into the block, call, repeat.
//
// Usage: unchanged_driver <channels> <blocks> <block_size> <taps>  -> one JSON line on stdout.
// This is the call shape a maintainer gets by flipping the alias and touching nothing else: one handle and one
// CUDA stream per channel, three small kernels and two B-sample copies per block. bench.py reports it as `e2e_modes.unchanged_driver`
// next to the batched calls, so the cost of NOT batching is on the record.
#include "../../include/neo_b200.hpp"

#include <cuda/std/mdspan>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

namespace stdex = cuda::std;
template<typename T>
using vec = stdex::mdspan<T, stdex::dextents<std::size_t, 1>>;
template<typename T>
using mat = stdex::mdspan<T, stdex::dextents<std::size_t, 2>>;
template<typename T>
using cube = stdex::mdspan<T, stdex::dextents<std::size_t, 3>>;

static auto noise(std::size_t n, unsigned seed) -> std::vector<float>
{
    auto rng  = std::mt19937{seed};
    auto dist = std::uniform_real_distribution<float>{-1.0F, 1.0F};
    auto v    = std::vector<float>(n);
    for (auto& x : v) { x = dist(rng); }
    return v;
}

// extra/cli/src/convolver.cpp:25-59 with `Convolver` swapped in; returns seconds of the channel/block loop (filter() included, as there)
template<typename Convolver>
static auto convolve(mat<float const> signal, cube<std::complex<float> const> partitions, mat<float> output, std::size_t block_size) -> double
{
    auto block_buffer = std::vector<float>(block_size);
    auto const start  = std::chrono::steady_clock::now();
    for (std::size_t channel = 0; channel < signal.extent(0); ++channel) {
        auto convolver = Convolver{};
        convolver.filter(mat<std::complex<float> const>{&partitions(channel, 0, 0), partitions.extent(1), partitions.extent(2)});
        for (std::size_t i = 0; i < output.extent(1); i += block_size) {
            std::fill(block_buffer.begin(), block_buffer.end(), 0.0F);
            auto const num_samples = std::min(output.extent(1) - i, block_size);
            std::copy_n(&signal(channel, i), num_samples, block_buffer.begin());
            convolver(vec<float>{block_buffer.data(), block_size});
            std::copy_n(block_buffer.begin(), num_samples, &output(channel, i));
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
}

// extra/benchmark/src/convolution.cpp:28-40: steady-state block loop of ONE convolver; returns seconds per block
template<typename Convolver>
static auto bench(mat<std::complex<float> const> filter, std::size_t block_size, std::size_t iterations) -> double
{
    auto convolver = Convolver{};
    convolver.filter(filter);
    auto const src = noise(block_size, 13);
    auto block     = src;
    for (std::size_t i = 0; i < 20; ++i) { convolver(vec<float>{block.data(), block_size}); }
    auto const start = std::chrono::steady_clock::now();
    for (std::size_t i = 0; i < iterations; ++i) {
        std::copy(src.begin(), src.end(), block.begin());
        convolver(vec<float>{block.data(), block_size});
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() / double(iterations);
}

auto main(int argc, char** argv) -> int
{
    if (argc != 5) {
        std::fprintf(stderr, "usage: %s <channels> <blocks> <block_size> <taps>\n", argv[0]);
        return 2;
    }
    auto const channels = std::size_t(std::atoll(argv[1]));
    auto const blocks   = std::size_t(std::atoll(argv[2]));
    auto const block    = std::size_t(std::atoll(argv[3]));
    auto const taps     = std::size_t(std::atoll(argv[4]));
    try {
        // impulse responses -> partitions [channels][P][B+1], as the CLI does with normalize_impulse + uniform_partition
        auto impulse = noise(channels * taps, 11);
        neo::b200::detail::check(neo_b200_normalize_impulse(impulse.data(), channels, taps, NEO_B200_F32, NEO_B200_HOST));
        auto const parts = neo_b200_num_partitions(taps, block);
        auto partitions  = std::vector<std::complex<float>>(channels * parts * (block + 1));
        neo::b200::uniform_partition(impulse.data(), channels, taps, block, partitions.data());
        auto const signal = noise(channels * blocks * block, 13);
        auto output       = std::vector<float>(signal.size());

        using Convolver    = neo::b200::upols_convolver<std::complex<float>>;
        auto const seconds = convolve<Convolver>(mat<float const>{signal.data(), channels, blocks * block},
                                                 cube<std::complex<float> const>{partitions.data(), channels, parts, block + 1},
                                                 mat<float>{output.data(), channels, blocks * block}, block);
        auto const per_block = bench<Convolver>(mat<std::complex<float> const>{partitions.data(), parts, block + 1}, block, 200);
        double energy = 0;
        for (auto v : output) { energy += double(v) * double(v); }
        std::printf("{\"channel_msamples_s\": %.4f, \"unit\": \"channel-Msamples/s\", \"seconds\": %.4f, \"channels\": %zu, \"blocks\": %zu, "
                    "\"steady_state_us_per_block_call\": %.2f, \"steady_state_channel_msamples_s\": %.4f, \"output_energy\": %.6e, "
                    "\"what\": \"extra/cli/src/convolver.cpp:25-59 loop (one neo::b200::upols_convolver per channel, filter() inside the "
                    "clock as there, one block per call, pageable std::vector memory) and extra/benchmark/src/convolution.cpp:28-40 "
                    "steady-state loop of one convolver\"}\n",
                    double(channels * blocks * block) / seconds / 1e6, seconds, channels, blocks, per_block * 1e6,
                    double(block) / per_block / 1e6, energy);
    } catch (std::exception const& e) {
        std::fprintf(stderr, "unchanged_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
