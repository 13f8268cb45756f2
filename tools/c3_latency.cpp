// c3_latency.cpp -- BASELINE config 3 (UPOLS stereo, B=512, 2^17-tap IR) through the C ABI from C++, the way a reference-side
// caller (extra/cli/src/convolver.cpp:42-55, one block per call) would: microseconds per block with HOST and DEVICE buffers.
// build: g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/c3_latency.cpp -L neo-dsp_b200 -lneo_b200 \
//        -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/neo-dsp_b200 -o tools/c3_latency
#include "neo_b200.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <random>
#include <vector>

int main()
{
    size_t const C = 2, B = 512, L = size_t(1) << 17, P = L / B;
    auto rng  = std::mt19937{11};
    auto dist = std::uniform_real_distribution<float>{-1.f, 1.f};
    auto ir   = std::vector<float>(C * L);
    for (auto& v : ir) { v = dist(rng) * 1e-2f; }

    for (size_t T : {size_t(1), size_t(16)}) {
        neo_b200_conv_config cfg{};
        cfg.kind = NEO_B200_UPOLS; cfg.dtype = NEO_B200_F32; cfg.topology = NEO_B200_DIAGONAL;
        cfg.outputs = cfg.inputs = C; cfg.block = B; cfg.partitions = P; cfg.max_blocks = T;
        neo_b200_conv* conv = nullptr;
        if (neo_b200_conv_create(&conv, &cfg) != 0 || neo_b200_conv_set_impulse(conv, ir.data(), L, NEO_B200_HOST) != 0) {
            std::printf("error: %s\n", neo_b200_last_error());
            return 1;
        }
        float* hbuf = nullptr;
        cudaMallocHost(&hbuf, C * T * B * sizeof(float));
        for (size_t i = 0; i < C * T * B; ++i) { hbuf[i] = dist(rng); }
        float* dbuf = nullptr;
        cudaMalloc(&dbuf, C * T * B * sizeof(float));
        cudaMemcpy(dbuf, hbuf, C * T * B * sizeof(float), cudaMemcpyHostToDevice);

        int const reps = 2000;
        for (int i = 0; i < 50; ++i) { neo_b200_conv_process(conv, hbuf, hbuf, T, NEO_B200_HOST); }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) { neo_b200_conv_process(conv, hbuf, hbuf, T, NEO_B200_HOST); }
        auto t1 = std::chrono::steady_clock::now();
        double const host_us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;

        for (int i = 0; i < 50; ++i) { neo_b200_conv_process(conv, dbuf, dbuf, T, NEO_B200_DEVICE); }
        neo_b200_conv_synchronize(conv);
        t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < reps; ++i) { neo_b200_conv_process(conv, dbuf, dbuf, T, NEO_B200_DEVICE); }
        neo_b200_conv_synchronize(conv);
        t1 = std::chrono::steady_clock::now();
        double const dev_us = std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;

        std::printf("{\"config\": \"C3 stereo B=512 L=2^17\", \"T\": %zu, \"host_call_us_per_block\": %.2f, \"device_us_per_block\": %.2f, "
                    "\"realtime_x_48k_host\": %.1f, \"realtime_x_48k_device\": %.1f}\n",
                    T, host_us / T, dev_us / T, (B * T / 48000.0) / (host_us * 1e-6), (B * T / 48000.0) / (dev_us * 1e-6));
        neo_b200_conv_destroy(conv);
        cudaFreeHost(hbuf);
        cudaFree(dbuf);
    }
    return 0;
}
