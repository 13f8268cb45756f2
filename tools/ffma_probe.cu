// ffma_probe.cu -- measures FP32 FMA issue rate on this GPU for the operand patterns the Toeplitz MAC can use:
//   (a) FFMA  acc[i] += x[i] * h        (h shared across 16 accumulators: operand-reuse friendly)
//   (b) FFMA  acc[i] += x[i] * y[i]     (three distinct registers every time)
//   (c) FFMA2 fma.rn.f32x2 acc2[i] += xx2[i] * h2  (two FMAs per instruction)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma_probe ffma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void k_shared(float* out, float h0)
{
    float acc[16], x[16];
    for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 1e-3f + i; x[i] = 1.0f + i * 1e-3f + (threadIdx.x + blockIdx.x) * 1e-6f; }
    float h = h0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[i] = fmaf(x[i], h, acc[i]); }
            h += 1e-7f;
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_distinct(float* out, float h0)
{
    float acc[16], x[16], y[16];
    for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 1e-3f + i; x[i] = 1.0f + i * 1e-3f; y[i] = h0 + i * 1e-4f + (threadIdx.x + blockIdx.x) * 1e-6f; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { acc[i] = fmaf(x[i], y[(i + r) & 15], acc[i]); }
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void fma2(float2& d, float2 a, float2 b)
{
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rd = *reinterpret_cast<unsigned long long*>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rd) : "l"(ra), "l"(rb));
    d = *reinterpret_cast<float2*>(&rd);
}

__global__ void k_ffma2(float* out, float h0)
{
    float2 acc[16], x[16];
    for (int i = 0; i < 16; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); x[i] = make_float2(1.0f + i * 1e-3f, 1.0f + (threadIdx.x + blockIdx.x) * 1e-6f); }
    float2 h = make_float2(h0, h0 * 0.5f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { fma2(acc[i], x[i], h); }
            h.x += 1e-7f;
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// complex MAC exactly as in the Toeplitz kernel: 16 accumulators, window x[], one h per pp
__global__ void k_cmac(float* out, float h0)
{
    float2 acc[16], win[23];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    for (int i = 0; i < 23; ++i) win[i] = make_float2(1.0f + i * 1e-3f, 0.5f + (threadIdx.x + blockIdx.x) * 1e-6f);
    float2 h = make_float2(h0, h0 * 0.5f);
    for (int it = 0; it < ITERS / 2; ++it) {
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                float2 const x = win[t - pp + 7];
                acc[t].x = fmaf(x.x, h.x, acc[t].x);
                acc[t].x = fmaf(-x.y, h.y, acc[t].x);
                acc[t].y = fmaf(x.x, h.y, acc[t].y);
                acc[t].y = fmaf(x.y, h.x, acc[t].y);
            }
            h.x += 1e-7f;
        }
    }
    float s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<typename K>
void run(char const* name, K kernel, double flops_per_thread, int threads)
{
    float* out;
    int const blocks = 148 * 8;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    kernel<<<blocks, threads>>>(out, 1.0f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) kernel<<<blocks, threads>>>(out, 1.0f);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double const tf = 5.0 * flops_per_thread * blocks * threads / (ms * 1e-3) / 1e12;
    printf("%-28s threads/CTA %4d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, threads, ms / 5, tf, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int threads : {128, 256}) {
        run("FFMA shared operand", k_shared, 2.0 * ITERS * 4 * 16, threads);
        run("FFMA 3 distinct regs", k_distinct, 2.0 * ITERS * 4 * 16, threads);
        run("FFMA2 (f32x2) shared operand", k_ffma2, 2.0 * ITERS * 2 * 16 * 2, threads);
        run("complex MAC 16x8 (kernel form)", k_cmac, 2.0 * (ITERS / 2) * 8 * 16 * 4, threads);
    }
    return 0;
}
