// fft_stream.cuh -- long real transforms (float32, N = 2^13 .. 2^14) as a PERSISTENT CTA fed by TMA bulk copies.
//
// Same arithmetic as r2c_kernel / c2r_kernel (fallback_rfft_plan.hpp:28-55 through the half-size complex transform), different data
// movement. A 4096- or 8192-point CTA needs a 35-70 KB exchange tile, so only 1-3 CTAs fit an SM and their phases (load the row,
// transform, store the row) do not overlap: ncu showed 60 % issue-active with the rest spent waiting on global loads. Here one CTA
// per SM slot walks rows b = blockIdx.x, blockIdx.x + gridDim.x, ... and
//   - row i+1 arrives by ONE cp.async.bulk (global -> shared, mbarrier completion) while row i is transformed,
//   - row i-1 leaves by ONE cp.async.bulk (shared -> global, bulk group) while row i is transformed,
// so per-thread global loads/stores and their address arithmetic leave the instruction stream (they become unit-stride LDS/STS) and
// HBM stays busy during the butterflies. Rows of M+1 complex (the reference's spectrum layout) start on odd 8-byte boundaries for odd
// b: the bulk copy then covers the 16-byte aligned M elements and the one remaining bin moves with a plain access.
#pragma once

#include "conv_kernels.cuh"  // tma:: helpers

namespace neo_b200 {

namespace tma {

// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, void const* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups have finished READING shared memory
template<int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template<int N>
__device__ __forceinline__ void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (the bulk store that follows)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma

template<int LOGM>
struct stream_cfg
{
    using F                        = cta_fft<float, LOGM, -1>;
    static constexpr int E         = F::E;
    static constexpr int TN        = F::TN;
    static constexpr int M         = F::M;
    static constexpr int THREADS   = TN;
    static constexpr size_t TILE   = (size_t(F::TILE) * sizeof(float2) + 15) / 16 * 16;
    static constexpr size_t ROW    = size_t(M) * sizeof(float2);            // 2M reals = M complex
    static constexpr size_t ROW2   = size_t(M + 2) * sizeof(float2);        // spectrum row with room for the odd-row shift
    static constexpr size_t SMEM   = TILE + ROW + ROW2 + 16;
    static constexpr int CTAS      = SMEM <= 113 * 1024 ? 2 : 1;            // resident CTAs per SM
};

// in: [batch][2M] reals, out: [batch][M+1] complex
template<int LOGM>
__global__ void __launch_bounds__(stream_cfg<LOGM>::THREADS, stream_cfg<LOGM>::CTAS)
    r2c_stream_kernel(float const* __restrict__ in, float2* __restrict__ out, float2 const* __restrict__ tw, float2 const* __restrict__ rtw,
                      size_t batch)
{
    using cfg = stream_cfg<LOGM>;
    using F   = cta_fft<float, LOGM, -1>;
    using C   = float2;
    constexpr int M = cfg::M, E = cfg::E, TN = cfg::TN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* const tile      = reinterpret_cast<C*>(smem_raw);
    C* const stage_in  = reinterpret_cast<C*>(smem_raw + cfg::TILE);
    C* const stage_out = reinterpret_cast<C*>(smem_raw + cfg::TILE + cfg::ROW);
    void* const bar    = smem_raw + cfg::TILE + cfg::ROW + cfg::ROW2;

    int const t = threadIdx.x;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    size_t b = blockIdx.x;
    if (t == 0 && b < batch) {
        tma::mbar_expect_tx(bar, unsigned(cfg::ROW));
        tma::bulk_g2s(stage_in, in + b * (2 * size_t(M)), unsigned(cfg::ROW), bar);
    }
    unsigned parity = 0;
    for (; b < batch; b += gridDim.x) {
        tma::mbar_wait(bar, parity);
        parity ^= 1U;
        C v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { v[e] = stage_in[t + e * TN]; }
        __syncthreads();  // the input stage is free again
        if (t == 0 && b + gridDim.x < batch) {
            tma::mbar_expect_tx(bar, unsigned(cfg::ROW));
            tma::bulk_g2s(stage_in, in + (b + gridDim.x) * (2 * size_t(M)), unsigned(cfg::ROW), bar);
        }

        F::run(v, tile, tw, t);

#pragma unroll
        for (int e = 0; e < E; ++e) { tile[t + e * TN] = v[e]; }  // Hermitian exchange, unit stride both ways: no padding
        if (t == 0) { tma::bulk_wait_read<0>(); }                  // the previous row's bulk store has read the output stage
        __syncthreads();

        int const shift = int((b * (size_t(M) + 1)) & 1U);  // odd rows start 8 bytes off a 16-byte boundary
        C* const row    = out + b * (size_t(M) + 1);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int const k = t + e * TN;
            if (k == 0) {
                C const dc = mk<float>(v[e].x + v[e].y, 0.f), ny = mk<float>(v[e].x - v[e].y, 0.f);
                if (shift == 0) {
                    stage_out[0] = dc;
                    row[M]       = ny;
                } else {
                    row[0]               = dc;
                    stage_out[M + shift] = ny;
                }
            } else {
                stage_out[k + shift] = r2c_post(v[e], tile[M - k], __ldg(rtw + k));
            }
        }
        tma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            // shift 0: bins [0, M) from stage[0..M); shift 1: bins [1, M] from stage[2..M+2)
            tma::bulk_s2g(row + shift, stage_out + 2 * shift, unsigned(cfg::ROW));
            tma::bulk_commit();
        }
    }
    if (t == 0) { tma::bulk_wait<0>(); }
}

// in: [batch][row_len] complex (first M+1 used), out: [batch][2M] reals, unnormalised
template<int LOGM>
__global__ void __launch_bounds__(stream_cfg<LOGM>::THREADS, stream_cfg<LOGM>::CTAS)
    c2r_stream_kernel(float2 const* __restrict__ in, size_t row_len, float* __restrict__ out, float2 const* __restrict__ tw,
                      float2 const* __restrict__ rtw, size_t batch)
{
    using cfg = stream_cfg<LOGM>;
    using F   = cta_fft<float, LOGM, +1>;
    using C   = float2;
    constexpr int M = cfg::M, E = cfg::E, TN = cfg::TN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* const tile      = reinterpret_cast<C*>(smem_raw);
    C* const stage_out = reinterpret_cast<C*>(smem_raw + cfg::TILE);              // M complex = 2M reals
    C* const stage_in  = reinterpret_cast<C*>(smem_raw + cfg::TILE + cfg::ROW);   // M + 2 slots: bin k sits at k + shift
    void* const bar    = smem_raw + cfg::TILE + cfg::ROW + cfg::ROW2;

    int const t = threadIdx.x;
    if (t == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    // bulk part of row b: the M bins starting at the first 16-byte aligned one; the remaining bin is fetched by thread 0
    auto const issue = [&](size_t row_index) {
        C const* const row = in + row_index * row_len;
        int const shift    = int((row_index * row_len) & 1U);
        tma::mbar_expect_tx(bar, unsigned(cfg::ROW));
        tma::bulk_g2s(stage_in + 2 * shift, row + shift, unsigned(cfg::ROW), bar);
    };
    size_t b = blockIdx.x;
    if (t == 0 && b < batch) { issue(b); }
    unsigned parity = 0;
    for (; b < batch; b += gridDim.x) {
        int const shift    = int((b * row_len) & 1U);
        C const* const row = in + b * row_len;
        if (t == 0) {  // the bin the bulk copy leaves out: k = M (shift 0) or k = 0 (shift 1); its slot is outside the bulk range
            if (shift == 0) { stage_in[M] = row[M]; }
            else { stage_in[1] = row[0]; }
        }
        tma::mbar_wait(bar, parity);
        parity ^= 1U;
        __syncthreads();  // thread 0's plain store is visible
        C v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int const k = t + e * TN;
            C const own = stage_in[k + shift];
            if (k == 0) {
                float const a0 = own.x, am = stage_in[M + shift].x;
                v[e] = mk<float>(a0 + am, a0 - am);
            } else {
                v[e] = c2r_pre(own, stage_in[M - k + shift], __ldg(rtw + k));
            }
        }
        __syncthreads();  // the input stage is free again
        if (t == 0 && b + gridDim.x < batch) { issue(b + gridDim.x); }

        F::run(v, tile, tw, t);

        if (t == 0) { tma::bulk_wait_read<0>(); }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) { stage_out[t + e * TN] = v[e]; }
        tma::fence_async_smem();
        __syncthreads();
        if (t == 0) {
            tma::bulk_s2g(out + b * (2 * size_t(M)), stage_out, unsigned(cfg::ROW));
            tma::bulk_commit();
        }
    }
    if (t == 0) { tma::bulk_wait<0>(); }
}

inline int stream_grid(size_t batch, int ctas_per_sm)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t const slots = size_t(sms) * size_t(ctas_per_sm);
    return int(std::min(batch, slots));
}

template<int LOGM>
int launch_r2c_stream(float const* in, float2* out, float2 const* tw, float2 const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = stream_cfg<LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = r2c_stream_kernel<LOGM>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    kernel<<<unsigned(stream_grid(batch, cfg::CTAS)), cfg::THREADS, cfg::SMEM, stream>>>(in, out, tw, rtw, batch);
    return check_launch("r2c_stream_kernel");
}

template<int LOGM>
int launch_c2r_stream(float2 const* in, size_t row_len, float* out, float2 const* tw, float2 const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = stream_cfg<LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = c2r_stream_kernel<LOGM>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    kernel<<<unsigned(stream_grid(batch, cfg::CTAS)), cfg::THREADS, cfg::SMEM, stream>>>(in, row_len, out, tw, rtw, batch);
    return check_launch("c2r_stream_kernel");
}

}  // namespace neo_b200
