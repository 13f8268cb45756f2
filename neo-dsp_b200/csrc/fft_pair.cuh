// fft_pair.cuh -- float32 real transforms of N = 2^16 points: ONE transform per PAIR of CTAs (thread-block cluster of two).
//
// The half-size complex transform (M = 2^15 points, 256 KB) does not fit one SM's shared memory, so the exchange tile of cta_fft is
// split across the two CTAs of a cluster (dsmem_tile2, fft_core.cuh): each of the 2 x 1024 threads owns 16 points in registers,
// element idx of the tile lives in CTA (idx >> 10) & 1, which makes the Stockham read pattern t + e*2048 local; only the scattered
// stores cross the cluster (st.shared::cluster) and every exchange ends in a cluster barrier. The transform then needs exactly one
// pass over HBM (read N reals, write N/2+1 bins) and no global scratch, where the four-step forms (fft_cluster.cuh, fft_large.cuh)
// go through an L2 scratch with two or three dependent phases.
// Reference semantics: fallback_rfft_plan.hpp:28-55 (unnormalised both ways).
#pragma once

#include "fft_kernels.cuh"

namespace neo_b200 {

template<int LOGM>
struct pair_cfg
{
    static constexpr int LOGC    = 10;  // 1024 threads per CTA
    static constexpr int THREADS = 1 << LOGC;
    static constexpr int M       = 1 << LOGM;
    static constexpr int E       = 16;
    static constexpr int TN      = M / E;  // threads per transform = 2 CTAs
    static_assert(TN == 2 * THREADS, "one transform per CTA pair");
    static constexpr size_t SMEM = size_t(padded<float>(M / 2) + 1) * sizeof(float2);
};

__device__ __forceinline__ unsigned pair_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ dsmem_tile2<float, 10> pair_tile(unsigned char* smem_raw, unsigned rank)
{
    unsigned const mine = static_cast<unsigned>(__cvta_generic_to_shared(smem_raw));
    unsigned peer;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(mine), "r"(rank ^ 1U));
    return {reinterpret_cast<float2*>(smem_raw), peer, rank};
}

template<int LOGM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair_cfg<LOGM>::THREADS, 1)
    r2c_pair_kernel(float const* __restrict__ in, float2* __restrict__ out, float2 const* __restrict__ tw, float2 const* __restrict__ rtw)
{
    using cfg = pair_cfg<LOGM>;
    using F   = cta_fft<float, LOGM, -1, 4>;
    static_assert(F::TN == cfg::TN && F::E == cfg::E, "geometry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned const rank = pair_rank();
    auto const sm       = pair_tile(smem_raw, rank);
    size_t const b      = blockIdx.x >> 1;
    int const t         = int(rank) * cfg::THREADS + int(threadIdx.x);

    float2 const* const src = reinterpret_cast<float2 const*>(in) + b * size_t(cfg::M);
    float2* const dst       = out + b * (size_t(cfg::M) + 1);
    float2 v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { v[e] = src[t + e * cfg::TN]; }
    tile_sync(sm);  // both CTAs are running before either touches the other's shared memory

    F::run(v, sm, tw, t);

#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { tile_store(sm, t + e * cfg::TN, v[e]); }  // own elements: local stores
    tile_sync(sm);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const k = t + e * cfg::TN;
        if (k == 0) {
            dst[0]      = mk<float>(v[e].x + v[e].y, 0.f);
            dst[cfg::M] = mk<float>(v[e].x - v[e].y, 0.f);
        } else {
            dst[k] = r2c_post(v[e], tile_load(sm, cfg::M - k), __ldg(rtw + k));
        }
    }
    tile_sync(sm);  // no CTA leaves while its partner may still read its half
}

template<int LOGM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair_cfg<LOGM>::THREADS, 1)
    c2r_pair_kernel(float2 const* __restrict__ in, size_t row_len, float* __restrict__ out, float2 const* __restrict__ tw,
                    float2 const* __restrict__ rtw)
{
    using cfg = pair_cfg<LOGM>;
    using F   = cta_fft<float, LOGM, +1, 4>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned const rank = pair_rank();
    auto const sm       = pair_tile(smem_raw, rank);
    size_t const b      = blockIdx.x >> 1;
    int const t         = int(rank) * cfg::THREADS + int(threadIdx.x);

    float2 const* const src = in + b * row_len;
    float2* const dst       = reinterpret_cast<float2*>(out) + b * size_t(cfg::M);
    float2 v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const k = t + e * cfg::TN;
        v[e]        = k == 0 ? mk<float>(src[0].x, src[cfg::M].x) : src[k];
    }
    tile_sync(sm);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { tile_store(sm, t + e * cfg::TN, v[e]); }
    tile_sync(sm);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const k = t + e * cfg::TN;
        if (k == 0) { v[e] = mk<float>(v[e].x + v[e].y, v[e].x - v[e].y); }
        else { v[e] = c2r_pre(v[e], tile_load(sm, cfg::M - k), __ldg(rtw + k)); }
    }
    tile_sync(sm);

    F::run(v, sm, tw, t);

#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { dst[t + e * cfg::TN] = v[e]; }
    tile_sync(sm);
}

template<int LOGM>
int launch_r2c_pair(float const* in, float2* out, float2 const* tw, float2 const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = pair_cfg<LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    if (batch > 0x3fffffffULL) { return fail(NEO_B200_ERR_UNSUPPORTED, "batch too large"); }
    auto kernel = r2c_pair_kernel<LOGM>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    kernel<<<unsigned(2 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, out, tw, rtw);
    return check_launch("r2c_pair_kernel");
}

template<int LOGM>
int launch_c2r_pair(float2 const* in, size_t row_len, float* out, float2 const* tw, float2 const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = pair_cfg<LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    if (batch > 0x3fffffffULL) { return fail(NEO_B200_ERR_UNSUPPORTED, "batch too large"); }
    auto kernel = c2r_pair_kernel<LOGM>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    kernel<<<unsigned(2 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, row_len, out, tw, rtw);
    return check_launch("c2r_pair_kernel");
}

}  // namespace neo_b200
