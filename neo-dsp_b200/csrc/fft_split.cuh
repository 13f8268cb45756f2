// fft_split.cuh -- real transforms of 2M = 4*M2 points where the half-size complex transform (M = 2*M2 points) is one
// decimation-in-frequency step too long for a single CTA (or too slow in one: the 8192-point CTA is issue-bound).
//
// Two CTAs per transform, no communication between them: CTA r (r = 0, 1) loads BOTH halves of the input (the second read of
// every line comes from L2), forms y_r[n] = (z[n] + (-1)^r z[n + M2]) * W_M^(r n) on the fly and runs the M2-point cta_fft,
// which yields the bins of parity r: Z[r + 2 k2]. The Hermitian split pairs Z[k] with Z[M - k]; M is even, so the partner has
// the same parity and lives in the same CTA at local index M2 - r - k2. HBM traffic stays one read + one write per datum;
// the price is 8-byte stores at a 16-byte stride (the sibling CTA fills the gaps in L2).
#pragma once

#include "fft_kernels.cuh"

namespace neo_b200 {

template<typename T, int LOGM2>
struct split_cfg
{
    using base                   = fft_cfg<T, LOGM2>;
    static constexpr int E       = base::E;
    static constexpr int TN      = base::TN;
    static constexpr int M2      = base::M;
    static constexpr int THREADS = TN;  // one half-transform per CTA (M2 >= 2^11)
    static constexpr size_t SMEM = size_t(base::TILE) * sizeof(cx<T>);
};

// in: [batch][4*M2] reals, out: [batch][2*M2 + 1] complex. tw: stage twiddles of the M2-point FFT; w_m[k] = exp(-2 pi i k / M)
// (= the M2-plan's split table); w_n1 = exp(-2 pi i / 2M) (one step of the real transform's twiddle)
template<typename T, int LOGM2>
__global__ void __launch_bounds__(split_cfg<T, LOGM2>::THREADS, fft_min_ctas<T, LOGM2, k_r2c>())
    r2c_split2_kernel(T const* __restrict__ in, cx<T>* __restrict__ out, cx<T> const* __restrict__ tw, cx<T> const* __restrict__ w_m,
                      cx<T> w_n1)
{
    using cfg = split_cfg<T, LOGM2>;
    using F   = cta_fft<T, LOGM2, -1>;
    using C   = cx<T>;
    constexpr int M2 = cfg::M2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);

    int const t      = threadIdx.x;
    int const r      = blockIdx.x & 1;
    size_t const b   = blockIdx.x >> 1;
    C const* const z = reinterpret_cast<C const*>(in) + b * (2 * size_t(M2));

    C v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const n = t + e * cfg::TN;
        C const a = z[n], c = z[n + M2];
        v[e] = r == 0 ? cadd(a, c) : cmul(csub(a, c), __ldg(w_m + n));
    }
    F::run(v, sm, tw, t);

#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { sm[t + e * cfg::TN] = v[e]; }  // unpadded: unit stride both ways
    __syncthreads();
    C* const row = out + b * (2 * size_t(M2) + 1);
    // pairs, as in r2c_kernel: bin k = r + 2 k2 (k2 < M2/2: the first half of the thread's points) and its partner M - k = r + 2 p2,
    // p2 = M2 - r - k2, come from one twiddle product
    constexpr int M = 2 * M2;
#pragma unroll
    for (int e = 0; e < cfg::E / 2; ++e) {
        int const k2 = t + e * cfg::TN;
        int const k  = r + 2 * k2;
        if (k == 0) {
            row[0] = mk<T>(v[e].x + v[e].y, T(0));
            row[M] = mk<T>(v[e].x - v[e].y, T(0));
        } else {
            C w = __ldg(w_m + k2);  // W_2M^(2 k2)
            if (r != 0) { w = cmul(w, w_n1); }
            C xk, xmk;
            r2c_post_pair(v[e], sm[M2 - r - k2], w, xk, xmk);
            row[k]     = xk;
            row[M - k] = xmk;
        }
    }
    if (r == 0 && t == 0) { row[M2] = cconj(v[cfg::E / 2]); }  // the self-paired bin k = M/2 (local index M2/2 of the even bins)
}

// in: [batch][row_len] complex (first 2*M2+1 used), out: [batch][4*M2] reals, unnormalised.
// w_n[k] = exp(-2 pi i k / 2M) for k < M2 (first half of the split table one size up)
template<typename T, int LOGM2>
__global__ void __launch_bounds__(split_cfg<T, LOGM2>::THREADS, fft_min_ctas<T, LOGM2, k_c2r>())
    c2r_split2_kernel(cx<T> const* __restrict__ in, size_t row_len, T* __restrict__ out, cx<T> const* __restrict__ tw,
                      cx<T> const* __restrict__ w_m, cx<T> const* __restrict__ w_n)
{
    using cfg = split_cfg<T, LOGM2>;
    using F   = cta_fft<T, LOGM2, +1>;
    using C   = cx<T>;
    constexpr int M2 = cfg::M2;
    constexpr int M  = 2 * M2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);

    int const t      = threadIdx.x;
    int const r      = blockIdx.x & 1;
    size_t const b   = blockIdx.x >> 1;
    C const* const x = in + b * row_len;

    C v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const k = t + e * cfg::TN;  // 0 <= k < M2
        C const wn  = __ldg(w_n + k);   // W_2M^k ; W_2M^(k+M2) = -i * W_2M^k
        C z0, z1;
        if (k == 0) {
            T const a0 = x[0].x, am = x[M].x;
            z0 = mk<T>(a0 + am, a0 - am);
        } else {
            z0 = c2r_pre(x[k], x[M - k], wn);
        }
        z1 = c2r_pre(x[k + M2], x[M2 - k], rot90<-1>(wn));
        // backward DIF step: y_r[k] = (Z[k] + (-1)^r Z[k + M2]) * conj(W_M^k)^r
        v[e] = r == 0 ? cadd(z0, z1) : cmulc(csub(z0, z1), __ldg(w_m + k));
    }
    F::run(v, sm, tw, t);
    C* const dst = reinterpret_cast<C*>(out) + b * size_t(M);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { dst[r + 2 * (t + e * cfg::TN)] = v[e]; }
}

// c2c of M = M1*M2 points (M1 = 2 or 4) as M1 CTAs per transform: CTA r forms y_r[n] = W_M^(r n) * sum_j x[j*M2 + n] W_M1^(j r)
// while loading (every CTA reads the whole input, M1-1 of those reads come from L2) and its M2-point FFT gives X[r + M1*k2].
// w_big[n] = exp(-2 pi i n / M), n < M2.
template<typename T, int LOGM2, int M1, int DIR>
__global__ void __launch_bounds__(split_cfg<T, LOGM2>::THREADS, fft_min_ctas<T, LOGM2, k_c2c>())
    c2c_split_kernel(cx<T> const* __restrict__ in, cx<T>* __restrict__ out, cx<T> const* __restrict__ tw, cx<T> const* __restrict__ w_big)
{
    using cfg = split_cfg<T, LOGM2>;
    using F   = cta_fft<T, LOGM2, DIR>;
    using C   = cx<T>;
    constexpr int M2 = cfg::M2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);

    int const t      = threadIdx.x;
    int const r      = blockIdx.x % M1;
    size_t const b   = blockIdx.x / M1;
    C const* const x = in + b * (size_t(M1) * M2);

    C v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) {
        int const n = t + e * cfg::TN;
        C y;
        if constexpr (M1 == 2) {
            C const a = x[n], c = x[n + M2];
            y = r == 0 ? cadd(a, c) : csub(a, c);
        } else {
            C const a = x[n], bb = x[n + M2], c = x[n + 2 * M2], d = x[n + 3 * M2];
            C const s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(bb, d), s3 = rot90<DIR>(csub(bb, d));
            y = r == 0 ? cadd(s0, s2) : r == 1 ? cadd(s1, s3) : r == 2 ? csub(s0, s2) : csub(s1, s3);
        }
        if (r != 0) {
            C const w1 = __ldg(w_big + n);
            C w        = w1;
            if (r >= 2) { w = cmul(w1, w1); }
            if (r == 3) { w = cmul(w, w1); }
            y = DIR < 0 ? cmul(y, w) : cmulc(y, w);
        }
        v[e] = y;
    }
    F::run(v, sm, tw, t);
    C* const dst = out + b * (size_t(M1) * M2);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { dst[r + M1 * (t + e * cfg::TN)] = v[e]; }
}

template<typename T, int LOGM2, int M1>
int launch_c2c_split(cx<T> const* in, cx<T>* out, cx<T> const* tw, cx<T> const* w_big, size_t batch, int direction, cudaStream_t stream)
{
    using cfg = split_cfg<T, LOGM2>;
    if (batch == 0) { return NEO_B200_OK; }
    if (direction < 0) {
        auto kernel = c2c_split_kernel<T, LOGM2, M1, -1>;
        NEO_TRY(enable_smem(kernel, cfg::SMEM));
        kernel<<<static_cast<unsigned>(M1 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, out, tw, w_big);
    } else {
        auto kernel = c2c_split_kernel<T, LOGM2, M1, +1>;
        NEO_TRY(enable_smem(kernel, cfg::SMEM));
        kernel<<<static_cast<unsigned>(M1 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, out, tw, w_big);
    }
    return check_launch("c2c_split_kernel");
}

template<typename T, int LOGM2>
int launch_r2c_split2(T const* in, cx<T>* out, cx<T> const* tw, cx<T> const* w_m, size_t batch, cudaStream_t stream)
{
    using cfg = split_cfg<T, LOGM2>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = r2c_split2_kernel<T, LOGM2>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    double const a = -3.14159265358979323846264338327950288 / double(size_t(2) << LOGM2);  // -2 pi / (2M), M = 2*M2
    cx<T> const w1 = mk<T>(T(std::cos(a)), T(std::sin(a)));
    kernel<<<static_cast<unsigned>(2 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, out, tw, w_m, w1);
    return check_launch("r2c_split2_kernel");
}

template<typename T, int LOGM2>
int launch_c2r_split2(cx<T> const* in, size_t row_len, T* out, cx<T> const* tw, cx<T> const* w_m, cx<T> const* w_n, size_t batch,
                      cudaStream_t stream)
{
    using cfg = split_cfg<T, LOGM2>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = c2r_split2_kernel<T, LOGM2>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    kernel<<<static_cast<unsigned>(2 * batch), cfg::THREADS, cfg::SMEM, stream>>>(in, row_len, out, tw, w_m, w_n);
    return check_launch("c2r_split2_kernel");
}

}  // namespace neo_b200
