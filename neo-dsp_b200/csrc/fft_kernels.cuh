// fft_kernels.cuh -- batched single-pass FFT kernels built on cta_fft: one transform (or G small ones) per CTA,
// one HBM read and one HBM write per datum. The load/store sides are policy objects so the convolver can fuse its
// window / FDL-insert / overlap-discard steps into the same kernels.
#pragma once

#include "common.hpp"
#include "fft_core.cuh"

#include <vector>

namespace neo_b200 {

// largest single-CTA transform (complex points, log2). Above it the four-step path in fft_large.cuh takes over.
template<typename T>
constexpr int max_cta_logm()
{
    return sizeof(T) == 4 ? 13 : 12;
}

template<typename T, int LOGM>
struct fft_cfg
{
    using F                      = cta_fft<T, LOGM, -1>;
    static constexpr int E       = F::E;
    static constexpr int TN      = F::TN;
    static constexpr int M       = F::M;
    static constexpr int G       = TN >= 64 ? 1 : 64 / TN;  // transforms per CTA
    static constexpr int THREADS = TN * G;
    static constexpr int TILE    = F::TILE;
    static constexpr size_t SMEM = size_t(G) * TILE * sizeof(cx<T>);
};

// Registers per thread the single-CTA kernels are held to (-> resident CTAs per SM). These kernels are issue/latency bound at
// 4 warps per scheduler, so occupancy buys more than registers do. Measured on B200, float, fraction of HBM peak at the
// compiler's own 108-120 registers vs capped (no spills at 72 and up; 64 spills and loses 5-13 points):
//   r2c  N=1024 .84 -> .91   2048 .84 -> .92   4096 .79 -> .91 (72) / .88 (85)   8192 .68 -> .75
//   c2r  N=1024 .80 -> .85   2048 .86 -> .82 (kept at 128)   4096 .77 -> .77   8192 .62 -> .69
enum fft_kind : int
{
    k_c2c = 0,
    k_r2c = 1,
    k_c2r = 2,
};

template<typename T>
constexpr int fft_regcap(int kind, int logm)
{
    if (sizeof(T) != 4) { return 128; }
#ifdef NEO_B200_C2R_CAP
    if (kind == k_c2r) { return NEO_B200_C2R_CAP; }
#endif
    // 2^13 points = 512 threads: 64 registers (no spills, checked with -Xptxas -v) let TWO CTAs share an SM, so one CTA's global
    // loads and barriers hide behind the other's butterflies; at 85 the SM held one CTA and its phases ran back to back
    if (logm == 13) { return 64; }
    // with the pairwise Hermitian pass the 2^10 .. 2^12-point kernels fit 64 registers without spills too (-Xptxas -v): 16 / 8 / 4
    // CTAs per SM. Measured against the round-1 caps (72-128) on one box: c2r N=2048 0.93 -> 0.99, N=8192 0.77 -> 0.79, the rest equal
    if (logm >= 10 && logm <= 12) { return 64; }
    if (kind == k_c2r) { return logm == 10 ? 128 : 85; }
    return logm == 11 ? 72 : 85;
}

template<typename T, int LOGM, int KIND>
constexpr int fft_min_ctas()
{
    int const n = (65536 / fft_regcap<T>(KIND, LOGM)) / fft_cfg<T, LOGM>::THREADS;
    return n > 0 ? n : 1;
}

// ---- twiddle tables (host, computed in double, rounded once) --------------------------------------------------------
template<typename T>
std::vector<cx<T>> make_stage_twiddles(int logm, int loge_forced = -1)
{
    int const loge = loge_forced >= 0 ? loge_forced : pick_loge<T>(logm);
    std::vector<cx<T>> lut(static_cast<size_t>(fft_twiddle_count(logm, loge)) + 1);
    if (loge == 0) { return lut; }
    int const r0 = fft_first_logr(logm, loge);
    int logns    = r0 > 0 ? r0 : loge;
    size_t off   = 0;
    int const r  = 1 << loge;
    while (logns < logm) {
        long const ns    = 1L << logns;
        double const den = static_cast<double>(ns) * r;
        for (int j = 0; j < loge; ++j) {
            for (long k = 0; k < ns; ++k) {
                double const a      = -2.0 * 3.14159265358979323846264338327950288 * double(1 << j) * double(k) / den;
                lut[off + j * ns + k] = mk<T>(T(std::cos(a)), T(std::sin(a)));
            }
        }
        off += static_cast<size_t>(loge) * ns;
        logns += loge;
    }
    return lut;
}

// W[k] = exp(-2 pi i k / (2M)), k < M: the real<->complex split twiddles
template<typename T>
std::vector<cx<T>> make_split_twiddles(int logm)
{
    size_t const m = size_t(1) << logm;
    std::vector<cx<T>> lut(m);
    for (size_t k = 0; k < m; ++k) {
        double const a = -3.14159265358979323846264338327950288 * double(k) / double(m);
        lut[k]         = mk<T>(T(std::cos(a)), T(std::sin(a)));
    }
    return lut;
}

// ---- c2c: IO policy provides load(b, j) / store(b, j, v) for interleaved or split-complex (SoA) data -----------------------
template<typename T>
struct c2c_interleaved_io
{
    cx<T> const* in;
    cx<T>* out;
    size_t m;
    __device__ __forceinline__ cx<T> load(size_t b, int j) const { return in[b * m + j]; }
    __device__ __forceinline__ void store(size_t b, int j, cx<T> v) const { out[b * m + j] = v; }
};

// neo::split_complex (complex/split_complex.hpp:10): separate real / imag planes, [batch][m] each
template<typename T>
struct c2c_split_io
{
    T const* re_in;
    T const* im_in;
    T* re_out;
    T* im_out;
    size_t m;
    __device__ __forceinline__ cx<T> load(size_t b, int j) const { return mk<T>(re_in[b * m + j], im_in[b * m + j]); }
    __device__ __forceinline__ void store(size_t b, int j, cx<T> v) const
    {
        re_out[b * m + j] = v.x;
        im_out[b * m + j] = v.y;
    }
};

// ---- Bluestein (dft_plan == fallback_dft_plan, fft/fallback/fallback_dft_plan.hpp:49-82) around a power-of-two c2c of m points:
// forward leg: a[j] = x[j] * w[j] zero padded (:57-58), transform, multiply by the transformed chirp filter B (:70-72);
// backward leg: transform back, scale by 1/m (:74-75), multiply by w and keep n points (:78).
template<typename T>
struct bluestein_fwd_io
{
    cx<T> const* x;      // [batch][n]
    cx<T>* a;            // [batch][m]
    cx<T> const* w;      // [n] chirp of this direction
    cx<T> const* bhat;   // [m] forward transform of the chirp filter
    size_t n, m;
    __device__ __forceinline__ cx<T> load(size_t b, int j) const
    {
        return size_t(j) < n ? cmul(x[b * n + j], w[j]) : mk<T>(T(0), T(0));
    }
    __device__ __forceinline__ void store(size_t b, int k, cx<T> v) const { a[b * m + k] = cmul(v, bhat[k]); }
};

template<typename T>
struct bluestein_bwd_io
{
    cx<T> const* a;      // [batch][m]
    cx<T>* out;          // [batch][n]
    cx<T> const* w;
    size_t n, m;
    T scale;             // 1 / m
    __device__ __forceinline__ cx<T> load(size_t b, int k) const { return a[b * m + k]; }
    __device__ __forceinline__ void store(size_t b, int j, cx<T> v) const
    {
        if (size_t(j) < n) { out[b * n + j] = cmul(mk<T>(v.x * scale, v.y * scale), w[j]); }
    }
};

// ---- type-2 DCT through one complex transform of the same size (fallback_dct2_plan, fft/dct.hpp:40-63, Makhoul):
// loads permute the real row (even samples ascending, odd samples descending from the top), stores keep Re(v * 2 exp(-i pi k / 2n)).
template<typename T>
struct dct2_io
{
    T const* in;         // [batch][n] reals
    T* out;              // [batch][n] reals
    cx<T> const* scale;  // [n]: 2 exp(-i pi k / (2n))
    size_t n;
    __device__ __forceinline__ cx<T> load(size_t b, int j) const
    {
        size_t const i   = size_t(j);
        size_t const src = 2 * i < n ? 2 * i : 2 * (n - 1 - i) + 1;
        return mk<T>(in[b * n + src], T(0));
    }
    __device__ __forceinline__ void store(size_t b, int k, cx<T> v) const { out[b * n + k] = v.x * scale[k].x - v.y * scale[k].y; }
};

template<typename T>
__global__ void __launch_bounds__(256) dct2_pre_kernel(T const* __restrict__ x, cx<T>* __restrict__ a, size_t n, size_t total)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    size_t const b = i / n, j = i - b * n;
    size_t const src = 2 * j < n ? 2 * j : 2 * (n - 1 - j) + 1;
    a[i] = mk<T>(x[b * n + src], T(0));
}

template<typename T>
__global__ void __launch_bounds__(256) dct2_post_kernel(cx<T> const* __restrict__ a, T* __restrict__ out, cx<T> const* __restrict__ scale,
                                                        size_t n, size_t total)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    cx<T> const s = scale[i % n];
    out[i]        = a[i].x * s.x - a[i].y * s.y;
}

// the same two steps as plain kernels, for padded sizes beyond the single-CTA transform
template<typename T>
__global__ void __launch_bounds__(256) bluestein_pre_kernel(cx<T> const* __restrict__ x, cx<T>* __restrict__ a, cx<T> const* __restrict__ w,
                                                            size_t n, size_t m, size_t total)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    size_t const b = i / m, j = i - b * m;
    a[i]           = j < n ? cmul(x[b * n + j], w[j]) : mk<T>(T(0), T(0));
}

template<typename T>
__global__ void __launch_bounds__(256) bluestein_mul_kernel(cx<T>* __restrict__ a, cx<T> const* __restrict__ bhat, size_t m, size_t total)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    a[i] = cmul(a[i], bhat[i % m]);
}

template<typename T>
__global__ void __launch_bounds__(256) bluestein_post_kernel(cx<T> const* __restrict__ a, cx<T>* __restrict__ out, cx<T> const* __restrict__ w,
                                                             size_t n, size_t m, T scale, size_t total)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) { return; }
    size_t const b = i / n, j = i - b * n;
    cx<T> const v  = a[b * m + j];
    out[i]         = cmul(mk<T>(v.x * scale, v.y * scale), w[j]);
}

template<typename T, int LOGM, int DIR, class IO>
__global__ void __launch_bounds__(fft_cfg<T, LOGM>::THREADS, fft_min_ctas<T, LOGM, k_c2c>())
    c2c_kernel(IO io, cx<T> const* __restrict__ tw, size_t batch)
{
    using cfg = fft_cfg<T, LOGM>;
    using F   = cta_fft<T, LOGM, DIR>;
    using C   = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);

    int const g     = threadIdx.x / cfg::TN;
    int const t     = threadIdx.x % cfg::TN;
    size_t const b  = size_t(blockIdx.x) * cfg::G + g;
    bool const live = b < batch;

    C v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { v[e] = live ? io.load(b, t + e * cfg::TN) : mk<T>(0, 0); }

    F::run(v, sm + g * cfg::TILE, tw, t);

    if (live) {
#pragma unroll
        for (int e = 0; e < cfg::E; ++e) { io.store(b, t + e * cfg::TN, v[e]); }
    }
}

// ---- r2c: IO policy provides load(b, j) -> z[j] and the spectrum stores ------------------------------------------------
template<typename T, int LOGM, class IO>
__global__ void __launch_bounds__(fft_cfg<T, LOGM>::THREADS, fft_min_ctas<T, LOGM, k_r2c>())
    r2c_kernel(IO io, cx<T> const* __restrict__ tw, cx<T> const* __restrict__ rtw, size_t batch)
{
    using cfg = fft_cfg<T, LOGM>;
    using F   = cta_fft<T, LOGM, -1>;
    using C   = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw) + (threadIdx.x / cfg::TN) * cfg::TILE;

    int const g     = threadIdx.x / cfg::TN;
    int const t     = threadIdx.x % cfg::TN;
    size_t const b  = size_t(blockIdx.x) * cfg::G + g;
    bool const live = b < batch;

    C v[cfg::E];
    typename IO::row_state row = io.open(live ? b : 0);
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { v[e] = live ? io.load(row, t + e * cfg::TN) : mk<T>(0, 0); }
    if (live) {
#pragma unroll
        for (int e = 0; e < cfg::E; ++e) { io.keep(row, t + e * cfg::TN, v[e]); }  // e.g. the convolver's next half-window
    }

    F::run(v, sm, tw, t);

    // Hermitian exchange: both sides are unit stride (k ascending on the store, M - k descending on the load), so the tile is used
    // UNPADDED here -- with the stage padding a half-warp's 16 partners straddle a pad boundary and collide two-way (ncu: 16 % of the
    // kernel's shared-memory wavefronts were bank conflicts).
    // A thread finishes PAIRS: for each of its bins k < M/2 (the first half of its points) it fetches Z[M-k] and produces X[k] and
    // X[M-k] from one twiddle product -- half the partner loads, half the twiddle loads and 10 of 24 flops per pair less than doing
    // every bin on its own.
    if constexpr (cfg::E >= 2) {
        constexpr int M = cfg::M, E = cfg::E, TN = cfg::TN;
#pragma unroll
        for (int e = 0; e < E; ++e) { sm[t + e * TN] = v[e]; }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int e = 0; e < E / 2; ++e) {
                int const k = t + e * TN;
                if (k == 0) {
                    io.store_edges(row, v[e].x + v[e].y, v[e].x - v[e].y);
                } else {
                    C xk, xmk;
                    r2c_post_pair(v[e], sm[M - k], __ldg(rtw + k), xk, xmk);
                    io.store(row, k, xk);
                    io.store(row, M - k, xmk);
                }
            }
            if (t == 0) { io.store(row, M / 2, cconj(v[E / 2])); }  // the self-paired bin: W^(M/2) = -i
        }
    } else {
        // one point per thread (M <= 4): every bin on its own
        if constexpr (LOGM > 0) {
            sm[t] = v[0];
            __syncthreads();
        }
        if (live) {
            int const k = t;
            if (k == 0) { io.store_edges(row, v[0].x + v[0].y, v[0].x - v[0].y); }
            else { io.store(row, k, r2c_post(v[0], sm[cfg::M - k], __ldg(rtw + k))); }
        }
    }
}

// ---- c2r: IO policy provides the spectrum loads and store(b, j, z[j]) -----------------------------------------------------
template<typename T, int LOGM, class IO>
__global__ void __launch_bounds__(fft_cfg<T, LOGM>::THREADS, fft_min_ctas<T, LOGM, k_c2r>())
    c2r_kernel(IO io, cx<T> const* __restrict__ tw, cx<T> const* __restrict__ rtw, size_t batch)
{
    using cfg = fft_cfg<T, LOGM>;
    using F   = cta_fft<T, LOGM, +1>;
    using C   = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw) + (threadIdx.x / cfg::TN) * cfg::TILE;

    int const g     = threadIdx.x / cfg::TN;
    int const t     = threadIdx.x % cfg::TN;
    size_t const b  = size_t(blockIdx.x) * cfg::G + g;
    bool const live = b < batch;

    C v[cfg::E];
    typename IO::row_state row = io.open(live ? b : 0);
    if constexpr (cfg::E >= 2) {
        // Hermitian pre-pass by PAIRS: a thread loads X[k] for the first half of its points (k < M/2) and the partners X[M-k], forms
        // Z[k] and Z[M-k] from one twiddle product (c2r_pre_pair), keeps Z[k] and hands Z[M-k] to its owner through the (unpadded,
        // unit-stride) tile. E global loads per thread instead of 2E, half the twiddle loads, 12 instead of 20 flops per pair.
        constexpr int M = cfg::M, E = cfg::E, TN = cfg::TN, HALF = E / 2;
        C own[HALF], mate[HALF];
#pragma unroll
        for (int e = 0; e < HALF; ++e) {
            int const k = t + e * TN;
            if (!live) {
                own[e] = mate[e] = mk<T>(0, 0);
            } else if (k == 0) {
                own[e]  = io.load_edges(row);    // (Re X[0], Re X[M])
                mate[e] = io.load(row, M / 2);   // the self-paired bin
            } else {
                own[e]  = io.load(row, k);
                mate[e] = io.load(row, M - k);
            }
        }
#pragma unroll
        for (int e = 0; e < HALF; ++e) {
            int const k = t + e * TN;
            if (k == 0) {
                v[e]    = mk<T>(own[e].x + own[e].y, own[e].x - own[e].y);
                v[HALF] = mk<T>(T(2) * mate[e].x, T(-2) * mate[e].y);  // Z[M/2] = 2 conj(X[M/2]); index M/2 = HALF * TN is this thread's
            } else {
                C zmk;
                c2r_pre_pair(own[e], mate[e], __ldg(rtw + k), v[e], zmk);
                sm[M - k] = zmk;
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = HALF; e < E; ++e) {
            if (!(e == HALF && t == 0)) { v[e] = sm[t + e * TN]; }
        }
        __syncthreads();
    } else {
        C const own = live ? io.load_edges(row) : mk<T>(0, 0);
        v[0]        = mk<T>(own.x + own.y, own.x - own.y);
    }

    F::run(v, sm, tw, t);

    if (live) {
#pragma unroll
        for (int e = 0; e < cfg::E; ++e) { io.store(row, t + e * cfg::TN, v[e]); }
    }
}

// ---- plain batched IO policies (the rfft plan) ---------------------------------------------------------------------------
template<typename T, int LOGM>
struct r2c_plain_io
{
    using C = cx<T>;
    T const* in;  // [batch][2M]
    C* out;       // [batch][M+1]
    struct row_state
    {
        C const* src;
        C* dst;
    };
    __device__ __forceinline__ row_state open(size_t b) const
    {
        return {reinterpret_cast<C const*>(in) + b * (size_t(1) << LOGM), out + b * ((size_t(1) << LOGM) + 1)};
    }
    __device__ __forceinline__ C load(row_state const& r, int j) const { return r.src[j]; }
    __device__ __forceinline__ void keep(row_state const&, int, C) const {}
    __device__ __forceinline__ void store(row_state const& r, int k, C x) const { r.dst[k] = x; }
    __device__ __forceinline__ void store_edges(row_state const& r, T dc, T nyq) const
    {
        r.dst[0]         = mk<T>(dc, T(0));
        r.dst[1 << LOGM] = mk<T>(nyq, T(0));
    }
};

template<typename T, int LOGM>
struct c2r_plain_io
{
    using C = cx<T>;
    C const* in;     // [batch][row_len], first M+1 used
    T* out;          // [batch][2M]
    size_t row_len;
    struct row_state
    {
        C const* src;
        C* dst;
    };
    __device__ __forceinline__ row_state open(size_t b) const
    {
        return {in + b * row_len, reinterpret_cast<C*>(out) + b * (size_t(1) << LOGM)};
    }
    __device__ __forceinline__ C load(row_state const& r, int k) const { return r.src[k]; }
    __device__ __forceinline__ C load_edges(row_state const& r) const
    {
        return mk<T>(r.src[0].x, r.src[1 << LOGM].x);
    }
    __device__ __forceinline__ void store(row_state const& r, int j, C z) const { r.dst[j] = z; }
};

// ---- launch helpers ----------------------------------------------------------------------------------------------------------
template<typename Kernel>
int enable_smem(Kernel kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        NEO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    }
    return NEO_B200_OK;
}

template<typename T, int LOGM, class IO>
int launch_r2c(IO const& io, cx<T> const* tw, cx<T> const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = fft_cfg<T, LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = r2c_kernel<T, LOGM, IO>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    size_t const grid = (batch + cfg::G - 1) / cfg::G;
    kernel<<<static_cast<unsigned>(grid), cfg::THREADS, cfg::SMEM, stream>>>(io, tw, rtw, batch);
    return check_launch("r2c_kernel");
}

template<typename T, int LOGM, class IO>
int launch_c2r(IO const& io, cx<T> const* tw, cx<T> const* rtw, size_t batch, cudaStream_t stream)
{
    using cfg = fft_cfg<T, LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = c2r_kernel<T, LOGM, IO>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    size_t const grid = (batch + cfg::G - 1) / cfg::G;
    kernel<<<static_cast<unsigned>(grid), cfg::THREADS, cfg::SMEM, stream>>>(io, tw, rtw, batch);
    return check_launch("c2r_kernel");
}

template<typename T, int LOGM, int DIR, class IO>
int launch_c2c_io(IO const& io, cx<T> const* tw, size_t batch, cudaStream_t stream)
{
    using cfg = fft_cfg<T, LOGM>;
    if (batch == 0) { return NEO_B200_OK; }
    auto kernel = c2c_kernel<T, LOGM, DIR, IO>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    size_t const grid = (batch + cfg::G - 1) / cfg::G;
    kernel<<<static_cast<unsigned>(grid), cfg::THREADS, cfg::SMEM, stream>>>(io, tw, batch);
    return check_launch("c2c_kernel");
}

template<typename T, int LOGM, int DIR>
int launch_c2c(cx<T> const* in, cx<T>* out, cx<T> const* tw, size_t batch, cudaStream_t stream)
{
    return launch_c2c_io<T, LOGM, DIR>(c2c_interleaved_io<T>{in, out, size_t(1) << LOGM}, tw, batch, stream);
}

// split <-> interleaved conversion for transforms that do not fit one CTA
template<typename T>
__global__ void __launch_bounds__(256) split_to_interleaved_kernel(T const* __restrict__ re, T const* __restrict__ im, cx<T>* __restrict__ out, size_t n)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = mk<T>(re[i], im[i]); }
}

template<typename T>
__global__ void __launch_bounds__(256) interleaved_to_split_kernel(cx<T> const* __restrict__ in, T* __restrict__ re, T* __restrict__ im, size_t n)
{
    size_t const i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        cx<T> const v = in[i];
        re[i] = v.x;
        im[i] = v.y;
    }
}

// device-resident twiddle tables of one transform size
template<typename T>
struct fft_tables
{
    int logm{-1};
    device_buffer stage;  // cta_fft stage twiddles
    device_buffer split;  // real<->complex split twiddles (only when requested)

    int build(int logm_, bool with_split, cudaStream_t stream)
    {
        logm          = logm_;
        auto const tw = make_stage_twiddles<T>(logm);
        NEO_TRY(stage.reserve(tw.size() * sizeof(cx<T>)));
        NEO_CUDA_TRY(cudaMemcpyAsync(stage.ptr, tw.data(), tw.size() * sizeof(cx<T>), cudaMemcpyHostToDevice, stream));
        if (with_split) {
            auto const sp = make_split_twiddles<T>(logm);
            NEO_TRY(split.reserve(sp.size() * sizeof(cx<T>)));
            NEO_CUDA_TRY(cudaMemcpyAsync(split.ptr, sp.data(), sp.size() * sizeof(cx<T>), cudaMemcpyHostToDevice, stream));
        }
        NEO_CUDA_TRY(cudaStreamSynchronize(stream));  // host vectors die here
        return NEO_B200_OK;
    }

    cx<T> const* tw() const { return stage.template as<cx<T>>(); }
    cx<T> const* rtw() const { return split.template as<cx<T>>(); }
};

// switch over the compile-time transform size
#define NEO_DISPATCH_LOGM(T, logm, ...) \
    switch (logm) { \
        case 0: { constexpr int LOGM = 0; __VA_ARGS__; } break; \
        case 1: { constexpr int LOGM = 1; __VA_ARGS__; } break; \
        case 2: { constexpr int LOGM = 2; __VA_ARGS__; } break; \
        case 3: { constexpr int LOGM = 3; __VA_ARGS__; } break; \
        case 4: { constexpr int LOGM = 4; __VA_ARGS__; } break; \
        case 5: { constexpr int LOGM = 5; __VA_ARGS__; } break; \
        case 6: { constexpr int LOGM = 6; __VA_ARGS__; } break; \
        case 7: { constexpr int LOGM = 7; __VA_ARGS__; } break; \
        case 8: { constexpr int LOGM = 8; __VA_ARGS__; } break; \
        case 9: { constexpr int LOGM = 9; __VA_ARGS__; } break; \
        case 10: { constexpr int LOGM = 10; __VA_ARGS__; } break; \
        case 11: { constexpr int LOGM = 11; __VA_ARGS__; } break; \
        case 12: { constexpr int LOGM = 12; __VA_ARGS__; } break; \
        case 13: { constexpr int LOGM = 13; __VA_ARGS__; } break; \
        default: break; \
    }

}  // namespace neo_b200
