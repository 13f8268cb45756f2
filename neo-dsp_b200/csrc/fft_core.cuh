// fft_core.cuh -- CTA-level power-of-two complex FFT held in registers, exchanged through shared memory.
//
// Replaces the arithmetic of the reference's radix-2 DIT plan (src/neo/fft/reference/c2c_dit2_plan.hpp:84-95,
// kernel/c2c_dit2.hpp:122-168) and its bit-reversal pass (bitrevorder.hpp:25-33) with a Stockham autosort
// decomposition: no permutation pass exists, every stage reads `x[j + q*M/r]` and writes
// `y[(j/Ns)*Ns*r + j%Ns + q*Ns]`, so data comes out in natural order.
//
// Layout contract: a transform of M = 2^LOGM complex points is owned by TN = M/E threads; thread t keeps
// E = 2^LOGE points v[e] = x[t + e*TN] before AND after run(). Between stages the points cross threads through a
// padded shared-memory tile (pad() below keeps both the scattered writes and the unit-stride reads conflict-free).
#pragma once

#include <cuda_runtime.h>

namespace neo_b200 {

template<typename T>
struct cx_of;
template<>
struct cx_of<float>
{
    using type = float2;
};
template<>
struct cx_of<double>
{
    using type = double2;
};
template<typename T>
using cx = typename cx_of<T>::type;

template<typename T>
__host__ __device__ __forceinline__ cx<T> mk(T re, T im)
{
    cx<T> r;
    r.x = re;
    r.y = im;
    return r;
}

template<typename C>
__device__ __forceinline__ C cadd(C a, C b)
{
    a.x += b.x;
    a.y += b.y;
    return a;
}

template<typename C>
__device__ __forceinline__ C csub(C a, C b)
{
    a.x -= b.x;
    a.y -= b.y;
    return a;
}

template<typename C>
__device__ __forceinline__ C cmul(C a, C b)
{
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// a * conj(b)
template<typename C>
__device__ __forceinline__ C cmulc(C a, C b)
{
    C r;
    r.x = a.x * b.x + a.y * b.y;
    r.y = a.y * b.x - a.x * b.y;
    return r;
}

template<typename C>
__device__ __forceinline__ C cconj(C a)
{
    a.y = -a.y;
    return a;
}

// multiply by exp(DIR * i*pi/2): forward (DIR=-1) is -i, backward is +i
template<int DIR, typename C>
__device__ __forceinline__ C rot90(C a)
{
    C r;
    if constexpr (DIR < 0) {
        r.x = a.y;
        r.y = -a.x;
    } else {
        r.x = -a.y;
        r.y = a.x;
    }
    return r;
}

// multiply by W16^EXP = exp(DIR * 2*pi*i*EXP/16), EXP known at compile time
template<int EXP, int DIR, typename C>
__device__ __forceinline__ C mul_w16(C a)
{
    using T          = decltype(a.x);
    constexpr int e  = ((EXP % 16) + 16) % 16;
    constexpr T c1   = T(0.92387953251128673848);
    constexpr T s1   = T(0.38268343236508978178);
    constexpr T h    = T(0.70710678118654752440);
    constexpr T cs[16] = {T(1), c1, h, s1, T(0), -s1, -h, -c1, T(-1), -c1, -h, -s1, T(0), s1, h, c1};
    constexpr T sn[16] = {T(0), s1, h, c1, T(1), c1, h, s1, T(0), -s1, -h, -c1, T(-1), -c1, -h, -s1};
    if constexpr (e == 0) {
        return a;
    } else if constexpr (e == 4) {
        return rot90<DIR>(a);
    } else if constexpr (e == 8) {
        a.x = -a.x;
        a.y = -a.y;
        return a;
    } else if constexpr (e == 12) {
        return rot90<-DIR>(a);
    } else {
        constexpr T wr = cs[e];
        constexpr T wi = (DIR < 0) ? -sn[e] : sn[e];
        C r;
        r.x = a.x * wr - a.y * wi;
        r.y = a.x * wi + a.y * wr;
        return r;
    }
}

// in-register DFTs of a compile-time radix over u[0..R)
template<int R, int DIR>
struct dft;

template<int DIR>
struct dft<1, DIR>
{
    template<typename C>
    static __device__ __forceinline__ void run(C*)
    {}
};

template<int DIR>
struct dft<2, DIR>
{
    template<typename C>
    static __device__ __forceinline__ void run(C* u)
    {
        C const a = u[0], b = u[1];
        u[0] = cadd(a, b);
        u[1] = csub(a, b);
    }
};

template<int DIR>
struct dft<4, DIR>
{
    template<typename C>
    static __device__ __forceinline__ void run(C* u)
    {
        C const t0 = cadd(u[0], u[2]);
        C const t1 = csub(u[0], u[2]);
        C const t2 = cadd(u[1], u[3]);
        C const t3 = rot90<DIR>(csub(u[1], u[3]));
        u[0] = cadd(t0, t2);
        u[1] = cadd(t1, t3);
        u[2] = csub(t0, t2);
        u[3] = csub(t1, t3);
    }
};

// R = R1*R2 (R1 = 4): n = R2*n1 + n2, k = k1 + R1*k2
//   A[n2][k1] = DFT_R1 over n1;  A *= W_R^(n2*k1);  X[k1 + R1*k2] = DFT_R2 over n2
template<int R, int DIR>
struct dft
{
    static constexpr int R1 = 4;
    static constexpr int R2 = R / 4;
    static_assert(R == 8 || R == 16, "radix");

    template<int N2, int K1, typename C>
    static __device__ __forceinline__ void twiddle_row(C (&a)[R2][R1])
    {
        if constexpr (K1 < R1) {
            a[N2][K1] = mul_w16<N2 * K1 * (16 / R), DIR>(a[N2][K1]);
            twiddle_row<N2, K1 + 1>(a);
        }
    }

    template<int N2, typename C>
    static __device__ __forceinline__ void twiddle_all(C (&a)[R2][R1])
    {
        if constexpr (N2 < R2) {
            twiddle_row<N2, 1>(a);
            twiddle_all<N2 + 1>(a);
        }
    }

    template<typename C>
    static __device__ __forceinline__ void run(C* u)
    {
        C a[R2][R1];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) { a[n2][n1] = u[R2 * n1 + n2]; }
            dft<R1, DIR>::run(a[n2]);
        }
        twiddle_all<1>(a);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            C b[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) { b[n2] = a[n2][k1]; }
            dft<R2, DIR>::run(b);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) { u[k1 + R1 * k2] = b[k2]; }
        }
    }
};

// points per thread for a transform of 2^logm complex points
template<typename T>
__host__ __device__ constexpr int pick_loge(int logm)
{
    if (logm <= 2) { return logm; }
    if (sizeof(T) == 4) { return logm >= 8 ? 4 : 3; }
    return 3;
}

// shared-memory index padding: one extra element per 128 bytes
template<typename T>
__host__ __device__ constexpr int pad_shift()
{
    return sizeof(T) == 4 ? 4 : 3;
}

template<typename T>
__host__ __device__ constexpr int padded(int i)
{
    return i + (i >> pad_shift<T>());
}

// stage list: an optional first stage of radix 2^(logm % loge) (no twiddles, Ns = 1), then logm/loge stages of
// radix 2^loge. The twiddle LUT holds, for every stage with Ns > 1, W_{Ns*r}^{q*k} for the powers of two q = 2^j only,
// at [j*Ns + k]; the other q are one to three complex products of those (loads, not FMAs, are the scarce resource).
__host__ __device__ constexpr int fft_first_logr(int logm, int loge) { return loge == 0 ? 0 : logm % loge; }

__host__ __device__ constexpr int fft_twiddle_offset(int logm, int loge, int logns_target)
{
    int off   = 0;
    int logns = 0;
    if (loge == 0) { return 0; }
    int const r0 = fft_first_logr(logm, loge);
    if (r0 > 0) { logns = r0; }
    else { logns = loge; }  // first full stage has Ns = 1: no table
    while (logns < logns_target) {
        off += loge << logns;
        logns += loge;
    }
    return off;
}

__host__ __device__ constexpr int fft_twiddle_count(int logm, int loge) { return fft_twiddle_offset(logm, loge, logm); }

// ---- the exchange tile behind cta_fft: overloads on the handle type ------------------------------------------------------------
// Plain pointer: one CTA's padded shared-memory tile.
template<typename C>
__device__ __forceinline__ void tile_store(C* sm, int idx, C v)
{
    sm[padded<decltype(v.x)>(idx)] = v;
}
template<typename C>
__device__ __forceinline__ C tile_load(C* sm, int idx)
{
    return sm[padded<decltype(sm->x)>(idx)];
}
template<typename C>
__device__ __forceinline__ void tile_sync(C*)
{
    __syncthreads();
}

// Two CTAs of a thread-block cluster share one transform: element idx lives in CTA (idx >> LOGC) & 1 (LOGC = log2 threads per
// CTA), so the Stockham read pattern t + e*TN (TN = 2 * threads per CTA) is always local and only stores cross the cluster
// (st.shared::cluster). Loads of arbitrary elements (the Hermitian partner) may be remote.
template<typename T, int LOGC>
struct dsmem_tile2
{
    cx<T>* local;     // this CTA's half, padded
    unsigned peer;    // shared::cluster address of the other CTA's half
    unsigned rank;    // %cluster_ctarank
    static __device__ __forceinline__ int offset(int idx) { return padded<T>(((idx >> (LOGC + 1)) << LOGC) | (idx & ((1 << LOGC) - 1))); }
};
template<typename T, int LOGC>
__device__ __forceinline__ void tile_store(dsmem_tile2<T, LOGC> sm, int idx, cx<T> v)
{
    static_assert(sizeof(T) == 4, "float only");
    int const off = dsmem_tile2<T, LOGC>::offset(idx);
    if (unsigned((idx >> LOGC) & 1) == sm.rank) {
        sm.local[off] = v;
    } else {
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(sm.peer + unsigned(off) * 8U), "f"(v.x), "f"(v.y) : "memory");
    }
}
template<typename T, int LOGC>
__device__ __forceinline__ cx<T> tile_load(dsmem_tile2<T, LOGC> sm, int idx)
{
    int const off = dsmem_tile2<T, LOGC>::offset(idx);
    if (unsigned((idx >> LOGC) & 1) == sm.rank) { return sm.local[off]; }
    cx<T> v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sm.peer + unsigned(off) * 8U) : "memory");
    return v;
}
template<typename T, int LOGC>
__device__ __forceinline__ void tile_sync(dsmem_tile2<T, LOGC>)
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// LOGE_FORCED >= 0 overrides the points-per-thread choice (the cluster kernels want 16 points per thread for short columns)
// WARP_PRIVATE: the transform's TN threads are exactly one warp and its tile is touched by no other warp, so the exchanges only
// need __syncwarp() (the cluster kernel's 512-point rows: no CTA-wide stall per stage).
template<typename T, int LOGM, int DIR, int LOGE_FORCED = -1, bool WARP_PRIVATE = false>
struct cta_fft
{
    template<typename SM>
    static __device__ __forceinline__ void sync(SM sm)
    {
        if constexpr (WARP_PRIVATE) { __syncwarp(); }
        else { tile_sync(sm); }
    }

    static constexpr int LOGE  = LOGE_FORCED >= 0 ? LOGE_FORCED : pick_loge<T>(LOGM);
    static constexpr int M     = 1 << LOGM;
    static constexpr int E     = 1 << LOGE;
    static constexpr int TN    = M / E;
    static constexpr int LOGR0 = fft_first_logr(LOGM, LOGE);
    // shared-memory elements one transform needs for its exchanges
    static constexpr int TILE = padded<T>(M) + 1;
    using C                   = cx<T>;

    // w[q] = W^(q*k): loaded when q is a power of two (row log2(q) of the stage table), else w[hi] * w[q - hi]
    template<int N>
    static __device__ __forceinline__ void stage_twiddle(int q, C (&w)[N], C const* __restrict__ row0, int ns)
    {
        int const hi = 1 << (31 - __clz(q));
        if (q == hi) { w[q] = row0[(31 - __clz(q)) * ns]; }  // plain load: the table may live in global (read-only path) or shared memory
        else { w[q] = cmul(w[hi], w[q - hi]); }
    }

    template<int LOGNS, int LOGR, typename SM>
    static __device__ __forceinline__ void stage(C (&v)[E], SM sm, C const* __restrict__ tw, int t)
    {
        constexpr int NS = 1 << LOGNS;
        constexpr int R  = 1 << LOGR;
        constexpr int BF = E / R;  // butterflies per thread
        constexpr bool last = (LOGNS + LOGR == LOGM);
        constexpr int off   = fft_twiddle_offset(LOGM, LOGE, LOGNS);

#pragma unroll
        for (int m = 0; m < BF; ++m) {
            int const j = t + m * TN;
            int const k = j & (NS - 1);
            C u[R];
#pragma unroll
            for (int q = 0; q < R; ++q) { u[q] = v[m + q * BF]; }
            if constexpr (NS > 1) {
                C w[R];
#pragma unroll
                for (int q = 1; q < R; ++q) {
                    stage_twiddle(q, w, tw + off + k, NS);
                    u[q] = (DIR < 0) ? cmul(u[q], w[q]) : cmulc(u[q], w[q]);
                }
            }
            dft<R, DIR>::run(u);
            if constexpr (last) {
#pragma unroll
                for (int q = 0; q < R; ++q) { v[m + q * BF] = u[q]; }
            } else {
                int const base = ((j >> LOGNS) << (LOGNS + LOGR)) + k;
#pragma unroll
                for (int q = 0; q < R; ++q) { tile_store(sm, base + (q << LOGNS), u[q]); }
            }
        }
        if constexpr (!last) {
            sync(sm);
#pragma unroll
            for (int e = 0; e < E; ++e) { v[e] = tile_load(sm, t + e * TN); }
            sync(sm);
        }
    }

    template<int LOGNS, typename SM>
    static __device__ __forceinline__ void full_stages(C (&v)[E], SM sm, C const* __restrict__ tw, int t)
    {
        if constexpr (LOGNS < LOGM) {
            stage<LOGNS, LOGE>(v, sm, tw, t);
            full_stages<LOGNS + LOGE>(v, sm, tw, t);
        }
    }

    // all threads of the CTA must call (barriers inside)
    template<typename SM>
    static __device__ __forceinline__ void run(C (&v)[E], SM sm, C const* __restrict__ tw, int t)
    {
        if constexpr (LOGM > 0) {
            if constexpr (LOGR0 > 0) { stage<0, LOGR0>(v, sm, tw, t); }
            full_stages<LOGR0>(v, sm, tw, t);
        }
    }
};

// ---- real <-> half-size complex (the packing trick the reference sketches in fft/experimental/rfft.hpp:152-203) ----
//
// r2c: x[2M] real, z[j] = x[2j] + i x[2j+1], Z = FFT_M(z). With Zc = conj(Z[M-k]) and W = exp(-2 pi i k / 2M):
//   X[k] = ((Z[k] + Zc) - i W (Z[k] - Zc)) / 2          0 <= k < M,   X[0] = Re Z0 + Im Z0,  X[M] = Re Z0 - Im Z0
// c2r (unnormalised backward, result scaled by 2M like fallback_rfft_plan.hpp:39-55):
//   Z[k] = (X[k] + Xc) + i conj(W) (X[k] - Xc),  Xc = conj(X[M-k]);  z = IFFT_M(Z);  x[2j] = Re z[j], x[2j+1] = Im z[j]
template<typename C>
__device__ __forceinline__ C r2c_post(C z, C zp, C w)
{
    using T = decltype(z.x);
    C const zc = cconj(zp);
    C const s  = cadd(z, zc);
    C const d  = csub(z, zc);
    C const wd = cmul(w, d);  // W * (Z - Zc); then multiply by -i: (x, y) -> (y, -x)
    C r;
    r.x = T(0.5) * (s.x + wd.y);
    r.y = T(0.5) * (s.y - wd.x);
    return r;
}

// Both members of a Hermitian pair from one twiddle product. With s = Z[k] + conj(Z[M-k]), d = Z[k] - conj(Z[M-k]) and
// W^(M-k) = -conj(W^k):   X[k] = (s - i W d) / 2,   X[M-k] = conj(s + i W d) / 2
template<typename C>
__device__ __forceinline__ void r2c_post_pair(C z, C zp, C w, C& xk, C& xmk)
{
    using T = decltype(z.x);
    C const zc = cconj(zp);
    C const s  = cadd(z, zc);
    C const d  = csub(z, zc);
    C const wd = cmul(w, d);
    T const hx = T(0.5) * s.x, hy = T(0.5) * s.y;
    xk.x  = fma(T(0.5), wd.y, hx);
    xk.y  = fma(T(-0.5), wd.x, hy);
    xmk.x = fma(T(-0.5), wd.y, hx);
    xmk.y = fma(T(-0.5), wd.x, -hy);
}

// c2r: Z[k] = s + i conj(W) d and Z[M-k] = conj(s - i conj(W) d) with s = X[k] + conj(X[M-k]), d = X[k] - conj(X[M-k])
template<typename C>
__device__ __forceinline__ void c2r_pre_pair(C x, C xp, C w, C& zk, C& zmk)
{
    C const xc = cconj(xp);
    C const s  = cadd(x, xc);
    C const d  = csub(x, xc);
    C const wd = cmulc(d, w);
    zk.x  = s.x - wd.y;
    zk.y  = s.y + wd.x;
    zmk.x = s.x + wd.y;
    zmk.y = wd.x - s.y;
}

template<typename C>
__device__ __forceinline__ C c2r_pre(C x, C xp, C w)
{
    C const xc = cconj(xp);
    C const s  = cadd(x, xc);
    C const d  = csub(x, xc);
    C const wd = cmulc(d, w);  // conj(W) * (X - Xc); then multiply by +i: (x, y) -> (-y, x)
    C r;
    r.x = s.x - wd.y;
    r.y = s.y + wd.x;
    return r;
}

}  // namespace neo_b200
