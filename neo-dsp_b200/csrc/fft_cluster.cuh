// fft_cluster.cuh -- real transforms of N = 2M points, M = M1 * 512 complex points with M1 = 16, 32, 64 (N = 2^14, 2^15, 2^16),
// as ONE persistent kernel: a thread-block cluster of CS = M1/8 CTAs owns a transform and runs the four-step algorithm with both
// passes fully coalesced; the matrix transposition between the passes goes through an L2-resident scratch (one row-major
// [M1][512] tile per cluster, double buffered) and a cluster barrier, not through HBM:
//   phase 1  every CTA takes 512/CS adjacent columns: loads (512/CS * 8 B contiguous per row), length-M1 column FFTs
//            (16 points per thread), twiddle W_M^(k1 n2), store to scratch
//   barrier.cluster (release/acquire)
//   phase 2  every CTA takes 8 rows (for r2c: 4 rows k1 and their mirrors M1-k1, so the Hermitian partner Z[M-k] is in the same
//            CTA), length-512 row FFTs, Hermitian split through shared memory, store X[k1 + M1 k2]
// HBM sees each datum once in and once out. The reference covers these sizes with its one radix-2 loop
// (fft/fallback/fallback_rfft_plan.hpp:28-55 over c2c_dit2_plan.hpp:84-95).
#pragma once

#include "fft_kernels.cuh"
#include "fft_large.cuh"

#include <cstdlib>

namespace neo_b200 {

template<int LOGM1>
struct cluster_cfg
{
    static constexpr int M1      = 1 << LOGM1;
    static constexpr int M2      = 512;
    static constexpr int M       = M1 * M2;
    static constexpr int CS      = M1 / 8;            // CTAs per cluster
    static constexpr int THREADS = 256;
    static constexpr int TN1     = M1 / 16;           // threads per column (16 points each)
    static constexpr int COLS    = THREADS / TN1;     // columns per CTA; COLS * CS == 512
    static constexpr int TN2     = 32;                // threads per 512-point row
    static constexpr int ROWS    = 8;                 // rows per CTA
    using FC                     = cta_fft<float, LOGM1, -1, 4>;
    using FR                     = cta_fft<float, 9, -1>;
    static constexpr int TILE1   = FC::TILE;
    static constexpr int TILE2   = FR::TILE;
    static constexpr int OUTS    = M2 * (ROWS + 1);   // phase-2 results transposed to [k2][row slot], padded to 9 per k2
    static constexpr int TILES0  = COLS * TILE1 > ROWS * TILE2 ? COLS * TILE1 : ROWS * TILE2;
    static constexpr int TILES   = TILES0 > OUTS ? TILES0 : OUTS;  // exchange tiles / output staging (elements)
    static constexpr int STAGE   = 16 * THREADS;                                               // one prefetched operand per point
    // r2c prefetches z[n]; c2r prefetches X[n] and X[M-n]
    // every twiddle table is copied into shared memory once per CTA: the cluster barrier's acquire invalidates L1, so tables left
    // in global memory would come from L2 again for every transform
    static constexpr int TAB_COL  = 0;                                   // column-FFT stage twiddles
    static constexpr int N_COL    = 4 * M1 + 8;
    static constexpr int TAB_ROW  = TAB_COL + N_COL;                     // 512-point row-FFT stage twiddles
    static constexpr int N_ROW    = 4 * (2 + 32) + 8;
    static constexpr int TAB_HI   = TAB_ROW + N_ROW;                     // W_M two-level tables
    static constexpr int N_HI     = 256;
    static constexpr int TAB_LO   = TAB_HI + N_HI;
    static constexpr int N_LO     = 256;
    static constexpr int TAB_K1   = TAB_LO + N_LO;                       // exp(-2 pi i k1 / N)
    static constexpr int TAB_1024 = TAB_K1 + M1;                         // exp(-2 pi i k2 / 1024)
    static constexpr int TAB_2M1  = TAB_1024 + 512;                      // exp(-2 pi i n1 / 2 M1)
    static constexpr int TAB_N2   = TAB_2M1 + M1;                        // exp(-2 pi i n2 / N)
    static constexpr int TABLES   = TAB_N2 + 512;
    __host__ __device__ static constexpr int stages(int direction, bool prefetch) { return prefetch ? (direction < 0 ? 1 : 2) : 0; }
    __host__ __device__ static constexpr size_t smem(int direction, bool prefetch) { return sizeof(float2) * size_t(TILES + stages(direction, prefetch) * STAGE + TABLES); }
};

struct cluster_tables
{
    float2 const* col_tw;   // stage twiddles of the length-M1 column FFT (16 points per thread)
    float2 const* row_tw;   // stage twiddles of the 512-point row FFT
    twiddle2_view<float> w_m;  // W_M^j, j < M (two-level)
    float2 const* w_n_k1;   // exp(-2 pi i k1 / N), k1 < M1
    float2 const* w_1024;   // exp(-2 pi i k2 / 1024), k2 < 512          (W_N^(M1 k2))
    float2 const* w_2m1;    // exp(-2 pi i n1 / (2 M1)), n1 < M1          (W_N^(512 n1))
    float2 const* w_n_n2;   // exp(-2 pi i n2 / N), n2 < 512
    int n_col, n_row, n_hi, n_lo;  // entries actually present in col_tw, row_tw, w_m.hi, w_m.lo
};

__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// 8-byte asynchronous global -> shared copy (LDGSTS); a thread only ever reads back what it copied itself
__device__ __forceinline__ void cp_async8(void* smem_dst, void const* gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ unsigned cluster_id_x()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}

__device__ __forceinline__ unsigned cluster_count_x()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}

// DIRECTION -1: r2c (in: [batch][2M] reals, out: [batch][M+1] complex); +1: c2r (in: [batch][row_len] complex, out: [batch][2M] reals)
template<int LOGM1, int DIRECTION, bool PREFETCH, int MINCTA>
__global__ void __launch_bounds__(256, MINCTA) rfft_cluster_kernel(void const* __restrict__ in_v, void* __restrict__ out_v, size_t row_len,
                                                          float2* __restrict__ scratch, cluster_tables tb, size_t batch)
{
    using cfg = cluster_cfg<LOGM1>;
    using C   = float2;
    constexpr int M1 = cfg::M1, M2 = cfg::M2, M = cfg::M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* const sm = reinterpret_cast<C*>(smem_raw);

    int const tid        = threadIdx.x;
    unsigned const rank  = cluster_rank();
    unsigned const cid   = cluster_id_x();
    unsigned const ncl   = cluster_count_x();
    C* const scratch_cl  = scratch + size_t(cid) * 2 * M;

    // phase-1 coordinates: column n2, 16 rows n1 = t1 + e*TN1
    int const g1 = tid % cfg::COLS;
    int const t1 = tid / cfg::COLS;
    int const n2 = int(rank) * cfg::COLS + g1;
    // phase-2 coordinates: row slot q (8 per CTA), k2 = t2 + e*32
    int const t2 = tid % cfg::TN2;
    int const q  = tid / cfg::TN2;
    int row, mate_slot;
    if constexpr (DIRECTION < 0) {
        int const pair = int(rank) * 4 + q / 2;  // pair 0 = rows (0, M1/2), pair p = rows (p, M1 - p)
        row            = (q & 1) == 0 ? pair : (pair == 0 ? M1 / 2 : M1 - pair);
        mate_slot      = pair == 0 ? q : (q ^ 1);
    } else {
        row       = int(rank) * cfg::ROWS + q;
        mate_slot = q;
    }

    // phase-1 operands of the NEXT transform are prefetched with cp.async into a private staging slot per (point, thread) while
    // the current transform sits in its barrier and row phase: the HBM latency leaves the dependency chain
    C* const stage0 = sm + cfg::TILES;
    C* const stage1 = stage0 + cfg::STAGE;  // c2r only: the Hermitian partners X[M-n]
    C* const tab    = stage0 + cfg::stages(DIRECTION, PREFETCH) * cfg::STAGE;
    {
        auto copy = [&](int at, C const* src, int n) {
            for (int i = tid; i < n; i += cfg::THREADS) { tab[at + i] = src[i]; }
        };
        copy(cfg::TAB_COL, tb.col_tw, tb.n_col);
        copy(cfg::TAB_ROW, tb.row_tw, tb.n_row);
        copy(cfg::TAB_HI, tb.w_m.hi, tb.n_hi);
        copy(cfg::TAB_LO, tb.w_m.lo, tb.n_lo);
        copy(cfg::TAB_K1, tb.w_n_k1, M1);
        copy(cfg::TAB_1024, tb.w_1024, 512);
        copy(cfg::TAB_2M1, tb.w_2m1, M1);
        copy(cfg::TAB_N2, tb.w_n_n2, 512);
        __syncthreads();
    }
    int const lo_bits = tb.w_m.lo_bits;
    auto w_m = [&](size_t j) {  // W_M^j, forward sign
        return cmul(tab[cfg::TAB_HI + int(j >> lo_bits)], tab[cfg::TAB_LO + int(j & ((size_t(1) << lo_bits) - 1))]);
    };
    auto prefetch = [&](size_t b) {
        if constexpr (DIRECTION < 0) {
            C const* const z = reinterpret_cast<C const*>(in_v) + b * size_t(M);
#pragma unroll
            for (int e = 0; e < 16; ++e) { cp_async8(stage0 + e * cfg::THREADS + tid, z + size_t(t1 + e * cfg::TN1) * M2 + n2); }
        } else {
            C const* const x = reinterpret_cast<C const*>(in_v) + b * row_len;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int const n = (t1 + e * cfg::TN1) * M2 + n2;
                cp_async8(stage0 + e * cfg::THREADS + tid, x + n);
                cp_async8(stage1 + e * cfg::THREADS + tid, x + (M - n));  // n = 0 reads X[M], the Nyquist bin it needs anyway
            }
        }
        cp_async_commit();
    };
    if constexpr (PREFETCH) {
        if (cid < batch) { prefetch(cid); }
    }

    int parity = 0;
    for (size_t b = cid; b < batch; b += ncl, parity ^= 1) {
        C* const a = scratch_cl + size_t(parity) * M;  // [M1][512]

        // ---------------- phase 1: columns ----------------
        C v[16];
        if constexpr (PREFETCH) { cp_async_wait_all(); }
        if constexpr (DIRECTION < 0) {
            C const* const zin = reinterpret_cast<C const*>(in_v) + b * size_t(M);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                v[e] = PREFETCH ? stage0[e * cfg::THREADS + tid] : __ldcs(zin + size_t(t1 + e * cfg::TN1) * M2 + n2);
            }
            cta_fft<float, LOGM1, -1, 4>::run(v, sm + g1 * cfg::TILE1, tab + cfg::TAB_COL, t1);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int const k1 = t1 + e * cfg::TN1;
                a[size_t(k1) * M2 + n2] = cmul(v[e], w_m(size_t(k1) * n2));
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int const n1 = t1 + e * cfg::TN1;
                int const n  = n1 * M2 + n2;
                C const* const xin = reinterpret_cast<C const*>(in_v) + b * row_len;
                C const own  = PREFETCH ? stage0[e * cfg::THREADS + tid] : __ldcs(xin + n);
                C const mate = PREFETCH ? stage1[e * cfg::THREADS + tid] : __ldcs(xin + (M - n));
                if (n == 0) {
                    v[e] = make_float2(own.x + mate.x, own.x - mate.x);  // (Re X[0], Re X[M])
                } else {
                    C const w = cmul(tab[cfg::TAB_2M1 + n1], tab[cfg::TAB_N2 + n2]);  // W_N^n
                    v[e]      = c2r_pre(own, mate, w);
                }
            }
            cta_fft<float, LOGM1, +1, 4>::run(v, sm + g1 * cfg::TILE1, tab + cfg::TAB_COL, t1);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int const k1 = t1 + e * cfg::TN1;
                a[size_t(k1) * M2 + n2] = cmulc(v[e], w_m(size_t(k1) * n2));
            }
        }
        if constexpr (PREFETCH) {
            if (b + ncl < batch) { prefetch(b + ncl); }
        }

        cluster_barrier();  // scratch tile complete and visible to the whole cluster

        // ---------------- phase 2: rows ----------------
        C* const tile = sm + q * cfg::TILE2;
#pragma unroll
        for (int e = 0; e < 16; ++e) { v[e] = __ldcg(a + size_t(row) * M2 + t2 + e * cfg::TN2); }
        // results leave through shared memory transposed to [k2][row slot]: a CTA owns only 8 rows, i.e. 32-64 contiguous bytes
        // per k2, so four (r2c) or eight (c2r) lanes write one full sector together instead of every lane its own 8 bytes
        // 512 bytes apart (measured: the strided form made phase 2 cost 70 % of the kernel)
        int slot_pos, k_row;
        if constexpr (DIRECTION < 0) {
            cta_fft<float, 9, -1, -1, true>::run(v, tile, tab + cfg::TAB_ROW, t2);  // one warp per row: warp-level exchanges
            // Hermitian split: Z[k], k = row + M1*k2, pairs with Z[M-k] = (mirror row, k2' below)
#pragma unroll
            for (int e = 0; e < 16; ++e) { tile[padded<float>(t2 + e * cfg::TN2)] = v[e]; }
            __syncthreads();
            C const* const mate = sm + mate_slot * cfg::TILE2;
            C const wk1         = tab[cfg::TAB_K1 + row];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                int const k2 = t2 + e * cfg::TN2;
                if (row + k2 == 0) {
                    C* const xrow = reinterpret_cast<C*>(out_v) + b * (size_t(M) + 1);
                    xrow[M]       = make_float2(v[e].x - v[e].y, 0.f);
                    v[e]          = make_float2(v[e].x + v[e].y, 0.f);
                } else {
                    int const k2m = row == 0 ? (M2 - k2) & (M2 - 1) : M2 - 1 - k2;
                    C const zp    = mate[padded<float>(k2m)];
                    v[e]          = r2c_post(v[e], zp, cmul(wk1, tab[cfg::TAB_1024 + k2]));
                }
            }
            __syncthreads();  // every partner has been read: the tiles become the output staging
            // slots 0-3: the four first rows (ascending), slots 4-7: their mirrors (ascending)
            slot_pos = (q & 1) == 0 ? q / 2 : 4 + (3 - q / 2);
        } else {
            cta_fft<float, 9, +1, -1, true>::run(v, tile, tab + cfg::TAB_ROW, t2);
            __syncthreads();
            slot_pos = q;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) { sm[(t2 + e * cfg::TN2) * (cfg::ROWS + 1) + slot_pos] = v[e]; }
        __syncthreads();
        {
            int const slot = tid & 7;
            if constexpr (DIRECTION < 0) {
                int const i    = slot < 4 ? slot : 3 - (slot - 4);
                int const pair = int(rank) * 4 + i;
                k_row          = slot < 4 ? pair : (pair == 0 ? M1 / 2 : M1 - pair);
            } else {
                k_row = int(rank) * cfg::ROWS + slot;
            }
            C* const dst = reinterpret_cast<C*>(out_v) + b * (DIRECTION < 0 ? size_t(M) + 1 : size_t(M));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int const k2 = (tid >> 3) + i * (cfg::THREADS / 8);
                dst[k_row + M1 * k2] = sm[k2 * (cfg::ROWS + 1) + slot];
            }
        }
        __syncthreads();  // staging is reused as exchange tiles by the next transform's phase 1
    }
}

// host side: tables + persistent launch
struct rfft_cluster_plan
{
    int logm1{0};
    device_buffer col_tw, row_tw, w_n_k1, w_1024, w_2m1, w_n_n2, scratch;
    twiddle2<float> w_m;
    int clusters[2]{0, 0};  // resident clusters for the forward / backward kernel

    template<typename F>
    static int upload(device_buffer& buf, size_t n, F f, cudaStream_t stream)
    {
        std::vector<float2> host(n);
        for (size_t i = 0; i < n; ++i) { host[i] = f(i); }
        NEO_TRY(buf.reserve(n * sizeof(float2)));
        NEO_CUDA_TRY(cudaMemcpyAsync(buf.ptr, host.data(), n * sizeof(float2), cudaMemcpyHostToDevice, stream));
        NEO_CUDA_TRY(cudaStreamSynchronize(stream));
        return NEO_B200_OK;
    }

    static float2 unit(double turns)  // exp(-2 pi i turns)
    {
        double const a = -2.0 * 3.14159265358979323846264338327950288 * turns;
        return make_float2(float(std::cos(a)), float(std::sin(a)));
    }

    int init(int order, cudaStream_t stream)
    {
        logm1            = order - 1 - 9;
        size_t const m1  = size_t(1) << logm1;
        double const n   = std::ldexp(1.0, order);
        auto const ctw   = make_stage_twiddles<float>(logm1, 4);
        auto const rtw   = make_stage_twiddles<float>(9);
        NEO_TRY(upload(col_tw, ctw.size(), [&](size_t i) { return ctw[i]; }, stream));
        NEO_TRY(upload(row_tw, rtw.size(), [&](size_t i) { return rtw[i]; }, stream));
        n_col = int(ctw.size());
        n_row = int(rtw.size());
        NEO_TRY(w_m.build(order - 1, stream));
        NEO_TRY(upload(w_n_k1, m1, [&](size_t i) { return unit(double(i) / n); }, stream));
        NEO_TRY(upload(w_1024, 512, [&](size_t i) { return unit(double(i) / 1024.0); }, stream));
        NEO_TRY(upload(w_2m1, m1, [&](size_t i) { return unit(double(i) / double(2 * m1)); }, stream));
        NEO_TRY(upload(w_n_n2, 512, [&](size_t i) { return unit(double(i) / n); }, stream));
        return NEO_B200_OK;
    }

    int n_col{0}, n_row{0};

    cluster_tables tables() const
    {
        return {col_tw.as<float2>(), row_tw.as<float2>(), w_m.view(), w_n_k1.as<float2>(), w_1024.as<float2>(), w_2m1.as<float2>(),
                w_n_n2.as<float2>(), n_col, n_row, int(w_m.hi.bytes / sizeof(float2)), int(w_m.lo.bytes / sizeof(float2))};
    }

    template<int LOGM1, int DIRECTION>
    int launch(void const* in, void* out, size_t row_len, size_t batch, cudaStream_t stream)
    {
        // measured (N = 2^16): cp.async prefetch of the next transform and 2, 3 or 4 resident CTAs per SM all land within 3 %
        // (1289-1329 us per GiB pass); the plain 3-CTA form ships
        return launch_variant<LOGM1, DIRECTION, false, 3>(in, out, row_len, batch, stream);
    }

    template<int LOGM1, int DIRECTION, bool PREFETCH, int MINCTA>
    int launch_variant(void const* in, void* out, size_t row_len, size_t batch, cudaStream_t stream)
    {
        using cfg   = cluster_cfg<LOGM1>;
        auto kernel = rfft_cluster_kernel<LOGM1, DIRECTION, PREFETCH, MINCTA>;
        NEO_TRY(enable_smem(kernel, cfg::smem(DIRECTION, PREFETCH)));
        cudaLaunchConfig_t lc{};
        lc.blockDim         = dim3(cfg::THREADS);
        lc.dynamicSmemBytes = cfg::smem(DIRECTION, PREFETCH);
        lc.stream           = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id               = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cfg::CS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        lc.attrs    = attr;
        lc.numAttrs = 1;
        int& resident = clusters[DIRECTION < 0 ? 0 : 1];
        if (resident == 0) {
            lc.gridDim = dim3(cfg::CS);  // occupancy query needs a grid that is a multiple of the cluster size
            int max_clusters = 0;
            NEO_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &lc));
            if (max_clusters < 1) { return fail(NEO_B200_ERR_CUDA, "no cluster of %d CTAs fits on this device", cfg::CS); }
            resident = max_clusters;
        }
        NEO_TRY(scratch.reserve(size_t(resident) * 2 * cfg::M * sizeof(float2)));
        size_t const use = std::min<size_t>(size_t(resident), batch);
        lc.gridDim       = dim3(unsigned(use * cfg::CS));
        cluster_tables const tb = tables();
        float2* const sc        = scratch.as<float2>();
        NEO_CUDA_TRY(cudaLaunchKernelEx(&lc, kernel, in, out, row_len, sc, tb, batch));
        return check_launch("rfft_cluster_kernel");
    }

    int forward(float const* in, float2* out, size_t batch, cudaStream_t stream)
    {
        if (batch == 0) { return NEO_B200_OK; }
        switch (logm1) {
            case 4: return launch<4, -1>(in, out, 0, batch, stream);
            case 5: return launch<5, -1>(in, out, 0, batch, stream);
            case 6: return launch<6, -1>(in, out, 0, batch, stream);
            default: return fail(NEO_B200_ERR_UNSUPPORTED, "cluster rfft: unsupported size");
        }
    }

    int backward(float2 const* in, size_t row_len, float* out, size_t batch, cudaStream_t stream)
    {
        if (batch == 0) { return NEO_B200_OK; }
        switch (logm1) {
            case 4: return launch<4, +1>(in, out, row_len, batch, stream);
            case 5: return launch<5, +1>(in, out, row_len, batch, stream);
            case 6: return launch<6, +1>(in, out, row_len, batch, stream);
            default: return fail(NEO_B200_ERR_UNSUPPORTED, "cluster rfft: unsupported size");
        }
    }
};

}  // namespace neo_b200
