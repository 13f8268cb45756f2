// conv_frame.cuh -- second partition level along BLOCK TIME ("frames") for calls of many blocks.
//
// The reference's per-bin work is itself a convolution along the block index n:
//     Y[n][k] = sum_p H[p][k] * X[n-p][k]        (uniform_partitioned_convolver.hpp:55-61 with fdl_index.hpp:28-31)
// A call of T blocks evaluates T outputs of that P-tap FIR for every bin. The Toeplitz kernels of conv_kernels.cuh do it directly
// (T*P complex FMAs per bin: FP32-bound at T >= 16). Here the same sum is evaluated by partitioned overlap-save ALONG n, which is
// the reference's own algorithm applied one level up: a frame = T consecutive blocks, L = 2T,
//     F[s][f][k]  = FFT_L over n of the spectra of frames (s-1, s)                        (frame spectrum, pushed into a ring)
//     G[q][f][k]  = FFT_L over m of ( H[qT + m][k], m < T ; 0, T <= m < L )              (filter, prepared once)
//     A[f][k]     = sum_q G[q][f][k] * F[s - q][f][k]                                     (Q = ceil(P/T) streamed rows)
//     Y[sT+t][k]  = (1/L) IFFT_L(A)[T + t]                                                (keep the last T: overlap-save)
// The MAC is then the T = 1 streaming kernel over Q rows of L*B bins: every byte used once, 2 * 2P*B complex per frame instead of
// T*P*B complex FMAs -- HBM-bound instead of FP32-bound. Results differ from the direct form only by rounding (tests: rel-L2 vs
// the reference block-by-block convolver <= 1e-5 float, <= 1e-12 double).
//
// Packed bin 0: a level-1 row stores (Re X[0], Re X[B]) in element 0. Both are REAL sequences along n, but their convolutions must
// not mix, so element 0 keeps only X[0] (imaginary part dropped on load) and the Nyquist sequence X[B][n] goes through the same
// transform into extra "Nyquist tiles" whose W columns are the L frame bins f. On the way back the Nyquist result is added as the
// imaginary part before the inverse frame transform (both results are real, so IFFT(A0 + i*AB) = Y0 + i*YB, the packed form).
//
// HBM layout (W = tile width, nt = B/W, tiles2 = nt*L + ceil(L/W)):
//   x1     [inputs][2T][B]                level-1 spectra of the previous and the current frame (two halves, ping-pong), row-major
//   fdl2   [inputs][tiles2][R2][W]        ring of frame spectra, R2 = ceil(partition_end / T); tile2 = tile*L + f, then Nyquist
//   filt2  [filters][tiles2][Q][W]        Q local second-level partitions
//   acc2   [S][outputs][tiles2][W]        MAC result (S partial planes, folded into plane 0)
// i.e. exactly the tile-major layout of conv_kernels.cuh with B' = tiles2*W bins per row, so fdl_mac_stream_kernel runs unchanged.
#pragma once

#include "conv_kernels.cuh"
#include "fft_wide.cuh"

#include <cstdlib>

// MAC staging of the 32-points-per-thread fused frame kernel: chunk = E / CHDIV points, NP stages outside the exchange tile (the others
// lie inside it). Both shapes that leave two CTAs resident at L = 512 were measured (C5 geometry, ms per fused step at T=256 Q=4 / T=256 Q=2 /
// T=512 Q=2): 4-point chunks, 2 + 4 stages: 7.29 / 4.44 / 9.09 (shipped); 8-point chunks, 1 + 2 stages (EXTRA="-DNEO_B200_FRAME32_CHDIV=4
// -DNEO_B200_FRAME32_NP=1"): 7.48 / 4.49 / 9.02 -- the chunk size is not what this form waits for.
#ifndef NEO_B200_FRAME32_CHDIV
#define NEO_B200_FRAME32_CHDIV 8
#endif
#ifndef NEO_B200_FRAME32_NP
#define NEO_B200_FRAME32_NP 2
#endif

namespace neo_b200 {

// bins (independent sequences) per CTA: enough adjacent bins to fill a 128-byte segment, and at least 128 threads
template<typename T, int LOGL>
constexpr int frame_default_logg()
{
    int const tn = (1 << LOGL) >> pick_loge<T>(LOGL);
    int g0       = sizeof(T) == 4 ? 4 : 3;
    while ((tn << g0) < 128) { ++g0; }
    return g0;
}

// LOGG: log2 bins per CTA (-1 = default); LOGE_F: log2 points per thread (-1 = pick_loge)
template<typename T, int LOGL, int LOGG = -1, int LOGE_F = -1>
struct frame_cfg
{
    using F                      = cta_fft<T, LOGL, -1, LOGE_F>;
    static constexpr int E       = F::E;
    static constexpr int TN      = F::TN;
    static constexpr int L       = F::M;
    static constexpr int G       = 1 << (LOGG >= 0 ? LOGG : frame_default_logg<T, LOGL>());
    static constexpr int THREADS = TN * G;
    static constexpr size_t SMEM = size_t(G) * F::TILE * sizeof(cx<T>);
};

// One L-point transform along block time per unit (a unit = one bin of one channel, or one channel's Nyquist sequence).
// io.open(unit) -> per-sequence state; io.load(state, n) / io.store(state, f, value).
template<typename T, int LOGL, int DIR, typename IO, int LOGG = -1>
__global__ void __launch_bounds__(frame_cfg<T, LOGL, LOGG>::THREADS) frame_fft_kernel(IO io, cx<T> const* __restrict__ tw, size_t units)
{
    using cfg = frame_cfg<T, LOGL, LOGG>;
    using F   = cta_fft<T, LOGL, DIR>;
    using C   = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const g       = threadIdx.x % cfg::G;
    int const t       = threadIdx.x / cfg::G;
    C* sm             = reinterpret_cast<C*>(smem_raw) + g * F::TILE;
    size_t const unit = size_t(blockIdx.x) * cfg::G + g;
    bool const live   = unit < units;
    auto const seq    = io.open(live ? unit : units - 1);  // ragged tail: every thread still runs the barriers

    C v[cfg::E];
#pragma unroll
    for (int e = 0; e < cfg::E; ++e) { v[e] = io.load(seq, t + e * cfg::TN); }
    F::run(v, sm, tw, t);
    if (live) {
#pragma unroll
        for (int e = 0; e < cfg::E; ++e) { io.store(seq, t + e * cfg::TN, v[e]); }
    }
}

__device__ __forceinline__ void frame_cp_async16(void* smem_dst, void const* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void frame_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N>
__device__ __forceinline__ void frame_cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct frame_geom
{
    int logb;    // log2 B
    int logw;    // log2 W
    int nt;      // B / W
    int frame;   // T
    int tiles2;  // nt * L + ceil(L / W)
};

// ---- forward: spectra of frames (s-1, s) -> frame spectrum in ring slot `slot` ----------------------------------------------------
template<typename T, bool NYQ>
struct frame_fwd_io
{
    using C = cx<T>;
    C const* x1;
    C* fdl2;
    frame_geom g;
    int new_half;  // which half of x1 holds the current frame
    int ring2, slot;
    size_t chan0;

    struct state
    {
        C const* src;
        C* dst;
        bool edge;
    };
    __device__ __forceinline__ state open(size_t unit) const
    {
        int const L         = 2 * g.frame;
        size_t const chan   = chan0 + (NYQ ? unit : unit >> g.logb);
        int const k         = NYQ ? 0 : int(unit & ((size_t(1) << g.logb) - 1));
        int const tile      = k >> g.logw;
        int const w         = k & ((1 << g.logw) - 1);
        C const* const src  = x1 + ((chan * size_t(L)) << g.logb) + k;
        size_t const tile2  = NYQ ? size_t(g.nt) * L : size_t(tile) * L;
        C* const dst        = fdl2 + (((chan * g.tiles2 + tile2) * ring2 + slot) << g.logw) + (NYQ ? 0 : w);
        return {src, dst, k == 0};
    }
    __device__ __forceinline__ C load(state const& s, int n) const
    {
        int const half = n < g.frame ? (new_half ^ 1) : new_half;
        int const row  = half * g.frame + (n & (g.frame - 1));
        C v            = s.src[size_t(row) << g.logb];
        if constexpr (NYQ) { return mk<T>(v.y, T(0)); }
        if (s.edge) { v.y = T(0); }
        return v;
    }
    __device__ __forceinline__ void store(state const& s, int f, C v) const
    {
        if constexpr (NYQ) {
            s.dst[((size_t(f >> g.logw) * ring2) << g.logw) + (f & ((1 << g.logw) - 1))] = v;
        } else {
            s.dst[(size_t(f) * ring2) << g.logw] = v;
        }
    }
};

// ---- inverse: MAC result -> level-1 spectra of the T blocks of this frame ([outputs][T][B], what conv_c2r_io reads) ---------------
template<typename T>
struct frame_inv_io
{
    using C = cx<T>;
    C const* acc2;
    C* y1;
    frame_geom g;
    T scale;  // 1 / L
    size_t chan0;

    struct state
    {
        C const* src;
        C const* nyq;
        C* dst;
    };
    __device__ __forceinline__ state open(size_t unit) const
    {
        int const L        = 2 * g.frame;
        size_t const chan  = chan0 + (unit >> g.logb);
        int const k        = int(unit & ((size_t(1) << g.logb) - 1));
        int const tile     = k >> g.logw;
        int const w        = k & ((1 << g.logw) - 1);
        C const* const row = acc2 + ((chan * g.tiles2) << g.logw);
        return {row + ((size_t(tile) * L) << g.logw) + w, k == 0 ? row + ((size_t(g.nt) * L) << g.logw) : nullptr,
                y1 + ((chan * g.frame) << g.logb) + k};
    }
    __device__ __forceinline__ C load(state const& s, int f) const
    {
        C v = s.src[size_t(f) << g.logw];
        if (s.nyq != nullptr) {  // + i * (Nyquist result)
            C const n = s.nyq[f];
            v.x -= n.y;
            v.y += n.x;
        }
        return v;
    }
    __device__ __forceinline__ void store(state const& s, int n, C v) const
    {
        if (n >= g.frame) { s.dst[size_t(n - g.frame) << g.logb] = mk<T>(v.x * scale, v.y * scale); }
    }
};

// ---- filter: level-1 partitions [filters][nt][parts][W] -> second-level partitions [filters][tiles2][Q][W] -------------------------
template<typename T, bool NYQ>
struct frame_filter_io
{
    using C = cx<T>;
    C const* h1;
    C* h2;
    frame_geom g;
    int parts, q_count;

    struct state
    {
        C const* src;
        C* dst;
        int rows;  // level-1 partitions this second-level partition covers (<= T)
        bool edge;
    };
    __device__ __forceinline__ state open(size_t unit) const
    {
        int const L       = 2 * g.frame;
        size_t const r    = NYQ ? unit : unit >> g.logb;
        int const k       = NYQ ? 0 : int(unit & ((size_t(1) << g.logb) - 1));
        size_t const f    = r / size_t(q_count);
        int const q       = int(r - f * size_t(q_count));
        int const tile    = k >> g.logw;
        int const w       = k & ((1 << g.logw) - 1);
        int const p0      = q * g.frame;
        int const left    = parts - p0;
        C const* const src = h1 + (((f * g.nt + tile) * size_t(parts) + p0) << g.logw) + w;
        size_t const tile2 = NYQ ? size_t(g.nt) * L : size_t(tile) * L;
        C* const dst       = h2 + (((f * g.tiles2 + tile2) * q_count + q) << g.logw) + (NYQ ? 0 : w);
        return {src, dst, left < g.frame ? left : g.frame, k == 0};
    }
    __device__ __forceinline__ C load(state const& s, int n) const
    {
        if (n >= s.rows) { return mk<T>(T(0), T(0)); }
        C v = s.src[size_t(n) << g.logw];
        if constexpr (NYQ) { return mk<T>(v.y, T(0)); }
        if (s.edge) { v.y = T(0); }
        return v;
    }
    __device__ __forceinline__ void store(state const& s, int f, C v) const
    {
        if constexpr (NYQ) {
            s.dst[((size_t(f >> g.logw) * q_count) << g.logw) + (f & ((1 << g.logw) - 1))] = v;
        } else {
            s.dst[(size_t(f) * q_count) << g.logw] = v;
        }
    }
};

template<typename T, int LOGL, int DIR, typename IO, int LOGG = -1>
int launch_frame_fft_g(IO const& io, cx<T> const* tw, size_t units, cudaStream_t stream)
{
    using cfg = frame_cfg<T, LOGL, LOGG>;
    if (units == 0) { return NEO_B200_OK; }
    auto kernel = frame_fft_kernel<T, LOGL, DIR, IO, LOGG>;
    NEO_TRY(enable_smem(kernel, cfg::SMEM));
    size_t const ctas = (units + cfg::G - 1) / cfg::G;
    if (ctas > 0x7fffffffULL) { return fail(NEO_B200_ERR_UNSUPPORTED, "frame transform grid too large"); }
    kernel<<<unsigned(ctas), cfg::THREADS, cfg::SMEM, stream>>>(io, tw, units);
    return check_launch("frame_fft_kernel");
}

// tuning knobs, read once per handle (conv_engine::init) and passed down, so a handle never changes form after its tables are built
struct frame_knobs
{
    int variant{-1};   // NEO_B200_FRAME_VARIANT: fused-kernel geometry at L >= 256 (float)
    bool async{true};  // NEO_B200_FRAME_NO_ASYNC clears it: MAC operands straight into registers
    bool pipelined{true};  // NEO_B200_FRAME_NO_PIPELINE clears it: one CTA per unit instead of the persistent software-pipelined form
    static frame_knobs from_env()
    {
        frame_knobs k;
        if (char const* v = std::getenv("NEO_B200_FRAME_VARIANT")) { k.variant = std::atoi(v); }
        k.async = std::getenv("NEO_B200_FRAME_NO_ASYNC") == nullptr;
        k.pipelined = std::getenv("NEO_B200_FRAME_NO_PIPELINE") == nullptr;
        return k;
    }
    bool eight_points(int logl, bool is_f32) const { return is_f32 && logl >= 8 && (variant == 2 || variant == 3 || variant == 4); }
    // 32 points per thread (float, L = 512 / 1024): two register stages and ONE exchange per frame transform, 256 / 512 threads
    bool wide_points(int logl, bool is_f32) const { return is_f32 && logl >= 9 && logl <= 10 && (variant == 5 || variant < 0); }
};

template<typename T, int LOGL, int DIR, typename IO>
int launch_frame_fft(IO const& io, cx<T> const* tw, size_t units, cudaStream_t stream, frame_knobs const& knobs = {})
{
    if constexpr (LOGL >= 8) {
        if (knobs.variant == 1) { return launch_frame_fft_g<T, LOGL, DIR, IO, frame_default_logg<T, LOGL>() - 1>(io, tw, units, stream); }
    }
    return launch_frame_fft_g<T, LOGL, DIR, IO, -1>(io, tw, units, stream);
}

// LOGL = log2(2T): frames of 2 ... 512 blocks
#define NEO_DISPATCH_LOGL(logl, ...)                                         \
    switch (logl) {                                                          \
        case 2: { constexpr int LOGL = 2; __VA_ARGS__ } break;               \
        case 3: { constexpr int LOGL = 3; __VA_ARGS__ } break;               \
        case 4: { constexpr int LOGL = 4; __VA_ARGS__ } break;               \
        case 5: { constexpr int LOGL = 5; __VA_ARGS__ } break;               \
        case 6: { constexpr int LOGL = 6; __VA_ARGS__ } break;               \
        case 7: { constexpr int LOGL = 7; __VA_ARGS__ } break;               \
        case 8: { constexpr int LOGL = 8; __VA_ARGS__ } break;               \
        case 9: { constexpr int LOGL = 9; __VA_ARGS__ } break;               \
        case 10: { constexpr int LOGL = 10; __VA_ARGS__ } break;             \
        default: break;                                                      \
    }

constexpr size_t k_max_frame_blocks = 512;

// ---- fused frame step for banks ----------------------------------------------------------------------------------------------------
// forward frame transform -> ring insert -> MAC over the Q second-level partitions -> inverse frame transform, all in ONE kernel with
// the L values of a bin held in registers between the steps (a thread owns frame bins f = t + e*TN of its bin k before AND after
// cta_fft::run, so the MAC needs no exchange). Against the three-kernel form this never writes or re-reads the MAC result and never
// re-reads the frame spectrum it has just produced: with S = the level-1 spectra of one frame (T rows of B bins per channel), a frame
// step moves 5 S (two frames read, the new ring slot written, T result rows written) + filter + older ring slots, instead of 11 S +
// filter + the whole ring. Diagonal topology, unsplit partition loop (sources == 1, splits == 1).
template<typename T, bool NYQ>
struct frame_fused_io
{
    using C = cx<T>;
    C const* x1;
    C* fdl2;
    C const* filt2;
    C* y1;        // main: [outputs][T][B]
    C* nyq_acc;   // [outputs][L]: written by the Nyquist launch, consumed by bin 0 of the main launch
    frame_geom g;
    int new_half;
    int ring2, slot;   // ring size and the slot this frame is written to
    int parts2, age0;  // local second-level partitions and the age of the first one
    T scale;           // 1 / L
    size_t chan0;
    // multi-device bank, push form: the T result rows of channel c go to the device that finishes c (owner = c / own_count), straight
    // into that device's inbox through a peer-mapped pointer (st.global over NVLink) -- the exchange of the partial spectra rides on
    // the stores of the kernel that produces them. owners <= 1: everything goes to y1.
    C* y1_owner[k_bank_max_shards];
    int owners, own_count;
};

// REGCAP: registers per thread the kernel is held to (resident CTAs = 65536 / REGCAP / threads, at least one)
template<typename T, int LOGL, int LOGG, int LOGE_F, int REGCAP>
constexpr int frame_min_ctas()
{
    int const n = 65536 / REGCAP / frame_cfg<T, LOGL, LOGG, LOGE_F>::THREADS;
    return n < 1 ? 1 : n;
}

// ASYNC: the MAC operands are staged through shared memory with cp.async, three chunks deep (two private stages plus the
// transform's exchange tile, idle during the MAC), so the bytes in flight per SM no longer depend on registers: the first two
// chunks are already on their way while the frame's spectra are loaded and transformed. Needs whole 128-byte groups of bins
// (tile width >= bins per CTA, no ragged tail) -- the launcher checks.
template<typename T, int LOGL, int LOGG, int LOGE_F>
struct frame_async_cfg
{
    using cfg                      = frame_cfg<T, LOGL, LOGG, LOGE_F>;
    // points per thread and chunk: half a thread's points; an eighth with 32 points per thread
    static constexpr int CH        = cfg::E >= 32 ? cfg::E / NEO_B200_FRAME32_CHDIV : cfg::E >= 2 ? cfg::E / 2 : 1;
    // stages outside the exchange tile (they can be filled while the forward transform runs): two. With 32 points per thread tile +
    // private stages must stay under half an SM's shared memory, so that two CTAs are resident
    static constexpr int NP        = cfg::E >= 32 ? NEO_B200_FRAME32_NP : 2;
    static constexpr int NH        = cfg::E / CH;                                                  // chunks per partition
    static constexpr int OPERAND   = CH * cfg::TN * cfg::G * int(sizeof(cx<T>));                   // bytes of one operand of a chunk
    static constexpr int STAGE     = 2 * OPERAND;                                                  // <= the exchange tile
    static constexpr int NS        = NP + int(cfg::SMEM) / STAGE;                                  // stages: the private ones + those the idle tile holds
    static constexpr size_t SMEM   = cfg::SMEM + NP * size_t(STAGE);
    static constexpr int PER_THREAD = OPERAND / 16 / cfg::THREADS;                                 // 16-byte pieces per thread
    static constexpr int ROW_PIECES = cfg::G * int(sizeof(cx<T>)) / 16;                            // pieces per group of bins
    static_assert(STAGE <= int(cfg::SMEM), "a stage must fit the exchange tile");
    static_assert(PER_THREAD * 16 * cfg::THREADS == OPERAND, "whole pieces per thread");
};

template<typename T, int LOGL, int LOGG, int LOGE_F, int REGCAP, bool NYQ, bool ASYNC = false>
__global__ void __launch_bounds__(frame_cfg<T, LOGL, LOGG, LOGE_F>::THREADS, frame_min_ctas<T, LOGL, LOGG, LOGE_F, REGCAP>())
    frame_fused_kernel(frame_fused_io<T, NYQ> io, cx<T> const* __restrict__ tw, size_t units)
{
    using cfg = frame_cfg<T, LOGL, LOGG, LOGE_F>;
    using C   = cx<T>;
    constexpr int E = cfg::E, TN = cfg::TN, L = cfg::L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const gi      = threadIdx.x % cfg::G;
    int const t       = threadIdx.x / cfg::G;
    C* sm             = reinterpret_cast<C*>(smem_raw) + gi * cfg::F::TILE;
    size_t const unit0 = size_t(blockIdx.x) * cfg::G + gi;
    bool const live   = unit0 < units;
    size_t const unit = live ? unit0 : units - 1;

    frame_geom const& g = io.g;
    int const logw      = g.logw;
    int const wmask     = (1 << logw) - 1;
    size_t const chan   = io.chan0 + (NYQ ? unit : unit >> g.logb);
    int const k         = NYQ ? 0 : int(unit & ((size_t(1) << g.logb) - 1));
    int const tile      = k >> logw;
    int const w         = k & wmask;
    size_t const tile2  = NYQ ? size_t(g.nt) * L : size_t(tile) * L;
    C const* const src  = io.x1 + ((chan * size_t(L)) << g.logb) + k;
    C* const ring       = io.fdl2 + (((chan * g.tiles2 + tile2) * io.ring2) << logw) + (NYQ ? 0 : w);
    C const* const filt = io.filt2 + (((chan * g.tiles2 + tile2) * io.parts2) << logw) + (NYQ ? 0 : w);
    // element (frame bin f, row r) of a [..][rows][W] run of tiles
    auto const at = [&](int f, int r, int rows) -> size_t {
        if constexpr (NYQ) { return ((size_t(f >> logw) * rows + r) << logw) + size_t(f & wmask); }
        return (size_t(f) * rows + r) << logw;
    };

    int slot0 = (io.slot - io.age0) % io.ring2;  // ring slot paired with local partition 0
    slot0 += slot0 < 0 ? io.ring2 : 0;
    bool const newest_in_regs = io.age0 == 0;  // partition 0 pairs with the frame spectrum this kernel has just computed

    // ---- ASYNC: chunk c = (partition q = c / NH, half eh = c % NH of the thread's points); stage = c % 3 ----
    using ac = frame_async_cfg<T, LOGL, LOGG, LOGE_F>;
    auto const stage_ptr = [&](int c) -> unsigned char* {  // stages NP .. NS-1 lie in the exchange tile
        int const st = c % ac::NS;
        return st >= ac::NP ? smem_raw + (st - ac::NP) * ac::STAGE : smem_raw + cfg::SMEM + st * ac::STAGE;
    };
    int const nchunks = io.parts2 * ac::NH;
    // Warp-private staging: a warp copies exactly the rows its own threads will read (its TW = 32/G values of t, all CH points u of
    // the chunk), so the MAC loop needs __syncwarp() only. Lane -> (row, 16-byte piece) is fixed; a chunk moves every pointer by a
    // constant, so the per-copy address work is one multiply-add.
    constexpr int TW    = 32 / cfg::G;                  // values of t per warp
    constexpr int RP    = ac::ROW_PIECES;               // 16-byte pieces per row of G bins
    constexpr int KSTEP = 16 / int(sizeof(C));          // u advances by KSTEP per copy of a lane
    static_assert(!ASYNC || (cfg::G <= 32 && (32 / RP) == KSTEP * TW), "warp-private staging geometry");
    int const lane      = int(threadIdx.x) & 31;
    int const lrow      = lane / RP;
    int const u0        = lrow / (TW > 0 ? TW : 1);
    int const t_mine    = (int(threadIdx.x) >> 5) * TW + lrow % (TW > 0 ? TW : 1);  // the t whose row this lane copies
    int const f0        = t_mine + u0 * TN;              // frame bin of the lane's first copy of a chunk with eh = 0
    size_t const eo     = size_t(lane % RP) * KSTEP;     // element offset inside the group of bins
    C const* const h0   = filt - gi + eo + at(f0, 0, io.parts2);
    C const* const x0   = ring - gi + eo + at(f0, 0, io.ring2);
    size_t const h_step = at(KSTEP * TN, 0, io.parts2);  // elements between a lane's consecutive copies
    size_t const x_step = at(KSTEP * TN, 0, io.ring2);
    size_t const h_half = at(ac::CH * TN, 0, io.parts2); // elements between the halves eh of a partition
    size_t const x_half = at(ac::CH * TN, 0, io.ring2);
    unsigned const dst0 = unsigned((u0 * TN + t_mine) * cfg::G * int(sizeof(C)) + (lane % RP) * 16);  // byte offset inside an operand
    constexpr unsigned dst_step = unsigned(KSTEP * TN * cfg::G * int(sizeof(C)));
    auto const issue = [&](int c) {
        if constexpr (ASYNC) {
            if (c < nchunks) {
                int const q  = c / ac::NH;
                int const eh = c - q * ac::NH;
                int sl       = slot0 - q;
                sl += sl < 0 ? io.ring2 : 0;
                unsigned char* const dst_h = stage_ptr(c) + dst0;
                C const* const hp          = h0 + size_t(eh) * h_half + (size_t(q) << logw);
                C const* const xp          = x0 + size_t(eh) * x_half + (size_t(sl) << logw);
                bool const want_x          = !(q == 0 && newest_in_regs);
#pragma unroll
                for (int i = 0; i < ac::PER_THREAD; ++i) {
                    frame_cp_async16(dst_h + i * dst_step, hp + i * h_step);
                    if (want_x) { frame_cp_async16(dst_h + ac::OPERAND + i * dst_step, xp + i * x_step); }
                }
            }
            frame_cp_async_commit();
        }
    };
    // the frame's own spectra first (they head the dependency chain), then the first two MAC chunks behind them in the queue
    C v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        // n = t + e * TN; the first E/2 points are the previous frame (n < T), the rest the current one
        int const half = e < E / 2 ? (io.new_half ^ 1) : io.new_half;
        int const row  = half * g.frame + t + (e % (E / 2)) * TN;
        v[e]           = src[size_t(row) << g.logb];
    }
    if constexpr (ASYNC) {
#pragma unroll
        for (int c = 0; c < ac::NP; ++c) { issue(c); }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if constexpr (NYQ) { v[e] = mk<T>(v[e].y, T(0)); }
        else if (k == 0) { v[e].y = T(0); }
    }
    cta_fft<T, LOGL, -1, LOGE_F>::run(v, sm, tw, t);
    if constexpr (ASYNC) {  // the exchange tile is free until the inverse transform
#pragma unroll
        for (int c = ac::NP; c < ac::NS; ++c) { issue(c); }
    }
    if (live) {
#pragma unroll
        for (int e = 0; e < E; ++e) { ring[at(t + e * TN, io.slot, io.ring2)] = v[e]; }
    }

    // MAC over the local partitions q (row q of the filter pairs with the frame spectrum of age0 + q frames ago)
    if constexpr (ASYNC && E >= 32) {
        // 32 points per thread: the sum is formed IN PLACE (partition 0 overwrites the frame spectrum it has just consumed), so one
        // array of E values is live, not two, and the kernel keeps to 128 registers. Same products in the same order as below.
        for (int q = 0; q < io.parts2; ++q) {
            bool const first = q == 0;
            bool const own   = first && newest_in_regs;
#pragma unroll
            for (int eh = 0; eh < ac::NH; ++eh) {
                int const c = q * ac::NH + eh;
                frame_cp_async_wait<ac::NS - 1>();
                __syncwarp();
                C const* const sh = reinterpret_cast<C const*>(stage_ptr(c));
                C const* const sx = reinterpret_cast<C const*>(stage_ptr(c) + ac::OPERAND);
#pragma unroll
                for (int u = 0; u < ac::CH; ++u) {
                    int const e = eh * ac::CH + u;
                    C const h   = sh[(u * TN + t) * cfg::G + gi];
                    C const x   = own ? v[e] : sx[(u * TN + t) * cfg::G + gi];
                    C acc       = first ? mk<T>(T(0), T(0)) : v[e];
                    acc.x       = ::fma(x.x, h.x, acc.x);
                    acc.x       = ::fma(-x.y, h.y, acc.x);
                    acc.y       = ::fma(x.x, h.y, acc.y);
                    acc.y       = ::fma(x.y, h.x, acc.y);
                    v[e]        = acc;
                }
                __syncwarp();
                issue(c + ac::NS);
            }
        }
        frame_cp_async_wait<0>();
        __syncthreads();  // stage 2 is the exchange tile the inverse transform is about to write
    } else if constexpr (ASYNC) {
        C a[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { a[e] = mk<T>(T(0), T(0)); }
        for (int q = 0; q < io.parts2; ++q) {
            bool const own = q == 0 && newest_in_regs;
#pragma unroll
            for (int eh = 0; eh < ac::NH; ++eh) {
                int const c = q * ac::NH + eh;
                frame_cp_async_wait<ac::NS - 1>();  // chunks are committed in order, one group each: chunk c has landed
                __syncwarp();              // ... for every lane of this warp, which copied all the rows the warp reads
                C const* const sh = reinterpret_cast<C const*>(stage_ptr(c));
                C const* const sx = reinterpret_cast<C const*>(stage_ptr(c) + ac::OPERAND);
#pragma unroll
                for (int u = 0; u < ac::CH; ++u) {
                    int const e = eh * ac::CH + u;
                    C const h   = sh[(u * TN + t) * cfg::G + gi];
                    C const x   = own ? v[e] : sx[(u * TN + t) * cfg::G + gi];
                    a[e].x      = ::fma(x.x, h.x, a[e].x);
                    a[e].x      = ::fma(-x.y, h.y, a[e].x);
                    a[e].y      = ::fma(x.x, h.y, a[e].y);
                    a[e].y      = ::fma(x.y, h.x, a[e].y);
                }
                __syncwarp();  // every lane is done with this warp's rows of the stage
                issue(c + ac::NS);
            }
        }
        frame_cp_async_wait<0>();
        __syncthreads();  // stage 2 is the exchange tile the inverse transform is about to write
#pragma unroll
        for (int e = 0; e < E; ++e) { v[e] = a[e]; }
    } else {
        int q = 0;
        if (newest_in_regs) {  // the newest frame spectrum is still in registers
#pragma unroll
            for (int e = 0; e < E; ++e) { v[e] = cmul(v[e], filt[at(t + e * TN, 0, io.parts2)]); }
            q = 1;
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) { v[e] = mk<T>(T(0), T(0)); }
        }
        constexpr int CH = E < 8 ? E : 8;
        for (; q < io.parts2; ++q) {
            int sl = slot0 - q;
            sl += sl < 0 ? io.ring2 : 0;
#pragma unroll
            for (int e0 = 0; e0 < E; e0 += CH) {
                C x[CH], h[CH];
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    int const f = t + (e0 + u) * TN;
                    h[u]        = __ldcs(filt + at(f, q, io.parts2));
                    x[u]        = __ldcs(ring + at(f, sl, io.ring2));
                }
#pragma unroll
                for (int u = 0; u < CH; ++u) {
                    v[e0 + u].x = ::fma(x[u].x, h[u].x, v[e0 + u].x);
                    v[e0 + u].x = ::fma(-x[u].y, h[u].y, v[e0 + u].x);
                    v[e0 + u].y = ::fma(x[u].x, h[u].y, v[e0 + u].y);
                    v[e0 + u].y = ::fma(x[u].y, h[u].x, v[e0 + u].y);
                }
            }
        }
    }

    if constexpr (NYQ) {
        if (live) {
#pragma unroll
            for (int e = 0; e < E; ++e) { io.nyq_acc[chan * L + t + e * TN] = v[e]; }
        }
    } else {
        if (k == 0) {  // + i * (Nyquist result): the inverse transform then yields the packed pair (Re Y[0], Re Y[B])
#pragma unroll
            for (int e = 0; e < E; ++e) {
                C const n = io.nyq_acc[chan * L + t + e * TN];
                v[e].x -= n.y;
                v[e].y += n.x;
            }
        }
        cta_fft<T, LOGL, 1, LOGE_F>::run(v, sm, tw, t);
        if (live) {
            C* dst = io.y1 + ((chan * g.frame) << g.logb) + k;
            if (io.owners > 1) {
                int const o = int(chan) / io.own_count;
                dst         = io.y1_owner[o] + (((chan - size_t(o) * io.own_count) * g.frame) << g.logb) + k;
            }
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int const n = t + e * TN;
                if (n >= g.frame) { dst[size_t(n - g.frame) << g.logb] = mk<T>(v[e].x * io.scale, v[e].y * io.scale); }
            }
        }
    }
}

// ---- the same frame step as a PERSISTENT, software-pipelined CTA -------------------------------------------------------------------
// ncu on frame_fused_kernel with few second-level partitions (Q = 2: one partition shard of a two-way split, or T = 512 on one device):
// 48 % issue-active, 55 % of DRAM peak -- per unit (G bins of one channel) the two frame transforms need ~10.8k issue cycles and the
// unit's bytes ~15.3k cycles of HBM time, and with one 512-thread CTA per SM (128 registers x 512 threads, ~200 KB of shared memory)
// they run one after the other. Here one CTA per SM walks units u = blockIdx.x, blockIdx.x + gridDim.x, ... and keeps HBM busy
// during the transforms:
//   - the two frames' level-1 spectra of unit u+1 are fetched with cp.async into a staging buffer while unit u's forward transform
//     runs (they used to be the first thing a CTA waited for);
//   - the first two MAC chunks of unit u+1 are fetched while unit u's inverse transform runs.
// Same arithmetic, same operand order, same results as frame_fused_kernel. MAC chunks are a quarter of a thread's points (CHDIV = 4)
// so that tile + two stages + the spectra staging buffer fit 227 KB.
template<typename T, int LOGL, int LOGG, int LOGE_F>
struct frame_pipe_cfg
{
    using cfg                        = frame_cfg<T, LOGL, LOGG, LOGE_F>;
    static constexpr int CH          = cfg::E >= 4 ? cfg::E / 4 : 1;
    static constexpr int NH          = cfg::E / CH;
    static constexpr int OPERAND     = CH * cfg::TN * cfg::G * int(sizeof(cx<T>));
    static constexpr int STAGE       = 2 * OPERAND;
    static constexpr int XSTAGE      = cfg::L * cfg::G * int(sizeof(cx<T>));                          // 2T rows of G bins
    static constexpr size_t SMEM     = cfg::SMEM + 2 * size_t(STAGE) + size_t(XSTAGE);
    static constexpr int PER_THREAD  = OPERAND / 16 / cfg::THREADS;
    static constexpr int ROW_PIECES  = cfg::G * int(sizeof(cx<T>)) / 16;
    static constexpr int X_PER_THREAD = XSTAGE / 16 / cfg::THREADS;
    static constexpr bool OK = STAGE <= int(cfg::SMEM) && PER_THREAD >= 1 && PER_THREAD * 16 * cfg::THREADS == OPERAND
                            && X_PER_THREAD * 16 * cfg::THREADS == XSTAGE && SMEM <= 227 * 1024 && cfg::THREADS <= 1024;
};

template<typename T, int LOGL, int LOGG, int LOGE_F, int REGCAP>
__global__ void __launch_bounds__(frame_cfg<T, LOGL, LOGG, LOGE_F>::THREADS, frame_min_ctas<T, LOGL, LOGG, LOGE_F, REGCAP>())
    frame_fused_pipelined_kernel(frame_fused_io<T, false> io, cx<T> const* __restrict__ tw, size_t units)
{
    using cfg = frame_cfg<T, LOGL, LOGG, LOGE_F>;
    using pc  = frame_pipe_cfg<T, LOGL, LOGG, LOGE_F>;
    using C   = cx<T>;
    constexpr int E = cfg::E, TN = cfg::TN, L = cfg::L, G = cfg::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int const gi = threadIdx.x % G;
    int const t  = threadIdx.x / G;
    C* const sm  = reinterpret_cast<C*>(smem_raw) + gi * cfg::F::TILE;
    unsigned char* const xstage = smem_raw + cfg::SMEM + 2 * pc::STAGE;

    frame_geom const& g = io.g;
    int const logw      = g.logw;
    int const wmask     = (1 << logw) - 1;
    auto const at       = [&](int f, int r, int rows) -> size_t { return (size_t(f) * rows + r) << logw; };
    int slot0           = (io.slot - io.age0) % io.ring2;  // ring slot paired with local partition 0
    slot0 += slot0 < 0 ? io.ring2 : 0;
    bool const newest_in_regs = io.age0 == 0;
    int const nchunks         = io.parts2 * pc::NH;
    auto const stage_ptr      = [&](int c) -> unsigned char* {  // stage 2 is the exchange tile
        int const st = c % 3;
        return st == 2 ? smem_raw : smem_raw + cfg::SMEM + st * pc::STAGE;
    };

    // what a unit (G adjacent bins of one channel; the first of them is `unit`) addresses
    struct where
    {
        size_t chan;
        int k0;          // first bin of the group
        C const* src;    // x1 row 0, bin k0
        C* ring;         // ring element (frame bin 0, slot 0), bin k0
        C const* filt;   // filter element (frame bin 0, partition 0), bin k0
    };
    auto const locate = [&](size_t unit) -> where {
        size_t const chan  = io.chan0 + (unit >> g.logb);
        int const k0       = int(unit & ((size_t(1) << g.logb) - 1));
        size_t const tile2 = size_t(k0 >> logw) * L;
        int const w0       = k0 & wmask;
        return {chan, k0, io.x1 + ((chan * size_t(L)) << g.logb) + k0, io.fdl2 + (((chan * g.tiles2 + tile2) * io.ring2) << logw) + w0,
                io.filt2 + (((chan * g.tiles2 + tile2) * io.parts2) << logw) + w0};
    };

    // warp-private MAC staging, exactly as in frame_fused_kernel (a warp copies the rows its own threads read)
    constexpr int TW    = 32 / G;
    constexpr int RP    = pc::ROW_PIECES;
    constexpr int KSTEP = 16 / int(sizeof(C));
    static_assert(G <= 32 && (32 / RP) == KSTEP * TW, "warp-private staging geometry");
    int const lane      = int(threadIdx.x) & 31;
    int const lrow      = lane / RP;
    int const u0        = lrow / TW;
    int const t_mine    = (int(threadIdx.x) >> 5) * TW + lrow % TW;
    int const f0        = t_mine + u0 * TN;
    size_t const eo     = size_t(lane % RP) * KSTEP;
    unsigned const dst0 = unsigned((u0 * TN + t_mine) * G * int(sizeof(C)) + (lane % RP) * 16);
    constexpr unsigned dst_step = unsigned(KSTEP * TN * G * int(sizeof(C)));
    auto const issue = [&](where const& u, int c) {
        if (c < nchunks) {
            int const q  = c / pc::NH;
            int const eh = c - q * pc::NH;
            int sl       = slot0 - q;
            sl += sl < 0 ? io.ring2 : 0;
            unsigned char* const dst_h = stage_ptr(c) + dst0;
            C const* const hp          = u.filt + eo + at(f0 + eh * pc::CH * TN, q, io.parts2);
            C const* const xp          = u.ring + eo + at(f0 + eh * pc::CH * TN, sl, io.ring2);
            bool const want_x          = !(q == 0 && newest_in_regs);
#pragma unroll
            for (int i = 0; i < pc::PER_THREAD; ++i) {
                frame_cp_async16(dst_h + i * dst_step, hp + at(i * KSTEP * TN, 0, io.parts2));
                if (want_x) { frame_cp_async16(dst_h + pc::OPERAND + i * dst_step, xp + at(i * KSTEP * TN, 0, io.ring2)); }
            }
        }
        frame_cp_async_commit();
    };
    // the two frames' level-1 spectra of a unit -> staging [frame bin n][G bins]; piece p = (row n, 16 bytes of the row's G bins)
    auto const fetch_x1 = [&](where const& u) {
#pragma unroll
        for (int i = 0; i < pc::X_PER_THREAD; ++i) {
            int const p    = int(threadIdx.x) + i * cfg::THREADS;
            int const n    = p / RP;
            int const half = n < g.frame ? (io.new_half ^ 1) : io.new_half;
            int const row  = half * g.frame + (n & (g.frame - 1));
            frame_cp_async16(xstage + size_t(p) * 16, u.src + (size_t(row) << g.logb) + size_t(p % RP) * KSTEP);
        }
        frame_cp_async_commit();
    };

    size_t const stride = size_t(gridDim.x) * G;
    size_t unit         = size_t(blockIdx.x) * G;
    if (unit >= units) { return; }
    where cur = locate(unit);
    fetch_x1(cur);
    issue(cur, 0);
    issue(cur, 1);
    for (; unit < units; unit += stride) {
        bool const more  = unit + stride < units;
        where const next = more ? locate(unit + stride) : cur;
        int const k      = cur.k0 + gi;

        frame_cp_async_wait<0>();  // this unit's spectra and its first two MAC chunks have landed
        __syncthreads();
        C v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { v[e] = reinterpret_cast<C const*>(xstage)[(t + e * TN) * G + gi]; }
        __syncthreads();           // the staging buffer is free: the next unit's spectra travel during this unit's forward transform
        if (more) { fetch_x1(next); }
        else { frame_cp_async_commit(); }
        if (k == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) { v[e].y = T(0); }
        }
        cta_fft<T, LOGL, -1, LOGE_F>::run(v, sm, tw, t);
        issue(cur, 2);  // the exchange tile is free until the inverse transform
        C* const ring = cur.ring + gi;
#pragma unroll
        for (int e = 0; e < E; ++e) { ring[at(t + e * TN, io.slot, io.ring2)] = v[e]; }

        C a[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { a[e] = mk<T>(T(0), T(0)); }
        for (int q = 0; q < io.parts2; ++q) {
            bool const own = q == 0 && newest_in_regs;
#pragma unroll
            for (int eh = 0; eh < pc::NH; ++eh) {
                int const c = q * pc::NH + eh;
                frame_cp_async_wait<2>();  // groups complete in order: chunk c (and everything older) has landed
                __syncwarp();
                C const* const sh = reinterpret_cast<C const*>(stage_ptr(c));
                C const* const sx = reinterpret_cast<C const*>(stage_ptr(c) + pc::OPERAND);
#pragma unroll
                for (int u = 0; u < pc::CH; ++u) {
                    int const e = eh * pc::CH + u;
                    C const h   = sh[(u * TN + t) * G + gi];
                    C const x   = own ? v[e] : sx[(u * TN + t) * G + gi];
                    a[e].x      = ::fma(x.x, h.x, a[e].x);
                    a[e].x      = ::fma(-x.y, h.y, a[e].x);
                    a[e].y      = ::fma(x.x, h.y, a[e].y);
                    a[e].y      = ::fma(x.y, h.x, a[e].y);
                }
                __syncwarp();
                issue(cur, c + 3);
            }
        }
        frame_cp_async_wait<0>();
        __syncthreads();  // every warp is done with the stages; stage 2 is the exchange tile the inverse transform writes
        if (more) {       // the next unit's first two chunks travel during this unit's inverse transform
            issue(next, 0);
            issue(next, 1);
        }
        if (k == 0) {  // + i * (Nyquist result): the inverse transform then yields the packed pair (Re Y[0], Re Y[B])
#pragma unroll
            for (int e = 0; e < E; ++e) {
                C const n = io.nyq_acc[cur.chan * L + t + e * TN];
                a[e].x -= n.y;
                a[e].y += n.x;
            }
        }
        cta_fft<T, LOGL, 1, LOGE_F>::run(a, sm, tw, t);
        C* dst = io.y1 + ((cur.chan * g.frame) << g.logb) + k;
        if (io.owners > 1) {
            int const o = int(cur.chan) / io.own_count;
            dst         = io.y1_owner[o] + (((cur.chan - size_t(o) * io.own_count) * g.frame) << g.logb) + k;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int const n = t + e * TN;
            if (n >= g.frame) { dst[size_t(n - g.frame) << g.logb] = mk<T>(a[e].x * io.scale, a[e].y * io.scale); }
        }
        cur = next;
    }
}

// PIPE: the persistent software-pipelined form may be taken (instantiated for the default geometries only)
template<typename T, int LOGL, int LOGG, int LOGE_F, int REGCAP, bool NYQ, bool PIPE = false>
int launch_frame_fused_g(frame_fused_io<T, NYQ> const& io, cx<T> const* tw, size_t units, cudaStream_t stream, bool async, bool pipelined = false)
{
    using cfg = frame_cfg<T, LOGL, LOGG, LOGE_F>;
    if constexpr (cfg::THREADS > 1024) { return fail(NEO_B200_ERR_UNSUPPORTED, "frame kernel variant needs %d threads", cfg::THREADS); }
    else {
        if (units == 0) { return NEO_B200_OK; }
        if constexpr (PIPE && !NYQ && LOGL >= 7 && sizeof(T) == 4) {
            using pc = frame_pipe_cfg<T, LOGL, LOGG, LOGE_F>;
            if constexpr (pc::OK) {
                // measured (C5 geometry, ms per fused step, pipelined / one CTA per unit): T=256 Q=4 7.26 / 6.57, T=128 Q=8 6.33 / 5.91,
                // T=256 Q=2 5.05 / 5.21, T=512 Q=2 11.07 / 12.47 -- the smaller MAC chunks cost more than the prefetch buys while the
                // MAC stream dominates; with one or two partitions per unit the transforms dominate and the prefetch pays
                if (async && pipelined && io.parts2 <= 2 && (1 << io.g.logw) >= cfg::G && units % cfg::G == 0) {
                    auto kernel = frame_fused_pipelined_kernel<T, LOGL, LOGG, LOGE_F, REGCAP>;
                    NEO_TRY(enable_smem(kernel, pc::SMEM));
                    int dev = 0, sms = 148, per_sm = 1;
                    cudaGetDevice(&dev);
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                    NEO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, cfg::THREADS, pc::SMEM));
                    size_t const resident = size_t(sms) * size_t(std::max(1, per_sm));  // persistent: one CTA per resident slot
                    kernel<<<unsigned(std::min(units / cfg::G, resident)), cfg::THREADS, pc::SMEM, stream>>>(io, tw, units);
                    return check_launch("frame_fused_pipelined_kernel");
                }
            }
        }
        if constexpr (!NYQ && LOGL >= 6) {
            using ac = frame_async_cfg<T, LOGL, LOGG, LOGE_F>;
            if (async && (1 << io.g.logw) >= cfg::G && units % cfg::G == 0 && ac::SMEM <= 227 * 1024) {
                auto kernel = frame_fused_kernel<T, LOGL, LOGG, LOGE_F, REGCAP, NYQ, true>;
                NEO_TRY(enable_smem(kernel, ac::SMEM));
                kernel<<<unsigned(units / cfg::G), cfg::THREADS, ac::SMEM, stream>>>(io, tw, units);
                return check_launch("frame_fused_kernel");
            }
        }
        auto kernel = frame_fused_kernel<T, LOGL, LOGG, LOGE_F, REGCAP, NYQ>;
        NEO_TRY(enable_smem(kernel, cfg::SMEM));
        size_t const ctas = (units + cfg::G - 1) / cfg::G;
        if (ctas > 0x7fffffffULL) { return fail(NEO_B200_ERR_UNSUPPORTED, "frame grid too large"); }
        kernel<<<unsigned(ctas), cfg::THREADS, cfg::SMEM, stream>>>(io, tw, units);
        return check_launch("frame_fused_kernel");
    }
}

// tw: stage twiddles of the transform's default points per thread; tw8: of the 8-points-per-thread variants (only when knobs ask);
// tw32: of the 32-points-per-thread form (float, L = 512 / 1024)
template<typename T, int LOGL, bool NYQ>
int launch_frame_fused(frame_fused_io<T, NYQ> const& io, cx<T> const* tw, cx<T> const* tw8, cx<T> const* tw32, size_t units,
                       cudaStream_t stream, frame_knobs const& knobs)
{
    constexpr int g0 = frame_default_logg<T, LOGL>();
    bool const a     = knobs.async;
    if constexpr (sizeof(T) == 4 && LOGL >= 9 && LOGL <= 10) {
        // 32 points per thread: two register stages and ONE exchange per frame transform, the sum over the partitions formed in place,
        // 16 bins x L/32 threads per CTA -- two resident CTAs at L = 512, so one unit's transforms run under the other's MAC stream.
        // measured (C5 geometry, ms per fused step, this form / 16 points per thread): T=256 Q=2 4.44 / 5.05, T=512 Q=2 9.09 / 11.1,
        // T=256 Q=4 7.29 / 6.53 (its MAC chunks are a quarter of the size: the longer the MAC stream, the more that costs)
        bool const wins = LOGL >= 10 || io.parts2 <= 2;
        if (knobs.variant == 5 || (knobs.variant < 0 && a && wins && tw32 != nullptr)) {
            return launch_frame_fused_g<T, LOGL, 4, 5, 128, NYQ>(io, tw32, units, stream, a);
        }
    }
    if constexpr (sizeof(T) == 4 && LOGL >= 8) {
        constexpr int g16 = LOGL >= 10 ? 3 : 4;  // 16 points per thread: 2^(LOGL-4) threads per bin, <= 512 threads per CTA
        switch (knobs.variant) {
            case 1: return launch_frame_fused_g<T, LOGL, g16 - 1, -1, 128, NYQ>(io, tw, units, stream, a);
            case 2: return launch_frame_fused_g<T, LOGL, LOGL >= 10 ? 3 : 4, 3, 64, NYQ>(io, tw8, units, stream, a);
            case 3: return launch_frame_fused_g<T, LOGL, LOGL >= 10 ? 2 : 3, 3, 64, NYQ>(io, tw8, units, stream, a);
            case 4: return launch_frame_fused_g<T, LOGL, LOGL >= 10 ? 2 : 3, 3, 96, NYQ>(io, tw8, units, stream, a);
            default: return launch_frame_fused_g<T, LOGL, g16, -1, 128, NYQ, true>(io, tw, units, stream, a, knobs.pipelined);
        }
    } else if constexpr (frame_cfg<T, LOGL, g0>::THREADS > 512) {
        return launch_frame_fused_g<T, LOGL, g0 - 1, -1, 128, NYQ, true>(io, tw, units, stream, a, knobs.pipelined);
    } else {
        return launch_frame_fused_g<T, LOGL, g0, -1, 128, NYQ, true>(io, tw, units, stream, a, knobs.pipelined);
    }
}

// ---- frame-level MAC for banks in the three-kernel form (diagonal topology, no partition split) --------------------------------
// Same arithmetic and operand order as fdl_mac_stream_kernel (row p = 0 first), but a frame-level row loop is short (Q = P/T rows,
// 4 at T = 256), so one thread walks `ncols` columns of the row, 128 vectors apart, as ONE software-pipelined stream of
// ncols * Q (filter, spectrum) pairs: 8 pairs = 256 bytes in flight per thread whatever Q is, and 1/ncols as many CTAs.
template<typename T>
__global__ void __launch_bounds__(k_mac_threads)
    frame_mac_kernel(cx<T> const* __restrict__ fdl, cx<T> const* __restrict__ filter, cx<T>* __restrict__ acc, mac_geom g, int ncols)
{
    using MV          = mac_vec<T>;
    using V           = typename MV::type;
    constexpr int UNR = k_mac_unroll;
    int const out      = blockIdx.y + g.out0;
    int const row_vec  = g.m / MV::VEC;
    int const logtv    = g.logw - (MV::VEC == 2 ? 1 : 0);  // log2(vectors per tile row)
    int const tv_mask  = (1 << logtv) - 1;
    int const c_first  = blockIdx.x * ncols * k_mac_threads + threadIdx.x;  // this thread's columns: c_first + i * 128
    int const nc       = c_first < row_vec ? min(ncols, (row_vec - c_first + k_mac_threads - 1) / k_mac_threads) : 0;
    int const total    = nc * g.parts;

    V const* const fbase = reinterpret_cast<V const*>(filter) + ((size_t(out) * g.nt * g.parts) << logtv);
    V const* const xbase = reinterpret_cast<V const*>(fdl) + ((size_t(out) * g.nt * g.ring) << logtv);
    V* const abase       = reinterpret_cast<V*>(acc) + size_t(out) * row_vec;
    int slot0            = (g.wp - g.age0) % g.ring;  // ring slot paired with local partition 0
    slot0 += slot0 < 0 ? g.ring : 0;

    V a      = MV::zero();
    int li = 0, lp = 0;  // load cursor: column i, row p
    int ci = 0, cp = 0;  // accumulate cursor
    for (int it = 0; it < total; it += UNR) {
        V x[UNR], h[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (it + u < total) {
                int const c      = c_first + li * k_mac_threads;
                int const tile   = c >> logtv;
                int const within = c & tv_mask;
                int slot         = slot0 - lp;
                slot += slot < 0 ? g.ring : 0;
                h[u] = ld_stream(fbase + ((size_t(tile) * g.parts + lp) << logtv) + within);
                x[u] = ld_stream(xbase + ((size_t(tile) * g.ring + slot) << logtv) + within);
                if (++lp == g.parts) {
                    lp = 0;
                    ++li;
                }
            } else {
                h[u] = MV::zero();
                x[u] = MV::zero();
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (it + u < total) {
                MV::cfma(a, x[u], h[u]);
                if (++cp == g.parts) {
                    abase[c_first + ci * k_mac_threads] = a;
                    a  = MV::zero();
                    cp = 0;
                    ++ci;
                }
            }
        }
    }
}

}  // namespace neo_b200
