// firfilt_fast.cu -- batched firfilt_crcf for up to 256 taps, instantiated for tap capacities 64 (BASELINE config #2:
// 63 taps, 1024 streams), 128 and 256.
//
//   y[s][n] = scale * sum_k h[k] x[s][n-k]                 (src/filter/fir/firfilt.rs:241-245, :267-278)
//
// Bound: FP32 pipe, not HBM: 2 * 63 lane-FMAs per 16 bytes moved, so at 128 lanes/clk/SM the
// ceiling is ~285 Gsamples/s = 70 % of the HBM roofline.  The kernel therefore spends its instruction
// budget on FFMA2 and almost nothing else:
//   * one CTA = one 4096-output tile of one stream; the tile plus its 63-sample history is staged once
//     in shared memory (coalesced loads), index-padded by one sample per 16 so that the per-thread
//     walk below is bank-conflict free;
//   * thread t owns 16 consecutive outputs: it walks its 79 input samples newest-first (the reference's
//     accumulation order: the VecDeque is newest-at-front) and feeds each sample into the <= 16
//     accumulators it belongs to: one packed FFMA2 (complex sample x broadcast real tap) per MAC;
//   * taps are kernel PARAMETERS (constant bank), so they cost no registers and no loads;
//   * results go back through the same shared tile for fully coalesced stores.
#include "common.cuh"

#include <algorithm>

namespace yg {

namespace {

constexpr int kR = 16;                       // outputs per thread
constexpr int kThreads = 256;
constexpr int kTile = kThreads * kR;         // 4096 outputs per CTA

template <int kH>                            // tap capacity (taps are zero-padded up to it)
struct FirTaps { float h[kH]; };

__device__ __forceinline__ int pad(int g) { return g + (g >> 4); }

// One input sample (relative index II inside the thread's 79-sample span) feeds every output r it
// belongs to through tap k = r + 63 - II.  Template recursion forces the full static unroll that keeps
// every tap a constant-bank operand and every accumulator a fixed register.
template <int kH, int II>
__device__ __forceinline__ void fir_step(float2 (&acc)[kR], const float2* tile, int g0, const FirTaps<kH>& taps)
{
    const float2 v = tile[pad(g0 + II)];
#pragma unroll
    for (int r = 0; r < kR; r++) {
        constexpr int kbase = (kH - 1) - II;
        const int k = r + kbase;
        if (k >= 0 && k < kH) acc[r] = __ffma2_rn(v, make_float2(taps.h[k], taps.h[k]), acc[r]);
    }
}
// samples II = HI-1 down to LO (newest first), split in halves so that the template depth stays logarithmic
template <int kH, int LO, int HI>
__device__ __forceinline__ void fir_steps(float2 (&acc)[kR], const float2* tile, int g0, const FirTaps<kH>& taps)
{
    if constexpr (HI - LO == 1) {
        fir_step<kH, LO>(acc, tile, g0, taps);
    } else {
        constexpr int MID = (LO + HI) / 2;
        fir_steps<kH, MID, HI>(acc, tile, g0, taps);
        fir_steps<kH, LO, MID>(acc, tile, g0, taps);
    }
}

template <int kH>
__global__ void __launch_bounds__(kThreads, 4)
k_firfilt_fast(const FirTaps<kH> taps, float scale, const float2* __restrict__ hist, int Hlen,
               const float2* __restrict__ x, float2* __restrict__ y, long long n, long long t_begin)
{
    constexpr int kIn = kTile + kH - 1;          // samples staged per tile
    constexpr int kPadded = kIn + (kIn >> 4) + 1;
    __shared__ float2 tile[kPadded];
    const int t = threadIdx.x;
    const long long s = blockIdx.y;                                                   // stream
    const long long n0 = t_begin + (long long)blockIdx.x * kTile;                     // first output of the tile
    const float2* xs = x + s * n;
    const float2* hs = hist + s * Hlen;

    // stage samples n0 - (kH-1) .. n0 + 4095 (zeros outside the stream, history for negative indices) with
    // asynchronous 8-byte copies (LDGSTS): all ~17 per thread are in flight at once, no registers held
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    if (n0 >= kH - 1 && n0 + kTile <= n) {                                            // interior tile: no edge logic
        const float2* src0 = xs + (n0 - (kH - 1));
#pragma unroll
        for (int i = 0; i < (kIn + kThreads - 1) / kThreads; i++) {
            const int g = t + i * kThreads;
            if (g < kIn)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile_s + 8u * (uint32_t)pad(g)), "l"(src0 + g) : "memory");
        }
    } else {
#pragma unroll
        for (int i = 0; i < (kIn + kThreads - 1) / kThreads; i++) {
            const int g = t + i * kThreads;
            if (g < kIn) {
                const long long gi = n0 - (kH - 1) + g;
                const float2* src = xs;
                uint32_t bytes = 0;                                   // 0 => the 8 destination bytes are zero-filled
                if (gi >= 0) { if (gi < n) { src = xs + gi; bytes = 8; } }
                else if (gi >= -(long long)Hlen) { src = hs + (Hlen + gi); bytes = 8; }
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(tile_s + 8u * (uint32_t)pad(g)), "l"(src), "r"(bytes) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float2 acc[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) acc[r] = make_float2(0.f, 0.f);
    // output r of this thread needs samples g = 16 t + r - k + 63; sample ii = g - 16 t feeds output r
    // through tap k = r + 63 - ii.  Newest sample first => k ascending for every output.
    const int g0 = kR * t;
    fir_steps<kH, 0, kR + kH - 1>(acc, tile, g0, taps);
    __syncthreads();                                         // everyone is done reading the inputs
#pragma unroll
    for (int r = 0; r < kR; r++) tile[pad(g0 + r)] = make_float2(acc[r].x * scale, acc[r].y * scale);
    __syncthreads();
    float2* ys = y + s * n + n0;
#pragma unroll
    for (int i = 0; i < kR; i++) {
        const int o = t + i * kThreads;
        if (n0 + o < n) __stcs(&ys[o], tile[pad(o)]);
    }
}

}  // namespace

namespace {
template <int kH>
int32_t launch_h(const float* h, size_t h_len, float scale, const float2* hist, long long Hlen, const float2* x,
                 float2* y, long long n, long long n_streams, cudaStream_t st, long long t_begin)
{
    FirTaps<kH> taps;
    for (int k = 0; k < kH; k++) taps.h[k] = (k < (int)h_len) ? h[k] : 0.0f;
    const long long tiles = (n - t_begin + kTile - 1) / kTile;        // outputs [t_begin, n) of every stream
    if (tiles > 0x7fffffffLL) return fail(YG_ERANGE, "too many tiles for one launch");
    for (long long s0 = 0; s0 < n_streams; s0 += 65535) {                  // grid.y carries the stream index
        const long long ns = std::min<long long>(65535, n_streams - s0);
        k_firfilt_fast<kH><<<dim3((unsigned)tiles, (unsigned)ns), kThreads, 0, st>>>(taps, scale, hist + s0 * Hlen, (int)Hlen,
                                                                                    x + s0 * n, y + s0 * n, n, t_begin);
        YG_LAUNCH_CHECK();
    }
    return YG_OK;
}
}  // namespace

bool firfilt_fast_supported(size_t h_len) { return h_len >= 1 && h_len <= 256; }

// Outputs [t_begin, n) of every stream (t_begin > 0: the samples before it are read from x itself -- the tail the
// tensor-core kernel leaves when n is not a whole number of its blocks).
int32_t firfilt_fast_launch(const float* h, size_t h_len, float scale, const float2* hist, long long Hlen, const float2* x,
                            float2* y, long long n, long long n_streams, cudaStream_t st, long long t_begin)
{
    if (t_begin >= n) return YG_OK;
    if (h_len <= 64) return launch_h<64>(h, h_len, scale, hist, Hlen, x, y, n, n_streams, st, t_begin);
    if (h_len <= 128) return launch_h<128>(h, h_len, scale, hist, Hlen, x, y, n, n_streams, st, t_begin);
    return launch_h<256>(h, h_len, scale, hist, Hlen, x, y, n, n_streams, st, t_begin);
}

}  // namespace yg
