// firpfbch_fast.cu -- fused firpfbch_crcf (critically sampled) ANALYSIS kernel for sm_100a, M = 64,
// p = 2m <= 16 taps per branch, many independent streams (BASELINE config #5).
//
//   X_q[pos] = sum_n h[(M-1-pos) + nM] s[(q-n)M + pos]       (thread `pos` owns input position pos)
//   y_q      = DFT_forward(X_q), unscaled                    (X[M-1-b] = V[b])
//
// Same machine as the firpfbch2 kernel (firpfbch2_fast.cu), re-cut for small M: one persistent
// warp-specialised CTA per SM processes FOUR streams at once in batches of 16 frames:
//   TMA bulk copies (one 8 KB copy per stream per batch, issued by four different lanes)
//   FIR role (warps 0-7): thread (stream slot, pos): 32-entry register ring of the branch's samples,
//     one packed FFMA2 (complex sample x broadcast real tap) per MAC, X written to smem
//   FFT role (warps 8-15): 8 threads per frame PAIR, radix-8 x radix-8 DFT of both frames at once in
//     packed (even, odd) lanes; the forward transform is the backward one with re/im swapped on the way
//     in and out (free); exchange through an XOR-swizzled tile in place; 64-byte coalesced stores.
// The four streams of a CTA are independent pipelines (per-stream mbarriers) sharing the input stages.
#include "firpfbch_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kM = 64;
constexpr int kSlots = 4;                        // streams per CTA
constexpr int kBatch = 16;                       // frames per batch (per stream)
constexpr int kFirThreads = 256;
constexpr int kThreads = 512;
constexpr int kInStreamBytes = kBatch * kM * 8;  // 8 KB of input per stream per batch
constexpr int kInStageBytes = kSlots * kInStreamBytes;
constexpr int kRowBytes = kM * 8 + 32;           // 544: X row of one frame, padded so consecutive pairs hit different banks
constexpr int kXStreamBytes = kBatch * kRowBytes;
constexpr int kXBufBytes = kSlots * kXStreamBytes;
constexpr int kSmemBytes = 2 * kInStageBytes + 2 * kXBufBytes + 256;
constexpr int kMbInFull = 0;      // [2]        4 issuing lanes (arrive.expect_tx each)
constexpr int kMbInFree = 2;      // [2]        8 FIR warps drained the stage
constexpr int kMbXFull = 4;       // [2][4][2]  the 2 FIR warps of a stream wrote frames 8h..8h+7
constexpr int kMbXFree = 20;      // [2][4]     the 2 FFT warps of a stream drained the buffer

struct PfbParams {
    const float2* hist;       // [n_streams][Hlen]
    long long Hlen;           // (p-1) * 64
    const float2* x;          // [n_streams][n_frames * 64]
    float2* y;                // [n_streams][n_frames * 64]
    long long n_frames;
    int n_groups;             // groups of 4 streams handled here
    int batches_per_group;
    const float* taps;        // [64][p]  h[(63 - pos) + 64 n]
    const float2* twid;       // [8][8]   e^{+j 2 pi n2 k1 / 64}
};

__device__ __forceinline__ constexpr int dr8(int k) { return ((k & 1) << 2) | (k >> 1); }

// 8-point backward DFT in packed (even, odd) lanes; X[k] is left at v[dr8(k)].
__device__ __forceinline__ void dft8(C2 (&v)[8])
{
    constexpr float r2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const C2 s = cadd(v[b], v[4 + b]), d = csub(v[b], v[4 + b]);
        v[b] = s;
        v[4 + b] = d;
    }
    dft4(v[0], v[1], v[2], v[3]);                                   // k1 = 0
    {                                                               // k1 = 1: W8^1, W8^2 = j, W8^3, 1/sqrt2 folded
        const C2 a0 = v[4], p1 = w8u(v[5]), t2 = v[6], p3 = w8u3(v[7]);
        const C2 s0 = caddj(a0, t2), d0 = csubj(a0, t2);
        const C2 sU = cadd(p1, p3), dU = csub(p1, p3);
        v[4] = cfma(sU, r2, s0); v[6] = cfma(sU, -r2, s0); v[5] = cfmaj(dU, r2, d0); v[7] = cfmaj(dU, -r2, d0);
    }
}

template <int kTaps>
__device__ __forceinline__ void fir_role(const PfbParams& p, uint32_t smem, uint32_t mbar, int L0, int L1)
{
    const int j = threadIdx.x;
    const int slot = j >> 6;                      // stream slot 0..3 (2 warps each)
    const int pos = j & 63;
    const bool issuer = pos == 0;                 // one lane per stream slot issues that stream's copies

    float T[kTaps];
#pragma unroll
    for (int n = 0; n < kTaps; n++) T[n] = __ldg(&p.taps[pos * kTaps + n]);

    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);

    const uint32_t in0 = smem;
    const uint32_t xb0 = smem + 2 * kInStageBytes;
    const long long stream_len = p.n_frames * kM;

    auto frames_in_batch = [&](int k) {
        const long long left = p.n_frames - (long long)k * kBatch;
        return (int)(left < kBatch ? left : kBatch);
    };
    auto issue_load = [&](int L, int st) {
        const int group = L / p.batches_per_group, k = L - group * p.batches_per_group;
        const uint32_t bytes = (uint32_t)frames_in_batch(k) * kM * 8;
        const uint32_t bar = mbar + 8 * (kMbInFull + st);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(in0 + st * kInStageBytes + slot * kInStreamBytes,
                    p.x + (long long)(group * kSlots + slot) * stream_len + (long long)k * kBatch * kM, bytes, bar);
    };

    if (issuer) {
        issue_load(L0, 0);
        if (L0 + 1 < L1) issue_load(L0 + 1, 1);
    }

    int group = L0 / p.batches_per_group;
    int k = L0 - group * p.batches_per_group;     // batch index inside the stream

    auto do_batch = [&](auto par_tag, int L) {
        constexpr int PAR = decltype(par_tag)::value;
        const int lb = L - L0;
        const int st = PAR;
        const long long s = (long long)group * kSlots + slot;          // this thread's stream
        if (k == 0 || lb == 0) {
            // (re)prime the window with u[q0 - i], i = 1..p-1, from history / earlier samples
            const long long q0 = (long long)k * kBatch;
#pragma unroll
            for (int i = 1; i < kTaps; i++) {
                const long long t = (q0 - i) * kM + pos;               // sample index inside the stream
                W[(16 * PAR - i) & 31] = (t >= 0) ? __ldg(&p.x[s * stream_len + t]) : __ldg(&p.hist[s * p.Hlen + p.Hlen + t]);
            }
        }
        mbar_wait(mbar + 8 * (kMbInFull + st), (uint32_t)((lb >> 1) & 1));
        const uint32_t in = in0 + st * kInStageBytes + slot * kInStreamBytes + pos * 8;
#pragma unroll
        for (int r = 0; r < kBatch; r++) W[16 * PAR + r] = lds64(in + r * (kM * 8));
        __syncwarp();
        if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbInFree + st));
        if (lb >= 2) mbar_wait(mbar + 8 * (kMbXFree + 4 * PAR + slot), (uint32_t)(((lb >> 1) - 1) & 1));
        const uint32_t xout = xb0 + PAR * kXBufBytes + slot * kXStreamBytes + pos * 8;
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = kTaps - 1; n >= 0; n--)                       // oldest sample first
                acc = fma2(W[(16 * PAR + r - n) & 31], f2(T[n]), acc);
            sts64(xout + r * kRowBytes, acc);
            if ((r & 7) == 7) {
                __syncwarp();
                if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbXFull + 8 * PAR + 2 * slot + (r >> 3)));
            }
        }
        if (issuer && L + 2 < L1) {                                    // all 8 FIR warps drained this stage?
            mbar_wait(mbar + 8 * (kMbInFree + st), (uint32_t)((lb >> 1) & 1));
            issue_load(L + 2, st);
        }
        if (++k == p.batches_per_group) { k = 0; group++; }
    };

    for (int L = L0; L < L1; L += 2) {
        do_batch(std::integral_constant<int, 0>{}, L);
        if (L + 1 < L1) do_batch(std::integral_constant<int, 1>{}, L + 1);
    }
}

__device__ __forceinline__ void fft_role(const PfbParams& p, uint32_t smem, uint32_t mbar, int L0, int L1)
{
    const int tid = threadIdx.x - kFirThreads;
    const int w = tid >> 5;
    const int slot = w >> 1;                      // stream slot
    const int half = w & 1;                       // frames 8*half .. 8*half+7 of the batch
    const int pr = half * 4 + ((tid >> 3) & 3);   // frame pair inside the batch (frames 2pr, 2pr+1)
    const int t = tid & 7;                        // n2 in pass 1, k1 in pass 2
    const uint32_t xb0 = smem + 2 * kInStageBytes;
    const long long stream_len = p.n_frames * kM;

    float twr[8], twi[8];
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        const float2 tw = __ldg(&p.twid[t * 8 + kk]);
        twr[kk] = tw.x;
        twi[kk] = tw.y;
    }

    int group = L0 / p.batches_per_group;
    int k = L0 - group * p.batches_per_group;
    for (int L = L0; L < L1; L++) {
        const int lb = L - L0;
        const int b = lb & 1;
        const uint32_t rows = xb0 + b * kXBufBytes + slot * kXStreamBytes + (2 * pr) * kRowBytes;   // two X rows = exchange tile
        mbar_wait(mbar + 8 * (kMbXFull + 8 * b + 2 * slot + half), (uint32_t)((lb >> 1) & 1));

        C2 v[8];
        // pass 1: thread n2 = t gathers X[8 n1 + n2] of both frames; re/im swapped => forward transform
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) {
            const float2 e = lds64(rows + (8 * n1 + t) * 8);
            const float2 o = lds64(rows + kRowBytes + (8 * n1 + t) * 8);
            v[n1].re = make_float2(e.y, o.y);
            v[n1].im = make_float2(e.x, o.x);
        }
        dft8(v);
        __syncwarp();                    // the 8 lanes of the pair have read both rows
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) {
            C2 z = v[dr8(k1)];
            if (k1 > 0) z = cmulw(z, twr[k1], twi[k1]);
            sts128(rows + (((t << 3) | (k1 ^ t)) << 4), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) {
            const float4 q4 = lds128(rows + (((n2 << 3) | (t ^ n2)) << 4));
            v[n2].re = make_float2(q4.x, q4.y);
            v[n2].im = make_float2(q4.z, q4.w);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbXFree + 4 * b + slot));
        dft8(v);
        const long long q = (long long)k * kBatch + 2 * pr;            // even frame of the pair
        if (q < p.n_frames) {
            float2* ye = p.y + (long long)(group * kSlots + slot) * stream_len + q * kM + t;
            const bool odd_ok = q + 1 < p.n_frames;
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
                const C2 z = v[dr8(k2)];
                __stcs(ye + 8 * k2, make_float2(z.im.x, z.re.x));       // swap back
                if (odd_ok) __stcs(ye + kM + 8 * k2, make_float2(z.im.y, z.re.y));
            }
        }
        if (++k == p.batches_per_group) { k = 0; group++; }
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch_analysis_fused(const PfbParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + 2 * kInStageBytes + 2 * kXBufBytes;

    const long long n_batches = (long long)p.n_groups * p.batches_per_group;
    const int L0 = (int)((n_batches * blockIdx.x) / gridDim.x);
    const int L1 = (int)((n_batches * (blockIdx.x + 1)) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(mbar + 8 * (kMbInFull + i), kSlots);
            mbar_init(mbar + 8 * (kMbInFree + i), 8);
        }
        for (int i = 0; i < 16; i++) mbar_init(mbar + 8 * (kMbXFull + i), 2);
        for (int i = 0; i < 8; i++) mbar_init(mbar + 8 * (kMbXFree + i), 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (L0 >= L1) return;
    if (threadIdx.x < kFirThreads) fir_role<kTaps>(p, smem, mbar, L0, L1);
    else fft_role(p, smem, mbar, L0, L1);
}

template <int kTaps>
int32_t launch_t(const FirpfbchFastPlan& plan, const PfbParams& p, cudaStream_t st)
{
    // idempotent and cheap (a few microseconds); doing it per launch keeps the code free of shared mutable state
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch_analysis_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long n_batches = (long long)p.n_groups * p.batches_per_group;
    const int grid = (int)std::min<long long>(plan.n_sm, n_batches);
    k_firpfbch_analysis_fused<kTaps><<<grid, kThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

// =================================================================== synthesis
//   U_q = IDFT_unnorm(X_q);   y[qM + i] = sum_n h[i + nM] U_{q-n}[i]            
// Roles swapped: TMA -> FFT role (8 threads per frame pair, backward DFT, U written to smem) -> FIR role
// (thread (stream slot, column i): 32-entry register ring of the column's U values, one output per frame).
// U is never materialised in HBM, so every stream segment starts with a WARM-UP batch: the 16 frames before
// it are transformed again with the outputs suppressed (frames before the call come from the 16-frame input
// history kept per stream).  The four stream slots are independent pipelines with their own mbarriers.
constexpr int kSMbInFull = 0;     // [2][4]     TMA transaction barrier per stage and slot
constexpr int kSMbInFree = 8;     // [2][4]     the 2 FFT warps of the slot drained the stage
constexpr int kSMbUFull = 16;     // [2][4][2]  FFT warp (slot, half) wrote U frames 8h..8h+7
constexpr int kSMbUFree = 32;     // [2][4]     the 2 FIR warps of the slot drained the buffer
constexpr int kSynSmemBytes = 2 * kInStageBytes + 2 * kXBufBytes + 40 * 8 + 64;

struct PfbSynParams {
    const float2* hist;       // [n_streams][hist_frames * 64] input history, oldest first
    long long hist_frames;    // >= 16
    const float2* x;          // [n_streams][n_frames * 64]
    float2* y;                // [n_streams][n_frames * 64]
    long long n_frames;
    int n_groups;
    int batches_per_group;
    const float* taps;        // [64][p]  h[i + 64 n]
    const float2* twid;       // [8][8]   e^{+j 2 pi n2 k1 / 64}
};

// The sequence of work items of a CTA, walked identically by every role: real batches L0..L1-1 of the
// linearised (group, batch) space, each stream segment preceded by one warm-up item (batch k-1, suppressed).
struct Walk {
    int L, L1, nbg, group, k, it;
    bool warm;
    __device__ Walk(int L0, int L1_, int nbg_) : L(L0), L1(L1_), nbg(nbg_), group(L0 / nbg_), k(L0 - (L0 / nbg_) * nbg_), it(0), warm(true) {}
    __device__ bool done() const { return L >= L1; }
    __device__ int batch() const { return warm ? k - 1 : k; }          // -1: the 16 history frames
    __device__ void next()
    {
        it++;
        if (warm) { warm = false; return; }
        L++;
        if (++k == nbg) { k = 0; group++; warm = true; }
    }
};

template <int kTaps>
__device__ __forceinline__ void syn_fir_role(const PfbSynParams& p, uint32_t smem, uint32_t mbar, int L0, int L1)
{
    const int j = threadIdx.x;
    const int slot = j >> 6, col = j & 63;
    float T[kTaps];
#pragma unroll
    for (int n = 0; n < kTaps; n++) T[n] = __ldg(&p.taps[col * kTaps + n]);
    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);
    const uint32_t ub0 = smem + 2 * kInStageBytes + slot * kXStreamBytes + col * 8;
    const long long stream_len = p.n_frames * kM;

    Walk w(L0, L1, p.batches_per_group);
    auto do_item = [&](auto par_tag) {
        constexpr int PAR = decltype(par_tag)::value;
        const uint32_t ph = (uint32_t)((w.it >> 1) & 1);
        const uint32_t ub = ub0 + PAR * kXBufBytes;
        const long long q0 = (long long)w.batch() * kBatch;
        float2* ys = p.y + (long long)(w.group * kSlots + slot) * stream_len + q0 * kM + col;
        const bool emit = !w.warm;
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            if ((r & 7) == 0) mbar_wait(mbar + 8 * (kSMbUFull + 8 * PAR + 2 * slot + (r >> 3)), ph);
            W[16 * PAR + r] = lds64(ub + r * kRowBytes);
            if (r == kBatch - 1) {
                __syncwarp();
                if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kSMbUFree + 4 * PAR + slot));
            }
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = kTaps - 1; n >= 0; n--) acc = fma2(W[(16 * PAR + r - n) & 31], f2(T[n]), acc);
            if (emit && q0 + r < p.n_frames) __stcs(ys + r * kM, acc);
        }
        w.next();
    };
    while (!w.done()) {
        do_item(std::integral_constant<int, 0>{});
        if (!w.done()) do_item(std::integral_constant<int, 1>{});
    }
}

__device__ __forceinline__ void syn_fft_role(const PfbSynParams& p, uint32_t smem, uint32_t mbar, int L0, int L1)
{
    const int tid = threadIdx.x - kFirThreads;
    const int wv = tid >> 5;
    const int slot = wv >> 1, half = wv & 1;
    const int pr = half * 4 + ((tid >> 3) & 3);   // frame pair inside the batch
    const int t = tid & 7;
    const bool issuer = (tid & 63) == 0;          // first lane of the slot's first FFT warp
    const uint32_t in0 = smem + slot * kInStreamBytes;
    const uint32_t ub0 = smem + 2 * kInStageBytes + slot * kXStreamBytes + (2 * pr) * kRowBytes;
    const long long stream_len = p.n_frames * kM;

    float twr[8], twi[8];
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
        const float2 tw = __ldg(&p.twid[t * 8 + kk]);
        twr[kk] = tw.x;
        twi[kk] = tw.y;
    }

    auto issue = [&](const Walk& it_w) {           // the slot's 16 frames of this item: one bulk copy
        const int st = it_w.it & 1;
        const long long s = (long long)it_w.group * kSlots + slot;
        const int kb = it_w.batch();
        const float2* src;
        uint32_t bytes = kBatch * kM * 8;
        if (kb < 0) src = p.hist + (s * p.hist_frames + p.hist_frames - kBatch) * kM;
        else {
            src = p.x + s * stream_len + (long long)kb * kBatch * kM;
            const long long left = p.n_frames - (long long)kb * kBatch;
            if (left < kBatch) bytes = (uint32_t)left * kM * 8;
        }
        const uint32_t bar = mbar + 8 * (kSMbInFull + 4 * st + slot);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(in0 + st * kInStageBytes, src, bytes, bar);
    };

    Walk w(L0, L1, p.batches_per_group);
    Walk ahead = w;                                // the item two steps ahead, for prefetch
    if (issuer) {
        issue(ahead);
        ahead.next();
        if (!ahead.done()) issue(ahead);
        ahead.next();
    }
    while (!w.done()) {
        const int b = w.it & 1;
        const uint32_t ph = (uint32_t)((w.it >> 1) & 1);
        const uint32_t in = in0 + b * kInStageBytes + (2 * pr) * (kM * 8);
        const uint32_t rows = ub0 + b * kXBufBytes;
        mbar_wait(mbar + 8 * (kSMbInFull + 4 * b + slot), ph);
        C2 v[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) {
            const float2 e = lds64(in + (8 * n1 + t) * 8);
            const float2 o = lds64(in + kM * 8 + (8 * n1 + t) * 8);
            v[n1].re = make_float2(e.x, o.x);
            v[n1].im = make_float2(e.y, o.y);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kSMbInFree + 4 * b + slot));
        dft8(v);
#pragma unroll
        for (int k1 = 1; k1 < 8; k1++) v[dr8(k1)] = cmulw(v[dr8(k1)], twr[k1], twi[k1]);
        if (w.it >= 2) mbar_wait(mbar + 8 * (kSMbUFree + 4 * b + slot), ph ^ 1);     // FIR role drained U[b] of item it-2
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) {
            const C2 z = v[dr8(k1)];
            sts128(rows + (((t << 3) | (k1 ^ t)) << 4), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 8; n2++) {
            const float4 q4 = lds128(rows + (((n2 << 3) | (t ^ n2)) << 4));
            v[n2].re = make_float2(q4.x, q4.y);
            v[n2].im = make_float2(q4.z, q4.w);
        }
        dft8(v);
        __syncwarp();                               // exchange tile consumed before U overwrites the rows
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            const C2 z = v[dr8(k2)];
            sts64(rows + (t + 8 * k2) * 8, make_float2(z.re.x, z.im.x));
            sts64(rows + kRowBytes + (t + 8 * k2) * 8, make_float2(z.re.y, z.im.y));
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kSMbUFull + 8 * b + 2 * slot + half));
        // prefetch the item two steps ahead into the stage this item has just drained
        if (issuer && !ahead.done()) {
            mbar_wait(mbar + 8 * (kSMbInFree + 4 * b + slot), ph);
            issue(ahead);
        }
        if (issuer) ahead.next();
        w.next();
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch_synthesis_fused(const PfbSynParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + 2 * kInStageBytes + 2 * kXBufBytes;
    const long long n_batches = (long long)p.n_groups * p.batches_per_group;
    const int L0 = (int)((n_batches * blockIdx.x) / gridDim.x);
    const int L1 = (int)((n_batches * (blockIdx.x + 1)) / gridDim.x);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; i++) {
            mbar_init(mbar + 8 * (kSMbInFull + i), 1);
            mbar_init(mbar + 8 * (kSMbInFree + i), 2);
            mbar_init(mbar + 8 * (kSMbUFree + i), 2);
        }
        for (int i = 0; i < 16; i++) mbar_init(mbar + 8 * (kSMbUFull + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (L0 >= L1) return;
    if (threadIdx.x < kFirThreads) syn_fir_role<kTaps>(p, smem, mbar, L0, L1);
    else syn_fft_role(p, smem, mbar, L0, L1);
}

template <int kTaps>
int32_t launch_syn_t(const FirpfbchFastPlan& plan, const PfbSynParams& p, cudaStream_t st)
{
    // idempotent and cheap (a few microseconds); doing it per launch keeps the code free of shared mutable state
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch_synthesis_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSynSmemBytes));
    const long long n_batches = (long long)p.n_groups * p.batches_per_group;
    const int grid = (int)std::min<long long>(plan.n_sm, n_batches);
    k_firpfbch_synthesis_fused<kTaps><<<grid, kThreads, kSynSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

}  // namespace

int32_t firpfbch_fast_plan(FirpfbchFastPlan& plan, int32_t type, uint32_t M, uint32_t p, const float* h)
{
    plan.supported = false;
    plan.p = p;
    if (M != (uint32_t)kM) return YG_OK;
    plan.type = type;
    if (p < 2 || p > 16 || (p & 1)) return YG_OK;        // instantiated: p = 2, 4, ..., 16
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    std::vector<float> taps((size_t)kM * p);
    for (int pos = 0; pos < kM; pos++)
        for (uint32_t n = 0; n < p; n++)
            taps[(size_t)pos * p + n] = (type == YG_ANALYZER) ? h[(kM - 1 - pos) + n * kM] : h[pos + n * kM];
    std::vector<float2> tw(64);
    for (int n2 = 0; n2 < 8; n2++)
        for (int k1 = 0; k1 < 8; k1++) {
            const double a = 2.0 * M_PI * (double)(n2 * k1) / 64.0;
            tw[n2 * 8 + k1] = make_float2((float)cos(a), (float)sin(a));
        }
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.supported = true;
    return YG_OK;
}

void firpfbch_fast_release(FirpfbchFastPlan& plan)
{
    if (plan.d_taps) cudaFree(plan.d_taps);
    if (plan.d_twid) cudaFree(plan.d_twid);
    plan.d_taps = plan.d_twid = nullptr;
    plan.supported = false;
}

int32_t firpfbch_fast_launch(const FirpfbchFastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                             long long n_frames, long long n_streams, cudaStream_t st)
{
    if (!plan.supported || plan.type != YG_ANALYZER) return fail(YG_EINTERNAL, "fused firpfbch kernel not available for this geometry");
    if (n_streams % kSlots) return fail(YG_EINTERNAL, "fused firpfbch kernel takes groups of 4 streams");
    if ((((uintptr_t)x) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    PfbParams p;
    p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
    p.n_frames = n_frames;
    p.n_groups = (int)(n_streams / kSlots);
    p.batches_per_group = (int)((n_frames + kBatch - 1) / kBatch);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    if ((long long)p.n_groups * p.batches_per_group > 0x7fffffffLL) return fail(YG_ERANGE, "too many batches for one launch");
    switch (plan.p) {
        case 2: return launch_t<2>(plan, p, st);
        case 4: return launch_t<4>(plan, p, st);
        case 6: return launch_t<6>(plan, p, st);
        case 8: return launch_t<8>(plan, p, st);
        case 10: return launch_t<10>(plan, p, st);
        case 12: return launch_t<12>(plan, p, st);
        case 14: return launch_t<14>(plan, p, st);
        case 16: return launch_t<16>(plan, p, st);
        default: return fail(YG_EINTERNAL, "fused firpfbch kernel not instantiated for p = %u", plan.p);
    }
}

// Synthesis: hist = [n_streams][hist_frames * 64] INPUT history (hist_frames >= 16); n_streams a multiple of 4.
int32_t firpfbch_fast_synth_launch(const FirpfbchFastPlan& plan, const float2* hist, long long hist_frames, const float2* x,
                                   float2* y, long long n_frames, long long n_streams, cudaStream_t st)
{
    if (!plan.supported || plan.type != YG_SYNTHESIZER) return fail(YG_EINTERNAL, "fused firpfbch synthesis kernel not available");
    if (n_streams % kSlots) return fail(YG_EINTERNAL, "fused firpfbch kernel takes groups of 4 streams");
    if (hist_frames < kBatch) return fail(YG_EINTERNAL, "input history too short for the warm-up batch");
    if ((((uintptr_t)x) & 15) != 0 || (((uintptr_t)hist) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    PfbSynParams p;
    p.hist = hist; p.hist_frames = hist_frames; p.x = x; p.y = y;
    p.n_frames = n_frames;
    p.n_groups = (int)(n_streams / kSlots);
    p.batches_per_group = (int)((n_frames + kBatch - 1) / kBatch);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    if ((long long)p.n_groups * p.batches_per_group > 0x3fffffffLL) return fail(YG_ERANGE, "too many batches for one launch");
    switch (plan.p) {
        case 2: return launch_syn_t<2>(plan, p, st);
        case 4: return launch_syn_t<4>(plan, p, st);
        case 6: return launch_syn_t<6>(plan, p, st);
        case 8: return launch_syn_t<8>(plan, p, st);
        case 10: return launch_syn_t<10>(plan, p, st);
        case 12: return launch_syn_t<12>(plan, p, st);
        case 14: return launch_syn_t<14>(plan, p, st);
        case 16: return launch_syn_t<16>(plan, p, st);
        default: return fail(YG_EINTERNAL, "fused firpfbch kernel not instantiated for p = %u", plan.p);
    }
}

}  // namespace yg
