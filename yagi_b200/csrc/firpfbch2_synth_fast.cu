// firpfbch2_synth_fast.cu -- fused firpfbch2 synthesis kernel for sm_100a (M = 256, m = 1..7).
//
//   y[k M/2 + i] = sum_{l < 4m} h[i + l M/2] u_{k-l}[(i + (k&1) M/2) mod M],   u_k = 1/2 IDFT_unnorm(X_k)
//
// Mirror image of the analysis kernel: one persistent warp-specialised CTA per SM walks a
// contiguous slab of frames through a ring of twelve 16 KB shared-memory chunks (4 frame pairs each):
//
//   TMA bulk copy (16 KB per chunk, mbarrier tx; 8 chunks kept in flight)  ->  X_E | X_O per pair region
//   FFT role (warps 8-15): 16 threads per frame pair, radix-16 x radix-16 backward DFT of both
//     frames at once in packed (even, odd) lanes, exchange in place in the region, then the
//     region is overwritten with U as {reE, imE, reO, imO} per column
//   FIR role (warps 0-7): thread j owns COLUMN j of U: the last 4m frames of that column live in a
//     32-entry register ring; one packed FFMA2 per tap (complex sample x broadcast real tap);
//     columns j < M/2 emit on even frames, columns j >= M/2 on odd frames (the same code runs for
//     both halves: the upper half keeps its ring one frame behind through per-thread load offsets)
//
// U is never materialised in HBM, so a slab cannot read its filter history: every CTA starts
// 32 frames early (the IFFTs of those frames are recomputed, outputs suppressed; 0.2 % extra
// work).  For the same reason the object's state is the last 32 INPUT frames (`prefix`).
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kM = 256;
constexpr int kM2 = 128;
constexpr int kFirThreads = 256;
constexpr int kThreads = 512;
constexpr int kRegionBytes = 4096;               // X_E|X_O -> XOR-swizzled 16x16 exchange tile -> U_E|U_O, all in place
constexpr int kChunkPairs = 4;                   // frame pairs per pipeline chunk (= 8 frames)
constexpr int kChunkFrames = 2 * kChunkPairs;
constexpr int kChunkBytes = kChunkPairs * kRegionBytes;   // 17408
constexpr int kRing = 12;                        // chunk buffers in flight: ~4 loading, 4 in the FFT role, 1-2 in the FIR role
constexpr int kLookahead = 6;                    // chunks between a FIR warp's position and the chunk it prefetches (swept 4..9: profiles/)
constexpr int kPeriodFrames = 32;                // register-ring period of the FIR role = 4 chunks
constexpr int kSmemBytes = kRing * kChunkBytes + 3 * kRing * 8 + 32;
constexpr int kMbXFull = 0;                      // [12] TMA transaction barriers
constexpr int kMbUFull = kRing;                  // [12] the 2 FFT warps of the chunk have written U
constexpr int kMbFree = 2 * kRing;               // [12] 8 FIR warps have drained the chunk

struct SynthParams {
    const float2* prefix;     // 32 input frames preceding x[0] of the call
    const float2* x;          // first input frame of the call
    float2* y;                // first output sample of the call
    long long f0;             // first frame handled here (even global parity)
    long long n_periods;      // periods of 32 frames handled here
    const float* taps;        // [256][4m] 0.5 * h[(j & 127) + l * 128]
    const float2* twid;       // [16][16] e^{+j 2 pi n2 k1 / 256}
};

// Slab-local chunk c covers frames f0 - 32 + 8c .. +7 (the first four chunks are the warm-up period).
__device__ __forceinline__ void issue_chunk(const SynthParams& p, uint32_t smem, uint32_t mbar, int c, int b)
{
    const long long k0 = p.f0 - kPeriodFrames + (long long)c * kChunkFrames;      // relative to x[0] of the call
    const uint32_t bar = mbar + 8 * (kMbXFull + b);
    const uint32_t dst = smem + b * kChunkBytes;
    mbar_expect_tx(bar, kChunkFrames * kM * 8);
    if (k0 & 1) {
        // odd call-relative start (a leading frame went to the generic kernel): a pair may straddle
        // the prefix | x boundary, so copy frame by frame
#pragma unroll 1
        for (int f = 0; f < kChunkFrames; f++) {
            const long long k = k0 + f;
            const float2* src = (k < 0) ? p.prefix + (k + kPeriodFrames) * kM : p.x + k * kM;
            tma_load_1d(dst + (f >> 1) * kRegionBytes + (f & 1) * 2048, src, kM * 8, bar);
        }
    } else {
        // regions are dense, so the whole chunk is ONE 16 KB bulk copy (a single issuing thread
        // sustains only ~3 bulk copies per microsecond: profiles/r01_tma_bench.log)
        const float2* src = (k0 < 0) ? p.prefix + (k0 + kPeriodFrames) * kM : p.x + k0 * kM;
        tma_load_1d(dst, src, kChunkFrames * kM * 8, bar);
    }
}

template <int kTaps>                              // 4m
__device__ __forceinline__ void fir_role(const SynthParams& p, uint32_t smem, uint32_t mbar, int n_chunks)
{
    const int j = threadIdx.x;                    // column of U
    const bool hi = j >= kM2;                     // warp-uniform
    const int i = j & (kM2 - 1);                  // output index inside a frame

    float T[kTaps];
#pragma unroll
    for (int l = 0; l < kTaps; l++) T[l] = __ldg(&p.taps[j * kTaps + l]);

    // Register ring: for the lower half slot (k mod 32) holds frame k of this column, for the
    // upper half slot (k mod 32) holds frame k-1, so "the window ending at slot 2s" is the window
    // of an even frame below and of an odd frame above, with identical static indices.
    float2 W[32];
#pragma unroll
    for (int s = 0; s < 32; s++) W[s] = make_float2(0.f, 0.f);

    // per-thread byte offsets inside a pair region: U_E plane at +0, U_O plane at +2048, 8 B per column
    const uint32_t col = (uint32_t)j * 8;
    const int off1 = hi ? (-kRegionBytes + 2048) : 0;   // slot 2s   : lo = E of region s, hi = O of region s-1
    const int off2 = hi ? 0 : 2048;                     // slot 2s+1 : lo = O of region s, hi = E of region s
    const int n_out = (int)p.n_periods * kPeriodFrames; // frames of this slab that are really emitted
    float2* const ybase = p.y + p.f0 * kM2 + i;         // output i of slab frame 0

    // `rel` = frame index relative to the first real frame of the slab (the warm-up period is -32..-1)
    auto output = [&](int slot_newest, int rel) {
        // two banks (even / odd l), each oldest first, then added (upstream y0 + y1)
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = kTaps - 1; l >= 0; l--) {
            const float2 w = W[(slot_newest - l) & 31];
            if (l & 1) a1 = fma2(w, f2(T[l]), a1);
            else a0 = fma2(w, f2(T[l]), a0);
        }
        const float2 r = add2(a0, a1);
        if ((unsigned)rel < (unsigned)n_out) __stcs(ybase + (long long)rel * kM2, r);
    };

    // Prefetch: chunk c + lookahead is issued by lane 0 of one FIR warp when that warp starts chunk c.
    // Issue is tied to CONSUMPTION so a fixed number of chunks stays in flight per SM.
    constexpr int la = kLookahead;
    if (j == 0)
        for (int c = 0; c < la && c < n_chunks; c++) issue_chunk(p, smem, mbar, c, c);

    uint32_t prev_last = 0;                       // O plane of the last region of the previous chunk (upper half)
    int b = 0;                                    // ring slot of the current chunk
    uint32_t phase = 0;                           // parity of the ring generation of the current chunk
    for (int c0 = 0; c0 < n_chunks; c0 += 4) {    // one register-ring period = 4 chunks
        const int rel0 = c0 * kChunkFrames - kPeriodFrames - (hi ? 1 : 0);
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int c = c0 + cc;
            const uint32_t base = smem + b * kChunkBytes + col;
            if ((j & 31) == 0 && (j >> 5) == (c & 7) && c + la < n_chunks) {
                int bn = b + la;                                          // slot of chunk c + la, last used by chunk c + la - 12
                uint32_t ph = phase;
                if (bn >= kRing) { bn -= kRing; ph ^= 1; }
                if (c + la >= kRing) mbar_wait(mbar + 8 * (kMbFree + bn), ph ^ 1);
                issue_chunk(p, smem, mbar, c + la, bn);
            }
            mbar_wait(mbar + 8 * (kMbUFull + b), phase);
#pragma unroll
            for (int q = 0; q < kChunkPairs; q++) {
                const int s = cc * kChunkPairs + q;                       // step inside the period, static
                if (q > 0) W[(2 * s) & 31] = lds64(base + q * kRegionBytes + off1);
                else if (!hi) W[(2 * s) & 31] = lds64(base);
                else if (c > 0) W[(2 * s) & 31] = lds64(prev_last);
                W[(2 * s + 1) & 31] = lds64(base + q * kRegionBytes + off2);
                if (q == 0 && c > 0) {                                    // the previous chunk is fully drained now
                    __syncwarp();
                    if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbFree + (b == 0 ? kRing - 1 : b - 1)));
                }
                output(2 * s, rel0 + 2 * s);
            }
            prev_last = base + (kChunkPairs - 1) * kRegionBytes + 2048;
            if (++b == kRing) { b = 0; phase ^= 1; }
        }
    }
    // the upper half is one frame behind: emit the last odd frame of the slab (slot 32 = 0)
    if (hi) {
        W[0] = lds64(prev_last);
        output(0, n_out - 1);
    }
}

__device__ __forceinline__ void fft_role(const SynthParams& p, uint32_t smem, uint32_t mbar, int n_chunks)
{
    const int tid = threadIdx.x - kFirThreads;
    const int g = tid >> 4;              // 0..15: chunk (g >> 2) of the super-round, region (g & 3) of the chunk
    const int t = tid & 15;              // n2 in pass 1, k1 in pass 2

    float twr[16], twi[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float2 w = __ldg(&p.twid[t * 16 + k]);
        twr[k] = w.x;
        twi[k] = w.y;
    }

    int b = g >> 2;                       // ring slot of this warp pair's current chunk
    uint32_t phase = 0;
    for (int c = (g >> 2); c < n_chunks; c += 4) {
        const uint32_t region = smem + b * kChunkBytes + (g & 3) * kRegionBytes;
        mbar_wait(mbar + 8 * (kMbXFull + b), phase);

        C2 v[16];
        // pass 1: thread n2 = t gathers X[16 n1 + n2] of both frames
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const float2 e = lds64(region + (16 * n1 + t) * 8);
            const float2 o = lds64(region + 2048 + (16 * n1 + t) * 8);
            v[n1].re = make_float2(e.x, o.x);
            v[n1].im = make_float2(e.y, o.y);
        }
        dft16(v);
        __syncwarp();                    // all 16 lanes of the group have read X
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            C2 z = v[dr4(k1)];
            if (k1 > 0) z = cmulw(z, twr[k1], twi[k1]);
            sts128(region + (((t << 4) | (k1 ^ t)) << 4), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));   // row n2, swizzled column
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const float4 q = lds128(region + (((n2 << 4) | (t ^ n2)) << 4));
            v[n2].re = make_float2(q.x, q.y);
            v[n2].im = make_float2(q.z, q.w);
        }
        dft16(v);
        __syncwarp();                    // exchange tile fully consumed before U overwrites it
        // U[k1 + 16 k2]: even frame plane at +0, odd frame plane at +2048; the 1/2 scale lives in the taps
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const C2 z = v[dr4(k2)];
            sts64(region + (t + 16 * k2) * 8, make_float2(z.re.x, z.im.x));
            sts64(region + 2048 + (t + 16 * k2) * 8, make_float2(z.re.y, z.im.y));
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbUFull + b));
        b += 4;
        if (b >= kRing) { b -= kRing; phase ^= 1; }
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch2_synthesis_fused(const SynthParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + kRing * kChunkBytes;

    // slab of periods [r0, r1) of the call, plus one warm-up period in front
    const long long r0 = (p.n_periods * blockIdx.x) / gridDim.x;
    const long long r1 = (p.n_periods * (blockIdx.x + 1)) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRing; i++) {
            mbar_init(mbar + 8 * (kMbXFull + i), 1);
            mbar_init(mbar + 8 * (kMbUFull + i), 2);
            mbar_init(mbar + 8 * (kMbFree + i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (r0 >= r1) return;

    SynthParams q = p;
    q.f0 = p.f0 + r0 * kPeriodFrames;
    q.n_periods = r1 - r0;
    const int n_chunks = (int)(q.n_periods + 1) * 4;       // warm-up period + real periods
    if (threadIdx.x < kFirThreads) fir_role<kTaps>(q, smem, mbar, n_chunks);
    else fft_role(q, smem, mbar, n_chunks);
}

template <int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const SynthParams& p, cudaStream_t st)
{
    // idempotent and cheap (a few microseconds); doing it per launch keeps the code free of shared mutable state
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_synthesis_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const int grid = (int)std::min<long long>(plan.n_sm, p.n_periods);
    k_firpfbch2_synthesis_fused<kTaps><<<grid, kThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

}  // namespace

int32_t firpfbch2_synth_fast_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h)
{
    p.supported = false;
    p.M = M;
    p.m = m;
    if (M != (uint32_t)kM) return YG_OK;
    if (m < 1 || m > 7) return YG_OK;         // the 32-slot ring holds the 4m-frame window plus the frame being loaded
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    p.n_sm = prop.multiProcessorCount;

    const int kTaps = 4 * (int)m;
    std::vector<float> taps((size_t)kM * kTaps);
    for (int j = 0; j < kM; j++)
        for (int l = 0; l < kTaps; l++) taps[(size_t)j * kTaps + l] = 0.5f * h[(j & (kM2 - 1)) + l * kM2];
    std::vector<float2> tw(256);
    for (int n2 = 0; n2 < 16; n2++)
        for (int k1 = 0; k1 < 16; k1++) {
            const double a = 2.0 * M_PI * (double)(n2 * k1) / 256.0;
            tw[n2 * 16 + k1] = make_float2((float)cos(a), (float)sin(a));
        }
    YG_CUDA(cudaMalloc(&p.d_taps, taps.size() * sizeof(float)));
    YG_CUDA(yg::memcpy_sync(p.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&p.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(p.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    p.min_frames = 64;
    p.supported = true;
    return YG_OK;
}

int32_t firpfbch2_synth_fast_launch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, float2* y,
                                    size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "fused synthesis kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames % kPeriodFrames) return fail(YG_EINTERNAL, "fused synthesis kernel needs a multiple of 32 frames");
    if ((((uintptr_t)x) & 15) != 0 || (((uintptr_t)prefix) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    SynthParams p;
    p.prefix = prefix; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_periods = (long long)(n_frames / kPeriodFrames);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    switch (plan.m) {
        case 1: return launch_t<4>(plan, p, st);
        case 2: return launch_t<8>(plan, p, st);
        case 3: return launch_t<12>(plan, p, st);
        case 4: return launch_t<16>(plan, p, st);
        case 5: return launch_t<20>(plan, p, st);
        case 6: return launch_t<24>(plan, p, st);
        case 7: return launch_t<28>(plan, p, st);
        default: return fail(YG_EINTERNAL, "fused synthesis kernel not instantiated for m = %u", plan.m);
    }
}

}  // namespace yg
