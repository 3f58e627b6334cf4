// firpfbch2_synth_fast.cu -- fused firpfbch2 synthesis kernel for sm_100a (M = 256, m = 1..7).
//
//   y[k M/2 + i] = sum_{l < 4m} h[i + l M/2] u_{k-l}[(i + (k&1) M/2) mod M],   u_k = 1/2 IDFT_unnorm(X_k)
//
// Mirror image of the analysis kernel: one persistent warp-specialised CTA per SM walks a
// contiguous slab of frames through three rotating shared-memory buffers of 16 frame pairs:
//
//   TMA bulk copies (4 KB per frame pair, mbarrier tx)  ->  X_E | X_O in a pair region
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
constexpr int kPairs = 16;                       // frame pairs per round (= 32 frames)
constexpr int kFrames = 2 * kPairs;
constexpr int kFirThreads = 256;
constexpr int kThreads = 512;
constexpr int kRegionBytes = 16 * 17 * 16;       // 4352: X_E|X_O (4096) -> padded exchange -> U (4096)
constexpr int kBufBytes = kPairs * kRegionBytes; // 69632
constexpr int kNumBufs = 3;
constexpr int kSmemBytes = kNumBufs * kBufBytes + 128;
constexpr int kMbXFull = 0;      // [3] TMA transaction barriers
constexpr int kMbUFull = 3;      // [3] 8 FFT warps have written U
constexpr int kMbFree = 6;       // [3] 8 FIR warps have drained the buffer

struct SynthParams {
    const float2* prefix;     // 32 input frames preceding x[0] of the call
    const float2* x;          // first input frame of the call
    float2* y;                // first output sample of the call
    long long f0;             // first frame handled here (even parity, multiple-of-32 count follows)
    long long n_rounds;       // rounds of 32 frames handled here
    const float* taps;        // [256][4m] 0.5 * h[(j & 127) + l * 128]
    const float2* twid;       // [16][16] e^{+j 2 pi n2 k1 / 256}
};

template <int kTaps>                              // 4m
__device__ __forceinline__ void fir_role(const SynthParams& p, uint32_t smem, uint32_t mbar,
                                         long long round_begin, long long round_end)
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

    // per-thread byte offsets inside a pair region {reE, imE, reO, imO} x 256 columns
    const uint32_t col = (uint32_t)j * 16;
    const int off1 = hi ? (-kRegionBytes + 8) : 0;    // slot 2s   : lo = E of region s, hi = O of region s-1
    const int off2 = hi ? 0 : 8;                      // slot 2s+1 : lo = O of region s, hi = E of region s

    auto output = [&](int slot_newest, long long frame) {
        // two banks (even / odd l), each oldest first, then added (upstream y0 + y1)
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = kTaps - 1; l >= 0; l--) {
            const float2 w = W[(slot_newest - l) & 31];
            if (l & 1) a1 = fma2(w, f2(T[l]), a1);
            else a0 = fma2(w, f2(T[l]), a0);
        }
        const float2 r = add2(a0, a1);
        if (frame >= p.f0 && frame < p.f0 + p.n_rounds * kFrames) __stcs(p.y + frame * kM2 + i, r);
    };

    for (long long round = round_begin; round < round_end; round++) {
        const long long lr = round - round_begin;
        const int b = (int)(lr % kNumBufs);
        const uint32_t base = smem + b * kBufBytes + col;
        const long long frame0 = p.f0 + (round - 1) * kFrames;        // round 0 of the call is the warm-up round
        mbar_wait(mbar + 8 * (kMbUFull + b), (uint32_t)((lr / kNumBufs) & 1));
#pragma unroll
        for (int s = 0; s < kPairs; s++) {
            if (s > 0 || !hi) W[(2 * s) & 31] = lds64(base + s * kRegionBytes + off1);
            W[(2 * s + 1) & 31] = lds64(base + s * kRegionBytes + off2);
            output(2 * s, frame0 + 2 * s - (hi ? 1 : 0));
        }
        if (hi) W[0] = lds64(base + (kPairs - 1) * kRegionBytes + 8);   // frame 31 -> slot 32 = 0 of the next round
        __syncwarp();
        if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbFree + b));
    }
    // the upper half is one frame behind: emit the last odd frame of the slab
    if (hi) output(0, p.f0 + (round_end - 1) * kFrames - 1);
}

__device__ __forceinline__ void fft_role(const SynthParams& p, uint32_t smem, uint32_t mbar,
                                         long long round_begin, long long round_end)
{
    const int tid = threadIdx.x - kFirThreads;
    const int g = tid >> 4;              // frame pair within the round
    const int t = tid & 15;              // n2 in pass 1, k1 in pass 2

    float twr[16], twi[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float2 w = __ldg(&p.twid[t * 16 + k]);
        twr[k] = w.x;
        twi[k] = w.y;
    }

    // frame index of the first frame of a round, relative to x[0] of the call (may be negative: prefix)
    auto round_frame0 = [&](long long round) { return p.f0 + (round - 1) * kFrames; };
    auto issue_load = [&](long long round) {
        const int b = (int)((round - round_begin) % kNumBufs);
        const long long k0 = round_frame0(round);
        mbar_expect_tx(mbar + 8 * (kMbXFull + b), kFrames * kM * 8);
        // one 2 KB bulk copy per frame: frames before the call come from the 32-frame prefix
#pragma unroll 1
        for (int f = 0; f < kFrames; f++) {
            const long long k = k0 + f;
            const float2* src = (k < 0) ? p.prefix + (k + kFrames) * kM : p.x + k * kM;
            tma_load_1d(smem + b * kBufBytes + (f >> 1) * kRegionBytes + (f & 1) * 2048, src, kM * 8,
                        mbar + 8 * (kMbXFull + b));
        }
    };

    if (tid == 0) issue_load(round_begin);

    for (long long round = round_begin; round < round_end; round++) {
        const long long lr = round - round_begin;
        const int b = (int)(lr % kNumBufs);
        const uint32_t region = smem + b * kBufBytes + g * kRegionBytes;

        // keep one round of input in flight: buffer of round+1 was last used by round-2
        if (tid == 0 && round + 1 < round_end) {
            const long long lr1 = lr + 1;
            if (lr1 >= kNumBufs) mbar_wait(mbar + 8 * (kMbFree + (int)(lr1 % kNumBufs)), (uint32_t)(((lr1 / kNumBufs) - 1) & 1));
            issue_load(round + 1);
        }
        mbar_wait(mbar + 8 * (kMbXFull + b), (uint32_t)((lr / kNumBufs) & 1));

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
            sts128(region + (t * 17 + k1) * 16, make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const float4 q = lds128(region + (n2 * 17 + t) * 16);
            v[n2].re = make_float2(q.x, q.y);
            v[n2].im = make_float2(q.z, q.w);
        }
        dft16(v);
        __syncwarp();                    // exchange tile fully consumed before U overwrites it
        // U[k1 + 16 k2] as {reE, imE, reO, imO}; the 1/2 scale lives in the taps
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const C2 z = v[dr4(k2)];
            sts128(region + (t + 16 * k2) * 16, make_float4(z.re.x, z.im.x, z.re.y, z.im.y));
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbUFull + b));
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch2_synthesis_fused(const SynthParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + kNumBufs * kBufBytes;

    // slab of real rounds [r0, r1) plus one warm-up round in front: local rounds [r0, r1 + 1) where
    // local round R covers frames f0 + (R - 1) * 32 ...; CTA c therefore gets its own f0.
    const long long r0 = (p.n_rounds * blockIdx.x) / gridDim.x;
    const long long r1 = (p.n_rounds * (blockIdx.x + 1)) / gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kNumBufs; i++) {
            mbar_init(mbar + 8 * (kMbXFull + i), 1);
            mbar_init(mbar + 8 * (kMbUFull + i), 8);
            mbar_init(mbar + 8 * (kMbFree + i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (r0 >= r1) return;

    // re-base the parameters on this CTA's slab: its frames are [f0 + 32 r0, f0 + 32 r1)
    SynthParams q = p;
    q.f0 = p.f0 + r0 * kFrames;
    q.n_rounds = r1 - r0;
    // local rounds 0 .. n_rounds (inclusive of the warm-up round 0)
    if (threadIdx.x < kFirThreads) fir_role<kTaps>(q, smem, mbar, 0, q.n_rounds + 1);
    else fft_role(q, smem, mbar, 0, q.n_rounds + 1);
}

template <int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const SynthParams& p, cudaStream_t st)
{
    static bool attr_done[64] = {};
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
        YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_synthesis_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_done[dev & 63] = true;
    }
    const int grid = (int)std::min<long long>(plan.n_sm, p.n_rounds);
    k_firpfbch2_synthesis_fused<kTaps><<<grid, kThreads, kSmemBytes, st>>>(p);
    YG_CUDA(cudaGetLastError());
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
    YG_CUDA(cudaMemcpy(p.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&p.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(cudaMemcpy(p.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    p.min_frames = 64;
    p.supported = true;
    return YG_OK;
}

int32_t firpfbch2_synth_fast_launch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, float2* y,
                                    size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "fused synthesis kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames % kFrames) return fail(YG_EINTERNAL, "fused synthesis kernel needs a multiple of 32 frames");
    if ((((uintptr_t)x) & 15) != 0 || (((uintptr_t)prefix) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    SynthParams p;
    p.prefix = prefix; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_rounds = (long long)(n_frames / kFrames);
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
