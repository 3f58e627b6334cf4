// firpfbch2_small.cu -- fused firpfbch2 ANALYSIS kernel for small M (M = 64 and M = 128), m = 1..8, sm_100a.
//
// The M = 256 kernel (firpfbch2_fast.cu) gives one thread to every polyphase branch; with 64 branches that
// would leave three quarters of the FIR role idle.  A single stream has no other parallelism than time, so
// each CTA works on 256/M independent time slabs of the same stream at once ("slots", M FIR threads each),
// every slab primed with its own (4m-1) M/2-sample history exactly like a time shard.
// Per slot the pipeline is the M = 256 one: TMA bulk copy of 16 frame pairs (8 KB) -> FIR role (register
// ring, packed FFMA2 for the even and odd frame of a pair, the upper half's even taps delayed one slot) ->
// V[pair][branch] in smem -> FFT role (8 threads per frame pair, radix-8 x radix-8 backward DFT in packed
// (even, odd) lanes, XOR-swizzled exchange in place) -> 64-byte coalesced stores.  M = 128 uses radix-16 then
// two radix-8 per thread.  Slots are independent pipelines with their own mbarriers.
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kPairsPerBatch = 16;
constexpr int kFirThreads = 256;
constexpr int kThreads = 512;
constexpr int kInStageBytes = kFirThreads * kPairsPerBatch * 8;      // 32 KB of input per batch (all slots)
constexpr int kVBufBytes = kFirThreads * kPairsPerBatch * 16;        // 64 KB of V per batch (all slots)
constexpr int kSmemBytes = 2 * kInStageBytes + 2 * kVBufBytes + 56 * 8 + 64;
constexpr int kMbInFull = 0;      // [2][4]     TMA transaction barrier per stage and slot
constexpr int kMbInFree = 8;      // [2][4]     the FIR warps of the slot drained the stage
constexpr int kMbVFull = 16;      // [2][4][4]  the FIR warps of the slot wrote regions 4g..4g+3
constexpr int kMbVFree = 48;      // [2][4]     the FFT warps of the slot drained the buffer

template <int kM>
struct Geo {
    static constexpr int M2 = kM / 2;
    static constexpr int slots = kFirThreads / kM;                   // time slabs per CTA
    static constexpr int warps_per_slot = kM / 32;                   // in each role
    static constexpr int in_slot_bytes = kPairsPerBatch * kM * 8;
    static constexpr int region_bytes = kM * 16;                     // V of one pair {reE, reO, imE, imO} x M = its exchange tile
    static constexpr int v_slot_bytes = kPairsPerBatch * region_bytes;
};

struct SmallParams {
    const float2* hist;       // Hlen samples preceding x[0] of the call
    long long Hlen;
    const float2* x;
    float2* y;
    long long f0;             // first frame handled here (even global parity)
    long long n_pairs;        // frame pairs handled here
    int n_slabs;              // = gridDim.x * slots
    const float2* taps;       // [M][2m+1] (even, odd) tap pairs, 1/M folded in
    const float2* twid;       // [8][M/8] e^{+j 2 pi n2 k1 / M}
};

__device__ __forceinline__ constexpr int dr8(int k) { return ((k & 1) << 2) | (k >> 1); }

__device__ __forceinline__ void dft8(C2 (&v)[8])
{
    constexpr float r2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const C2 s = cadd(v[b], v[4 + b]), d = csub(v[b], v[4 + b]);
        v[b] = s;
        v[4 + b] = d;
    }
    dft4(v[0], v[1], v[2], v[3]);
    {
        const C2 a0 = v[4], p1 = w8u(v[5]), t2 = v[6], p3 = w8u3(v[7]);
        const C2 s0 = caddj(a0, t2), d0 = csubj(a0, t2);
        const C2 sU = cadd(p1, p3), dU = csub(p1, p3);
        v[4] = cfma(sU, r2, s0); v[6] = cfma(sU, -r2, s0); v[5] = cfmaj(dU, r2, d0); v[7] = cfmaj(dU, -r2, d0);
    }
}

// batches [b0, b1) of slab `slab`, in units of 16-pair batches over the n_pairs of the launch
__device__ __forceinline__ void slab_range(const SmallParams& p, int slab, long long& b0, long long& b1)
{
    const long long n_batches = (p.n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    b0 = (n_batches * slab) / p.n_slabs;
    b1 = (n_batches * (slab + 1)) / p.n_slabs;
}

template <int kM, int kTaps>
__device__ __forceinline__ void fir_role(const SmallParams& p, uint32_t smem, uint32_t mbar)
{
    using G = Geo<kM>;
    constexpr int kM2 = G::M2, kSlots = G::slots, kInSlotBytes = G::in_slot_bytes, kRegionBytes = G::region_bytes, kVSlotBytes = G::v_slot_bytes;
    constexpr int kHist = kTaps - 1;
    const int j = threadIdx.x;
    const int slot = j / kM, br = j % kM;
    const int pos = (br < kM2) ? (kM2 - 1 - br) : (kM + kM2 - 1 - br);
    const bool issuer = br == 0;
    long long b0, b1;
    slab_range(p, blockIdx.x * kSlots + slot, b0, b1);
    if (b0 >= b1) return;
    const float2* xf = p.x + p.f0 * kM2;
    const long long call_off = p.f0 * kM2;

    float2 T[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; i++) T[i] = __ldg(&p.taps[br * kTaps + i]);
    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);
    const long long q0 = b0 * kPairsPerBatch;
#pragma unroll
    for (int i = 1; i <= kHist; i++) {
        const long long ta = (q0 - i) * kM + pos + call_off;            // relative to x[0] of the call
        float2 v;
        if (ta >= 0) v = __ldg(&p.x[ta]);
        else if (p.Hlen + ta >= 0) v = __ldg(&p.hist[p.Hlen + ta]);
        else v = make_float2(0.f, 0.f);
        W[(32 - i) & 31] = v;
    }

    const uint32_t in0 = smem + slot * kInSlotBytes;
    const uint32_t vb0 = smem + 2 * kInStageBytes + slot * kVSlotBytes + br * 16;

    auto issue_load = [&](long long batch) {
        const int st = (int)((batch - b0) & 1);
        long long np = p.n_pairs - batch * kPairsPerBatch;
        if (np > kPairsPerBatch) np = kPairsPerBatch;
        const uint32_t bytes = (uint32_t)(np * kM * 8);
        const uint32_t bar = mbar + 8 * (kMbInFull + 4 * st + slot);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(in0 + st * kInStageBytes, xf + batch * (long long)(kPairsPerBatch * kM), bytes, bar);
    };
    if (issuer) {
        issue_load(b0);
        if (b0 + 1 < b1) issue_load(b0 + 1);
    }

    auto do_batch = [&](auto par_tag, long long batch) {
        constexpr int PAR = decltype(par_tag)::value;
        const long long lb = batch - b0;
        const uint32_t ph = (uint32_t)((lb >> 1) & 1);
        mbar_wait(mbar + 8 * (kMbInFull + 4 * PAR + slot), ph);
        const uint32_t in = in0 + PAR * kInStageBytes + pos * 8;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) W[16 * PAR + r] = lds64(in + r * (kM * 8));
        __syncwarp();
        if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbInFree + 4 * PAR + slot));
        if (lb >= 2) mbar_wait(mbar + 8 * (kMbVFree + 4 * PAR + slot), ph ^ 1);
        const uint32_t vout = vb0 + PAR * kVBufBytes;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) {
            float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {
                const float2 w = W[(16 * PAR + r - i) & 31];
                are = fma2(T[i], f2(w.x), are);
                aim = fma2(T[i], f2(w.y), aim);
            }
            sts128(vout + r * kRegionBytes, make_float4(are.x, are.y, aim.x, aim.y));
            if ((r & 3) == 3) {
                __syncwarp();
                if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbVFull + 16 * PAR + 4 * slot + (r >> 2)));
            }
        }
        if (issuer && batch + 2 < b1) {
            mbar_wait(mbar + 8 * (kMbInFree + 4 * PAR + slot), ph);
            issue_load(batch + 2);
        }
    };
    for (long long batch = b0; batch < b1; batch += 2) {
        do_batch(std::integral_constant<int, 0>{}, batch);
        if (batch + 1 < b1) do_batch(std::integral_constant<int, 1>{}, batch + 1);
    }
}

template <int kM>
__device__ __forceinline__ void fft_role(const SmallParams& p, uint32_t smem, uint32_t mbar)
{
    using G = Geo<kM>;
    constexpr int kSlots = G::slots, kRegionBytes = G::region_bytes, kVSlotBytes = G::v_slot_bytes;
    constexpr int kR1 = kM / 8;                   // first-pass radix: 8 (M = 64) or 16 (M = 128); 8 threads per pair
    const int tid = threadIdx.x - kFirThreads;
    const int slot = tid / kM;
    const int t = tid & 7;
    long long b0, b1;
    slab_range(p, blockIdx.x * kSlots + slot, b0, b1);
    if (b0 >= b1) return;

    float twr[kR1], twi[kR1];
#pragma unroll
    for (int kk = 0; kk < kR1; kk++) {
        const float2 tw = __ldg(&p.twid[t * kR1 + kk]);
        twr[kk] = tw.x;
        twi[kk] = tw.y;
    }
    const uint32_t vb0 = smem + 2 * kInStageBytes + slot * kVSlotBytes;

    for (long long batch = b0; batch < b1; batch++) {
        const long long lb = batch - b0;
        const int b = (int)(lb & 1);
        const uint32_t ph = (uint32_t)((lb >> 1) & 1);
        if constexpr (kM == 64) {
            const int wv = (tid >> 5) & 1;            // warp of the slot: regions 4wv..4wv+3, then 8+4wv..
            const int sub = (tid >> 3) & 3;
#pragma unroll
            for (int round = 0; round < 2; round++) {
                const int g = 2 * round + wv;
                const int pr = 4 * g + sub;
                const uint32_t region = vb0 + b * kVBufBytes + pr * kRegionBytes;
                mbar_wait(mbar + 8 * (kMbVFull + 16 * b + 4 * slot + g), ph);
                C2 v[8];
#pragma unroll
                for (int n1 = 0; n1 < 8; n1++) {
                    const float4 q4 = lds128(region + (8 * n1 + t) * 16);
                    v[n1].re = make_float2(q4.x, q4.y);
                    v[n1].im = make_float2(q4.z, q4.w);
                }
                dft8(v);
                __syncwarp();
#pragma unroll
                for (int k1 = 0; k1 < 8; k1++) {
                    C2 z = v[dr8(k1)];
                    if (k1 > 0) z = cmulw(z, twr[k1], twi[k1]);
                    sts128(region + (((t << 3) | (k1 ^ t)) << 4), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
                }
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 8; n2++) {
                    const float4 q4 = lds128(region + (((n2 << 3) | (t ^ n2)) << 4));
                    v[n2].re = make_float2(q4.x, q4.y);
                    v[n2].im = make_float2(q4.z, q4.w);
                }
                if (round == 1) {
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbVFree + 4 * b + slot));
                }
                dft8(v);
                const long long pair = batch * kPairsPerBatch + pr;
                if (pair < p.n_pairs) {
                    float2* ye = p.y + (p.f0 + 2 * pair) * (long long)kM + t;
#pragma unroll
                    for (int k2 = 0; k2 < 8; k2++) {
                        const C2 z = v[dr8(k2)];
                        __stcs(ye + 8 * k2, make_float2(z.re.x, z.im.x));
                        __stcs(ye + kM + 8 * k2, make_float2(z.re.y, z.im.y));
                    }
                }
            }
        } else {
            // M = 128 = 16 x 8: thread n2 = t runs a radix-16 over n1, then (as k1 = t and t + 8) two radix-8 over n2
            const int pr = (tid % kM) >> 3;           // pair inside the batch; its warp covers regions 4g..4g+3
            const uint32_t region = vb0 + b * kVBufBytes + pr * kRegionBytes;
            mbar_wait(mbar + 8 * (kMbVFull + 16 * b + 4 * slot + (pr >> 2)), ph);
            C2 v[16];
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                const float4 q4 = lds128(region + (8 * n1 + t) * 16);
                v[n1].re = make_float2(q4.x, q4.y);
                v[n1].im = make_float2(q4.z, q4.w);
            }
            dft16(v);
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) {
                C2 z = v[dr4(k1)];
                if (k1 > 0) z = cmulw(z, twr[k1], twi[k1]);
                sts128(region + (((t << 4) | (k1 ^ t)) << 4), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));   // row n2, 16 columns
            }
            __syncwarp();
            C2 va[8], vb[8];
#pragma unroll
            for (int n2 = 0; n2 < 8; n2++) {
                const float4 qa = lds128(region + (((n2 << 4) | (t ^ n2)) << 4));
                const float4 qb = lds128(region + (((n2 << 4) | ((t ^ n2) + 8)) << 4));
                va[n2].re = make_float2(qa.x, qa.y); va[n2].im = make_float2(qa.z, qa.w);
                vb[n2].re = make_float2(qb.x, qb.y); vb[n2].im = make_float2(qb.z, qb.w);
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbVFree + 4 * b + slot));
            dft8(va);
            dft8(vb);
            const long long pair = batch * kPairsPerBatch + pr;
            if (pair < p.n_pairs) {
                float2* ye = p.y + (p.f0 + 2 * pair) * (long long)kM + t;
#pragma unroll
                for (int k2 = 0; k2 < 8; k2++) {
                    const C2 za = va[dr8(k2)], zb = vb[dr8(k2)];
                    __stcs(ye + 16 * k2, make_float2(za.re.x, za.im.x));
                    __stcs(ye + 16 * k2 + 8, make_float2(zb.re.x, zb.im.x));
                    __stcs(ye + kM + 16 * k2, make_float2(za.re.y, za.im.y));
                    __stcs(ye + kM + 16 * k2 + 8, make_float2(zb.re.y, zb.im.y));
                }
            }
        }
    }
}

template <int kM, int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch2_analysis_small(const SmallParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + 2 * kInStageBytes + 2 * kVBufBytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; i++) {
            mbar_init(mbar + 8 * (kMbInFull + i), 1);
            mbar_init(mbar + 8 * (kMbInFree + i), Geo<kM>::warps_per_slot);
            mbar_init(mbar + 8 * (kMbVFree + i), Geo<kM>::warps_per_slot);
        }
        for (int i = 0; i < 32; i++) mbar_init(mbar + 8 * (kMbVFull + i), Geo<kM>::warps_per_slot);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < kFirThreads) fir_role<kM, kTaps>(p, smem, mbar);
    else fft_role<kM>(p, smem, mbar);
}

template <int kM, int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, SmallParams p, cudaStream_t st)
{
    constexpr int kSlots = Geo<kM>::slots;
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_analysis_small<kM, kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long n_batches = (p.n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    const int grid = (int)std::min<long long>(plan.n_sm, (n_batches + kSlots - 1) / kSlots);
    p.n_slabs = grid * kSlots;
    k_firpfbch2_analysis_small<kM, kTaps><<<grid, kThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

}  // namespace

int32_t firpfbch2_small_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if ((M != 64 && M != 128) || m < 1 || m > 8) return YG_OK;
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    const int iM = (int)M, iM2 = iM / 2;
    const int kTaps = 2 * (int)m + 1, P = 2 * (int)m;
    std::vector<float2> taps((size_t)iM * kTaps);
    const float s = 1.0f / (float)iM;
    for (int j = 0; j < iM; j++)
        for (int i = 0; i < kTaps; i++) {
            float te = 0.f, to = 0.f;
            if (j < iM2) {
                if (i < P) { te = h[j + i * iM]; to = h[j + iM2 + i * iM]; }
            } else {
                if (i >= 1) te = h[j + (i - 1) * iM];
                if (i < P) to = h[j - iM2 + i * iM];
            }
            taps[(size_t)j * kTaps + i] = make_float2(te * s, to * s);
        }
    const int R1 = iM / 8;                                   // twiddles e^{+j 2 pi n2 k1 / M}, n2 < 8, k1 < M/8
    std::vector<float2> tw((size_t)8 * R1);
    for (int n2 = 0; n2 < 8; n2++)
        for (int k1 = 0; k1 < R1; k1++) {
            const double a = 2.0 * M_PI * (double)(n2 * k1) / (double)iM;
            tw[(size_t)n2 * R1 + k1] = make_float2((float)cos(a), (float)sin(a));
        }
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float2), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.min_frames = 256;
    plan.supported = true;
    return YG_OK;
}

namespace {
template <int kM>
int32_t launch_m(const Firpfbch2FastPlan& plan, const SmallParams& p, cudaStream_t st)
{
    switch (plan.m) {
        case 1: return launch_t<kM, 3>(plan, p, st);
        case 2: return launch_t<kM, 5>(plan, p, st);
        case 3: return launch_t<kM, 7>(plan, p, st);
        case 4: return launch_t<kM, 9>(plan, p, st);
        case 5: return launch_t<kM, 11>(plan, p, st);
        case 6: return launch_t<kM, 13>(plan, p, st);
        case 7: return launch_t<kM, 15>(plan, p, st);
        case 8: return launch_t<kM, 17>(plan, p, st);
        default: return fail(YG_EINTERNAL, "fused kernel not instantiated for m = %u", plan.m);
    }
}
}  // namespace

int32_t firpfbch2_small_launch(const Firpfbch2FastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                               size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "small-M fused kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames & 1) return fail(YG_EINTERNAL, "fused kernel needs an even number of frames");
    if (((uintptr_t)(x + f0 * (plan.M / 2)) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    SmallParams p;
    p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_pairs = (long long)(n_frames / 2);
    p.n_slabs = 0;
    p.taps = reinterpret_cast<const float2*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    return plan.M == 64 ? launch_m<64>(plan, p, st) : launch_m<128>(plan, p, st);
}

}  // namespace yg
