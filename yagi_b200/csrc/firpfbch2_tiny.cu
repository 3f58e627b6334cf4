// firpfbch2_tiny.cu -- fused firpfbch2 ANALYSIS kernel for tiny M (M = 8, 16, 32), m = 1..8, sm_100a.
//
// With 32 branches or fewer a whole channelizer fits in one warp, so the unit of work is a PAIR OF WARPS:
// FIR warp u (32 / M time slabs side by side, one polyphase branch per lane: the register-ring / packed-FFMA2
// arithmetic of the M = 256 kernel) and DFT warp u, meeting in the unit's own double-buffered 8 KB V tile with
// the unit's own mbarriers.  A CTA is eight independent units; nothing is shared between them.
//   input:  the 16 pairs of a slab-batch are one contiguous run of samples: the FIR warp copies its runs for the next
//           batch cooperatively with 16-byte cp.async.cg (zero-filled past the end of the slab), double-buffered;
//   V:      packed {re_e, re_o, im_e, im_o} per (frame pair, branch), branch index XOR-swizzled with the pair so
//           that both the row-wise writes and the column-wise reads are bank-conflict free;
//   DFT:    M <= 16: ONE thread transforms a frame pair entirely in registers (radix 8 / 16, packed even/odd
//           lanes); M = 32: two threads (16 x 2: radix 16, twiddle, one shuffle exchange, radix 2);
//   output: every thread holds 8 or 16 consecutive bins of its two frames; they go back through the pair's own region
//           of the V tile so that the warp writes the batch (one contiguous run per slab) with 512-byte stores.
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
constexpr int kRoleThreads = 256;
constexpr int kUnits = 8;                                        // (FIR warp, DFT warp) pairs per CTA
constexpr int kInStageBytes = kRoleThreads * kPairsPerBatch * 8;  // 32 KB: one batch of input (4 KB per FIR warp)
constexpr int kVUnitBytes = 32 * kPairsPerBatch * 16;             // 8 KB: one batch of V of one unit
constexpr int kVOff = 2 * kInStageBytes;
constexpr int kMbar = kVOff + 2 * kUnits * kVUnitBytes;           // per unit: vfull[2], vfree[2]
constexpr int kSmemBytes = kMbar + kUnits * 32;

struct TinyParams {
    const float2* hist;       // Hlen samples preceding x[0] of the call
    long long Hlen;
    const float2* x;
    float2* y;
    long long f0;             // first frame handled here (even global parity)
    long long n_pairs;        // frame pairs handled here
    const float2* taps;       // [M][2m+1] (even, odd) tap pairs, 1/M folded in
    const float2* twid;       // [M] e^{+j 2 pi k / M}
};

// branch index inside the V row of pair r
template <int kM>
__device__ __forceinline__ int swz(int b, int r) { return b ^ (kM == 32 ? 2 * (r & 3) : (r & (kM - 1))); }

// batches [b0, b1) of slab `sl`
__device__ __forceinline__ void slab_range(long long n_pairs, int n_slabs, int sl, long long& b0, long long& b1)
{
    const long long n_batches = (n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    b0 = (n_batches * sl) / n_slabs;
    b1 = (n_batches * (sl + 1)) / n_slabs;
}

template <int kM, int kTaps>
__device__ __forceinline__ void fir_role(const TinyParams& p, uint32_t smem, int unit, long long nbw)
{
    constexpr int kM2 = kM / 2, kSPW = 32 / kM, kHist = kTaps - 1;
    const int lane = threadIdx.x & 31;
    const int sw = lane / kM, br = lane % kM;
    const int pos = (br < kM2) ? (kM2 - 1 - br) : (kM + kM2 - 1 - br);
    long long b0, b1;
    slab_range(p.n_pairs, (int)gridDim.x * kUnits * kSPW, ((int)blockIdx.x * kUnits + unit) * kSPW + sw, b0, b1);
    const uint32_t mb = smem + kMbar + unit * 32;

    float2 T[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; i++) T[i] = __ldg(&p.taps[br * kTaps + i]);
    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);
    const long long call_off = p.f0 * kM2;                               // sample of pair 0 relative to x[0]
    const long long q0 = b0 * kPairsPerBatch;
#pragma unroll
    for (int i = 1; i <= kHist; i++) {
        const long long ta = (q0 - i) * kM + pos + call_off;
        float2 v = make_float2(0.f, 0.f);
        if (b0 < b1) {
            if (ta >= 0) v = __ldg(&p.x[ta]);
            else if (p.Hlen + ta >= 0) v = __ldg(&p.hist[p.Hlen + ta]);
        }
        W[(32 - i) & 31] = v;
    }

    // input staging, double-buffered: the 16 pairs of a slab-batch are one contiguous run of 16 M samples, so the
    // warp copies its 32 / M runs cooperatively (16-byte cp.async.cg, 8 per lane and batch, zero-filled past the end
    // of the slab) and every lane then picks the sample of its branch out of row r
    constexpr int kChunksPerSlab = kPairsPerBatch * kM / 2;             // 16-byte chunks of one slab-batch
    const uint32_t stage_w = smem + (threadIdx.x & ~31) * (kPairsPerBatch * 8);          // the warp's 4 KB of a stage
    const uint32_t stage_rd = stage_w + sw * (kChunksPerSlab * 16) + pos * 8;           // + r * kM * 8
    long long sq0[kSPW], sqe[kSPW];                                     // per slab of the warp: first pair, end
#pragma unroll
    for (int s2 = 0; s2 < kSPW; s2++) {
        long long c0, c1;
        slab_range(p.n_pairs, (int)gridDim.x * kUnits * kSPW, ((int)blockIdx.x * kUnits + unit) * kSPW + s2, c0, c1);
        sq0[s2] = c0 * kPairsPerBatch;
        sqe[s2] = min(c1 * kPairsPerBatch, p.n_pairs);
    }
    auto prefetch = [&](long long lb, int st) {
#pragma unroll
        for (int s2 = 0; s2 < kSPW; s2++) {
            const long long q = sq0[s2] + lb * kPairsPerBatch;         // first pair of the run
#pragma unroll
            for (int c0 = 0; c0 < kChunksPerSlab; c0 += 32) {
                const int c = c0 + lane;
                const bool ok = q + c / (kM / 2) < sqe[s2];
                const float2* src = ok ? p.x + (q * kM + call_off + 2 * c) : p.x;
                const uint32_t bytes = ok ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(stage_w + st * kInStageBytes + s2 * (kChunksPerSlab * 16) + c * 16),
                             "l"(src), "r"(bytes) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (nbw > 0) prefetch(0, 0);

    const uint32_t vrow0 = smem + kVOff + unit * kVUnitBytes + (sw * kPairsPerBatch) * (kM * 16);
    auto do_batch = [&](auto par_tag, long long lb) {
        constexpr int PAR = decltype(par_tag)::value;
        if (lb + 1 < nbw) {
            prefetch(lb + 1, PAR ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                                   // every lane's chunks of the stage have landed
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) W[16 * PAR + r] = lds64(stage_rd + PAR * kInStageBytes + r * (kM * 8));
        if (lb >= 2) mbar_wait(mb + 8 * (2 + PAR), (uint32_t)(((lb >> 1) - 1) & 1));    // the DFT warp has drained V[PAR]
        const uint32_t vrow = vrow0 + PAR * (kUnits * kVUnitBytes);
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) {
            float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {                      // oldest sample first
                const float2 w = W[(16 * PAR + r - i) & 31];
                are = fma2(T[i], f2(w.x), are);
                aim = fma2(T[i], f2(w.y), aim);
            }
            sts128(vrow + (r * kM + swz<kM>(br, r)) * 16, make_float4(are.x, are.y, aim.x, aim.y));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * PAR);
    };
    for (long long lb = 0; lb < nbw; lb += 2) {
        do_batch(std::integral_constant<int, 0>{}, lb);
        if (lb + 1 < nbw) do_batch(std::integral_constant<int, 1>{}, lb + 1);
    }
}

// The NB consecutive bins X[0..NB) of both frames of a pair go back into the pair's own region of the V tile as
// 16-byte chunks (frame e: chunks c0 .., frame o: chunks kM/2 + c0 ..), chunk index XOR-swizzled with the pair.
template <int kM, int NB>
__device__ __forceinline__ void stash_bins(const C2* X, uint32_t ptile, int sx, int c0)
{
#pragma unroll
    for (int k = 0; k < NB; k += 2) {
        sts128(ptile + (((c0 + k / 2) ^ sx) * 16), make_float4(X[k].re.x, X[k].im.x, X[k + 1].re.x, X[k + 1].im.x));
        sts128(ptile + (((kM / 2 + c0 + k / 2) ^ sx) * 16), make_float4(X[k].re.y, X[k].im.y, X[k + 1].re.y, X[k + 1].im.y));
    }
}

template <int kM>
__device__ __forceinline__ void dft_role(const TinyParams& p, uint32_t smem, int unit, long long nbw)
{
    constexpr int kSPW = 32 / kM;
    constexpr int kTPP = (kM == 32) ? 2 : 1;                            // threads per frame pair
    constexpr int kPPT = kSPW * kPairsPerBatch * kTPP / 32;             // frame pairs per thread and batch (2 at M = 8)
    constexpr int kNV = kM / kTPP;                                      // values per thread and pair
    const int lane = threadIdx.x & 31;
    const int tt = (kTPP == 2) ? (lane & 1) : 0;
    const uint32_t mb = smem + kMbar + unit * 32;
    const int slab0 = ((int)blockIdx.x * kUnits + unit) * kSPW;
    const int n_slabs = (int)gridDim.x * kUnits * kSPW;

    int pi[kPPT];                                                       // pair of the unit-batch: slab sw = pi / 16, pair r = pi % 16
#pragma unroll
    for (int pp = 0; pp < kPPT; pp++) pi[pp] = (kTPP == 2) ? (lane >> 1) : lane + 32 * pp;
    long long sq0[kSPW], sqe[kSPW];                                     // per slab of the warp: first pair, end
#pragma unroll
    for (int sw = 0; sw < kSPW; sw++) {
        long long b0, b1;
        slab_range(p.n_pairs, n_slabs, slab0 + sw, b0, b1);
        sq0[sw] = b0 * kPairsPerBatch;
        sqe[sw] = min(b1 * kPairsPerBatch, p.n_pairs);
    }
    float twr[16], twi[16];                                             // M = 32: W_32^{tt k1}
    if (kM == 32) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 w = __ldg(&p.twid[tt * k]);
            twr[k] = w.x;
            twi[k] = w.y;
        }
    }

    for (long long lb = 0; lb < nbw; lb++) {
        const int b = (int)(lb & 1);
        const uint32_t vtile = smem + kVOff + b * (kUnits * kVUnitBytes) + unit * kVUnitBytes;
        mbar_wait(mb + 8 * b, (uint32_t)((lb >> 1) & 1));               // the FIR warp has written V[b]
        C2 v[kPPT][kNV];
#pragma unroll
        for (int pp = 0; pp < kPPT; pp++) {
            const int r = pi[pp] % kPairsPerBatch;
#pragma unroll
            for (int n = 0; n < kNV; n++) v[pp][n] = ldc2(vtile + (pi[pp] * kM + swz<kM>(kTPP * n + tt, r)) * 16);
        }
        __syncwarp();                                                   // (M = 32: the pair's other thread has read too)
#pragma unroll
        for (int pp = 0; pp < kPPT; pp++) {
            const uint32_t ptile = vtile + (pi[pp] * kM) * 16;          // the pair's own region: only its thread(s) read it
            const int sx = pi[pp] & (kM - 1);
            if constexpr (kM == 8) {
                dft_r<8>(v[pp]);
                stash_bins<kM, 8>(v[pp], ptile, sx, 0);
            } else if constexpr (kM == 16) {
                dft_r<16>(v[pp]);
                stash_bins<kM, 16>(v[pp], ptile, sx, 0);
            } else {
                // n = 2 n1 + tt: radix 16 over n1, twiddle W_32^{tt k1}, then X[k1 + 16 k2] = A_0[k1] + (-1)^k2 A_1[k1];
                // thread tt keeps k2 = tt: bins 16 tt .. 16 tt + 15
                dft_r<16>(v[pp]);
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    C2 a = v[pp][k];
                    if (k > 0) a = cmulw(a, twr[k], twi[k]);
                    C2 o;
                    o.re.x = __shfl_xor_sync(0xffffffffu, a.re.x, 1); o.re.y = __shfl_xor_sync(0xffffffffu, a.re.y, 1);
                    o.im.x = __shfl_xor_sync(0xffffffffu, a.im.x, 1); o.im.y = __shfl_xor_sync(0xffffffffu, a.im.y, 1);
                    v[pp][k] = tt ? csub(o, a) : cadd(a, o);
                }
                stash_bins<kM, 16>(v[pp], ptile, sx, 8 * tt);
            }
        }
        __syncwarp();
        // cooperative read-out: the 32 frames of a slab are one contiguous run of output, 512 bytes per warp store
#pragma unroll
        for (int g0 = 0; g0 < 32 * kPairsPerBatch; g0 += 32) {
            const int g = g0 + lane;
            const int pj = g / kM, c = g % kM;                          // pair of the unit-batch, chunk inside the pair
            const int sw = pj / kPairsPerBatch;
            const long long q = sq0[sw] + lb * kPairsPerBatch + (pj % kPairsPerBatch);
            const float4 z = lds128(vtile + (pj * kM + (c ^ (pj & (kM - 1)))) * 16);
            if (q < sqe[sw]) {
                float2* dst = p.y + (p.f0 + 2 * q + (c >= kM / 2 ? 1 : 0)) * (long long)kM + 2 * (c % (kM / 2));
                __stcs(reinterpret_cast<float4*>(dst), z);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * (2 + b));                   // V[b] may be overwritten
    }
}

template <int kM, int kTaps>
__global__ void __launch_bounds__(2 * kRoleThreads, 1) k_firpfbch2_analysis_tiny(const TinyParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    constexpr int kSPW = 32 / kM;
    if (threadIdx.x == 0) {
        for (int q = 0; q < 4 * kUnits; q++) mbar_init(smem + kMbar + 8 * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int unit = (threadIdx.x >> 5) & (kUnits - 1);
    // both warps of a unit run the batch count of the unit's longest slab; shorter slabs pad (zero input, no output)
    long long nbw = 0;
    for (int sw = 0; sw < kSPW; sw++) {
        long long b0, b1;
        slab_range(p.n_pairs, (int)gridDim.x * kUnits * kSPW, ((int)blockIdx.x * kUnits + unit) * kSPW + sw, b0, b1);
        nbw = max(nbw, b1 - b0);
    }
    if (threadIdx.x < kRoleThreads) fir_role<kM, kTaps>(p, smem, unit, nbw);
    else dft_role<kM>(p, smem, unit, nbw);
}

template <int kM, int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const TinyParams& p, cudaStream_t st)
{
    constexpr int kSPW = 32 / kM;
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_analysis_tiny<kM, kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long n_batches = (p.n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    const int grid = (int)std::max<long long>(1, std::min<long long>(plan.n_sm, (n_batches + kUnits * kSPW - 1) / (kUnits * kSPW)));
    k_firpfbch2_analysis_tiny<kM, kTaps><<<grid, 2 * kRoleThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

template <int kM>
int32_t launch_m(const Firpfbch2FastPlan& plan, const TinyParams& p, cudaStream_t st)
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
        default: return fail(YG_EINTERNAL, "tiny-M kernel not instantiated for m = %u", plan.m);
    }
}

}  // namespace

int32_t firpfbch2_tiny_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if ((M != 8 && M != 16 && M != 32) || m < 1 || m > 8) return YG_OK;
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    // tap pairs (Te[i], To[i]) per branch, 1/M folded in: the layout of the M = 256 kernel (firpfbch2_fast.cu)
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
    std::vector<float2> tw(M);
    for (uint32_t k = 0; k < M; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float2), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.min_frames = 2048;
    plan.supported = true;
    return YG_OK;
}

int32_t firpfbch2_tiny_launch(const Firpfbch2FastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                              size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "tiny-M fused kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames & 1) return fail(YG_EINTERNAL, "fused kernel needs an even number of frames");
    if ((reinterpret_cast<uintptr_t>(y) & 15) != 0) return fail(YG_EVALUE, "output pointer must be 16-byte aligned");
    TinyParams p;
    p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_pairs = (long long)(n_frames / 2);
    p.taps = reinterpret_cast<const float2*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    switch (plan.M) {
        case 8: return launch_m<8>(plan, p, st);
        case 16: return launch_m<16>(plan, p, st);
        default: return launch_m<32>(plan, p, st);
    }
}

}  // namespace yg
