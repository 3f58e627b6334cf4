// firpfbch2_tiny_synth.cu -- fused firpfbch2 SYNTHESIS kernel for tiny M (M = 8, 16, 32), m = 1..7, sm_100a.
//
//   y[k M/2 + i] = sum_{l < 4m} h[i + l M/2] u_{k-l}[(i + (k&1) M/2) mod M],   u_k = 1/2 IDFT_unnorm(X_k)
//
// Mirror image of firpfbch2_tiny.cu: a CTA is eight independent units of two warps that meet in the unit's own
// double-buffered 8 KB U tile (own mbarriers).  Each unit walks 32 / M time slabs side by side, every slab starting
// one warm-up batch (32 frames) early with its outputs suppressed, so the object's state is the last 32 INPUT frames.
//   DFT warp:  the batch's input frames (two contiguous 4 KB / M-slab chunks... one contiguous run per slab) are
//              copied cooperatively with 16-byte cp.async, one batch ahead, into a swizzled tile from which ONE
//              thread per frame pair (two at M = 32) reads its frames conflict-free and transforms them in
//              registers on packed (even, odd) lanes; U goes to the tile as {re_e, re_o, im_e, im_o} per column;
//   overlap-add warp: lane = (slab, column j): the last 4m frames of its column in a 32-entry register ring, one
//              packed FFMA2 per tap; columns j < M/2 emit on even frames, j >= M/2 on odd ones (ring one frame
//              behind).
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kPairs = 16;                                   // frame pairs per batch and slab (= 32 frames)
constexpr int kRoleThreads = 256;
constexpr int kUnits = 8;
constexpr int kXUnitBytes = 32 * kPairs * 16;                // 8 KB: the input frames of one unit-batch
constexpr int kUUnitBytes = 32 * kPairs * 16;                // 8 KB: U of one unit-batch
constexpr int kUOff = kUnits * kXUnitBytes;
constexpr int kMbar = kUOff + 2 * kUnits * kUUnitBytes;      // per unit: ufull[2], ufree[2]
constexpr int kSmemBytes = kMbar + kUnits * 32;

struct TinySynthParams {
    const float2* prefix;     // the 32 input frames preceding x[0]
    const float2* x;          // input frames of the call, [frame][M]
    float2* y;                // output sample 0 of the call
    long long f0;             // first frame handled (even global parity)
    long long n_batches;      // output batches of 32 frames
    const float* taps;        // [M][4m]  0.5 * h[(j & (M/2-1)) + l * M/2]
    const float2* twid;       // [M] e^{+j 2 pi k / M}
};

template <int kM>
__device__ __forceinline__ int swz(int b, int r) { return b ^ (kM == 32 ? 2 * (r & 3) : (r & (kM - 1))); }

__device__ __forceinline__ void slab_range(long long n_batches, int n_slabs, int sl, long long& B0, long long& B1)
{
    B0 = (n_batches * sl) / n_slabs;
    B1 = (n_batches * (sl + 1)) / n_slabs;
}

template <int kM>
__device__ __forceinline__ void dft_role(const TinySynthParams& p, uint32_t smem, int unit, long long nbw)
{
    constexpr int kSPW = 32 / kM;
    constexpr int kTPP = (kM == 32) ? 2 : 1;                  // threads per frame pair
    constexpr int kPPT = kSPW * kPairs * kTPP / 32;           // frame pairs per thread and batch (2 at M = 8)
    constexpr int kNV = kM / kTPP;                            // values per thread and pair
    constexpr int kCPP = 2 * kM * 8 / 16;                     // 16-byte chunks per frame pair (M)
    constexpr int kCPS = kPairs * kCPP;                       // chunks per slab-batch (contiguous in global memory)
    const int lane = threadIdx.x & 31;
    const int tt = (kTPP == 2) ? (lane & 1) : 0;
    const uint32_t mb = smem + kMbar + unit * 32;
    const int slab0 = ((int)blockIdx.x * kUnits + unit) * kSPW;
    const int n_slabs = (int)gridDim.x * kUnits * kSPW;
    const uint32_t xtile = smem + unit * kXUnitBytes;

    long long B0s[kSPW], nbs[kSPW];                            // per slab of the warp: first batch, batches incl. warm-up
#pragma unroll
    for (int sw = 0; sw < kSPW; sw++) {
        long long B0, B1;
        slab_range(p.n_batches, n_slabs, slab0 + sw, B0, B1);
        B0s[sw] = B0;
        nbs[sw] = (B1 > B0) ? B1 - B0 + 1 : 0;
    }
    // cooperative copy of one batch: the 32 frames of slab sw are one contiguous run of kCPS 16-byte chunks (frames
    // vi .. vi + 31; an odd call-relative start makes one PAIR straddle prefix | x, never a frame); chunk c of the
    // run belongs to pair c / kCPP, position i = c % kCPP and lands in slot (pair, i ^ (pair % kCPP... )) of the tile
    auto fetch_x = [&](long long lb) {
#pragma unroll
        for (int sw = 0; sw < kSPW; sw++) {
            if (lb >= nbs[sw]) continue;
            const long long vi = p.f0 + 32 * (B0s[sw] + lb - 1);
#pragma unroll
            for (int c0 = 0; c0 < kCPS; c0 += 32) {
                const int c = c0 + lane;
                const int pr = c / kCPP, i = c % kCPP;        // pair of the slab-batch, chunk inside the pair
                const long long fr = vi + 2 * pr + (i >= kCPP / 2 ? 1 : 0);           // frame of this chunk
                const int ci = i % (kCPP / 2);                // chunk inside the frame
                const float2* src = ((fr < 0) ? p.prefix + (32 + fr) * kM : p.x + fr * kM) + 2 * ci;
                const int pi = sw * kPairs + pr;
                const uint32_t dst = xtile + (pi * kCPP + (i ^ (pi % kCPP))) * 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
    };
    int pi[kPPT];
#pragma unroll
    for (int pp = 0; pp < kPPT; pp++) pi[pp] = (kTPP == 2) ? (lane >> 1) : lane + 32 * pp;
    float twr[16], twi[16];                                   // M = 32: W_32^{tt k1}
    if (kM == 32) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 w = __ldg(&p.twid[tt * k]);
            twr[k] = w.x;
            twi[k] = w.y;
        }
    }

    if (nbw > 0) fetch_x(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (long long lb = 0; lb < nbw; lb++) {
        const int b = (int)(lb & 1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                        // every lane's chunks of the tile have landed
        C2 v[kPPT][kNV];
#pragma unroll
        for (int pp = 0; pp < kPPT; pp++) {
            // sample n of frame e lives in chunk n / 2 (half n % 2), of frame o in chunk kCPP / 2 + n / 2
            const uint32_t prow = xtile + (pi[pp] * kCPP) * 16;
            const int sx = pi[pp] % kCPP;
            if (kTPP == 1) {
#pragma unroll
                for (int c = 0; c < kM / 2; c++) {
                    const float4 e = lds128(prow + ((c ^ sx) * 16));
                    const float4 o = lds128(prow + (((kCPP / 2 + c) ^ sx) * 16));
                    v[pp][2 * c].re = make_float2(e.x, o.x);     v[pp][2 * c].im = make_float2(e.y, o.y);
                    v[pp][2 * c + 1].re = make_float2(e.z, o.z); v[pp][2 * c + 1].im = make_float2(e.w, o.w);
                }
            } else {
#pragma unroll
                for (int n = 0; n < kNV; n++) {              // n-th value of this thread: sample 2 n + tt
                    const float2 e = lds64(prow + ((n ^ sx) * 16) + tt * 8);
                    const float2 o = lds64(prow + (((kCPP / 2 + n) ^ sx) * 16) + tt * 8);
                    v[pp][n].re = make_float2(e.x, o.x);
                    v[pp][n].im = make_float2(e.y, o.y);
                }
            }
        }
        __syncwarp();                                        // tile consumed: refill it a batch ahead
        if (lb + 1 < nbw) fetch_x(lb + 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (lb >= 2) mbar_wait(mb + 8 * (2 + b), (uint32_t)(((lb >> 1) - 1) & 1));      // overlap-add warp drained U[b]
        const uint32_t utile = smem + kUOff + b * (kUnits * kUUnitBytes) + unit * kUUnitBytes;
#pragma unroll
        for (int pp = 0; pp < kPPT; pp++) {
            const int r = pi[pp] % kPairs;
            const uint32_t urow = utile + (pi[pp] * kM) * 16;
            if constexpr (kM == 8) {
                dft_r<8>(v[pp]);
#pragma unroll
                for (int k = 0; k < 8; k++) stc2(urow + swz<kM>(k, r) * 16, v[pp][k]);
            } else if constexpr (kM == 16) {
                dft_r<16>(v[pp]);
#pragma unroll
                for (int k = 0; k < 16; k++) stc2(urow + swz<kM>(k, r) * 16, v[pp][k]);
            } else {
                dft_r<16>(v[pp]);                            // n = 2 n1 + tt; then U[k1 + 16 k2] = A_0[k1] + (-1)^k2 A_1[k1]
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    C2 a = v[pp][k];
                    if (k > 0) a = cmulw(a, twr[k], twi[k]);
                    C2 o;
                    o.re.x = __shfl_xor_sync(0xffffffffu, a.re.x, 1); o.re.y = __shfl_xor_sync(0xffffffffu, a.re.y, 1);
                    o.im.x = __shfl_xor_sync(0xffffffffu, a.im.x, 1); o.im.y = __shfl_xor_sync(0xffffffffu, a.im.y, 1);
                    stc2(urow + swz<kM>(k + 16 * tt, r) * 16, tt ? csub(o, a) : cadd(a, o));
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * b);
    }
}

template <int kM, int kTaps>
__device__ __forceinline__ void wola_role(const TinySynthParams& p, uint32_t smem, int unit, long long nbw)
{
    constexpr int kM2 = kM / 2, kSPW = 32 / kM;
    const int lane = threadIdx.x & 31;
    const int sw = lane / kM, j = lane % kM;
    const bool hi = j >= kM2;
    const int i = j & (kM2 - 1);
    long long B0, B1;
    slab_range(p.n_batches, (int)gridDim.x * kUnits * kSPW, ((int)blockIdx.x * kUnits + unit) * kSPW + sw, B0, B1);
    const long long n_out = 32 * (B1 - B0);
    const long long nb = (B1 > B0) ? B1 - B0 + 1 : 0;          // batches of this slab, warm-up included
    const uint32_t mb = smem + kMbar + unit * 32;

    float T[kTaps];
#pragma unroll
    for (int l = 0; l < kTaps; l++) T[l] = __ldg(&p.taps[j * kTaps + l]);
    float2 W[32];
#pragma unroll
    for (int q = 0; q < 32; q++) W[q] = make_float2(0.f, 0.f);
    float2* yb = p.y + (p.f0 + 32 * B0) * (long long)kM2 + i;  // first real output frame of the slab
    float2 carry = make_float2(0.f, 0.f);                     // odd frame of the previous pair (upper half)

    auto window = [&](int newest) {                           // two banks (even / odd l), each oldest first, then added
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = kTaps - 1; l >= 0; l--) {
            const float2 w = W[(newest - l) & 31];
            if (l & 1) a1 = fma2(w, f2(T[l]), a1);
            else a0 = fma2(w, f2(T[l]), a0);
        }
        return add2(a0, a1);
    };
    for (long long lb = 0; lb < nbw; lb++) {
        const int b = (int)(lb & 1);
        const uint32_t urow = smem + kUOff + b * (kUnits * kUUnitBytes) + unit * kUUnitBytes + (sw * kPairs) * (kM * 16);
        mbar_wait(mb + 8 * b, (uint32_t)((lb >> 1) & 1));     // the DFT warp has written U[b]
#pragma unroll
        for (int ss = 0; ss < 16; ss++) {
            const float4 u = lds128(urow + (ss * kM + swz<kM>(j, ss)) * 16);
            // lower half: slots (2ss, 2ss+1) = frames (2q, 2q+1) of the slab; upper half: frames (2q-1, 2q)
            W[(2 * ss) & 31] = hi ? carry : make_float2(u.x, u.z);
            W[(2 * ss + 1) & 31] = hi ? make_float2(u.x, u.z) : make_float2(u.y, u.w);
            carry = make_float2(u.y, u.w);
            const float2 r = window(2 * ss);
            const long long rel = 2 * (16 * lb + ss) + (hi ? -1 : 0) - 32;   // frame relative to the slab's first real frame
            if (rel >= 0 && rel < n_out) __stcs(yb + rel * kM2, r);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * (2 + b));
        // last odd frame of THIS slab (a shorter slab of the warp keeps looping on padding batches): ends at slot 0
        if (hi && lb == nb - 1) {
            W[0] = carry;
            __stcs(yb + (n_out - 1) * kM2, window(0));
        }
    }
}

template <int kM, int kTaps>
__global__ void __launch_bounds__(2 * kRoleThreads, 1) k_firpfbch2_synthesis_tiny(const TinySynthParams p)
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
    long long nbw = 0;                                        // batches of the unit's longest slab, warm-up included
    for (int sw = 0; sw < kSPW; sw++) {
        long long B0, B1;
        slab_range(p.n_batches, (int)gridDim.x * kUnits * kSPW, ((int)blockIdx.x * kUnits + unit) * kSPW + sw, B0, B1);
        if (B1 > B0) nbw = max(nbw, B1 - B0 + 1);
    }
    if (threadIdx.x < kRoleThreads) wola_role<kM, kTaps>(p, smem, unit, nbw);
    else dft_role<kM>(p, smem, unit, nbw);
}

template <int kM, int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const TinySynthParams& p, cudaStream_t st)
{
    constexpr int kSPW = 32 / kM;
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_synthesis_tiny<kM, kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const int grid = (int)std::max<long long>(1, std::min<long long>(plan.n_sm, (p.n_batches + kUnits * kSPW - 1) / (kUnits * kSPW)));
    k_firpfbch2_synthesis_tiny<kM, kTaps><<<grid, 2 * kRoleThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

template <int kM>
int32_t launch_m(const Firpfbch2FastPlan& plan, const TinySynthParams& p, cudaStream_t st)
{
    switch (plan.m) {
        case 1: return launch_t<kM, 4>(plan, p, st);
        case 2: return launch_t<kM, 8>(plan, p, st);
        case 3: return launch_t<kM, 12>(plan, p, st);
        case 4: return launch_t<kM, 16>(plan, p, st);
        case 5: return launch_t<kM, 20>(plan, p, st);
        case 6: return launch_t<kM, 24>(plan, p, st);
        case 7: return launch_t<kM, 28>(plan, p, st);
        default: return fail(YG_EINTERNAL, "tiny-M synthesis kernel not instantiated for m = %u", plan.m);
    }
}

}  // namespace

int32_t firpfbch2_tiny_synth_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if ((M != 8 && M != 16 && M != 32) || m < 1 || m > 7) return YG_OK;
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    const int iM = (int)M, iM2 = iM / 2, kTaps = 4 * (int)m;
    std::vector<float> taps((size_t)iM * kTaps);
    for (int j = 0; j < iM; j++)
        for (int l = 0; l < kTaps; l++) taps[(size_t)j * kTaps + l] = 0.5f * h[(j & (iM2 - 1)) + l * iM2];
    std::vector<float2> tw(M);
    for (uint32_t k = 0; k < M; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.min_frames = 2048;
    plan.supported = true;
    return YG_OK;
}

int32_t firpfbch2_tiny_synth_launch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, float2* y,
                                    size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "tiny-M synthesis kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames % 32) return fail(YG_EINTERNAL, "tiny-M synthesis kernel needs a multiple of 32 frames");
    TinySynthParams p;
    p.prefix = prefix; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_batches = (long long)(n_frames / 32);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    switch (plan.M) {
        case 8: return launch_m<8>(plan, p, st);
        case 16: return launch_m<16>(plan, p, st);
        default: return launch_m<32>(plan, p, st);
    }
}

}  // namespace yg
