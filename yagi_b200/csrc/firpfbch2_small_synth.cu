// firpfbch2_small_synth.cu -- fused firpfbch2 synthesis kernel for small M (M = 64, 128; m = 1..7), sm_100a.
//
//   y[k M/2 + i] = sum_{l < 4m} h[i + l M/2] u_{k-l}[(i + (k&1) M/2) mod M],   u_k = 1/2 IDFT_unnorm(X_k)
//
// A 256-column CTA is too wide for one small-M stream, so every persistent CTA walks S = 256 / M independent
// time slabs side by side (the firpfbch2_small.cu idea), each with its own mbarriers: M / 32 warps of each of
// the two roles per slab, meeting in the slab's M columns of a double-buffered shared U batch of 16 frame pairs:
//
//   DFT role (warps 8-15): a team of T = M / 16 threads owns one frame pair of one slab per batch.  Its input
//     (2 x 16 samples per thread) is staged one batch ahead with cp.async.cg, 16 bytes shared by a lane pair.
//     Packed (even, odd) transform M = 16 x T: radix 16 in registers, twiddle, one XOR-swizzled exchange through
//     the team's own segment of the U row it is about to write, radix T, then U -> the same segment.
//   overlap-add role (warps 0-7): thread (slab, column j) keeps the last 4m frames of its column in a
//     32-entry register ring, one packed FFMA2 per tap; columns j < M/2 emit on even frames, j >= M/2 on odd
//     frames (the upper half keeps its ring one frame behind).
//
// U never touches global memory, so a slab cannot read its filter history: every slab starts one warm-up
// batch (32 frames) early, outputs suppressed; the object's state is the last 32 INPUT frames (`prefix`).
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kRoleThreads = 256;
constexpr int kPairs = 16;                                   // frame pairs per batch and slab (= 32 frames)
constexpr int kXStageBytes = kRoleThreads * 32 * 8;          // 64 KB: one pair (2 x 16 samples) per DFT thread column
constexpr int kURowBytes = kRoleThreads * 16 + 64;           // one pair row of U; +64 keeps two teams of a quarter-warp apart
constexpr int kUBufBytes = kPairs * kURowBytes;
constexpr int kMbar = kXStageBytes + 2 * kUBufBytes;         // per slab s < 4: ufull[s][2], ufree[s][2]
constexpr int kSmemBytes = kMbar + 4 * 32;

struct SmallSynthParams {
    const float2* prefix;     // the 32 input frames preceding x[0]
    const float2* x;          // input frames of the call, [frame][M]
    float2* y;                // output sample 0 of the call
    long long f0;             // first frame handled (even global parity)
    long long n_batches;      // output batches of 32 frames
    const float* taps;        // [M][4m]  0.5 * h[(j & (M/2-1)) + l * M/2]
    const float2* twid;       // [M] e^{+j 2 pi k / M}
};

// batches [B0, B1) of slab `sl` out of `n_slabs`
__device__ __forceinline__ void slab_range(long long n_batches, int n_slabs, int sl, long long& B0, long long& B1)
{
    B0 = (n_batches * sl) / n_slabs;
    B1 = (n_batches * (sl + 1)) / n_slabs;
}

template <int kM>
__device__ __forceinline__ void dft_role(const SmallSynthParams& p, uint32_t smem)
{
    constexpr int T = kM / 16, S = kRoleThreads / kM, I = 16 / T;
    const int lane = threadIdx.x & 31;
    const int dt = threadIdx.x - kRoleThreads;
    const int team = dt / T, tt = dt % T;
    const int s = team / kPairs, pr = team % kPairs;          // slab of the CTA, pair of the batch
    long long B0, B1;
    slab_range(p.n_batches, (int)gridDim.x * S, (int)blockIdx.x * S + s, B0, B1);
    const long long nb = (B1 > B0) ? B1 - B0 + 1 : 0;         // batches of this slab, warm-up included
    const uint32_t mb = smem + kMbar + s * 32;                // the slab's own barriers: slabs never wait for each other

    // staging: row 2 n1 + parity of this thread's column holds sample T n1 + tt of that frame of the pair; a lane
    // pair shares 16-byte copies (even lane: rows 0-15, odd lane: rows 16-31)
    const uint32_t xcol = smem + dt * 8;
    const uint32_t xcol_wr = smem + (dt & ~1) * 8 + (dt & 1) * (16 * kRoleThreads * 8);
    auto fetch_x = [&](long long lb) {
        // frames vi, vi + 1; with an odd call-relative start the pair can straddle the prefix | x boundary
        const long long vi = p.f0 + 32 * (B0 + lb - 1) + 2 * pr;
        const long long off = (tt & ~1) + (dt & 1) * (8 * T);
        const float2* se = ((vi < 0) ? p.prefix + (32 + vi) * kM : p.x + vi * kM) + off;
        const float2* so = ((vi + 1 < 0) ? p.prefix + (33 + vi) * kM : p.x + (vi + 1) * kM) + off;
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xcol_wr + (2 * n1) * (kRoleThreads * 8)), "l"(se + T * n1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xcol_wr + (2 * n1 + 1) * (kRoleThreads * 8)), "l"(so + T * n1) : "memory");
        }
    };
    float2 tw[16];                                            // W_M^{tt k1}
#pragma unroll
    for (int k = 0; k < 16; k++) tw[k] = __ldg(&p.twid[tt * k]);

    if (nb > 0) fetch_x(0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (long long lb = 0; lb < nb; lb++) {
        const int b = (int)(lb & 1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                        // the partner lane's rows have landed too
        C2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const float2 e = lds64(xcol + (2 * n1) * (kRoleThreads * 8)), o = lds64(xcol + (2 * n1 + 1) * (kRoleThreads * 8));
            v[n1].re = make_float2(e.x, o.x);
            v[n1].im = make_float2(e.y, o.y);
        }
        dft16(v);
        __syncwarp();                                        // every lane has consumed its staged samples:
        if (lb + 1 < nb) fetch_x(lb + 1);                    // refill the columns a whole batch ahead of use
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (lb >= 2) mbar_wait(mb + 8 * (2 + b), (uint32_t)(((lb >> 1) - 1) & 1));      // overlap-add role drained U[b]
        // The team's own segment of its U row (M entries of 16 bytes) doubles as the exchange tile: entry
        // (k1, n2) at slot k1 T + (n2 ^ (k1 % T)), conflict-free for the writes (fixed k1) and the gathers.
        const uint32_t useg = smem + kXStageBytes + b * kUBufBytes + pr * kURowBytes + (s * kM) * 16;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            C2 z = v[dr4(k)];
            if (k > 0) z = cmulw(z, tw[k].x, tw[k].y);
            stc2(useg + (k * T + (tt ^ (k % T))) * 16, z);
        }
        __syncwarp();
        // thread tt gathers, for its k1 = tt + T i, the T team-mates' values
#pragma unroll
        for (int i = 0; i < I; i++)
#pragma unroll
            for (int n2 = 0; n2 < T; n2++) v[i * T + n2] = ldc2(useg + ((tt + T * i) * T + (n2 ^ tt)) * 16);
        __syncwarp();                                        // tile consumed: it can take the final U
#pragma unroll
        for (int i = 0; i < I; i++) dft_r<T>(&v[i * T]);     // v[i T + k2] = U[tt + T i + 16 k2]
#pragma unroll
        for (int i = 0; i < I; i++)
#pragma unroll
            for (int k2 = 0; k2 < T; k2++) stc2(useg + (tt + T * i + 16 * k2) * 16, v[i * T + k2]);
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * b);
    }
}

template <int kM, int kTaps>
__device__ __forceinline__ void wola_role(const SmallSynthParams& p, uint32_t smem)
{
    constexpr int kM2 = kM / 2, S = kRoleThreads / kM;
    const int lane = threadIdx.x & 31;
    const int s = threadIdx.x / kM, j = threadIdx.x % kM;
    const bool hi = j >= kM2;                                 // warp-uniform (M/2 >= 32)
    const int i = j & (kM2 - 1);
    long long B0, B1;
    slab_range(p.n_batches, (int)gridDim.x * S, (int)blockIdx.x * S + s, B0, B1);
    const long long n_out = 32 * (B1 - B0);
    const long long nb = (B1 > B0) ? B1 - B0 + 1 : 0;          // batches of this slab, warm-up included
    const uint32_t mb = smem + kMbar + s * 32;

    float T[kTaps];
#pragma unroll
    for (int l = 0; l < kTaps; l++) T[l] = __ldg(&p.taps[j * kTaps + l]);
    float2 W[32];
#pragma unroll
    for (int q = 0; q < 32; q++) W[q] = make_float2(0.f, 0.f);
    float2* yb = p.y + (p.f0 + 32 * B0) * (long long)kM2 + i;  // first real output frame of the slab
    float2 carry = make_float2(0.f, 0.f);                     // odd frame of the previous pair (upper half)
    const uint32_t ucol = smem + kXStageBytes + threadIdx.x * 16;

    for (long long lb = 0; lb < nb; lb++) {
        const int b = (int)(lb & 1);
        mbar_wait(mb + 8 * b, (uint32_t)((lb >> 1) & 1));     // the slab's DFT warps have written U[b]
#pragma unroll
        for (int ss = 0; ss < 16; ss++) {
            const float4 u = lds128(ucol + b * kUBufBytes + ss * kURowBytes);
            // lower half: slots (2ss, 2ss+1) = frames (2q, 2q+1) of the slab; upper half: frames (2q-1, 2q)
            W[(2 * ss) & 31] = hi ? carry : make_float2(u.x, u.z);
            W[(2 * ss + 1) & 31] = hi ? make_float2(u.x, u.z) : make_float2(u.y, u.w);
            carry = make_float2(u.y, u.w);
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int l = kTaps - 1; l >= 0; l--) {
                const float2 w = W[(2 * ss - l) & 31];
                if (l & 1) a1 = fma2(w, f2(T[l]), a1);
                else a0 = fma2(w, f2(T[l]), a0);
            }
            const long long rel = 2 * (16 * lb + ss) + (hi ? -1 : 0) - 32;   // frame relative to the slab's first real frame
            if (rel >= 0 && rel < n_out) __stcs(yb + rel * kM2, add2(a0, a1));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * (2 + b));
    }
    if (hi && nb > 0) {                                       // last odd frame of the slab: window ends at slot 0
        W[0] = carry;
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = kTaps - 1; l >= 0; l--) {
            const float2 w = W[(0 - l) & 31];
            if (l & 1) a1 = fma2(w, f2(T[l]), a1);
            else a0 = fma2(w, f2(T[l]), a0);
        }
        __stcs(yb + (n_out - 1) * kM2, add2(a0, a1));
    }
}

template <int kM, int kTaps>
__global__ void __launch_bounds__(2 * kRoleThreads, 1) k_firpfbch2_synthesis_small(const SmallSynthParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    constexpr int S = kRoleThreads / kM;
    if (threadIdx.x == 0) {
        const uint32_t mb = smem + kMbar;
        for (int q = 0; q < 4 * S; q++) mbar_init(mb + 8 * q, kM / 32);  // per slab: one arrival per warp of the other role
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < kRoleThreads) wola_role<kM, kTaps>(p, smem);
    else dft_role<kM>(p, smem);
}

template <int kM, int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const SmallSynthParams& p, cudaStream_t st)
{
    constexpr int S = kRoleThreads / kM;
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_synthesis_small<kM, kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const int grid = (int)std::max<long long>(1, std::min<long long>(plan.n_sm, (p.n_batches + S - 1) / S));
    k_firpfbch2_synthesis_small<kM, kTaps><<<grid, 2 * kRoleThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

template <int kM>
int32_t launch_m(const Firpfbch2FastPlan& plan, const SmallSynthParams& p, cudaStream_t st)
{
    switch (plan.m) {
        case 1: return launch_t<kM, 4>(plan, p, st);
        case 2: return launch_t<kM, 8>(plan, p, st);
        case 3: return launch_t<kM, 12>(plan, p, st);
        case 4: return launch_t<kM, 16>(plan, p, st);
        case 5: return launch_t<kM, 20>(plan, p, st);
        case 6: return launch_t<kM, 24>(plan, p, st);
        case 7: return launch_t<kM, 28>(plan, p, st);
        default: return fail(YG_EINTERNAL, "small-M synthesis kernel not instantiated for m = %u", plan.m);
    }
}

}  // namespace

int32_t firpfbch2_small_synth_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if ((M != 64 && M != 128) || m < 1 || m > 7) return YG_OK;
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
    plan.min_frames = 256;
    plan.supported = true;
    return YG_OK;
}

int32_t firpfbch2_small_synth_launch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, float2* y,
                                     size_t f0, size_t n_frames, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "small-M synthesis kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames % 32) return fail(YG_EINTERNAL, "small-M synthesis kernel needs a multiple of 32 frames");
    SmallSynthParams p;
    p.prefix = prefix; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_batches = (long long)(n_frames / 32);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    return plan.M == 64 ? launch_m<64>(plan, p, st) : launch_m<128>(plan, p, st);
}

}  // namespace yg
