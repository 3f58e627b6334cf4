// firpfbch2_fast.cu -- fused firpfbch2 analysis kernel for sm_100a (M = 256).
//
// One persistent, warp-specialised CTA per SM walks a contiguous slab of frame pairs:
//
//   TMA bulk copy          FIR role (warps 0-7)                FFT role (warps 8-15)
//   global -> smem  ---->  thread j owns polyphase branch j:   16 threads per frame PAIR:
//   (32 KB / batch,        its 2m(+1) most recent samples      radix-16 x radix-16 backward DFT
//    mbarrier tx)          live in a 32-entry register ring;   of both frames of the pair at once
//                          packed FFMA2 computes the even and  (SoA (even,odd) float2 lanes, packed
//                          odd frame of a pair together  --->  FADD2/FMUL2/FFMA2), exchange through
//                          V[pair][branch] in smem (STS.128)   padded smem, 128 B coalesced stores.
//
// Why this shape (DESIGN.md has the numbers): per input sample the path moves 24 B of HBM but
// needs ~105 FP32 lane-ops, so on B200 the FP32 pipe (128 lanes/clk/SM) is as close a limit as
// HBM.  Windows and taps therefore never touch shared memory, every FP32 instruction is a packed
// f32x2 op (half the issue slots), and the two roles overlap through double-buffered smem.
//
// Ownership view: branch j's window holds {s[t] : t = (M/2-1-j) mod M};
// even frames dot it with sub-filter j, odd frames with sub-filter (j + M/2) mod M, and the
// result always lands in X[j] -- the circular shift costs nothing.  Branches j >= M/2 receive
// their new sample on the odd frame of a pair, so their even-frame taps are delayed by one
// slot: both halves run the same (2m+1)-tap code with per-thread tap tables.
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kM = 256;                 // channels
constexpr int kM2 = 128;
constexpr int kPairsPerBatch = 16;      // frame pairs per pipeline batch (= 32 frames)
constexpr int kFirThreads = 256;
constexpr int kFftThreads = 256;
constexpr int kThreads = kFirThreads + kFftThreads;
constexpr int kInStageBytes = kPairsPerBatch * kM * 8;            // 32 KB of input per batch
constexpr int kRegionBytes = 16 * 17 * 16;                        // 4352: padded 16x16 exchange, 16 B units
constexpr int kVBufBytes = kPairsPerBatch * kRegionBytes;         // 69632
constexpr int kSmemBytes = 2 * kInStageBytes + 2 * kVBufBytes + 128;
// mbarrier slots (8 B each) after the data buffers
constexpr int kMbInFull = 0;     // [2]    TMA transaction barriers, one per input stage
constexpr int kMbInFree = 2;     // [2]    8 FIR warps have drained the stage          
constexpr int kMbVFull = 4;      // [2][4] 8 FIR warps have written regions 4g..4g+3   
constexpr int kMbVFree = 12;     // [2]    8 FFT warps have drained the buffer         

struct FastParams {
    const float2* hist;       // Hlen samples preceding x[0] of the call
    long long Hlen;
    const float2* x;          // first sample of the call
    float2* y;                // first output frame of the call
    long long f0;             // first frame handled here (even global parity)
    long long n_pairs;        // frame pairs handled here
    const float2* taps;       // [256][kTaps] (even, odd) tap pairs, 1/M folded in
    const float2* twid;       // [16][16] e^{+j 2 pi n2 k1 / 256}
    float2* hist_new;         // if non-null: receives the last Hlen samples of (hist ++ x[0 .. n_new)), the
    long long n_new;          //   object's state after the call (folds the k_update_hist launch into this one)
};

template <int kTaps>                     // 2m + 1
__device__ __forceinline__ void fir_role(const FastParams& p, uint32_t smem, uint32_t mbar,
                                         long long batch_begin, long long batch_end)
{
    constexpr int kHist = kTaps - 1;     // 2m
    const int j = threadIdx.x;           // branch / window index
    const int pos = (j < kM2) ? (kM2 - 1 - j) : (kM + kM2 - 1 - j);     // sample slot inside a 256-sample block
    const float2* xf = p.x + p.f0 * kM2;                               // sample 0 of pair 0

    float2 T[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; i++) T[i] = __ldg(&p.taps[j * kTaps + i]);
    pdl_wait();                          // the input stream and the history may come from the previous kernel

    // register ring of 32 samples: slot (q mod 32) holds u_j[q]
    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);

    // history of the first batch of this slab: u_j[q0 - kHist .. q0 - 1], straight from global
    const long long q0 = batch_begin * kPairsPerBatch;
    const long long call_off = p.f0 * kM2;                             // xf[0] == x[call_off]
    // ring slots are static only relative to the batch parity, so the first batch must start on
    // an even batch index *within this slab*: local batch index lb = batch - batch_begin.
#pragma unroll
    for (int i = 1; i <= kHist; i++) {
        const long long t = (q0 - i) * kM + pos;                       // relative to xf
        const long long ta = t + call_off;                             // relative to x[0] of the call
        float2 v;
        if (ta >= 0) v = __ldg(&p.x[ta]);
        else if (p.Hlen + ta >= 0) v = __ldg(&p.hist[p.Hlen + ta]);
        else v = make_float2(0.f, 0.f);
        W[(32 - i) & 31] = v;                                          // slot of u[q0 - i] with q0 -> slot 0
    }

    const uint32_t in_stage0 = smem;                                     // + st * kInStageBytes
    const uint32_t vbuf0 = smem + 2 * kInStageBytes;                     // + b * kVBufBytes

    auto issue_load = [&](long long batch) {
        const int st = (int)((batch - batch_begin) & 1);
        long long np = p.n_pairs - batch * kPairsPerBatch;
        if (np > kPairsPerBatch) np = kPairsPerBatch;
        const uint32_t bytes = (uint32_t)(np * kM * 8);
        mbar_expect_tx(mbar + 8 * st, bytes);
        tma_load_1d(in_stage0 + st * kInStageBytes, xf + batch * (long long)(kPairsPerBatch * kM), bytes, mbar + 8 * st);
    };

    if (j == 0) {
        issue_load(batch_begin);
        if (batch_begin + 1 < batch_end) issue_load(batch_begin + 1);
    }

    // one batch with compile-time ring parity PAR
    auto do_batch = [&](auto par_tag, long long batch) {
        constexpr int PAR = decltype(par_tag)::value;
        const long long lb = batch - batch_begin;
        const int st = PAR;
        mbar_wait(mbar + 8 * st, (uint32_t)((lb >> 1) & 1));
        const uint32_t in = in_stage0 + st * kInStageBytes + pos * 8;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) W[16 * PAR + r] = lds64(in + r * (kM * 8));
        __syncwarp();
        if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbInFree + st));      // this warp has drained the stage
        if (lb >= 2) mbar_wait(mbar + 8 * (kMbVFree + PAR), (uint32_t)(((lb >> 1) - 1) & 1));   // FFT role released V[PAR]
        const uint32_t vout = vbuf0 + PAR * kVBufBytes + j * 16;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) {
            float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {                      // oldest sample first
                const float2 w = W[(16 * PAR + r - i) & 31];
                are = fma2(T[i], f2(w.x), are);
                aim = fma2(T[i], f2(w.y), aim);
            }
            sts128(vout + r * kRegionBytes, make_float4(are.x, are.y, aim.x, aim.y));
            if ((r & 3) == 3) {                                         // regions 4g..4g+3 of this warp are written
                __syncwarp();
                if ((j & 31) == 0) mbar_arrive(mbar + 8 * (kMbVFull + 4 * PAR + (r >> 2)));
            }
        }
        // Prefetch batch+2 into the stage just drained.  Issuing a bulk copy stalls its thread for ~0.3 us
        // (profiles/r01_tma_bench.log), so the eight FIR warps take turns instead of always taxing warp 0.
        if ((j & 31) == 0 && (j >> 5) == (int)(lb & 7) && batch + 2 < batch_end) {
            mbar_wait(mbar + 8 * (kMbInFree + st), (uint32_t)((lb >> 1) & 1));   // all 8 FIR warps drained this stage
            issue_load(batch + 2);
        }
    };

    for (long long batch = batch_begin; batch < batch_end; batch += 2) {
        do_batch(std::integral_constant<int, 0>{}, batch);
        if (batch + 1 < batch_end) do_batch(std::integral_constant<int, 1>{}, batch + 1);
    }
}

__device__ __forceinline__ void fft_role(const FastParams& p, uint32_t smem, uint32_t mbar, long long batch_begin, long long batch_end)
{
    const int tid = threadIdx.x - kFirThreads;
    const int g = tid >> 4;              // frame pair within the batch
    const int t = tid & 15;              // n2 in pass 1, k1 in pass 2
    const uint32_t vbuf0 = smem + 2 * kInStageBytes;

    float twr[16], twi[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const float2 w = __ldg(&p.twid[t * 16 + k]);
        twr[k] = w.x;
        twi[k] = w.y;
    }
    pdl_wait();                          // nothing is written before the previous kernel in the stream has completed
    // State hand-off folded into this launch: the FFT-role threads of the last CTA have nothing to do until the
    // first V regions are written, so they copy the tail of the input stream into the other history buffer.
    if (p.hist_new != nullptr && blockIdx.x == gridDim.x - 1) {
        for (long long i = tid; i < p.Hlen; i += kFftThreads) {
            const long long ts = p.n_new - p.Hlen + i;
            p.hist_new[i] = (ts >= 0) ? __ldg(&p.x[ts]) : __ldg(&p.hist[p.Hlen + ts]);
        }
    }

    for (long long batch = batch_begin; batch < batch_end; batch++) {
        const int b = (int)((batch - batch_begin) & 1);
        const uint32_t region = vbuf0 + b * kVBufBytes + g * kRegionBytes;
        const long long pair = batch * kPairsPerBatch + g;
        // staggered start: wait only for the four regions this warp pair consumes
        mbar_wait(mbar + 8 * (kMbVFull + 4 * b + (tid >> 6)), (uint32_t)(((batch - batch_begin) >> 1) & 1));

        C2 v[16];
        // pass 1: thread n2 = t gathers X[16 n1 + n2], n1 = 0..15
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const float4 q = lds128(region + (16 * n1 + t) * 16);
            v[n1].re = make_float2(q.x, q.y);
            v[n1].im = make_float2(q.z, q.w);
        }
        dft16(v);
        __syncwarp();                    // all 16 lanes of the group have read the region
        // twiddle by W256^{n2 k1} and write row n2 of the padded exchange tile
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            C2 z = v[dr4(k1)];
            if (k1 > 0) z = cmulw(z, twr[k1], twi[k1]);
            sts128(region + (t * 17 + k1) * 16, make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
        }
        __syncwarp();
        // pass 2: thread k1 = t gathers Z[n2][k1], n2 = 0..15
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) {
            const float4 q = lds128(region + (n2 * 17 + t) * 16);
            v[n2].re = make_float2(q.x, q.y);
            v[n2].im = make_float2(q.z, q.w);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(mbar + 8 * (kMbVFree + b));   // V[b] may be overwritten by the FIR role
        dft16(v);
        if (pair < p.n_pairs) {
            float2* ye = p.y + (p.f0 + 2 * pair) * (long long)kM + t;     // even frame of the pair
            float2* yo = ye + kM;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                const C2 z = v[dr4(k2)];
                __stcs(ye + 16 * k2, make_float2(z.re.x, z.im.x));
                __stcs(yo + 16 * k2, make_float2(z.re.y, z.im.y));
            }
        }
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_firpfbch2_analysis_fused(const FastParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const uint32_t mbar = smem + 2 * kInStageBytes + 2 * kVBufBytes;      // two 8-byte mbarriers

    const long long n_batches = (p.n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    const long long batch_begin = (n_batches * blockIdx.x) / gridDim.x;
    const long long batch_end = (n_batches * (blockIdx.x + 1)) / gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(mbar + 8 * (kMbInFull + 0), 1);
        mbar_init(mbar + 8 * (kMbInFull + 1), 1);
        for (int i = 0; i < 2; i++) mbar_init(mbar + 8 * (kMbInFree + i), 8);
        for (int i = 0; i < 8; i++) mbar_init(mbar + 8 * (kMbVFull + i), 8);
        for (int i = 0; i < 2; i++) mbar_init(mbar + 8 * (kMbVFree + i), 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();             // the next kernel in the stream may start its prologue as SMs free up
    if (batch_begin >= batch_end) return;       // never taken: the grid has at most one CTA per batch

    if (threadIdx.x < kFirThreads) fir_role<kTaps>(p, smem, mbar, batch_begin, batch_end);
    else fft_role(p, smem, mbar, batch_begin, batch_end);
}

template <int kTaps>
int32_t launch_t(const Firpfbch2FastPlan& plan, const FastParams& p, cudaStream_t st)
{
    // idempotent and cheap (a few microseconds); doing it per launch keeps the code free of shared mutable state
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch2_analysis_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long n_batches = (p.n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    const int grid = (int)std::min<long long>(plan.n_sm, n_batches);
    // programmatic dependent launch: back-to-back calls overlap this kernel's prologue with the previous one's tail
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = plan.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YG_CUDA(cudaLaunchKernelEx(&cfg, k_firpfbch2_analysis_fused<kTaps>, p));
    count_launch();
    return YG_OK;
}

}  // namespace

int32_t firpfbch2_fast_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h)
{
    p.supported = false;
    p.M = M;
    p.m = m;
    if (M != (uint32_t)kM) return YG_OK;
    if (m < 1 || m > 8) return YG_OK;         // instantiated tap counts 2m+1 = 3..17 (ring of 32 holds 16 + 2m)
    const int kTaps = 2 * (int)m + 1;
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;       // sm_100a only
    p.n_sm = prop.multiProcessorCount;

    // tap pairs (Te[i], To[i]) per branch, 1/M folded in (exact: M is a power of two)
    //   j <  M/2 : Te[i] = h[j + iM]          To[i] = h[j + M/2 + iM]      (i < 2m), tap 2m = 0
    //   j >= M/2 : Te[i] = h[j + (i-1)M] (i>=1, Te[0] = 0)   To[i] = h[j - M/2 + iM] (i < 2m), To[2m] = 0
    std::vector<float2> taps((size_t)kM * kTaps);
    const float s = 1.0f / (float)kM;
    const int P = 2 * (int)m;
    for (int j = 0; j < kM; j++) {
        for (int i = 0; i < kTaps; i++) {
            float te = 0.f, to = 0.f;
            if (j < kM2) {
                if (i < P) { te = h[j + i * kM]; to = h[j + kM2 + i * kM]; }
            } else {
                if (i >= 1) te = h[j + (i - 1) * kM];
                if (i < P) to = h[j - kM2 + i * kM];
            }
            taps[(size_t)j * kTaps + i] = make_float2(te * s, to * s);
        }
    }
    std::vector<float2> tw(256);
    for (int n2 = 0; n2 < 16; n2++)
        for (int k1 = 0; k1 < 16; k1++) {
            const double a = 2.0 * M_PI * (double)(n2 * k1) / 256.0;
            tw[n2 * 16 + k1] = make_float2((float)cos(a), (float)sin(a));
        }
    YG_CUDA(cudaMalloc(&p.d_taps, taps.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(p.d_taps, taps.data(), taps.size() * sizeof(float2), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&p.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(p.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    p.min_frames = 64;
    p.supported = true;
    const char* e = getenv("YG_PDL");     // debugging knob: YG_PDL=0 launches with full stream serialization
    p.pdl = !(e && e[0] == '0');
    return YG_OK;
}

void firpfbch2_fast_release(Firpfbch2FastPlan& p)
{
    if (p.d_taps) cudaFree(p.d_taps);
    if (p.d_twid) cudaFree(p.d_twid);
    if (p.d_scratch) cudaFree(p.d_scratch);
    if (p.d_flags) cudaFree(p.d_flags);
    p.d_taps = p.d_twid = p.d_scratch = p.d_flags = nullptr;
    p.n_groups = 0;
    p.supported = false;
}

int32_t firpfbch2_fast_launch(const Firpfbch2FastPlan& plan, const float2* hist, long long Hlen, const float2* x,
                              float2* y, size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new, long long n_new)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "fused kernel not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames & 1) return fail(YG_EINTERNAL, "fused kernel needs an even number of frames");
    if (((uintptr_t)(x + f0 * kM2) & 15) != 0) return fail(YG_EVALUE, "input pointer must be 16-byte aligned");
    FastParams p;
    p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
    p.f0 = (long long)f0;
    p.n_pairs = (long long)(n_frames / 2);
    p.taps = reinterpret_cast<const float2*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    p.hist_new = hist_new;
    p.n_new = n_new;
    switch (plan.m) {
        case 1: return launch_t<3>(plan, p, st);
        case 2: return launch_t<5>(plan, p, st);
        case 3: return launch_t<7>(plan, p, st);
        case 4: return launch_t<9>(plan, p, st);
        case 5: return launch_t<11>(plan, p, st);
        case 6: return launch_t<13>(plan, p, st);
        case 7: return launch_t<15>(plan, p, st);
        case 8: return launch_t<17>(plan, p, st);
        default: return fail(YG_EINTERNAL, "fused kernel not instantiated for m = %u", plan.m);
    }
}

}  // namespace yg
