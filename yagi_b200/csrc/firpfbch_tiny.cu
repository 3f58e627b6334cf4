// firpfbch_tiny.cu -- fused firpfbch_crcf (critically sampled) analysis and synthesis kernels for tiny M
// (M = 8, 16, 32), p = 2m <= 16 taps per branch, many independent streams, sm_100a.
//
//   analysis:   X_q[pos] = sum_n h[(M-1-pos) + nM] s[(q-n)M + pos],   y_q = DFT_forward(X_q)
//   synthesis:  U_q = IDFT_unnorm(X_q),   y[qM + i] = sum_n h[i + nM] U_{q-n}[i]          
//
// The firpfbch2_tiny.cu machine with streams in place of time slabs: a CTA is eight independent units of two
// warps, each unit walks a contiguous range of (group of 32/M streams, batch of 16 frames) items and its two
// warps meet in the unit's own double-buffered 4 KB tile (own mbarriers).  One warp is the FIR role (one branch /
// column per lane, 32-entry register ring, one packed FFMA2 per tap), the other transforms frame PAIRS in packed
// (even, odd) lanes -- one thread per pair entirely in registers for M <= 16, two threads and a shuffle exchange
// for M = 32 (the forward transform is the backward one with re/im swapped on the way in and out).  Tiles hold
// (frame, stream, column) entries of 8 bytes with the 16-byte chunk index XOR-swizzled by the frame pair, so
// row-wise and column-wise accesses are both conflict-free and all global traffic is 16-byte and coalesced.
// Synthesis never materialises U, so every run of batches of a stream group starts with a warm-up item: the 16
// frames before it (from the per-stream input history for the first batch), outputs suppressed.
#include "firpfbch_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kBatch = 16;                                   // frames per item and stream
constexpr int kRoleThreads = 256;
constexpr int kUnits = 8;
constexpr int kTileBytes = kBatch * 32 * 8;                  // 4 KB: 16 frames x (32/M streams x M columns) x 8 B
constexpr int kStageOff = 0;                                 // [2][unit]  input staging
constexpr int kTileOff = 2 * kUnits * kTileBytes;            // [2][unit]  X / U tile
constexpr int kMbar = 4 * kUnits * kTileBytes;               // per unit: full[2], free[2]
constexpr int kSmemBytes = kMbar + kUnits * 32;

struct TinyPfbParams {
    const float2* hist;       // analysis: [n_streams][Hlen] samples; synthesis: [n_streams][hist_frames * M] input frames
    long long Hlen;           // analysis: (p-1) * M;  synthesis: hist_frames * M (hist_frames >= 16)
    const float2* x;          // [n_streams][n_frames * M]
    float2* y;                // [n_streams][n_frames * M]
    long long n_frames;
    int n_groups;             // groups of 32 / M streams
    int batches_per_group;
    const float* taps;        // [M][p]
    const float2* twid;       // [M] e^{+j 2 pi k / M}
};

// byte offset of entry (frame row r, stream sw of the unit, column b) inside a tile
template <int kM>
__device__ __forceinline__ uint32_t tile_off(int r, int sw, int b)
{
    return (uint32_t)(((r * 32 + sw * kM) * 8) + ((((b >> 1) ^ ((r >> 1) & (kM / 2 - 1)))) * 16) + (b & 1) * 8);
}

// The items of a unit in order.  Synthesis inserts a warm-up item in front of every run of batches of one group.
struct Walk {
    int L, L1, nbg, group, k, it;
    bool warm;
    __device__ Walk(int L0, int L1_, int nbg_, bool warmups)
        : L(L0), L1(L1_), nbg(nbg_), group(L0 / nbg_), k(L0 - (L0 / nbg_) * nbg_), it(0), warm(warmups) {}
    __device__ bool done() const { return L >= L1; }
    __device__ void next(bool warmups)
    {
        it++;
        if (warm) { warm = false; return; }
        L++;
        if (++k == nbg) { k = 0; group++; warm = warmups; }
    }
    // first frame of the item (negative: input history)
    __device__ long long frame0() const { return (long long)k * kBatch - (warm ? kBatch : 0); }
};

// cooperative copy of one item's frames (16 per stream, one contiguous run each) into a swizzled tile
template <int kM>
__device__ __forceinline__ void fetch_item(const TinyPfbParams& p, const Walk& w, uint32_t tile, int lane, long long hist_frames)
{
    constexpr int kSPW = 32 / kM, kCPF = kM / 2, kCPS = kBatch * kCPF;          // 16-byte chunks per frame / per stream-item
    const long long stream_len = p.n_frames * kM;
    const long long f0 = w.frame0();
#pragma unroll
    for (int sw = 0; sw < kSPW; sw++) {
        const long long s = (long long)w.group * kSPW + sw;
#pragma unroll
        for (int c0 = 0; c0 < kCPS; c0 += 32) {
            const int c = c0 + lane;
            const int r = c / kCPF, ci = c % kCPF;
            const long long f = f0 + r;
            const float2* src = p.x;
            uint32_t bytes = 0;                                                     // past the end of the stream: zero-fill
            if (f >= 0) { if (f < p.n_frames) { src = p.x + s * stream_len + f * kM + 2 * ci; bytes = 16; } }
            else { src = p.hist + s * p.Hlen + (hist_frames + f) * kM + 2 * ci; bytes = 16; }
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile + tile_off<kM>(r, sw, 2 * ci)), "l"(src), "r"(bytes) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// the transforming warp's view of a tile: thread -> (stream sw, frame pair pr, half tt), `act` = it owns a pair
template <int kM>
struct PairMap {
    static constexpr int kSPW = 32 / kM, kTPP = (kM == 32) ? 2 : 1, kNV = kM / kTPP;
    int sw, pr, tt;
    bool act;
    __device__ explicit PairMap(int lane)
    {
        const int idx = (kTPP == 2) ? (lane >> 1) : lane;                         // pair index inside the item
        tt = (kTPP == 2) ? (lane & 1) : 0;
        act = idx < kSPW * 8;
        const int j = idx % (kSPW * 8);
        sw = j / 8;
        pr = j % 8;
    }
};

// load the pair's two frames, transform (backward DFT; kSwap: forward via re/im swap), leave bin k (natural order,
// kNV bins starting at bin0) in v[k]
template <int kM, bool kSwap>
__device__ __forceinline__ void transform_pair(C2 (&v)[PairMap<kM>::kNV], const PairMap<kM>& m, uint32_t tile, const float* twr, const float* twi)
{
    constexpr int kTPP = PairMap<kM>::kTPP;
    if constexpr (kTPP == 1) {
#pragma unroll
        for (int c = 0; c < kM / 2; c++) {
            const float4 e = lds128(tile + tile_off<kM>(2 * m.pr, m.sw, 2 * c));
            const float4 o = lds128(tile + tile_off<kM>(2 * m.pr + 1, m.sw, 2 * c));
            if (kSwap) {
                v[2 * c].re = make_float2(e.y, o.y);     v[2 * c].im = make_float2(e.x, o.x);
                v[2 * c + 1].re = make_float2(e.w, o.w); v[2 * c + 1].im = make_float2(e.z, o.z);
            } else {
                v[2 * c].re = make_float2(e.x, o.x);     v[2 * c].im = make_float2(e.y, o.y);
                v[2 * c + 1].re = make_float2(e.z, o.z); v[2 * c + 1].im = make_float2(e.w, o.w);
            }
        }
        dft_r<kM>(v);
    } else {
#pragma unroll
        for (int n = 0; n < 16; n++) {                                            // sample 2 n + tt
            const float2 e = lds64(tile + tile_off<kM>(2 * m.pr, m.sw, 2 * n + m.tt));
            const float2 o = lds64(tile + tile_off<kM>(2 * m.pr + 1, m.sw, 2 * n + m.tt));
            v[n].re = kSwap ? make_float2(e.y, o.y) : make_float2(e.x, o.x);
            v[n].im = kSwap ? make_float2(e.x, o.x) : make_float2(e.y, o.y);
        }
        dft_r<16>(v);                                                              // then X[k1 + 16 k2] = A_0[k1] + (-1)^k2 A_1[k1]
#pragma unroll
        for (int k = 0; k < 16; k++) {
            C2 a = v[k];
            if (k > 0) a = cmulw(a, twr[k], twi[k]);
            C2 o;
            o.re.x = __shfl_xor_sync(0xffffffffu, a.re.x, 1); o.re.y = __shfl_xor_sync(0xffffffffu, a.re.y, 1);
            o.im.x = __shfl_xor_sync(0xffffffffu, a.im.x, 1); o.im.y = __shfl_xor_sync(0xffffffffu, a.im.y, 1);
            v[k] = m.tt ? csub(o, a) : cadd(a, o);                                 // thread tt keeps bins 16 tt .. 16 tt + 15
        }
    }
}

// write the pair's bins back into its two rows of a tile (kSwap: undo the forward-transform swap)
template <int kM, bool kSwap>
__device__ __forceinline__ void store_pair(const C2 (&v)[PairMap<kM>::kNV], const PairMap<kM>& m, uint32_t tile)
{
    constexpr int kNV = PairMap<kM>::kNV;
    const int bin0 = (PairMap<kM>::kTPP == 2) ? 16 * m.tt : 0;
#pragma unroll
    for (int k = 0; k < kNV; k += 2) {
        const C2 a = v[k], b = v[k + 1];
        const float4 e = kSwap ? make_float4(a.im.x, a.re.x, b.im.x, b.re.x) : make_float4(a.re.x, a.im.x, b.re.x, b.im.x);
        const float4 o = kSwap ? make_float4(a.im.y, a.re.y, b.im.y, b.re.y) : make_float4(a.re.y, a.im.y, b.re.y, b.im.y);
        sts128(tile + tile_off<kM>(2 * m.pr, m.sw, bin0 + k), e);
        sts128(tile + tile_off<kM>(2 * m.pr + 1, m.sw, bin0 + k), o);
    }
}

// ================================================================================ analysis
template <int kM, int kTaps>
__device__ __forceinline__ void ana_fir_role(const TinyPfbParams& p, uint32_t smem, int unit, int L0, int L1)
{
    constexpr int kSPW = 32 / kM;
    const int lane = threadIdx.x & 31;
    const int sw = lane / kM, pos = lane % kM;
    const uint32_t mb = smem + kMbar + unit * 32;
    const long long stream_len = p.n_frames * kM;

    float T[kTaps];
#pragma unroll
    for (int n = 0; n < kTaps; n++) T[n] = __ldg(&p.taps[pos * kTaps + n]);
    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);

    Walk w(L0, L1, p.batches_per_group, false), ahead = w;
    fetch_item<kM>(p, ahead, smem + kStageOff + unit * kTileBytes, lane, 0);
    ahead.next(false);

    auto do_item = [&](auto par_tag) {
        constexpr int PAR = decltype(par_tag)::value;
        const long long s = (long long)w.group * kSPW + sw;
        if (w.k == 0 || w.it == 0) {                                    // (re)prime the window with u[q0 - i], i = 1 .. p-1
            const long long q0 = (long long)w.k * kBatch;
#pragma unroll
            for (int i = 1; i < kTaps; i++) {
                const long long t = (q0 - i) * kM + pos;
                W[(16 * PAR - i) & 31] = (t >= 0) ? __ldg(&p.x[s * stream_len + t]) : __ldg(&p.hist[s * p.Hlen + p.Hlen + t]);
            }
        }
        if (!ahead.done()) {
            fetch_item<kM>(p, ahead, smem + kStageOff + ((PAR ^ 1) * kUnits + unit) * kTileBytes, lane, 0);
            ahead.next(false);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                                   // every lane's chunks of the stage have landed
        const uint32_t stage = smem + kStageOff + (PAR * kUnits + unit) * kTileBytes;
#pragma unroll
        for (int r = 0; r < kBatch; r++) W[16 * PAR + r] = lds64(stage + tile_off<kM>(r, sw, pos));
        if (w.it >= 2) mbar_wait(mb + 8 * (2 + PAR), (uint32_t)(((w.it >> 1) - 1) & 1));     // the DFT warp has drained X[PAR]
        const uint32_t xt = smem + kTileOff + (PAR * kUnits + unit) * kTileBytes;
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = kTaps - 1; n >= 0; n--)                        // oldest sample first
                acc = fma2(W[(16 * PAR + r - n) & 31], f2(T[n]), acc);
            sts64(xt + tile_off<kM>(r, sw, pos), acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * PAR);
        w.next(false);
    };
    while (!w.done()) {
        do_item(std::integral_constant<int, 0>{});
        if (!w.done()) do_item(std::integral_constant<int, 1>{});
    }
}

template <int kM>
__device__ __forceinline__ void ana_dft_role(const TinyPfbParams& p, uint32_t smem, int unit, int L0, int L1)
{
    constexpr int kSPW = 32 / kM;
    const int lane = threadIdx.x & 31;
    const PairMap<kM> m(lane);
    const uint32_t mb = smem + kMbar + unit * 32;
    const long long stream_len = p.n_frames * kM;
    float twr[16], twi[16];
    if (kM == 32) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 t = __ldg(&p.twid[m.tt * k]);
            twr[k] = t.x;
            twi[k] = t.y;
        }
    }
    for (Walk w(L0, L1, p.batches_per_group, false); !w.done(); w.next(false)) {
        const int b = w.it & 1;
        const uint32_t xt = smem + kTileOff + (b * kUnits + unit) * kTileBytes;
        mbar_wait(mb + 8 * b, (uint32_t)((w.it >> 1) & 1));               // the FIR warp has written X[b]
        C2 v[PairMap<kM>::kNV];
        transform_pair<kM, true>(v, m, xt, twr, twi);
        __syncwarp();                                                   // every thread has read its rows
        if (m.act) store_pair<kM, true>(v, m, xt);                       // the bins go back into the pair's own rows
        __syncwarp();
        // cooperative read-out: the 16 frames of a stream are one contiguous run of output
#pragma unroll
        for (int c0 = 0; c0 < 32 * kBatch / 2; c0 += 32) {
            const int c = c0 + lane;                                     // 16-byte chunk of the item, stream-major
            const int sw = c / (kBatch * kM / 2), r = (c / (kM / 2)) % kBatch, ci = c % (kM / 2);
            const long long f = (long long)w.k * kBatch + r;
            const float4 z = lds128(xt + tile_off<kM>(r, sw, 2 * ci));
            if (f < p.n_frames)
                __stcs(reinterpret_cast<float4*>(p.y + ((long long)w.group * kSPW + sw) * stream_len + f * kM + 2 * ci), z);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * (2 + b));                   // X[b] may be overwritten
    }
}

// ================================================================================ synthesis
template <int kM>
__device__ __forceinline__ void syn_dft_role(const TinyPfbParams& p, uint32_t smem, int unit, int L0, int L1)
{
    const int lane = threadIdx.x & 31;
    const PairMap<kM> m(lane);
    const uint32_t mb = smem + kMbar + unit * 32;
    const long long hist_frames = p.Hlen / kM;
    const uint32_t stage = smem + kStageOff + unit * kTileBytes;          // single buffer: refilled as soon as it is read
    float twr[16], twi[16];
    if (kM == 32) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const float2 t = __ldg(&p.twid[m.tt * k]);
            twr[k] = t.x;
            twi[k] = t.y;
        }
    }
    Walk w(L0, L1, p.batches_per_group, true), ahead = w;
    fetch_item<kM>(p, ahead, stage, lane, hist_frames);
    ahead.next(true);
    for (; !w.done(); w.next(true)) {
        const int b = w.it & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        C2 v[PairMap<kM>::kNV];
        transform_pair<kM, false>(v, m, stage, twr, twi);
        __syncwarp();                                                   // stage consumed: refill it an item ahead
        if (!ahead.done()) {
            fetch_item<kM>(p, ahead, stage, lane, hist_frames);
            ahead.next(true);
        }
        if (w.it >= 2) mbar_wait(mb + 8 * (2 + b), (uint32_t)(((w.it >> 1) - 1) & 1));      // the FIR warp has drained U[b]
        if (m.act) store_pair<kM, false>(v, m, smem + kTileOff + (b * kUnits + unit) * kTileBytes);
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * b);
    }
}

template <int kM, int kTaps>
__device__ __forceinline__ void syn_fir_role(const TinyPfbParams& p, uint32_t smem, int unit, int L0, int L1)
{
    constexpr int kSPW = 32 / kM;
    const int lane = threadIdx.x & 31;
    const int sw = lane / kM, i = lane % kM;
    const uint32_t mb = smem + kMbar + unit * 32;
    const long long stream_len = p.n_frames * kM;

    float T[kTaps];
#pragma unroll
    for (int n = 0; n < kTaps; n++) T[n] = __ldg(&p.taps[i * kTaps + n]);
    float2 W[32];
#pragma unroll
    for (int q = 0; q < 32; q++) W[q] = make_float2(0.f, 0.f);

    Walk w(L0, L1, p.batches_per_group, true);
    auto do_item = [&](auto par_tag) {
        constexpr int PAR = decltype(par_tag)::value;
        const uint32_t ut = smem + kTileOff + (PAR * kUnits + unit) * kTileBytes;
        mbar_wait(mb + 8 * PAR, (uint32_t)((w.it >> 1) & 1));             // the DFT warp has written U[PAR]
        float2* yb = p.y + ((long long)w.group * kSPW + sw) * stream_len + (long long)w.k * kBatch * kM + i;
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            W[16 * PAR + r] = lds64(ut + tile_off<kM>(r, sw, i));
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = kTaps - 1; n >= 0; n--)                        // oldest frame first
                acc = fma2(W[(16 * PAR + r - n) & 31], f2(T[n]), acc);
            if (!w.warm && (long long)w.k * kBatch + r < p.n_frames) __stcs(yb + (long long)r * kM, acc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(mb + 8 * (2 + PAR));
        w.next(true);
    };
    while (!w.done()) {
        do_item(std::integral_constant<int, 0>{});
        if (!w.done()) do_item(std::integral_constant<int, 1>{});
    }
}

template <int kM, int kTaps, bool kSynth>
__global__ void __launch_bounds__(2 * kRoleThreads, 1) k_firpfbch_tiny(const TinyPfbParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    if (threadIdx.x == 0) {
        for (int q = 0; q < 4 * kUnits; q++) mbar_init(smem + kMbar + 8 * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int unit = (threadIdx.x >> 5) & (kUnits - 1);
    const long long n_items = (long long)p.n_groups * p.batches_per_group;
    const long long ui = (long long)blockIdx.x * kUnits + unit, nu = (long long)gridDim.x * kUnits;
    const int L0 = (int)((n_items * ui) / nu), L1 = (int)((n_items * (ui + 1)) / nu);
    if (L0 >= L1) return;
    if (kSynth) {
        if (threadIdx.x < kRoleThreads) syn_fir_role<kM, kTaps>(p, smem, unit, L0, L1);
        else syn_dft_role<kM>(p, smem, unit, L0, L1);
    } else {
        if (threadIdx.x < kRoleThreads) ana_fir_role<kM, kTaps>(p, smem, unit, L0, L1);
        else ana_dft_role<kM>(p, smem, unit, L0, L1);
    }
}

template <int kM, int kTaps, bool kSynth>
int32_t launch_t(int n_sm, const TinyPfbParams& p, cudaStream_t st)
{
    YG_CUDA(cudaFuncSetAttribute(k_firpfbch_tiny<kM, kTaps, kSynth>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long n_items = (long long)p.n_groups * p.batches_per_group;
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_sm, (n_items + kUnits - 1) / kUnits));
    k_firpfbch_tiny<kM, kTaps, kSynth><<<grid, 2 * kRoleThreads, kSmemBytes, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

template <int kM, bool kSynth>
int32_t launch_p(uint32_t taps, int n_sm, const TinyPfbParams& p, cudaStream_t st)
{
    switch (taps) {
        case 2: return launch_t<kM, 2, kSynth>(n_sm, p, st);
        case 4: return launch_t<kM, 4, kSynth>(n_sm, p, st);
        case 6: return launch_t<kM, 6, kSynth>(n_sm, p, st);
        case 8: return launch_t<kM, 8, kSynth>(n_sm, p, st);
        case 10: return launch_t<kM, 10, kSynth>(n_sm, p, st);
        case 12: return launch_t<kM, 12, kSynth>(n_sm, p, st);
        case 14: return launch_t<kM, 14, kSynth>(n_sm, p, st);
        case 16: return launch_t<kM, 16, kSynth>(n_sm, p, st);
        default: return fail(YG_EINTERNAL, "tiny firpfbch kernel not instantiated for p = %u", taps);
    }
}

}  // namespace

int32_t firpfbch_tiny_plan(FirpfbchFastPlan& plan, int32_t type, uint32_t M, uint32_t p, const float* h)
{
    plan.supported = false;
    plan.p = p;
    plan.M = M;
    if (M != 8 && M != 16 && M != 32) return YG_OK;
    plan.type = type;
    if (p < 2 || p > 16 || (p & 1)) return YG_OK;        // instantiated: p = 2, 4, ..., 16
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    const int iM = (int)M;
    std::vector<float> taps((size_t)iM * p);
    for (int pos = 0; pos < iM; pos++)
        for (uint32_t n = 0; n < p; n++)
            taps[(size_t)pos * p + n] = (type == YG_ANALYZER) ? h[(iM - 1 - pos) + n * iM] : h[pos + n * iM];
    std::vector<float2> tw(M);
    for (uint32_t k = 0; k < M; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.supported = true;
    return YG_OK;
}

// hist: analysis [n_streams][Hlen = (p-1) M] samples; synthesis [n_streams][Hlen = hist_frames M] input frames
// (hist_frames >= 16).  n_streams must be a multiple of 32 / M; x, y and hist 16-byte aligned.
int32_t firpfbch_tiny_launch(const FirpfbchFastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                             long long n_frames, long long n_streams, cudaStream_t st)
{
    if (!plan.supported) return fail(YG_EINTERNAL, "tiny firpfbch kernel not available for this geometry");
    const int spw = 32 / (int)plan.M;
    if (n_streams % spw) return fail(YG_EINTERNAL, "tiny firpfbch kernel takes groups of %d streams", spw);
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(hist)) & 15) != 0)
        return fail(YG_EVALUE, "pointers must be 16-byte aligned");
    const bool synth = plan.type == YG_SYNTHESIZER;
    if (synth && Hlen < (long long)kBatch * plan.M) return fail(YG_EINTERNAL, "input history too short for the warm-up batch");
    TinyPfbParams p;
    p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
    p.n_frames = n_frames;
    p.n_groups = (int)(n_streams / spw);
    p.batches_per_group = (int)((n_frames + kBatch - 1) / kBatch);
    p.taps = reinterpret_cast<const float*>(plan.d_taps);
    p.twid = reinterpret_cast<const float2*>(plan.d_twid);
    if ((long long)p.n_groups * p.batches_per_group > 0x3fffffffLL) return fail(YG_ERANGE, "too many batches for one launch");
    switch (plan.M) {
        case 8: return synth ? launch_p<8, true>(plan.p, plan.n_sm, p, st) : launch_p<8, false>(plan.p, plan.n_sm, p, st);
        case 16: return synth ? launch_p<16, true>(plan.p, plan.n_sm, p, st) : launch_p<16, false>(plan.p, plan.n_sm, p, st);
        default: return synth ? launch_p<32, true>(plan.p, plan.n_sm, p, st) : launch_p<32, false>(plan.p, plan.n_sm, p, st);
    }
}

}  // namespace yg
