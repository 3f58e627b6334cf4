// firpfbch2_large.cu -- firpfbch2 analysis and synthesis for large power-of-two M (512 .. 4096; M = 1024 is
// BASELINE config #4), sm_100a.
//
// One SM cannot hold M >= 512 register-resident branch windows plus the transform.  Three implementations live here:
//
//  * SINGLE-SM (s1k / s1ks: M = 1024 with m <= 4 = BASELINE config #4; s5k / s5ks: M = 512 with m <= 7): one CTA per SM
//    and no exchange at all -- taps in registers, windows in a shared-memory input ring (analysis) / running output sums in
//    registers over a shared-memory frame ring transformed in place (synthesis).  Described where they are defined.

//  * FUSED (k_large_fused / k_large_synth_fused, one cooperative launch per call): groups of G = M / 256
//    persistent CTAs; in every CTA warps 0-7 own 256 polyphase branches (the register-window FIR / overlap-add
//    arithmetic of the M = 256 kernels) and warps 8-15 transform frame pairs in packed (even, odd) form as teams
//    of M / 16 threads (16 x 16 x R, three passes, two shared exchanges).  The two roles of a group meet in a
//    ring of 16-pair batches in global memory that stays in L2, guarded by two counters per slot.  HBM sees the
//    algorithmic 24 B per sample; the bound is L2 throughput (56 B per sample through it).
//  * TWO-STAGE (k_large_fir + k_large_fft / k_large_fft + k_large_wola per 96 MB chunk, chained through L2):
//    takes what the fused kernels do not -- fewer than 32 leftover frames, device buffers that are not 16-byte
//    aligned, devices without cooperative launch.  Stage B is a warp-level DFT: 32 x 32 with a radix-32 in
//    registers and one XOR-swizzled shared exchange for M = 1024, one decimation-in-frequency step on top for
//    2048 / 4096 (2 / 4 warps per frame), 16 x 32 on frame pairs for 512.
#include "firpfbch2_fast.cuh"
#include "fused_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kFirThreads = 256;                 // branches per CTA (M / 256 CTAs cover a frame)
constexpr int kPairsPerBatch = 16;

struct LargeParams {
    const float2* hist;       // Hlen samples preceding x[0] of the call
    long long Hlen;
    const float2* x;
    float2* y;
    long long f0;             // first frame handled (even global parity)
    long long pair_begin, pair_end;      // frame pairs of this launch, relative to f0
    int slabs;                // CTAs along the pair axis
    int M;                    // channels (power of two, multiple of 256)
    const float2* taps;       // [M][2m+1] (even, odd) tap pairs, 1/M folded in
    const float2* twid;       // [M] e^{+j 2 pi k / M}
    float2* hist_new = nullptr;          // single-SM kernel only: if non-null it also writes the object's next state, the
    long long n_new = 0;                 //   last Hlen samples of (hist ++ x[0 .. n_new)), like the M = 256 kernel
};

// ------------------------------------------------------------------ stage A: branch FIRs
template <int kTaps>
__global__ void __launch_bounds__(kFirThreads, 2) k_large_fir(const LargeParams p)
{
    constexpr int kHist = kTaps - 1;
    const int kM = p.M, kM2 = p.M >> 1;
    const int j = blockIdx.y * kFirThreads + threadIdx.x;               // branch
    const int pos = (j < kM2) ? (kM2 - 1 - j) : (kM + kM2 - 1 - j);       // sample slot inside an M-sample block
    const long long n_pairs = p.pair_end - p.pair_begin;
    const long long n_batches = (n_pairs + kPairsPerBatch - 1) / kPairsPerBatch;
    const long long b0 = (n_batches * blockIdx.x) / p.slabs, b1 = (n_batches * (blockIdx.x + 1)) / p.slabs;
    if (b0 >= b1) return;

    float2 T[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; i++) T[i] = __ldg(&p.taps[j * kTaps + i]);

    const long long call_off = p.f0 * kM2;                               // sample index of pair 0 relative to x[0]
    auto sample = [&](long long q) {                                     // u_j[q], q relative to f0
        const long long ta = q * kM + pos + call_off;
        if (ta >= 0) return __ldg(&p.x[ta]);
        if (p.Hlen + ta >= 0) return __ldg(&p.hist[p.Hlen + ta]);
        return make_float2(0.f, 0.f);
    };

    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);
    const long long q_first = p.pair_begin + b0 * kPairsPerBatch;
#pragma unroll
    for (int i = 1; i <= kHist; i++) W[(32 - i) & 31] = sample(q_first - i);

    float2 Pf[kPairsPerBatch];                                           // next batch, in flight
#pragma unroll
    for (int r = 0; r < kPairsPerBatch; r++) Pf[r] = (q_first + r < p.pair_end) ? sample(q_first + r) : make_float2(0.f, 0.f);

    auto do_batch = [&](auto par_tag, long long batch) {
        constexpr int PAR = decltype(par_tag)::value;
        const long long q0 = p.pair_begin + batch * kPairsPerBatch;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) W[16 * PAR + r] = Pf[r];
        if (batch + 1 < b1) {
#pragma unroll
            for (int r = 0; r < kPairsPerBatch; r++) {
                const long long q = q0 + kPairsPerBatch + r;
                Pf[r] = (q < p.pair_end) ? sample(q) : make_float2(0.f, 0.f);
            }
        }
        float2* yb = p.y + (p.f0 + 2 * q0) * (long long)kM + j;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) {
            float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {
                const float2 w = W[(16 * PAR + r - i) & 31];
                are = fma2(T[i], f2(w.x), are);
                aim = fma2(T[i], f2(w.y), aim);
            }
            if (q0 + r < p.pair_end) {
                yb[(long long)(2 * r) * kM] = make_float2(are.x, aim.x);          // even frame, stays in L2 for stage B
                yb[(long long)(2 * r + 1) * kM] = make_float2(are.y, aim.y);      // odd frame
            }
        }
    };
    for (long long batch = b0; batch < b1; batch += 2) {
        do_batch(std::integral_constant<int, 0>{}, batch);
        if (batch + 1 < b1) do_batch(std::integral_constant<int, 1>{}, batch + 1);
    }
}

// ------------------------------------------------------------------ stage B: 1024-point DFT per frame
// Single-frame (re, im) arithmetic: a complex add is one packed FADD2.
__device__ __forceinline__ float2 xadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 xsub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 xaddj(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }   // a + j b
__device__ __forceinline__ float2 xsubj(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }   // a - j b
__device__ __forceinline__ float2 xmul(float2 a, float wr, float wi)
{
    const float2 t = __fmul2_rn(a, make_float2(wr, wr));                // one packed multiply + two FMAs (was 2 + 2)
    return make_float2(fmaf(-a.y, wi, t.x), fmaf(a.x, wi, t.y));
}
__device__ __forceinline__ void xdft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 s0 = xadd(a0, a2), d0 = xsub(a0, a2), s1 = xadd(a1, a3), d1 = xsub(a1, a3);
    a0 = xadd(s0, s1); a2 = xsub(s0, s1); a1 = xaddj(d0, d1); a3 = xsubj(d0, d1);
}
// 16-point backward DFT of v[o], v[o+1], ...; X[k] left at v[o + dr4(k)]
template <int O>
__device__ __forceinline__ void xdft16(float2 (&v)[32])
{
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#pragma unroll
    for (int b = 0; b < 4; b++) xdft4(v[O + b], v[O + 4 + b], v[O + 8 + b], v[O + 12 + b]);
    v[O + 5] = xmul(v[O + 5], c1, s1);   v[O + 9] = xmul(v[O + 9], r2, r2);    v[O + 13] = xmul(v[O + 13], s1, c1);
    v[O + 6] = xmul(v[O + 6], r2, r2);   v[O + 10] = make_float2(-v[O + 10].y, v[O + 10].x);   v[O + 14] = xmul(v[O + 14], -r2, r2);
    v[O + 7] = xmul(v[O + 7], s1, c1);   v[O + 11] = xmul(v[O + 11], -r2, r2); v[O + 15] = xmul(v[O + 15], -c1, -s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) xdft4(v[O + 4 * k1], v[O + 4 * k1 + 1], v[O + 4 * k1 + 2], v[O + 4 * k1 + 3]);
}
// 32-point backward DFT: X[k] left at v[(k & 1) * 16 + dr4(k >> 1)]
__device__ __forceinline__ constexpr int dr32(int k) { return ((k & 1) << 4) | dr4(k >> 1); }
__device__ __forceinline__ void xdft32(float2 (&v)[32], const float2* w32 /* e^{+j 2 pi b / 32}, b < 16, shared */)
{
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const float2 s = xadd(v[b], v[16 + b]), d = xsub(v[b], v[16 + b]);
        v[b] = s;
        if (b == 0) v[16] = d;
        else if (b == 8) v[24] = make_float2(-d.y, d.x);                // W32^8 = j
        else { const float2 w = w32[b]; v[16 + b] = xmul(d, w.x, w.y); }
    }
    xdft16<0>(v);
    xdft16<16>(v);
}

// the same with e^{+j 2 pi b / 32} as immediates instead of 16 broadcast loads per transform
__device__ __forceinline__ void xdft32c(float2 (&v)[32])
{
    constexpr float kC[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                              0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.f, -0.19509032201612825f,
                              -0.38268343236508977f, -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                              -0.92387953251128674f, -0.98078528040323043f};
    constexpr float kS[16] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f, 0.70710678118654752f,
                              0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f, 1.f, 0.98078528040323043f,
                              0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                              0.38268343236508977f, 0.19509032201612825f};
#pragma unroll
    for (int b = 0; b < 16; b++) {
        const float2 s = xadd(v[b], v[16 + b]), d = xsub(v[b], v[16 + b]);
        v[b] = s;
        if (b == 0) v[16] = d;
        else if (b == 8) v[24] = make_float2(-d.y, d.x);                // W32^8 = j
        else v[16 + b] = xmul(d, kC[b], kS[b]);
    }
    xdft16<0>(v);
    xdft16<16>(v);
}

constexpr int kFftWarps = 8;                     // warps per CTA, one 1024-point transform each
constexpr int kFftSmem = kFftWarps * 1024 * 8 + 16 * 8;

// ---- warp-level transforms.  M = 1024 S (S = 1, 2, 4): one decimation-in-frequency step splits a frame into S
// interleaved 1024-point transforms, X[S k + r] = DFT1024{ (sum_q x[1024 q + n] W_S^{q r}) W_M^{n r} }[k]; a warp takes
// one (frame, r) item: 32 x 32 with a radix-32 in registers and one XOR-swizzled 8 KB shared exchange tile.
template <int S>
__device__ __forceinline__ void fft_load(float2 (&v)[32], const float2* src, int r, const float2* __restrict__ twid, int lane)
{
    float2 ws[S];                                                        // W_S^{q r}
#pragma unroll
    for (int q = 0; q < S; q++) ws[q] = __ldg(&twid[((q * r) & (S - 1)) * 1024]);
    // lane n2 gathers x[32 n1 + n2]: 256-byte coalesced rows, L2 hits (written by the FIR stage)
#pragma unroll
    for (int n1 = 0; n1 < 32; n1++) {
        const int n = 32 * n1 + lane;
        float2 a = src[n];
        if (S > 1) {
#pragma unroll
            for (int q = 1; q < S; q++) {
                const float2 z = src[1024 * q + n];
                if (S == 2) a = r ? xsub(a, z) : xadd(a, z);
                else a = xadd(a, xmul(z, ws[q].x, ws[q].y));
            }
            const float2 w = __ldg(&twid[n * r]);
            a = xmul(a, w.x, w.y);
        }
        v[n1] = a;
    }
}

template <int S>
__device__ __forceinline__ void fft_compute(float2 (&v)[32], uint32_t tile, const float2* w32, const float2* __restrict__ twid, int lane)
{
    xdft32(v, w32);
    // twiddle by W1024^{n2 k1}, write row n2 of the swizzled tile
#pragma unroll
    for (int k1 = 0; k1 < 32; k1++) {
        float2 z = v[dr32(k1)];
        if (k1 > 0) { const float2 w = __ldg(&twid[lane * k1 * S]); z = xmul(z, w.x, w.y); }
        sts64(tile + (((lane << 5) | (k1 ^ lane)) << 3), z);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(tile + (((n2 << 5) | (lane ^ n2)) << 3));
    __syncwarp();
    xdft32(v, w32);
}

template <int S>
__device__ __forceinline__ void fft_store(const float2 (&v)[32], float2* fr, int r, int lane, int streaming_store)
{
    if (streaming_store) {
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) __stcs(fr + S * (lane + 32 * k2) + r, v[dr32(k2)]);
    } else {                                                             // keep U in L2 for the overlap-add stage
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) fr[S * (lane + 32 * k2) + r] = v[dr32(k2)];
    }
}

// M = 512 = 16 x 32: one warp per PAIR of frames.  Pass 1: lane n2 runs the 16-point transforms over n1 of both
// frames; pass 2: lane (frame, k1) runs one 32-point transform over n2.  Same 8 KB swizzled tile per warp.
__device__ __forceinline__ void fft512_load(float2 (&v)[32], const float2* src0, const float2* src1, int lane)
{
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = src0[32 * n1 + lane];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[16 + n1] = src1[32 * n1 + lane];
}

__device__ __forceinline__ void fft512_compute(float2 (&v)[32], uint32_t tile, const float2* w32, const float2* __restrict__ twid, int lane)
{
    xdft16<0>(v);
    xdft16<16>(v);
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            float2 z = v[16 * h + dr4(k1)];
            if (k1 > 0) { const float2 w = __ldg(&twid[lane * k1]); z = xmul(z, w.x, w.y); }
            sts64(tile + (((h << 9) | (lane << 4) | (k1 ^ (lane & 15))) << 3), z);
        }
    __syncwarp();
    const int h = lane >> 4, k1 = lane & 15;
#pragma unroll
    for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(tile + (((h << 9) | (n2 << 4) | (k1 ^ (n2 & 15))) << 3));
    __syncwarp();
    xdft32(v, w32);
}

// lane (h, k1) holds X_h[k1 + 16 k2]; `fr` = frame h of the pair (nullptr: frame beyond the end)
__device__ __forceinline__ void fft512_store(const float2 (&v)[32], float2* fr, int lane, int streaming_store)
{
    if (!fr) return;
    fr += lane & 15;
    if (streaming_store) {
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) __stcs(fr + 16 * k2, v[dr32(k2)]);
    } else {
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) fr[16 * k2] = v[dr32(k2)];
    }
}

// ------------------------------------------------------------------ stage B kernels (two-stage path)
// Frame f of the launch is read from the virtual stream (prefix ++ x) at index v0 + f (negative indices
// live in the 32-frame prefix) and written to dst frame f.  The analysis path calls it in place
// (x == dst, v0 == 0): the S warps of a frame (neighbours in one CTA) meet on a named barrier between
// their loads and their stores.  The synthesis path transforms input frames into the U scratch.
template <int S>
__global__ void __launch_bounds__(kFftWarps * 32, 2) k_large_fft(const float2* prefix, const float2* x, long long v0,
                                                                 float2* dst, long long n_frames,
                                                                 const float2* __restrict__ twid, int streaming_store)
{
    constexpr int kM = 1024 * S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* w32 = reinterpret_cast<float2*>(smem_raw + kFftWarps * 1024 * 8);
    if (threadIdx.x < 16) w32[threadIdx.x] = __ldg(&twid[threadIdx.x * 32 * S]);
    __syncthreads();
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int r = wrp & (S - 1);                                         // residue this warp produces
    const uint32_t tile = smem_u32(smem_raw) + wrp * 8192;
    const long long n_items = n_frames * S;
    for (long long it = (long long)blockIdx.x * kFftWarps + wrp; it < n_items; it += (long long)gridDim.x * kFftWarps) {
        const long long f = it / S;
        const long long vi = v0 + f;
        const float2* src = (vi < 0) ? prefix + (32 + vi) * kM : x + vi * kM;
        float2 v[32];
        fft_load<S>(v, src, r, twid, lane);
        fft_compute<S>(v, tile, w32, twid, lane);
        if (S > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + wrp / S), "r"(S * 32) : "memory");
        fft_store<S>(v, dst + f * kM, r, lane, streaming_store);
    }
}

__global__ void __launch_bounds__(kFftWarps * 32, 2) k_large_fft512(const float2* prefix, const float2* x, long long v0,
                                                                    float2* dst, long long n_frames,
                                                                    const float2* __restrict__ twid, int streaming_store)
{
    constexpr int kM = 512;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* w32 = reinterpret_cast<float2*>(smem_raw + kFftWarps * 1024 * 8);
    if (threadIdx.x < 16) w32[threadIdx.x] = __ldg(&twid[threadIdx.x * 16]);
    __syncthreads();
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const uint32_t tile = smem_u32(smem_raw) + wrp * 8192;
    const long long n_pairs = (n_frames + 1) >> 1;
    for (long long pr = (long long)blockIdx.x * kFftWarps + wrp; pr < n_pairs; pr += (long long)gridDim.x * kFftWarps) {
        const float2* src[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const long long vi = v0 + min(2 * pr + h, n_frames - 1);     // an odd tail recomputes the last frame
            src[h] = (vi < 0) ? prefix + (32 + vi) * kM : x + vi * kM;
        }
        float2 v[32];
        fft512_load(v, src[0], src[1], lane);
        fft512_compute(v, tile, w32, twid, lane);
        const long long f = 2 * pr + (lane >> 4);
        fft512_store(v, f < n_frames ? dst + f * kM : nullptr, lane, streaming_store);
    }
}

// ------------------------------------------------------------------ fused kernel: one launch, groups of M/256 CTAs
// A GROUP of G = M / 256 persistent CTAs walks a contiguous slab of 16-pair batches.  In every CTA warps 0-7
// are the FIR role of stage A for 256 of the M branches, warps 8-15 transform whole frames (stage B): the
// 32 frames of a batch are 8 G warp items (M = 512: 16 frame pairs; M = 1024 S: 32 S (frame, residue) items),
// one per DFT warp of the group.  The branch sums cross the group through a small per-group ring of V batches
// in global memory that never leaves L2 (kSlots x 32 frames x 8 M bytes per group, 18 MB for the whole grid),
// guarded by two monotone counters per slot: full (one release-add per FIR warp of the group per use) and
// free (one per DFT warp, after its loads of the slot have landed).  HBM sees the algorithmic 24 B per input
// sample; the launch is cooperative so that every member of a group is resident while the others wait for it.
constexpr int kSlots = 4;                          // swept 2 / 4 / 8: 2 starves the consumer, 8 spills the ring out of L2
constexpr int kInStageBytes = kPairsPerBatch * kFirThreads * 8;      // 32 KB: one batch of input for 256 branches
constexpr int kDftSmem = 256 * 17 * 16;                                // 256 DFT threads x one padded 16-entry row
constexpr int kVStageBytes = 256 * 16 * 16;                            // next pair's V, one 16-entry column per DFT thread
constexpr int kFusedSmem = kDftSmem + kVStageBytes + 2 * kInStageBytes;
constexpr int kFlagStride = 32;                  // 128 B of counters per group: full[kSlots], free[kSlots]

struct FusedParams {
    LargeParams base;                            // pair_begin .. pair_end: whole batches
    float2* scratch;                             // [n_groups][kSlots][16 pairs][M] packed (re_e, re_o, im_e, im_o)
    unsigned* flags;                             // [n_groups][kFlagStride], zero at launch
    int n_groups;
};

// The V ring is written with st.global.cg and read with ld.global.cg (L2 on both sides), so the waiters poll
// with relaxed loads: an acquire load would flush the SM's L1 (CCTL.IVALL) on every poll for nothing.
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// producer side: the warp's V stores, then one release-add by lane 0 (the grid-sync idiom: barrier, fence, atomic)
__device__ __forceinline__ void warp_publish(unsigned* p, int lane)
{
    __syncwarp();
    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// consumer side: every lane has USED the values it loaded from the slot (so the loads have landed); nothing to flush
__device__ __forceinline__ void warp_retire(unsigned* p, int lane)
{
    __syncwarp();
    if (lane == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// The polls are relaxed.  The PTX memory model would want an acquire after the counter has been seen, to pair with the
// producer's red.release before this warp reads the ring.  One `fence.acq_rel.gpu` per batch and warp was built and
// measured (-DYG_LARGE_ACQUIRE_FENCE=1, profiles/r02_large_fence_ab.log): it costs 10-11 % (M = 1024, 2^24: 0.146 ->
// 0.163 ms; M = 512: 0.477 -> 0.537 ms) because at gpu scope it also invalidates the SM's L1 (CCTL.IVALL), where the
// twiddles live.  It stays compiled out: the ring is written with st.global.cg and read with cp.async.cg / ld.global.cg
// (L2 on both sides, nothing stale to invalidate), and the reads are issued only after the poll loop's exit branch has
// resolved on the loaded value -- the SM does not issue loads past an unresolved branch -- so the counter load is
// ordered before them on this hardware.
#ifndef YG_LARGE_ACQUIRE_FENCE
#define YG_LARGE_ACQUIRE_FENCE 0
#endif
__device__ __forceinline__ void acquire_after_poll()
{
#if YG_LARGE_ACQUIRE_FENCE
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
__device__ __forceinline__ void warp_wait(const unsigned* p, unsigned target, int lane)
{
    if (lane == 0) {
        unsigned spins = 0;
        while (ld_relaxed_gpu(p) < target) {
            if (++spins > (1u << 27)) __trap();  // seconds without progress: fail the launch rather than hang the GPU
        }
        acquire_after_poll();
    }
    __syncwarp();
}
// A counter read costs an L2 round trip (a microsecond under load), so each role reads the counter it will need
// for its NEXT batch one batch early (peek) and only falls back to polling when that early value was too small.
__device__ __forceinline__ unsigned warp_peek(const unsigned* p, int lane) { return lane == 0 ? ld_relaxed_gpu(p) : 0u; }
__device__ __forceinline__ void warp_wait_seen(unsigned seen, const unsigned* p, unsigned target, int lane)
{
    if (lane == 0) {
        if (seen < target) {
            unsigned spins = 0;
            while (ld_relaxed_gpu(p) < target) {
                if (++spins > (1u << 27)) __trap();
            }
        }
        acquire_after_poll();
    }
    __syncwarp();
}

template <int kTaps>
__device__ __forceinline__ void fused_fir_role(const FusedParams& fp, int group, int c, long long b0, long long b1,
                                               uint32_t smem_in)
{
    const LargeParams& p = fp.base;
    constexpr int kHist = kTaps - 1;
    const int kM = p.M, kM2 = p.M >> 1;
    const int lane = threadIdx.x & 31;
    const int j = c * kFirThreads + threadIdx.x;                          // branch
    const int pos = (j < kM2) ? (kM2 - 1 - j) : (kM + kM2 - 1 - j);
    const unsigned n_dft_warps = (unsigned)(kFftWarps * (kM / kFirThreads));
    unsigned* flags = fp.flags + group * kFlagStride;
    float4* sbase = reinterpret_cast<float4*>(fp.scratch) + (long long)group * kSlots * kPairsPerBatch * kM + j;

    float2 T[kTaps];
#pragma unroll
    for (int i = 0; i < kTaps; i++) T[i] = __ldg(&p.taps[j * kTaps + i]);

    const long long call_off = p.f0 * kM2;
    auto sample = [&](long long q) {
        const long long ta = q * kM + pos + call_off;
        if (ta >= 0) return __ldg(&p.x[ta]);
        if (p.Hlen + ta >= 0) return __ldg(&p.hist[p.Hlen + ta]);
        return make_float2(0.f, 0.f);
    };

    float2 W[32];
#pragma unroll
    for (int i = 0; i < 32; i++) W[i] = make_float2(0.f, 0.f);
    const long long q_first = p.pair_begin + b0 * kPairsPerBatch;
#pragma unroll
    for (int i = 1; i <= kHist; i++) W[(32 - i) & 31] = sample(q_first - i);

    // input staging (double-buffered, cp.async.cg so that the stream does not wash the twiddles out of L1):
    // branches j and j + 1 read neighbouring samples (pos falls as j rises, pos(even tid) is odd), so a lane pair
    // shares 16-byte copies -- the even lane fetches rows 0-7 of the batch, the odd lane rows 8-15 -- and each
    // lane then finds its own sample in slot tid ^ 1 of the row.
    const uint32_t stage_wr = smem_in + (threadIdx.x & ~1) * 8 + (threadIdx.x & 1) * (8 * kFirThreads * 8);
    const uint32_t stage_rd = smem_in + (threadIdx.x ^ 1) * 8;
    const float2* xs = p.x + (p.pair_begin * (long long)kM + (pos & ~1) + call_off)       // 16-byte aligned chunk of the pair
                       + (threadIdx.x & 1) * (8 * (long long)kM);
    auto prefetch = [&](long long batch, int st) {
        const float2* src = xs + batch * (long long)(kPairsPerBatch * kM);
#pragma unroll
        for (int r = 0; r < 8; r++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_wr + st * kInStageBytes + r * (kFirThreads * 8)),
                         "l"(src + (long long)r * kM) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(b0, 0);
    unsigned seen_free = 0;                                               // early read of this batch's free counter

    auto do_batch = [&](auto par_tag, long long batch) {
        constexpr int PAR = decltype(par_tag)::value;
        const long long lb = batch - b0;
        const int slot = (int)(lb % kSlots);
        const unsigned use = (unsigned)(lb / kSlots);
        const unsigned seen_now = seen_free;
        seen_free = warp_peek(flags + kSlots + (int)((lb + 1) % kSlots), lane);  // in flight across this batch's arithmetic
        if (batch + 1 < b1) {
            prefetch(batch + 1, PAR ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                                     // the partner lane's half of the rows has landed too
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) W[16 * PAR + r] = lds64(stage_rd + PAR * kInStageBytes + r * (kFirThreads * 8));
        if (use) warp_wait_seen(seen_now, flags + kSlots + slot, use * n_dft_warps, lane);   // the slot's previous batch has been read
        float4* vb = sbase + (long long)slot * kPairsPerBatch * kM;
#pragma unroll
        for (int r = 0; r < kPairsPerBatch; r++) {
            float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {
                const float2 w = W[(16 * PAR + r - i) & 31];
                are = fma2(T[i], f2(w.x), are);
                aim = fma2(T[i], f2(w.y), aim);
            }
            __stcg(vb + (long long)r * kM, make_float4(are.x, are.y, aim.x, aim.y));    // packed (even, odd) pair
        }
        warp_publish(flags + slot, lane);
    };
    for (long long batch = b0; batch < b1; batch += 2) {
        do_batch(std::integral_constant<int, 0>{}, batch);
        if (batch + 1 < b1) do_batch(std::integral_constant<int, 1>{}, batch + 1);
    }
}

// ---- DFT role of the fused kernel: frame PAIRS in packed (even, odd) form, M = 16 x 16 x R.
// A team of T = M / 16 threads owns one pair; every thread carries 16 packed values through three passes:
//   pass 1  thread n2:      16-point DFT over n1 of V[T n1 + n2], times W_M^{n2 k1}
//   pass 2  thread (b, k1): 16-point DFT over a of the values n2 = R a + b, times W_T^{b e}
//   pass 3  thread (g, k1): 16 / R radix-R DFTs over b for e in its slice  ->  X[k1 + 16 (e + 16 f)]
// with two exchanges through one shared region per team (17-entry padded rows for the first, k1-fastest for the
// second); 128-byte contiguous stores per half-warp and frame.
template <int R>
__device__ __forceinline__ void team_bar(int id)
{
    if (R == 2) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(16 * R) : "memory");
}

// passes 1 (twiddle + exchange) .. 3 of the team transform; on entry v[dr4(k1)] holds the 16-point DFT over n1,
// on exit v[ei R + f] = X[k1 + 16 (hi E + ei + 16 f)] with hi = tt >> 4, k1 = tt & 15, E = 16 / R
template <int R>
__device__ __forceinline__ void team_dft_tail(C2 (&v)[16], uint32_t region, int team, int tt, const float2* __restrict__ twid)
{
    constexpr int E = 16 / R;
    const int hi = tt >> 4, k1 = tt & 15;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        C2 z = v[dr4(k)];
        if (k > 0) { const float2 w = __ldg(&twid[tt * k]); z = cmulw(z, w.x, w.y); }
        stc2(region + (tt * 17 + k) * 16, z);
    }
    team_bar<R>(1 + team);
#pragma unroll
    for (int a = 0; a < 16; a++) v[a] = ldc2(region + ((R * a + hi) * 17 + k1) * 16);
    team_bar<R>(1 + team);
    dft16(v);
#pragma unroll
    for (int e = 0; e < 16; e++) {
        C2 z = v[dr4(e)];
        if (e > 0) { const float2 w = __ldg(&twid[16 * hi * e]); z = cmulw(z, w.x, w.y); }
        stc2(region + (((hi * 16 + e) * 16) + k1) * 16, z);
    }
    team_bar<R>(1 + team);
#pragma unroll
    for (int ei = 0; ei < E; ei++)
#pragma unroll
        for (int bb = 0; bb < R; bb++) v[ei * R + bb] = ldc2(region + (((bb * 16 + hi * E + ei) * 16) + k1) * 16);
    team_bar<R>(1 + team);                                               // region free for the next pair
#pragma unroll
    for (int ei = 0; ei < E; ei++) dft_r<R>(&v[ei * R]);
}

template <int R>
__device__ __forceinline__ void fused_dft_role_r(const FusedParams& fp, int group, int c, long long b0, long long b1,
                                                 uint32_t smem_dft, uint32_t smem_stage)
{
    constexpr int T = 16 * R, kM = 256 * R, E = 16 / R, G = R;
    const LargeParams& p = fp.base;
    const int lane = threadIdx.x & 31;
    const int dt = threadIdx.x - kFirThreads;                            // 0 .. 255
    const int team = dt / T, tt = dt % T;
    const int hi = tt >> 4, k1 = tt & 15;                                // pass 2: b = hi; pass 3: g = hi
    const int pr = c * (256 / T) + team;                                 // pair of the batch owned by this team
    const uint32_t region = smem_dft + team * (T * 17 * 16);
    const unsigned n_fir_warps = (unsigned)((kFirThreads / 32) * G);
    unsigned* flags = fp.flags + group * kFlagStride;
    const float4* sbase = reinterpret_cast<const float4*>(fp.scratch) + ((long long)group * kSlots * kPairsPerBatch + pr) * kM + tt;
    // V of the next batch streams into a private shared column of this thread (cp.async) while the current pair
    // is transformed
    const uint32_t vstage = smem_stage + dt * 16;
    auto fetch_v = [&](long long lb) {
        const float4* vs = sbase + (long long)(lb % kSlots) * kPairsPerBatch * kM;
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vstage + n1 * (256 * 16)), "l"(vs + T * n1) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    warp_wait(flags + 0, n_fir_warps, lane);
    fetch_v(0);
    unsigned seen_full = (b0 + 1 < b1) ? warp_peek(flags + 1 % kSlots, lane) : 0u;
    for (long long batch = b0; batch < b1; batch++) {
        const long long lb = batch - b0;
        const int slot = (int)(lb % kSlots);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        C2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) v[n1] = ldc2(vstage + n1 * (256 * 16));
        warp_retire(flags + kSlots + slot, lane);                        // every lane's copy of the slot has landed
        dft16(v);
        if (batch + 1 < b1) {                                            // own column has been consumed: refill it
            const unsigned seen_now = seen_full;
            if (batch + 2 < b1) seen_full = warp_peek(flags + (int)((lb + 2) % kSlots), lane);
            warp_wait_seen(seen_now, flags + (int)((lb + 1) % kSlots), (unsigned)((lb + 1) / kSlots + 1) * n_fir_warps, lane);
            fetch_v(lb + 1);
        }
        team_dft_tail<R>(v, region, team, tt, p.twid);
        float2* ye = p.y + (p.f0 + 2 * (p.pair_begin + batch * kPairsPerBatch + pr)) * (long long)kM + k1 + 16 * (hi * E);
        float2* yo = ye + kM;
#pragma unroll
        for (int ei = 0; ei < E; ei++)
#pragma unroll
            for (int f = 0; f < R; f++) {
                const C2 z = v[ei * R + f];
                __stcs(ye + 16 * (ei + 16 * f), make_float2(z.re.x, z.im.x));
                __stcs(yo + 16 * (ei + 16 * f), make_float2(z.re.y, z.im.y));
            }
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kFirThreads + kFftWarps * 32, 1) k_large_fused(const FusedParams fp)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = fp.base.M / kFirThreads;
    const int group = blockIdx.x / G, c = blockIdx.x % G;
    const long long n_batches = (fp.base.pair_end - fp.base.pair_begin) / kPairsPerBatch;
    const long long b0 = (n_batches * group) / fp.n_groups, b1 = (n_batches * (group + 1)) / fp.n_groups;
    if (threadIdx.x < kFirThreads) fused_fir_role<kTaps>(fp, group, c, b0, b1, smem_u32(smem_raw) + kDftSmem + kVStageBytes);
    else if (fp.base.M == 512) fused_dft_role_r<2>(fp, group, c, b0, b1, smem_u32(smem_raw), smem_u32(smem_raw) + kDftSmem);
    else if (fp.base.M == 1024) fused_dft_role_r<4>(fp, group, c, b0, b1, smem_u32(smem_raw), smem_u32(smem_raw) + kDftSmem);
    else if (fp.base.M == 2048) fused_dft_role_r<8>(fp, group, c, b0, b1, smem_u32(smem_raw), smem_u32(smem_raw) + kDftSmem);
    else fused_dft_role_r<16>(fp, group, c, b0, b1, smem_u32(smem_raw), smem_u32(smem_raw) + kDftSmem);
}

template <int kTaps>
int32_t launch_fused(const FusedParams& fp, cudaStream_t st)
{
    void* args[] = {const_cast<FusedParams*>(&fp)};              // the shared-memory attribute was set by plan_fused
    const int G = fp.base.M / kFirThreads;
    YG_CUDA(cudaLaunchCooperativeKernel((const void*)k_large_fused<kTaps>, dim3((unsigned)(G * fp.n_groups)),
                                        dim3(kFirThreads + kFftWarps * 32), args, (size_t)kFusedSmem, st));
    count_launch();
    return YG_OK;
}

// ------------------------------------------------------------------ single-SM fused analysis kernel: M = 1024, m <= 4
// The kernels above split a frame over G CTAs because M register-resident branch windows do not fit one SM, and pay for
// it with the V exchange through L2.  For M = 1024 with at most 9 taps per branch the frame fits one SM after all if the
// WINDOWS stay in shared memory and only the TAPS are register-resident:
//   * input ring: 16 frame pairs (128 KB, eight 16 KB TMA bulk copies of two pairs each); a pair stays in the ring until
//     the windows of eight later pairs have read it, so the ring IS the window storage;
//   * FIR role (warps 0-7): thread t owns the taps of branches t, t + 256, t + 512, t + 768 (72 registers); per batch of
//     two pairs and per branch it reads a transient 10-sample window from the ring (conflict-free LDS.64, 20 B per
//     output sample), runs the packed even/odd-frame FFMA2 dot products of K1 and stores the four frames' V values
//     UNPACKED (one 8 KB region per frame, double-buffered: 64 KB);
//   * DFT role (warps 8-15): one warp per frame, the 32 x 32 warp transform of the two-stage path (radix 32 in
//     registers, ONE XOR-swizzled exchange, done in place in the frame's V region), 256-byte coalesced stores.
// HBM and L2 see the algorithmic 24 B per sample; shared memory moves ~56 B per output sample (K1: 40).
#ifndef YG_S1K_STAMPS
#define YG_S1K_STAMPS 0        // debugging build (tools/build_variant.sh stamps firpfbch2_large.cu -DYG_S1K_STAMPS=1): the first,
#endif                         // middle and last CTA of k_m1024_fused print globaltimer stamps of their pipeline
#if YG_S1K_STAMPS
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define YG_STAMP(arr, i) do { (arr)[i] = gtime(); } while (0)
#else
#define YG_STAMP(arr, i) do { } while (0)
#endif
namespace s1k {
constexpr int kM = 1024, kM2 = 512;
constexpr int kBP = 2;                                   // frame pairs per batch
constexpr int kStages = 8;                               // ring of 8 batches = 16 pairs
constexpr int kStageBytes = kBP * kM * 8;                // 16 KB
constexpr int kRowBytes = kM * 8;                        // one pair of input = one frame of V = 8 KB
constexpr int kVBufBytes = 2 * kBP * kRowBytes;          // 4 frames
constexpr int kOffV = kStages * kStageBytes;             // 131072
constexpr int kOffTw = kOffV + 2 * kVBufBytes;           // W1024^{lane k1} as [k1 / 2][lane] float4 (k1 even, k1 odd): 8 KB
constexpr int kOffBar = kOffTw + 8192;
constexpr int kBarInFull = 0;                            // [8] TMA transaction barriers
constexpr int kBarInFree = 8;                            // [8] the 8 FIR warps no longer need the stage
constexpr int kBarVFull = 16;                            // [2] the 8 FIR warps have written the buffer
constexpr int kBarVFree = 18;                            // [2] its 4 DFT warps have read it
constexpr int kBarHist = 20;                             // the part of the slab's eight history pairs that is copied from x
constexpr int kSmem = kOffBar + 21 * 8;
constexpr int kThreads = 512;
constexpr int kMaxTaps = 9;
// column of branch t + 256 k inside a pair of input, plus t: pos(j) = (511 - j) mod 1024
__host__ __device__ constexpr int pos_k(int k) { return k == 0 ? 511 : k == 1 ? 255 : k == 2 ? 1023 : 767; }

template <int kTaps>
__device__ __forceinline__ void fir_role(const LargeParams& p, uint32_t smem, long long b0, long long b1, volatile unsigned long long* stamps)
{
    constexpr int kHist = kTaps - 1;                     // pairs of history a window reaches back
    const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
    const uint32_t bar = smem + kOffBar;

    const long long call_off = p.f0 * kM2;
    const long long q_first = p.pair_begin + b0 * kBP;   // first pair of the slab, relative to f0
    const float2* xb = p.x + call_off + q_first * kM;    // its first sample
    const long long nb = b1 - b0;

    auto issue_load = [&](long long lb) {                // local batch lb -> stage (lb + 4) mod 8
        const int st = (int)((lb + 4) & 7);
        mbar_expect_tx(bar + 8 * (kBarInFull + st), kStageBytes);
        tma_load_1d(smem + st * kStageBytes, xb + lb * (long long)(kBP * kM), kStageBytes, bar + 8 * (kBarInFull + st));
    };
    // The eight pairs before the slab (ring rows 0-7): whatever of them lies inside x (all of it for every slab but the
    // call's first) comes in with bulk copies as well, the rest -- history buffer or zeros -- with plain loads.
    const long long ta0 = (q_first - 8) * kM + call_off; // call-relative sample index of ring row 0 (a multiple of 512)
    const int n_head = ta0 >= 0 ? 0 : (int)min(-ta0, (long long)(8 * kM));
    if (t == 0) {                                        // first of all: get the copies going
        pdl_wait();                                      // x and the history may come from the previous kernel
        YG_STAMP(stamps, 0);                             // the clock of the stamps starts when the previous kernel is done
        if (n_head < 8 * kM) {
            const uint32_t bytes = (uint32_t)(8 * kM - n_head) * 8;
            mbar_expect_tx(bar + 8 * kBarHist, bytes);
            for (uint32_t off = 0; off < bytes; off += 16384)
                tma_load_1d(smem + n_head * 8 + off, p.x + (ta0 + n_head) + off / 8, min(16384u, bytes - off), bar + 8 * kBarHist);
        }
        for (long long lb = 0; lb < 4 && lb < nb; lb++) issue_load(lb);
    }

    float2 T[4][kTaps];                                  // taps of branches t + 256 k
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < kTaps; i++) T[k][i] = __ldg(&p.taps[(t + 256 * k) * kTaps + i]);
    pdl_wait();

    for (int i0 = 0; i0 < n_head; i0 += 8 * 256) {       // eight loads in flight per thread
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = i0 + u * 256 + t;
            const long long ta = ta0 + idx;
            v[u] = make_float2(0.f, 0.f);
            if (idx < n_head && p.Hlen + ta >= 0) v[u] = __ldg(&p.hist[p.Hlen + ta]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = i0 + u * 256 + t;
            if (idx < n_head) sts64(smem + idx * 8, v[u]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // these rows are overwritten by bulk copies later
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (t == 0) YG_STAMP(stamps, 1);                    // taps loaded, head rows stored
    if (n_head < 8 * kM) mbar_wait(bar + 8 * kBarHist, 0);
    if (t == 0) YG_STAMP(stamps, 2);                    // history rows landed

    const uint32_t ring_t = smem - t * 8;
    const uint32_t v_t = smem + kOffV + t * 8;

    auto do_batch = [&](auto ph_tag, long long lb) {
        constexpr int PH = decltype(ph_tag)::value;      // lb mod 8: every ring row below is a compile-time constant
        constexpr int ST = (PH + 4) & 7;
        constexpr int BUF = PH & 1;
        mbar_wait(bar + 8 * (kBarInFull + ST), (uint32_t)((lb >> 3) & 1));       // the batch's own two pairs have landed
        if (t == 0 && lb == 0) YG_STAMP(stamps, 3);     // first batch landed
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float2 w[kTaps + 1];                         // u[q0 - kHist .. q0 + 1]
#pragma unroll
            for (int i = 0; i <= kTaps; i++) w[i] = lds64(ring_t + ((2 * PH + 8 - kHist + i) & 15) * kRowBytes + pos_k(k) * 8);
            // accumulators are (re, im) of one frame: complex sample x broadcast real tap, so that a frame's value is
            // one aligned register pair for the 8-byte store (K1 packs (even, odd) instead and stores both frames at once)
            float2 e0 = make_float2(0.f, 0.f), o0 = e0, e1 = e0, o1 = e0;
#pragma unroll
            for (int i = kTaps - 1; i >= 0; i--) {       // oldest sample first, as K1
                e0 = fma2(w[kHist - i], f2(T[k][i].x), e0);
                o0 = fma2(w[kHist - i], f2(T[k][i].y), o0);
                e1 = fma2(w[kHist + 1 - i], f2(T[k][i].x), e1);
                o1 = fma2(w[kHist + 1 - i], f2(T[k][i].y), o1);
            }
            if (k == 0 && lb >= 2) mbar_wait(bar + 8 * (kBarVFree + BUF), (uint32_t)(((lb >> 1) - 1) & 1));
            const uint32_t vo = v_t + BUF * kVBufBytes + k * (256 * 8);
            sts64(vo + 0 * kRowBytes, e0);                               // pair 0, even frame
            sts64(vo + 1 * kRowBytes, o0);                               //         odd frame
            sts64(vo + 2 * kRowBytes, e1);                               // pair 1
            sts64(vo + 3 * kRowBytes, o1);
        }
        __syncwarp();
        if (t == 0 && lb == 0) YG_STAMP(stamps, 4);     // first V batch written
        if (t == 0 && lb == nb - 1) YG_STAMP(stamps, 6); // last V batch written
        if (lane == 0) {
            mbar_arrive(bar + 8 * (kBarVFull + BUF));
            mbar_arrive(bar + 8 * (kBarInFree + PH));    // rows 2 lb, 2 lb + 1 (stage lb mod 8) are out of every later window
            if (wrp == PH && lb + 4 < nb) {              // the FIR warps take turns refilling that stage with batch lb + 4
                mbar_wait(bar + 8 * (kBarInFree + PH), (uint32_t)((lb >> 3) & 1));
                issue_load(lb + 4);
            }
        }
    };
    for (long long lb = 0; lb < nb; lb += 8) {
        do_batch(std::integral_constant<int, 0>{}, lb);
        if (lb + 1 < nb) do_batch(std::integral_constant<int, 1>{}, lb + 1);
        if (lb + 2 < nb) do_batch(std::integral_constant<int, 2>{}, lb + 2);
        if (lb + 3 < nb) do_batch(std::integral_constant<int, 3>{}, lb + 3);
        if (lb + 4 < nb) do_batch(std::integral_constant<int, 4>{}, lb + 4);
        if (lb + 5 < nb) do_batch(std::integral_constant<int, 5>{}, lb + 5);
        if (lb + 6 < nb) do_batch(std::integral_constant<int, 6>{}, lb + 6);
        if (lb + 7 < nb) do_batch(std::integral_constant<int, 7>{}, lb + 7);
    }
}

__device__ __forceinline__ void dft_role(const LargeParams& p, const unsigned char* smem_raw, uint32_t smem, long long b0, long long b1,
                                         volatile unsigned long long* stamps)
{
    const int dt = threadIdx.x - 256, lane = dt & 31, dw = dt >> 5;
    const int buf = dw >> 2, fi = dw & 3;                // this warp's V buffer and frame (pair 0 even, odd, pair 1 even, odd)
    const uint32_t bar = smem + kOffBar;
    const uint32_t frame = smem + kOffV + buf * kVBufBytes + fi * kRowBytes;
    const uint32_t twt = smem + kOffTw + lane * 16;
    const long long nb = b1 - b0;
    pdl_wait();                                          // nothing is written before the previous kernel has completed
    // state hand-off folded into this launch (as in the M = 256 kernel): every CTA copies a slice of the next state
    // while its DFT warps wait for the first V buffer
    if (p.hist_new != nullptr) {
        const long long per = (p.Hlen + gridDim.x - 1) / gridDim.x;
        const long long i1 = min(p.Hlen, per * (long long)(blockIdx.x + 1));
        for (long long i = per * blockIdx.x + dt; i < i1; i += 256) {
            const long long ts = p.n_new - p.Hlen + i;
            p.hist_new[i] = (ts >= 0) ? __ldg(&p.x[ts]) : __ldg(&p.hist[p.Hlen + ts]);
        }
    }
    for (long long lb = buf; lb < nb; lb += 2) {
        mbar_wait(bar + 8 * (kBarVFull + buf), (uint32_t)((lb >> 1) & 1));
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) v[n1] = lds64(frame + (32 * n1 + lane) * 8);
        xdft32c(v);
        __syncwarp();                                    // every lane has read the frame: exchange in place
        // inter-pass twiddles from the transposed shared table: a gather twid[lane * k1] from global costs one L1 tag
        // lookup per distinct line, ~700 LSU cycles per frame, which made this role the bottleneck (0.56 -> see DESIGN)
#pragma unroll
        for (int a = 0; a < 16; a++) {
            const float4 w = lds128(twt + a * 512);
            float2 z0 = v[dr32(2 * a)], z1 = v[dr32(2 * a + 1)];
            if (a > 0) z0 = xmul(z0, w.x, w.y);
            z1 = xmul(z1, w.z, w.w);
            sts64(frame + (((lane << 5) | ((2 * a) ^ lane)) << 3), z0);
            sts64(frame + (((lane << 5) | ((2 * a + 1) ^ lane)) << 3), z1);
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(frame + (((n2 << 5) | (lane ^ n2)) << 3));
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + 8 * (kBarVFree + buf));
        xdft32c(v);
        const long long q = p.pair_begin + (b0 + lb) * kBP + (fi >> 1);
        float2* fr = p.y + (p.f0 + 2 * q + (fi & 1)) * (long long)kM + lane;
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) __stcs(fr + 32 * k2, v[dr32(k2)]);
        if (dt == 0 && lb == 0) YG_STAMP(stamps, 5);    // first frame stored
    }
#if YG_S1K_STAMPS
    if (dt == 128 * ((nb - 1) & 1) && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 || blockIdx.x == gridDim.x - 1)) {
        const unsigned long long t7 = gtime(), t0 = stamps[0];
        printf("cta %3d nb %3lld: taps %6llu hist %6llu batch0 %6llu V0 %6llu out0 %6llu lastV %6llu end %6llu ns\n", (int)blockIdx.x, nb,
               stamps[1] - t0, stamps[2] - t0, stamps[3] - t0, stamps[4] - t0, stamps[5] - t0, stamps[6] - t0, t7 - t0);
    }
#endif
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_m1024_fused(const LargeParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
#if YG_S1K_STAMPS
    __shared__ unsigned long long stamps_s[8];
    volatile unsigned long long* stamps = stamps_s;
#else
    volatile unsigned long long* stamps = nullptr;
#endif
    const long long n_batches = (p.pair_end - p.pair_begin) / kBP;
    const long long b0 = (n_batches * blockIdx.x) / gridDim.x, b1 = (n_batches * (blockIdx.x + 1)) / gridDim.x;
    if (threadIdx.x == 0) {
        const uint32_t bar = smem + kOffBar;
        for (int i = 0; i < 8; i++) mbar_init(bar + 8 * (kBarInFull + i), 1);
        for (int i = 0; i < 8; i++) mbar_init(bar + 8 * (kBarInFree + i), 8);
        for (int i = 0; i < 2; i++) mbar_init(bar + 8 * (kBarVFull + i), 8);
        for (int i = 0; i < 2; i++) mbar_init(bar + 8 * (kBarVFree + i), 4);
        mbar_init(bar + 8 * kBarHist, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // (filling this table inside the DFT role, off the FIR role's start-up path, made the kernels 3-4 % slower: same-box A/B)
    for (int i = threadIdx.x; i < 1024; i += kThreads) {              // entry i: k1 = 2 (i >> 6) + (i & 1), lane = (i >> 1) & 31
        const int k1 = 2 * (i >> 6) + (i & 1), ln = (i >> 1) & 31;
        sts64(smem + kOffTw + i * 8, __ldg(&p.twid[ln * k1]));
    }
    __syncthreads();
    pdl_launch_dependents();
    if (b0 >= b1) return;                                // never taken: the grid has at most one CTA per batch
    if (threadIdx.x < 256) fir_role<kTaps>(p, smem, b0, b1, stamps);
    else dft_role(p, smem_raw, smem, b0, b1, stamps);
}

template <int kTaps>
int32_t launch(const Firpfbch2FastPlan& plan, const LargeParams& p, cudaStream_t st)
{
    const long long n_batches = (p.pair_end - p.pair_begin) / kBP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>(plan.n_sm, n_batches));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = plan.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YG_CUDA(cudaLaunchKernelEx(&cfg, k_m1024_fused<kTaps>, p));
    count_launch();
    return YG_OK;
}

// plan time, on the object's device: the kernel instance may use kSmem bytes of dynamic shared memory
inline int32_t prepare(uint32_t m)
{
    switch (m) {
        case 1: YG_CUDA(cudaFuncSetAttribute(k_m1024_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        case 2: YG_CUDA(cudaFuncSetAttribute(k_m1024_fused<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        case 3: YG_CUDA(cudaFuncSetAttribute(k_m1024_fused<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        default: YG_CUDA(cudaFuncSetAttribute(k_m1024_fused<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
    }
    return YG_OK;
}
}  // namespace s1k

// ------------------------------------------------------------------ single-SM fused analysis kernel: M = 512, m <= 7
// s1k re-cut for half the frame length: a pair of input is 4 KB, so the same 128 KB ring holds 32 pairs (16 of history,
// enough for 15 taps per branch), a batch is FOUR pairs (still one 16 KB bulk copy, still 32 KB of V), a FIR thread owns
// the taps of TWO branches (t, t + 256; 60 registers at m = 7) and walks the batch as two half-batches of two pairs, and a
// DFT warp takes a PAIR of frames: 16 x 32 -- lane n2 runs the 16-point transforms over n1 of both frames, one
// XOR-swizzled exchange in place in the pair's 8 KB of V, lane (frame, k1) runs one 32-point transform over n2.
namespace s5k {
constexpr int kM = 512, kM2 = 256;
constexpr int kBP = 4;                                   // frame pairs per batch
constexpr int kStages = 8;                               // ring of 8 batches = 32 pairs
constexpr int kStageBytes = kBP * kM * 8;                // 16 KB
constexpr int kRowBytes = kM * 8;                        // one pair of input = one frame of V = 4 KB
constexpr int kHistPairs = 16;                           // ring rows 0-15 hold the pairs before the slab
constexpr int kVBufBytes = 2 * kBP * kRowBytes;          // 8 frames
constexpr int kOffV = kStages * kStageBytes;             // 131072
constexpr int kOffTw = kOffV + 2 * kVBufBytes;           // W512^{lane k1} as [k1 / 2][lane] float4 (k1 even, k1 odd): 4 KB
constexpr int kOffBar = kOffTw + 4096;
constexpr int kBarInFull = 0;                            // [8] TMA transaction barriers
constexpr int kBarInFree = 8;                            // [8] the 8 FIR warps no longer need the stage
constexpr int kBarVFull = 16;                            // [2] the 8 FIR warps have written the buffer
constexpr int kBarVFree = 18;                            // [2] its 4 DFT warps have read it
constexpr int kBarHist = 20;                             // the part of the slab's history pairs that is copied from x
constexpr int kSmem = kOffBar + 21 * 8;
constexpr int kThreads = 512;
constexpr int kMaxM = 7;
// column of branch t + 256 k inside a pair of input, plus t: pos(j) = (255 - j) mod 512
__host__ __device__ constexpr int pos_k(int k) { return k == 0 ? 255 : 511; }

template <int kTaps>
__device__ __forceinline__ void fir_role(const LargeParams& p, uint32_t smem, long long b0, long long b1)
{
    constexpr int kHist = kTaps - 1;                     // pairs of history a window reaches back (<= 14)
    const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
    const uint32_t bar = smem + kOffBar;

    const long long call_off = p.f0 * kM2;
    const long long q_first = p.pair_begin + b0 * kBP;   // first pair of the slab, relative to f0
    const float2* xb = p.x + call_off + q_first * kM;    // its first sample
    const long long nb = b1 - b0;

    auto issue_load = [&](long long lb) {                // local batch lb -> stage (lb + 4) mod 8
        const int st = (int)((lb + 4) & 7);
        mbar_expect_tx(bar + 8 * (kBarInFull + st), kStageBytes);
        tma_load_1d(smem + st * kStageBytes, xb + lb * (long long)(kBP * kM), kStageBytes, bar + 8 * (kBarInFull + st));
    };
    // the sixteen pairs before the slab (ring rows 0-15): bulk copies for what lies inside x, plain loads for the rest
    constexpr int kHead = kHistPairs * kM;               // 8192 samples
    const long long ta0 = (q_first - kHistPairs) * kM + call_off;
    const int n_head = ta0 >= 0 ? 0 : (int)min(-ta0, (long long)kHead);
    if (t == 0) {
        pdl_wait();                                      // x and the history may come from the previous kernel
        if (n_head < kHead) {
            const uint32_t bytes = (uint32_t)(kHead - n_head) * 8;
            mbar_expect_tx(bar + 8 * kBarHist, bytes);
            for (uint32_t off = 0; off < bytes; off += 16384)
                tma_load_1d(smem + n_head * 8 + off, p.x + (ta0 + n_head) + off / 8, min(16384u, bytes - off), bar + 8 * kBarHist);
        }
        for (long long lb = 0; lb < 4 && lb < nb; lb++) issue_load(lb);
    }

    float2 T[2][kTaps];                                  // taps of branches t, t + 256
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
        for (int i = 0; i < kTaps; i++) T[k][i] = __ldg(&p.taps[(t + 256 * k) * kTaps + i]);
    pdl_wait();

    for (int i0 = 0; i0 < n_head; i0 += 8 * 256) {       // eight loads in flight per thread
        float2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = i0 + u * 256 + t;
            const long long ta = ta0 + idx;
            v[u] = make_float2(0.f, 0.f);
            if (idx < n_head && p.Hlen + ta >= 0) v[u] = __ldg(&p.hist[p.Hlen + ta]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = i0 + u * 256 + t;
            if (idx < n_head) sts64(smem + idx * 8, v[u]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // these rows are overwritten by bulk copies later
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (n_head < kHead) mbar_wait(bar + 8 * kBarHist, 0);

    const uint32_t ring_t = smem - t * 8;
    const uint32_t v_t = smem + kOffV + t * 8;

    auto do_batch = [&](auto ph_tag, long long lb) {
        constexpr int PH = decltype(ph_tag)::value;      // lb mod 8: every ring row below is a compile-time constant
        constexpr int ST = (PH + 4) & 7;
        constexpr int BUF = PH & 1;
        mbar_wait(bar + 8 * (kBarInFull + ST), (uint32_t)((lb >> 3) & 1));       // the batch's own four pairs have landed
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                float2 w[kTaps + 1];                     // u[q0 - kHist .. q0 + 1], q0 = 4 lb + 2 half
#pragma unroll
                for (int i = 0; i <= kTaps; i++)
                    w[i] = lds64(ring_t + ((4 * PH + 2 * half + kHistPairs - kHist + i) & 31) * kRowBytes + pos_k(k) * 8);
                float2 e0 = make_float2(0.f, 0.f), o0 = e0, e1 = e0, o1 = e0;
#pragma unroll
                for (int i = kTaps - 1; i >= 0; i--) {   // oldest sample first, as K1
                    e0 = fma2(w[kHist - i], f2(T[k][i].x), e0);
                    o0 = fma2(w[kHist - i], f2(T[k][i].y), o0);
                    e1 = fma2(w[kHist + 1 - i], f2(T[k][i].x), e1);
                    o1 = fma2(w[kHist + 1 - i], f2(T[k][i].y), o1);
                }
                if (half == 0 && k == 0 && lb >= 2) mbar_wait(bar + 8 * (kBarVFree + BUF), (uint32_t)(((lb >> 1) - 1) & 1));
                const uint32_t vo = v_t + BUF * kVBufBytes + (4 * half) * kRowBytes + k * (256 * 8);
                sts64(vo + 0 * kRowBytes, e0);                           // pair 2 half: even frame
                sts64(vo + 1 * kRowBytes, o0);                           //              odd frame
                sts64(vo + 2 * kRowBytes, e1);                           // pair 2 half + 1
                sts64(vo + 3 * kRowBytes, o1);
            }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bar + 8 * (kBarVFull + BUF));
            mbar_arrive(bar + 8 * (kBarInFree + PH));    // the pairs of batch lb - 4 (stage lb mod 8) are out of every later window
            if (wrp == PH && lb + 4 < nb) {              // the FIR warps take turns refilling that stage with batch lb + 4
                mbar_wait(bar + 8 * (kBarInFree + PH), (uint32_t)((lb >> 3) & 1));
                issue_load(lb + 4);
            }
        }
    };
    for (long long lb = 0; lb < nb; lb += 8) {
        do_batch(std::integral_constant<int, 0>{}, lb);
        if (lb + 1 < nb) do_batch(std::integral_constant<int, 1>{}, lb + 1);
        if (lb + 2 < nb) do_batch(std::integral_constant<int, 2>{}, lb + 2);
        if (lb + 3 < nb) do_batch(std::integral_constant<int, 3>{}, lb + 3);
        if (lb + 4 < nb) do_batch(std::integral_constant<int, 4>{}, lb + 4);
        if (lb + 5 < nb) do_batch(std::integral_constant<int, 5>{}, lb + 5);
        if (lb + 6 < nb) do_batch(std::integral_constant<int, 6>{}, lb + 6);
        if (lb + 7 < nb) do_batch(std::integral_constant<int, 7>{}, lb + 7);
    }
}

// 512-point backward DFTs of the two frames in `tile` (frame h at tile + 4096 h, natural order), in place as far as the
// exchange goes; on return lane (h, k1) = (lane >> 4, lane & 15) holds X_h[k1 + 16 k2] in v[dr32(k2)].  `twt` = this
// lane's column of the shared twiddle table; `on_read` runs when the tile has been read for the last time.
template <typename F>
__device__ __forceinline__ void pair_dft512(float2 (&v)[32], uint32_t tile, uint32_t twt, int lane, F on_read)
{
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) v[16 * h + n1] = lds64(tile + h * 4096 + (32 * n1 + lane) * 8);
    xdft16<0>(v);
    xdft16<16>(v);
    __syncwarp();                                        // every lane has read both frames: exchange in place
#pragma unroll
    for (int a = 0; a < 8; a++) {
        const float4 w = lds128(twt + a * 512);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float2 z0 = v[16 * h + dr4(2 * a)], z1 = v[16 * h + dr4(2 * a + 1)];
            if (a > 0) z0 = xmul(z0, w.x, w.y);
            z1 = xmul(z1, w.z, w.w);
            sts64(tile + (((h << 9) | (lane << 4) | ((2 * a) ^ (lane & 15))) << 3), z0);
            sts64(tile + (((h << 9) | (lane << 4) | ((2 * a + 1) ^ (lane & 15))) << 3), z1);
        }
    }
    __syncwarp();
    const int h = lane >> 4, k1 = lane & 15;
#pragma unroll
    for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(tile + (((h << 9) | (n2 << 4) | (k1 ^ (n2 & 15))) << 3));
    __syncwarp();
    on_read();
    xdft32c(v);
}

__device__ __forceinline__ void dft_role(const LargeParams& p, uint32_t smem, long long b0, long long b1)
{
    const int dt = threadIdx.x - 256, lane = dt & 31, dw = dt >> 5;
    const int buf = dw >> 2, fi = dw & 3;                // this warp's V buffer and frame pair of the batch
    const uint32_t bar = smem + kOffBar;
    const uint32_t tile = smem + kOffV + buf * kVBufBytes + fi * (2 * kRowBytes);
    const uint32_t twt = smem + kOffTw + lane * 16;
    const long long nb = b1 - b0;
    pdl_wait();                                          // nothing is written before the previous kernel has completed
    if (p.hist_new != nullptr) {                         // every CTA copies a slice of the object's next state
        const long long per = (p.Hlen + gridDim.x - 1) / gridDim.x;
        const long long i1 = min(p.Hlen, per * (long long)(blockIdx.x + 1));
        for (long long i = per * blockIdx.x + dt; i < i1; i += 256) {
            const long long ts = p.n_new - p.Hlen + i;
            p.hist_new[i] = (ts >= 0) ? __ldg(&p.x[ts]) : __ldg(&p.hist[p.Hlen + ts]);
        }
    }
    for (long long lb = buf; lb < nb; lb += 2) {
        mbar_wait(bar + 8 * (kBarVFull + buf), (uint32_t)((lb >> 1) & 1));
        float2 v[32];
        pair_dft512(v, tile, twt, lane, [&] { if (lane == 0) mbar_arrive(bar + 8 * (kBarVFree + buf)); });
        const long long q = p.pair_begin + (b0 + lb) * kBP + fi;
        float2* fr = p.y + (p.f0 + 2 * q + (lane >> 4)) * (long long)kM + (lane & 15);
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) __stcs(fr + 16 * k2, v[dr32(k2)]);
    }
}

__device__ __forceinline__ void fill_twiddles(uint32_t smem_tw, const float2* __restrict__ twid)
{
    for (int i = threadIdx.x; i < 512; i += kThreads) {               // entry i: k1 = 2 (i >> 6) + (i & 1), lane = (i >> 1) & 31
        const int k1 = 2 * (i >> 6) + (i & 1), ln = (i >> 1) & 31;
        sts64(smem_tw + i * 8, __ldg(&twid[ln * k1]));
    }
}

template <int kTaps>
__global__ void __launch_bounds__(kThreads, 1) k_m512_fused(const LargeParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const long long n_batches = (p.pair_end - p.pair_begin) / kBP;
    const long long b0 = (n_batches * blockIdx.x) / gridDim.x, b1 = (n_batches * (blockIdx.x + 1)) / gridDim.x;
    if (threadIdx.x == 0) {
        const uint32_t bar = smem + kOffBar;
        for (int i = 0; i < 8; i++) mbar_init(bar + 8 * (kBarInFull + i), 1);
        for (int i = 0; i < 8; i++) mbar_init(bar + 8 * (kBarInFree + i), 8);
        for (int i = 0; i < 2; i++) mbar_init(bar + 8 * (kBarVFull + i), 8);
        for (int i = 0; i < 2; i++) mbar_init(bar + 8 * (kBarVFree + i), 4);
        mbar_init(bar + 8 * kBarHist, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    fill_twiddles(smem + kOffTw, p.twid);
    __syncthreads();
    pdl_launch_dependents();
    if (b0 >= b1) return;                                // never taken: the grid has at most one CTA per batch
    if (threadIdx.x < 256) fir_role<kTaps>(p, smem, b0, b1);
    else dft_role(p, smem, b0, b1);
}

template <int kTaps>
int32_t launch(const Firpfbch2FastPlan& plan, const LargeParams& p, cudaStream_t st)
{
    const long long n_batches = (p.pair_end - p.pair_begin) / kBP;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>(plan.n_sm, n_batches));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = plan.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YG_CUDA(cudaLaunchKernelEx(&cfg, k_m512_fused<kTaps>, p));
    count_launch();
    return YG_OK;
}

template <int kTaps>
int32_t prepare_one() { YG_CUDA(cudaFuncSetAttribute(k_m512_fused<kTaps>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); return YG_OK; }

// per tap count: f(std::integral_constant<int, 2 m + 1>) for m = 1 .. 7
template <typename F>
int32_t for_m(uint32_t m, F f)
{
    switch (m) {
        case 1: return f(std::integral_constant<int, 3>{});
        case 2: return f(std::integral_constant<int, 5>{});
        case 3: return f(std::integral_constant<int, 7>{});
        case 4: return f(std::integral_constant<int, 9>{});
        case 5: return f(std::integral_constant<int, 11>{});
        case 6: return f(std::integral_constant<int, 13>{});
        case 7: return f(std::integral_constant<int, 15>{});
        default: return fail(YG_EINTERNAL, "single-SM M = 512 kernel not instantiated for m = %u", m);
    }
}
inline int32_t prepare(uint32_t m) { return for_m(m, [](auto tag) { return prepare_one<decltype(tag)::value>(); }); }
}  // namespace s5k

int32_t launch_fft(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, long long v0, float2* dst,
                   long long n_frames, int streaming, cudaStream_t st)
{
    const float2* tw = reinterpret_cast<const float2*>(plan.d_twid);
    if (plan.M >= 1024) {
        const long long items = n_frames * (plan.M >> 10);
        const int grid = (int)std::min<long long>((items + kFftWarps - 1) / kFftWarps, (long long)plan.n_sm * 2);
        if (plan.M == 1024) k_large_fft<1><<<grid, kFftWarps * 32, kFftSmem, st>>>(prefix, x, v0, dst, n_frames, tw, streaming);
        else if (plan.M == 2048) k_large_fft<2><<<grid, kFftWarps * 32, kFftSmem, st>>>(prefix, x, v0, dst, n_frames, tw, streaming);
        else k_large_fft<4><<<grid, kFftWarps * 32, kFftSmem, st>>>(prefix, x, v0, dst, n_frames, tw, streaming);
    } else if (plan.M == 512) {
        const long long items = (n_frames + 1) >> 1;
        const int grid = (int)std::min<long long>((items + kFftWarps - 1) / kFftWarps, (long long)plan.n_sm * 2);
        k_large_fft512<<<grid, kFftWarps * 32, kFftSmem, st>>>(prefix, x, v0, dst, n_frames, tw, streaming);
    } else {
        return fail(YG_EINTERNAL, "large-M path: no transform for M = %u", plan.M);
    }
    YG_LAUNCH_CHECK();
    return YG_OK;
}

// ------------------------------------------------------------------ synthesis stage C: weighted overlap-add
// Thread j owns COLUMN j of U (the firpfbch2_synth_fast.cu FIR role with global loads): 32-entry register
// ring over the last 4m frames, values prefetched 8 steps (16 frames) ahead through a 16-entry register
// queue.  Columns j < M/2 emit on even frames, columns j >= M/2 on odd frames (ring kept one frame behind).
struct WolaParams {
    const float2* U;          // U[0] = virtual frame c0 - 32 (the warm-up period), frames of kM columns
    float2* y;                // output sample 0 of frame c0
    long long n_frames;       // frames of this chunk (multiple of 32); U holds n_frames + 32 frames
    int slabs;
    int M;
    const float* taps;        // [M][4m]  0.5 * h[(j & (M/2-1)) + l * M/2]
};

template <int kTaps>
__global__ void __launch_bounds__(kFirThreads, 2) k_large_wola(const WolaParams p)
{
    const int kM = p.M, kM2 = p.M >> 1;
    const int j = blockIdx.y * kFirThreads + threadIdx.x;
    const bool hi = j >= kM2;
    const int i = j & (kM2 - 1);
    const long long n_periods = p.n_frames / 32;
    const long long r0 = (n_periods * blockIdx.x) / p.slabs, r1 = (n_periods * (blockIdx.x + 1)) / p.slabs;
    if (r0 >= r1) return;

    float T[kTaps];
#pragma unroll
    for (int l = 0; l < kTaps; l++) T[l] = __ldg(&p.taps[j * kTaps + l]);
    float2 W[32];
#pragma unroll
    for (int s = 0; s < 32; s++) W[s] = make_float2(0.f, 0.f);

    // slab-local frame index g: g = 0 is the first warm-up frame = U frame (r0 * 32); real output frames are
    // g in [32, 32 + n_out).  Lower half: slot k holds frame k; upper half: slot k holds frame k - 1.
    const int n_out = (int)(r1 - r0) * 32;
    const float2* Ub = p.U + (r0 * 32) * (long long)kM + j;
    float2* yb = p.y + (r0 * 32) * (long long)kM2 + i;
    const int o = hi ? -1 : 0;
    const int g_last = 32 + n_out - 1;                                   // last U frame that exists for this slab
    auto fetch = [&](int g) {                                            // frame g of the slab (clamped; g = -1 only feeds suppressed outputs)
        g = g < 0 ? 0 : (g > g_last ? g_last : g);
        return __ldg(Ub + (long long)g * kM);
    };
    float2 Pf[16];                                                       // steps s .. s+7 in flight
#pragma unroll
    for (int s = 0; s < 8; s++) { Pf[2 * s] = fetch(2 * s + o); Pf[2 * s + 1] = fetch(2 * s + 1 + o); }

    const int n_steps = (32 + n_out) / 2;
    for (int s0 = 0; s0 < n_steps; s0 += 16) {
#pragma unroll
        for (int ss = 0; ss < 16; ss++) {
            const int s = s0 + ss;
            W[(2 * ss) & 31] = Pf[(2 * ss) & 15];
            W[(2 * ss + 1) & 31] = Pf[(2 * ss + 1) & 15];
            Pf[(2 * ss) & 15] = fetch(2 * (s + 8) + o);
            Pf[(2 * ss + 1) & 15] = fetch(2 * (s + 8) + 1 + o);
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int l = kTaps - 1; l >= 0; l--) {
                const float2 w = W[(2 * ss - l) & 31];
                if (l & 1) a1 = fma2(w, f2(T[l]), a1);
                else a0 = fma2(w, f2(T[l]), a0);
            }
            const int rel = 2 * s + o - 32;                              // output frame relative to the slab's first real frame
            if ((unsigned)rel < (unsigned)n_out) __stcs(yb + (long long)rel * kM2, add2(a0, a1));
        }
    }
    if (hi) {                                                            // last odd frame of the slab: window ends at slot 0
        W[0] = Pf[0];
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = kTaps - 1; l >= 0; l--) {
            const float2 w = W[(0 - l) & 31];
            if (l & 1) a1 = fma2(w, f2(T[l]), a1);
            else a0 = fma2(w, f2(T[l]), a0);
        }
        __stcs(yb + (long long)(n_out - 1) * kM2, add2(a0, a1));
    }
}

// ------------------------------------------------------------------ fused synthesis kernel
// The mirror image of k_large_fused: in every group of G = M / 256 CTAs the DFT teams (warps 8-15) transform
// input frame pairs (staged from HBM with cp.async one batch ahead) and publish U batches of 16 pairs in the
// group's L2-resident ring; the overlap-add threads (warps 0-7, one branch each, the k_large_wola arithmetic)
// read their branch's column of the ring eight steps ahead of use.  A slab starts with one warm-up batch: the
// 32 input frames before its first output frame.
struct SynthFusedParams {
    const float2* prefix;     // the 32 input frames preceding x[0]
    const float2* x;          // input frames of the call, [frame][M]
    float2* y;                // output sample 0 of the call
    long long f0;             // first frame handled (even global parity)
    long long n_batches;      // output batches of 32 frames
    int M;
    const float* taps;        // [M][4m]
    const float2* twid;       // [M]
    float4* scratch;          // [n_groups][kSlots][16 pairs][M] packed (re_e, re_o, im_e, im_o)
    unsigned* flags;
    int n_groups;
};

constexpr int kXStageBytes = 256 * 32 * 8;                               // one pair per DFT thread column: 2 x 16 samples
constexpr int kUStageBytes = 256 * 16 * 16;                              // one batch of U per overlap-add thread column
constexpr int synth_fused_smem(int taps) { return kDftSmem + kXStageBytes + (taps > 16 ? kUStageBytes : 0); }

template <int R>
__device__ __forceinline__ void synth_dft_role_r(const SynthFusedParams& p, int group, int c, long long B0, long long B1,
                                                 uint32_t smem_dft, uint32_t smem_stage)
{
    constexpr int T = 16 * R, kM = 256 * R, E = 16 / R, G = R;
    const int lane = threadIdx.x & 31;
    const int dt = threadIdx.x - kFirThreads;
    const int team = dt / T, tt = dt % T;
    const int hi = tt >> 4, k1 = tt & 15;
    const int pr = c * (256 / T) + team;
    const uint32_t region = smem_dft + team * (T * 17 * 16);
    const unsigned n_wola_warps = (unsigned)((kFirThreads / 32) * G);
    unsigned* flags = p.flags + group * kFlagStride;
    float4* sbase = p.scratch + ((long long)group * kSlots * kPairsPerBatch + pr) * kM + k1 + 16 * (hi * E);
    const long long nb = B1 - B0 + 1;                                    // batches of this slab, warm-up included
    // X staging: row 2 n1 + parity of the column block holds sample T n1 + tt of that frame of the pair; a lane pair
    // shares 16-byte copies (even lane: rows 0-15, odd lane: rows 16-31)
    const uint32_t xstage = smem_stage + dt * 8;
    const uint32_t xstage_wr = smem_stage + (dt & ~1) * 8 + (dt & 1) * (16 * 256 * 8);
    auto fetch_x = [&](long long lb) {
        // frames vi, vi + 1; with an odd call-relative start the pair can straddle the prefix | x boundary
        const long long vi = p.f0 + 32 * (B0 + lb - 1) + 2 * pr;
        const long long off = (tt & ~1) + (dt & 1) * (8 * T);
        const float2* se = ((vi < 0) ? p.prefix + (32 + vi) * kM : p.x + vi * kM) + off;
        const float2* so = ((vi + 1 < 0) ? p.prefix + (33 + vi) * kM : p.x + (vi + 1) * kM) + off;
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xstage_wr + (2 * n1) * (256 * 8)), "l"(se + T * n1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xstage_wr + (2 * n1 + 1) * (256 * 8)), "l"(so + T * n1) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_x(0);
    unsigned seen_free = 0;
    for (long long lb = 0; lb < nb; lb++) {
        const int slot = (int)(lb % kSlots);
        const unsigned use = (unsigned)(lb / kSlots);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                                    // the partner lane's rows have landed too
        C2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const float2 e = lds64(xstage + (2 * n1) * (256 * 8)), o = lds64(xstage + (2 * n1 + 1) * (256 * 8));
            v[n1].re = make_float2(e.x, o.x);
            v[n1].im = make_float2(e.y, o.y);
        }
        dft16(v);
        if (lb + 1 < nb) fetch_x(lb + 1);                                // own column has been consumed: refill it
        const unsigned seen_now = seen_free;
        seen_free = warp_peek(flags + kSlots + (int)((lb + 1) % kSlots), lane);
        team_dft_tail<R>(v, region, team, tt, p.twid);
        if (use) warp_wait_seen(seen_now, flags + kSlots + slot, use * n_wola_warps, lane);  // the slot's previous batch has been read
        float4* ub = sbase + (long long)slot * kPairsPerBatch * kM;
#pragma unroll
        for (int ei = 0; ei < E; ei++)
#pragma unroll
            for (int f = 0; f < R; f++) {
                const C2 z = v[ei * R + f];
                __stcg(ub + 16 * (ei + 16 * f), make_float4(z.re.x, z.re.y, z.im.x, z.im.y));
            }
        warp_publish(flags + slot, lane);
    }
}

template <int kTaps>
__device__ __forceinline__ void synth_wola_role(const SynthFusedParams& p, int group, int c, long long B0, long long B1,
                                                uint32_t smem_stage)
{
    const int kM = p.M, kM2 = p.M >> 1;
    const int lane = threadIdx.x & 31;
    const int j = c * kFirThreads + threadIdx.x;
    const bool hi = j >= kM2;
    const int i = j & (kM2 - 1);
    const unsigned n_dft_warps = (unsigned)(8 * (kM / kFirThreads));
    unsigned* flags = p.flags + group * kFlagStride;
    const float4* sbase = p.scratch + (long long)group * kSlots * kPairsPerBatch * kM + j;
    const long long nb = B1 - B0 + 1;

    float T[kTaps];
#pragma unroll
    for (int l = 0; l < kTaps; l++) T[l] = __ldg(&p.taps[j * kTaps + l]);
    float2 W[32];
#pragma unroll
    for (int s = 0; s < 32; s++) W[s] = make_float2(0.f, 0.f);
    float2* yb = p.y + (p.f0 + 32 * B0) * (long long)kM2 + i;             // first real output frame of the slab
    const long long n_out = 32 * (B1 - B0);

    // U of this branch arrives eight steps (pairs) ahead of use: in registers while the taps leave room for them
    // (kTaps <= 16), otherwise through a private shared column filled half a batch at a time with cp.async
    constexpr bool kStageU = kTaps > 16;
    const uint32_t ustage = smem_stage + threadIdx.x * 16;               // + half * 32 KB + pair * 4 KB
    auto fetch = [&](long long q) {                                      // pair q of the slab, q = 16 lb + ss
        const long long lb = q >> 4;
        return __ldcg(sbase + ((long long)(lb % kSlots) * kPairsPerBatch + (q & 15)) * kM);
    };
    auto fetch8 = [&](long long lb, int half) {
        const float4* src = sbase + ((long long)(lb % kSlots) * kPairsPerBatch + 8 * half) * kM;
#pragma unroll
        for (int r = 0; r < 8; r++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ustage + half * (8 * 4096) + r * 4096),
                         "l"(src + (long long)r * kM) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float4 Pf[kStageU ? 1 : 8];
    warp_wait(flags + 0, n_dft_warps, lane);
    if (kStageU) fetch8(0, 0);
    else {
#pragma unroll
        for (int s = 0; s < 8; s++) Pf[s & (kStageU ? 0 : 7)] = fetch(s);
    }
    float2 carry = make_float2(0.f, 0.f);                                // odd frame of the previous pair (upper half)
    unsigned seen_full = (nb > 1) ? warp_peek(flags + 1 % kSlots, lane) : 0u;
    for (long long lb = 0; lb < nb; lb++) {
#pragma unroll
        for (int ss = 0; ss < 16; ss++) {
            const long long q = 16 * lb + ss;
            if (kStageU && ss == 0) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                fetch8(lb, 1);
            }
            if (ss == 8) {                                               // the prefetch is about to cross into the next batch
                if (kStageU) asm volatile("cp.async.wait_group 0;" ::: "memory");
                if (lb + 1 < nb) {
                    const unsigned seen_now = seen_full;
                    if (lb + 2 < nb) seen_full = warp_peek(flags + (int)((lb + 2) % kSlots), lane);
                    warp_wait_seen(seen_now, flags + (int)((lb + 1) % kSlots), (unsigned)((lb + 1) / kSlots + 1) * n_dft_warps, lane);
                    if (kStageU) fetch8(lb + 1, 0);
                }
            }
            float4 u;
            if (kStageU) u = lds128(ustage + ss * 4096);
            else {
                u = Pf[ss & (kStageU ? 0 : 7)];
                if (q + 8 < 16 * nb) Pf[ss & (kStageU ? 0 : 7)] = fetch(q + 8);
            }
            // lower half: slots (2ss, 2ss+1) = frames (2q, 2q+1); upper half: frames (2q-1, 2q)
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
            const long long rel = 2 * q + (hi ? -1 : 0) - 32;             // output frame relative to the slab's first real frame
            if (rel >= 0 && rel < n_out) __stcs(yb + rel * kM2, add2(a0, a1));
        }
        warp_retire(flags + kSlots + (int)(lb % kSlots), lane);          // every copy out of this batch's slot has landed
    }
    if (hi) {                                                            // last odd frame of the slab: window ends at slot 0
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

template <int kTaps>
__global__ void __launch_bounds__(kFirThreads + 256, 1) k_large_synth_fused(const SynthFusedParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = p.M / kFirThreads;
    const int group = blockIdx.x / G, c = blockIdx.x % G;
    const long long B0 = (p.n_batches * group) / p.n_groups, B1 = (p.n_batches * (group + 1)) / p.n_groups;
    if (B0 >= B1) return;
    const uint32_t smem = smem_u32(smem_raw);
    if (threadIdx.x < kFirThreads) synth_wola_role<kTaps>(p, group, c, B0, B1, smem + kDftSmem + kXStageBytes);
    else if (p.M == 512) synth_dft_role_r<2>(p, group, c, B0, B1, smem, smem + kDftSmem);
    else if (p.M == 1024) synth_dft_role_r<4>(p, group, c, B0, B1, smem, smem + kDftSmem);
    else if (p.M == 2048) synth_dft_role_r<8>(p, group, c, B0, B1, smem, smem + kDftSmem);
    else synth_dft_role_r<16>(p, group, c, B0, B1, smem, smem + kDftSmem);
}

template <int kTaps>
int32_t launch_synth_fused(const SynthFusedParams& p, cudaStream_t st)
{
    void* args[] = {const_cast<SynthFusedParams*>(&p)};          // the shared-memory attribute was set by plan_fused
    const int G = p.M / kFirThreads;
    YG_CUDA(cudaLaunchCooperativeKernel((const void*)k_large_synth_fused<kTaps>, dim3((unsigned)(G * p.n_groups)),
                                        dim3(kFirThreads + 256), args, (size_t)synth_fused_smem(kTaps), st));
    count_launch();
    return YG_OK;
}

// ------------------------------------------------------------------ single-SM fused synthesis kernel: M = 1024, m <= 4
// The mirror image of s1k.  Holding the overlap-add WINDOWS of 1024 columns on one SM would take 4m frames of U (128 KB
// at m = 4) on top of the frames in flight; holding the partial OUTPUT sums instead takes half of that and fits the
// register file.  With g_i[l] = h[i + l M/2] / 2 and c = i + (f odd ? M/2 : 0)
//       y_f[i] = sum_{l < 4m} g_i[l] u_{f-l}[c],
// so a value u_g[c] is read ONCE, when frame g arrives, and added into the 2m outputs of its column's parity that it
// reaches (taps of even lag from a frame of that parity, taps of odd lag from a frame of the other parity):
//   * frame ring: 6 stages of 4 input frames (192 KB); a stage is filled by TMA bulk copies, transformed IN PLACE by four
//     DFT warps (the 32 x 32 warp transform of s1k, one frame per warp) and then read once by the overlap-add role;
//   * overlap-add role (warps 0-7): thread t owns output samples i = t and t + 256: per i 2m running sums for the even
//     frames, 2m for the odd frames (64 registers at m = 4) and the 4m taps (32 registers); per pair of frames and i it
//     reads four values (conflict-free LDS.64, 16 B per output sample), issues 8m packed FFMA2 and emits two outputs.
//     The sums rotate through their registers by unrolling one period of 2m pairs; a slab starts one period early on
//     zeroed sums and drops that period's outputs (the warm-up frames come from the call or from the kept prefix).
// The order of the additions differs from the two-bank form of the other synthesis kernels (old to new across both
// banks instead of bank by bank), within the same f32 rounding bound.
namespace s1ks {
constexpr int kM = 1024, kM2 = 512;
constexpr int kFB = 4;                                   // frames per batch
constexpr int kFrameBytes = kM * 8;
constexpr int kStageBytes = kFB * kFrameBytes;           // 32 KB
constexpr int kStages = 6;
constexpr int kOffTw = kStages * kStageBytes;            // W1024^{lane k1}, the s1k table: 8 KB
constexpr int kOffBar = kOffTw + 8192;
constexpr int kBarInFull = 0;                            // [6] TMA transaction barriers
constexpr int kBarUFull = 6;                             // [6] the stage's four DFT warps have transformed it
constexpr int kBarFree = 12;                             // [6] the eight overlap-add warps have read it
constexpr int kOffSrc = kOffBar + 18 * 8;                // address of the slab's local frame 0 in x (refills never reach the prefix)
constexpr int kSmem = kOffSrc + 8;
constexpr int kThreads = 512;
constexpr int kMaxM = 4;

struct Params {
    const float2* prefix;     // the 32 input frames preceding x[0]
    const float2* x;          // input frames of the call, [frame][1024]
    float2* y;                // output sample 0 of the call
    long long f0;             // first frame handled (even global parity)
    long long n_batches;      // batches of 4 frames
    const float* taps;        // [1024][4m]  h[(j & 511) + l * 512] / 2
    const float2* twid;       // [1024] e^{+j 2 pi k / 1024}
    float2* hist_new;         // if non-null the kernel also writes the object's next state there: the last 32 frames of
    long long n_new;          //   (prefix ++ x[0 .. n_new))
};

template <int kL>                                        // kL = 2m: pairs of frames a sum collects
__device__ __forceinline__ void ola_role(const Params& p, uint32_t smem, long long b0, long long b1)
{
    constexpr int kWarm = kL / 2;                        // batches per period of kL pairs = warm-up batches
    const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
    const uint32_t bar = smem + kOffBar;

    const int nb = (int)(b1 - b0) + kWarm;               // local batches, warm-up period first
    const long long fr_base = p.f0 + (b0 - kWarm) * kFB; // call-relative frame of local batch 0 (>= -16)
    auto issue_load = [&](int lb, int stg) {
        const uint32_t fb = bar + 8 * (kBarInFull + stg);
        const uint32_t dst = smem + stg * kStageBytes;
        const long long fr = fr_base + (long long)lb * kFB;
        mbar_expect_tx(fb, kStageBytes);
        if (fr >= 0) tma_load_1d(dst, p.x + fr * kM, kStageBytes, fb);
        else if (fr + kFB <= 0) tma_load_1d(dst, p.prefix + (32 + fr) * kM, kStageBytes, fb);
        else
            for (int k = 0; k < kFB; k++)
                tma_load_1d(dst + k * kFrameBytes, fr + k >= 0 ? p.x + (fr + k) * kM : p.prefix + (32 + fr + k) * kM, kFrameBytes, fb);
    };
    if (t == 0) {                                        // first of all: get the copies going
        pdl_wait();                                      // x and the prefix may come from the previous kernel
        for (int lb = 0; lb < kStages && lb < nb; lb++) issue_load(lb, lb);
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(smem + kOffSrc), "l"(p.x + fr_base * kM) : "memory");
    }

    float A[2][kL], B[2][kL];                            // taps of even / odd lag of outputs t and t + 256
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int j = 0; j < kL; j++) {
            A[s][j] = __ldg(&p.taps[(t + 256 * s) * (2 * kL) + 2 * j]);
            B[s][j] = __ldg(&p.taps[(t + 256 * s) * (2 * kL) + 2 * j + 1]);
        }
    float2 E[2][kL], O[2][kL];                           // running sums of the even- and odd-frame outputs
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int j = 0; j < kL; j++) E[s][j] = O[s][j] = make_float2(0.f, 0.f);
    pdl_wait();                                          // nothing is written before the previous kernel has completed

    float2* yo = p.y + fr_base * kM2 + t;                // output of the current batch's first frame
    int st = 0;
    uint32_t ph = 0;
    for (int lb0 = 0; lb0 < nb; lb0 += kWarm) {
        const bool emit = lb0 > 0;                       // the first period only warms the sums up
#pragma unroll
        for (int bb = 0; bb < kWarm; bb++) {
            const int lb = lb0 + bb;
            if (lb >= nb) break;
            mbar_wait(bar + 8 * (kBarUFull + st), ph);
            const uint32_t fr = smem + st * kStageBytes + t * 8;
#pragma unroll
            for (int pq = 0; pq < 2; pq++) {
                const int pp = 2 * bb + pq;              // pair within the period: every register index below is a constant
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    const uint32_t a = fr + 2 * pq * kFrameBytes + s * 2048;
                    // outputs of the even frames g, g + 2, ...: lag 2j from u_g, lag 2j + 1 from u_{g+1} (to g + 2 + 2j)
                    const float2 ue = lds64(a);                                  // column i of the even frame
#pragma unroll
                    for (int j = 0; j < kL; j++) E[s][(pp + j) % kL] = fma2(ue, f2(A[s][j]), E[s][(pp + j) % kL]);
                    if (emit) __stcs(yo + 2 * pq * kM2 + 256 * s, E[s][pp % kL]);
                    const float2 uo = lds64(a + kFrameBytes);                    //            of the odd frame
#pragma unroll
                    for (int j = 0; j < kL - 1; j++) E[s][(pp + 1 + j) % kL] = fma2(uo, f2(B[s][j]), E[s][(pp + 1 + j) % kL]);
                    E[s][pp % kL] = mul2(uo, f2(B[s][kL - 1]));
                    // outputs of the odd frames g + 1, g + 3, ...: lag 2j + 1 from u_g, lag 2j from u_{g+1}
                    const float2 ve = lds64(a + 4096);                           // column i + 512 of the even frame
                    O[s][(pp + kL - 1) % kL] = mul2(ve, f2(B[s][kL - 1]));
#pragma unroll
                    for (int j = 0; j < kL - 1; j++) O[s][(pp + j) % kL] = fma2(ve, f2(B[s][j]), O[s][(pp + j) % kL]);
                    const float2 vo = lds64(a + kFrameBytes + 4096);             //                 of the odd frame
#pragma unroll
                    for (int j = 0; j < kL; j++) O[s][(pp + j) % kL] = fma2(vo, f2(A[s][j]), O[s][(pp + j) % kL]);
                    if (emit) __stcs(yo + (2 * pq + 1) * kM2 + 256 * s, O[s][pp % kL]);
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar + 8 * (kBarFree + st));
                if (wrp == (lb & 7) && lb + kStages < nb) {          // the overlap-add warps take turns refilling the stage
                    mbar_wait(bar + 8 * (kBarFree + st), ph);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the DFT warps wrote it with plain stores
                    // batch lb + 6 lies inside x (kStages > kWarm); its address comes from shared memory, not from
                    // registers every thread would have to keep alive across the loop
                    unsigned long long src;
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(src) : "r"(smem + kOffSrc) : "memory");
                    mbar_expect_tx(bar + 8 * (kBarInFull + st), kStageBytes);
                    tma_load_1d(smem + st * kStageBytes, reinterpret_cast<const char*>(src) + (long long)(lb + kStages) * kStageBytes,
                                kStageBytes, bar + 8 * (kBarInFull + st));
                }
            }
            yo += kFB * kM2;
            if (++st == kStages) { st = 0; ph ^= 1; }
        }
    }
}

__device__ __forceinline__ void dft_role(const Params& p, uint32_t smem, int nb)
{
    const int dt = threadIdx.x - 256, lane = dt & 31, dw = dt >> 5;
    const int grp = dw >> 2, fi = dw & 3;                // batches of this warp's parity, frame fi of each
    const uint32_t bar = smem + kOffBar;
    const uint32_t twt = smem + kOffTw + lane * 16;
    // state hand-off folded into this launch: every CTA copies a slice of the next state (the tail of the input stream)
    if (p.hist_new != nullptr) {
        constexpr long long kH = 32 * kM;
        pdl_wait();                                      // nothing is written before the previous kernel has completed
        const long long per = (kH / 2 + gridDim.x - 1) / gridDim.x;
        const long long i1 = min(kH / 2, per * (long long)(blockIdx.x + 1));
        for (long long i = per * blockIdx.x + dt; i < i1; i += 256) {            // 16 bytes per thread and turn
            const long long ts = p.n_new - kH + 2 * i;
            reinterpret_cast<float4*>(p.hist_new)[i] = __ldg(reinterpret_cast<const float4*>(ts >= 0 ? p.x + ts : p.prefix + (kH + ts)));
        }
    }
    int st = grp;
    uint32_t ph = 0;
    for (int lb = grp; lb < nb; lb += 2) {
        mbar_wait(bar + 8 * (kBarInFull + st), ph);
        const uint32_t frame = smem + st * kStageBytes + fi * kFrameBytes;
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) v[n1] = lds64(frame + (32 * n1 + lane) * 8);
        xdft32c(v);
        __syncwarp();                                    // every lane has read the frame: exchange in place
#pragma unroll
        for (int a = 0; a < 16; a++) {
            const float4 w = lds128(twt + a * 512);
            float2 z0 = v[dr32(2 * a)], z1 = v[dr32(2 * a + 1)];
            if (a > 0) z0 = xmul(z0, w.x, w.y);
            z1 = xmul(z1, w.z, w.w);
            sts64(frame + (((lane << 5) | ((2 * a) ^ lane)) << 3), z0);
            sts64(frame + (((lane << 5) | ((2 * a + 1) ^ lane)) << 3), z1);
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = lds64(frame + (((n2 << 5) | (lane ^ n2)) << 3));
        __syncwarp();
        xdft32c(v);
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) sts64(frame + (32 * k2 + lane) * 8, v[dr32(k2)]);      // U in natural order
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + 8 * (kBarUFull + st));
        st += 2;
        if (st >= kStages) { st -= kStages; ph ^= 1; }
    }
}

template <int kL>
__global__ void __launch_bounds__(kThreads, 1) k_m1024_synth_fused(const Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const long long b0 = (p.n_batches * blockIdx.x) / gridDim.x, b1 = (p.n_batches * (blockIdx.x + 1)) / gridDim.x;
    if (threadIdx.x == 0) {
        const uint32_t bar = smem + kOffBar;
        for (int i = 0; i < kStages; i++) {
            mbar_init(bar + 8 * (kBarInFull + i), 1);
            mbar_init(bar + 8 * (kBarUFull + i), 4);
            mbar_init(bar + 8 * (kBarFree + i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // (filling this table inside the DFT role, off the FIR role's start-up path, made the kernels 3-4 % slower: same-box A/B)
    for (int i = threadIdx.x; i < 1024; i += kThreads) {              // entry i: k1 = 2 (i >> 6) + (i & 1), lane = (i >> 1) & 31
        const int k1 = 2 * (i >> 6) + (i & 1), ln = (i >> 1) & 31;
        sts64(smem + kOffTw + i * 8, __ldg(&p.twid[ln * k1]));
    }
    __syncthreads();
    pdl_launch_dependents();
    if (b0 >= b1) return;                                // never taken: the grid has at most one CTA per batch
    if (threadIdx.x < 256) ola_role<kL>(p, smem, b0, b1);
    else dft_role(p, smem, (int)(b1 - b0) + kL / 2);
}

template <int kL>
int32_t launch(const Firpfbch2FastPlan& plan, const Params& p, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>(plan.n_sm, p.n_batches));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = plan.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YG_CUDA(cudaLaunchKernelEx(&cfg, k_m1024_synth_fused<kL>, p));
    count_launch();
    return YG_OK;
}

inline int32_t prepare(uint32_t m)
{
    switch (m) {
        case 1: YG_CUDA(cudaFuncSetAttribute(k_m1024_synth_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        case 2: YG_CUDA(cudaFuncSetAttribute(k_m1024_synth_fused<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        case 3: YG_CUDA(cudaFuncSetAttribute(k_m1024_synth_fused<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
        default: YG_CUDA(cudaFuncSetAttribute(k_m1024_synth_fused<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); break;
    }
    return YG_OK;
}
}  // namespace s1ks

// ------------------------------------------------------------------ single-SM fused synthesis kernel: M = 512, m <= 7
// s1ks re-cut for half the frame length: a stage of the frame ring holds EIGHT input frames (still 32 KB), each of its four
// DFT warps transforms a PAIR of frames in place (s5k::pair_dft512), and an overlap-add thread owns ONE output sample
// i = t with its 2m + 2m running sums (56 registers at m = 7) and 4m taps.  The sums rotate with a period of 2m pairs
// while a batch is four pairs, so one period of lcm(2m, 4) pairs is unrolled; the warm-up (ceil(2m / 4) batches before
// the slab, at most the 32 frames of kept prefix) is a run-time flag per batch.
namespace s5ks {
constexpr int kM = 512, kM2 = 256;
constexpr int kFB = 8;                                   // frames per batch
constexpr int kFrameBytes = kM * 8;
constexpr int kStageBytes = kFB * kFrameBytes;           // 32 KB
constexpr int kStages = 6;
constexpr int kOffTw = kStages * kStageBytes;            // the s5k twiddle table: 4 KB
constexpr int kOffBar = kOffTw + 4096;
constexpr int kBarInFull = 0;                            // [6] TMA transaction barriers
constexpr int kBarUFull = 6;                             // [6] the stage's four DFT warps have transformed it
constexpr int kBarFree = 12;                             // [6] the eight overlap-add warps have read it
constexpr int kOffSrc = kOffBar + 18 * 8;                // address of the slab's local frame 0 in x (refills never reach the prefix)
constexpr int kSmem = kOffSrc + 8;
constexpr int kThreads = 512;
constexpr int kMaxM = 7;
using Params = s1ks::Params;                             // taps: [512][4m]  h[(j & 255) + l * 256] / 2

template <int kL>                                        // kL = 2m: pairs of frames a sum collects
__device__ __forceinline__ void ola_role(const Params& p, uint32_t smem, long long b0, long long b1)
{
    constexpr int kWarm = (kL + 3) / 4;                  // warm-up batches (>= kL pairs)
    constexpr int kPer = (kL % 4 == 0 ? kL : kL % 4 == 2 ? 2 * kL : 4 * kL) / 4;      // batches per rotation period
    static_assert(kStages > kWarm, "refills must lie inside x");
    const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
    const uint32_t bar = smem + kOffBar;

    const int nb = (int)(b1 - b0) + kWarm;               // local batches, warm-up first
    const long long fr_base = p.f0 + (b0 - kWarm) * kFB; // call-relative frame of local batch 0 (>= -32)
    auto issue_load = [&](int lb, int stg) {
        const uint32_t fb = bar + 8 * (kBarInFull + stg);
        const uint32_t dst = smem + stg * kStageBytes;
        const long long fr = fr_base + (long long)lb * kFB;
        mbar_expect_tx(fb, kStageBytes);
        if (fr >= 0) tma_load_1d(dst, p.x + fr * kM, kStageBytes, fb);
        else if (fr + kFB <= 0) tma_load_1d(dst, p.prefix + (32 + fr) * kM, kStageBytes, fb);
        else
            for (int k = 0; k < kFB; k++)
                tma_load_1d(dst + k * kFrameBytes, fr + k >= 0 ? p.x + (fr + k) * kM : p.prefix + (32 + fr + k) * kM, kFrameBytes, fb);
    };
    if (t == 0) {                                        // first of all: get the copies going
        pdl_wait();                                      // x and the prefix may come from the previous kernel
        for (int lb = 0; lb < kStages && lb < nb; lb++) issue_load(lb, lb);
        asm volatile("st.shared.u64 [%0], %1;" ::"r"(smem + kOffSrc), "l"(p.x + fr_base * kM) : "memory");
    }

    float A[kL], B[kL];                                  // taps of even / odd lag of output t
#pragma unroll
    for (int j = 0; j < kL; j++) {
        A[j] = __ldg(&p.taps[t * (2 * kL) + 2 * j]);
        B[j] = __ldg(&p.taps[t * (2 * kL) + 2 * j + 1]);
    }
    float2 E[kL], O[kL];                                 // running sums of the even- and odd-frame outputs
#pragma unroll
    for (int j = 0; j < kL; j++) E[j] = O[j] = make_float2(0.f, 0.f);
    pdl_wait();                                          // nothing is written before the previous kernel has completed

    float2* yo = p.y + fr_base * kM2 + t;                // output of the current batch's first frame
    int st = 0;
    uint32_t ph = 0;
    for (int lb0 = 0; lb0 < nb; lb0 += kPer) {
#pragma unroll
        for (int bb = 0; bb < kPer; bb++) {
            const int lb = lb0 + bb;
            if (lb >= nb) break;
            const bool emit = lb >= kWarm;               // the first batches only warm the sums up
            mbar_wait(bar + 8 * (kBarUFull + st), ph);
            const uint32_t fr = smem + st * kStageBytes + t * 8;
#pragma unroll
            for (int pq = 0; pq < 4; pq++) {
                const int pp = (4 * bb + pq) % kL;       // pair within the rotation: every register index below is a constant
                const uint32_t a = fr + 2 * pq * kFrameBytes;
                // outputs of the even frames g, g + 2, ...: lag 2j from u_g, lag 2j + 1 from u_{g+1} (to g + 2 + 2j)
                const float2 ue = lds64(a);                                      // column i of the even frame
#pragma unroll
                for (int j = 0; j < kL; j++) E[(pp + j) % kL] = fma2(ue, f2(A[j]), E[(pp + j) % kL]);
                if (emit) __stcs(yo + 2 * pq * kM2, E[pp % kL]);
                const float2 uo = lds64(a + kFrameBytes);                        //            of the odd frame
#pragma unroll
                for (int j = 0; j < kL - 1; j++) E[(pp + 1 + j) % kL] = fma2(uo, f2(B[j]), E[(pp + 1 + j) % kL]);
                E[pp % kL] = mul2(uo, f2(B[kL - 1]));
                // outputs of the odd frames g + 1, g + 3, ...: lag 2j + 1 from u_g, lag 2j from u_{g+1}
                const float2 ve = lds64(a + kM2 * 8);                            // column i + 256 of the even frame
                O[(pp + kL - 1) % kL] = mul2(ve, f2(B[kL - 1]));
#pragma unroll
                for (int j = 0; j < kL - 1; j++) O[(pp + j) % kL] = fma2(ve, f2(B[j]), O[(pp + j) % kL]);
                const float2 vo = lds64(a + kFrameBytes + kM2 * 8);              //                 of the odd frame
#pragma unroll
                for (int j = 0; j < kL; j++) O[(pp + j) % kL] = fma2(vo, f2(A[j]), O[(pp + j) % kL]);
                if (emit) __stcs(yo + (2 * pq + 1) * kM2, O[pp % kL]);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar + 8 * (kBarFree + st));
                if (wrp == (lb & 7) && lb + kStages < nb) {          // the overlap-add warps take turns refilling the stage
                    mbar_wait(bar + 8 * (kBarFree + st), ph);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the DFT warps wrote it with plain stores
                    unsigned long long src;
                    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(src) : "r"(smem + kOffSrc) : "memory");
                    mbar_expect_tx(bar + 8 * (kBarInFull + st), kStageBytes);
                    tma_load_1d(smem + st * kStageBytes, reinterpret_cast<const char*>(src) + (long long)(lb + kStages) * kStageBytes,
                                kStageBytes, bar + 8 * (kBarInFull + st));
                }
            }
            yo += kFB * kM2;
            if (++st == kStages) { st = 0; ph ^= 1; }
        }
    }
}

__device__ __forceinline__ void dft_role(const Params& p, uint32_t smem, int nb)
{
    const int dt = threadIdx.x - 256, lane = dt & 31, dw = dt >> 5;
    const int grp = dw >> 2, fi = dw & 3;                // batches of this warp's parity, frame pair fi of each
    const uint32_t bar = smem + kOffBar;
    const uint32_t twt = smem + kOffTw + lane * 16;
    if (p.hist_new != nullptr) {                         // every CTA copies a slice of the next state (the tail of the input stream)
        constexpr long long kH = 32 * kM;
        pdl_wait();                                      // nothing is written before the previous kernel has completed
        const long long per = (kH / 2 + gridDim.x - 1) / gridDim.x;
        const long long i1 = min(kH / 2, per * (long long)(blockIdx.x + 1));
        for (long long i = per * blockIdx.x + dt; i < i1; i += 256) {            // 16 bytes per thread and turn
            const long long ts = p.n_new - kH + 2 * i;
            reinterpret_cast<float4*>(p.hist_new)[i] = __ldg(reinterpret_cast<const float4*>(ts >= 0 ? p.x + ts : p.prefix + (kH + ts)));
        }
    }
    int st = grp;
    uint32_t ph = 0;
    for (int lb = grp; lb < nb; lb += 2) {
        mbar_wait(bar + 8 * (kBarInFull + st), ph);
        const uint32_t tile = smem + st * kStageBytes + fi * (2 * kFrameBytes);
        float2 v[32];
        s5k::pair_dft512(v, tile, twt, lane, [] {});
        const uint32_t uo = tile + (lane >> 4) * kFrameBytes + (lane & 15) * 8;
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) sts64(uo + 16 * k2 * 8, v[dr32(k2)]);    // U in natural order
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + 8 * (kBarUFull + st));
        st += 2;
        if (st >= kStages) { st -= kStages; ph ^= 1; }
    }
}

template <int kL>
__global__ void __launch_bounds__(kThreads, 1) k_m512_synth_fused(const Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    const long long b0 = (p.n_batches * blockIdx.x) / gridDim.x, b1 = (p.n_batches * (blockIdx.x + 1)) / gridDim.x;
    if (threadIdx.x == 0) {
        const uint32_t bar = smem + kOffBar;
        for (int i = 0; i < kStages; i++) {
            mbar_init(bar + 8 * (kBarInFull + i), 1);
            mbar_init(bar + 8 * (kBarUFull + i), 4);
            mbar_init(bar + 8 * (kBarFree + i), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    s5k::fill_twiddles(smem + kOffTw, p.twid);
    __syncthreads();
    pdl_launch_dependents();
    if (b0 >= b1) return;                                // never taken: the grid has at most one CTA per batch
    if (threadIdx.x < 256) s5ks::ola_role<kL>(p, smem, b0, b1);           // (qualified: Params is s1ks's, so ADL would find both)
    else s5ks::dft_role(p, smem, (int)(b1 - b0) + (kL + 3) / 4);
}

template <int kL>
int32_t launch(const Firpfbch2FastPlan& plan, const Params& p, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)std::min<long long>(plan.n_sm, p.n_batches));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = plan.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    YG_CUDA(cudaLaunchKernelEx(&cfg, k_m512_synth_fused<kL>, p));
    count_launch();
    return YG_OK;
}

template <int kL>
int32_t prepare_one() { YG_CUDA(cudaFuncSetAttribute(k_m512_synth_fused<kL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); return YG_OK; }

// per filter length: f(std::integral_constant<int, 2 m>) for m = 1 .. 7
template <typename F>
int32_t for_m(uint32_t m, F f)
{
    switch (m) {
        case 1: return f(std::integral_constant<int, 2>{});
        case 2: return f(std::integral_constant<int, 4>{});
        case 3: return f(std::integral_constant<int, 6>{});
        case 4: return f(std::integral_constant<int, 8>{});
        case 5: return f(std::integral_constant<int, 10>{});
        case 6: return f(std::integral_constant<int, 12>{});
        case 7: return f(std::integral_constant<int, 14>{});
        default: return fail(YG_EINTERNAL, "single-SM M = 512 synthesis kernel not instantiated for m = %u", m);
    }
}
inline int32_t prepare(uint32_t m) { return for_m(m, [](auto tag) { return prepare_one<decltype(tag)::value>(); }); }
}  // namespace s5ks

template <int kTaps>
int32_t launch_wola(const WolaParams& p, cudaStream_t st)
{
    dim3 grid((unsigned)p.slabs, (unsigned)(p.M / kFirThreads));
    k_large_wola<kTaps><<<grid, kFirThreads, 0, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

template <int kTaps>
int32_t launch_fir(const LargeParams& p, cudaStream_t st)
{
    dim3 grid((unsigned)p.slabs, (unsigned)(p.M / kFirThreads));
    k_large_fir<kTaps><<<grid, kFirThreads, 0, st>>>(p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

}  // namespace

namespace {

bool large_geometry_ok(uint32_t M) { return M == 512 || M == 1024 || M == 2048 || M == 4096; }

int32_t plan_common(Firpfbch2FastPlan& plan, uint32_t M)
{
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    YG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return YG_OK;
    plan.n_sm = prop.multiProcessorCount;
    std::vector<float2> tw(M);
    for (uint32_t k = 0; k < M; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    YG_CUDA(cudaMalloc(&plan.d_twid, tw.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_twid, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    if (M == 1024) YG_CUDA(cudaFuncSetAttribute(k_large_fft<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmem));
    else if (M == 2048) YG_CUDA(cudaFuncSetAttribute(k_large_fft<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmem));
    else if (M == 512) YG_CUDA(cudaFuncSetAttribute(k_large_fft512, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmem));
    else if (M == 4096) YG_CUDA(cudaFuncSetAttribute(k_large_fft<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFftSmem));
    plan.min_frames = 64;
    plan.supported = true;
    return YG_OK;
}

// the fused analysis kernel's per-group V ring and counters
// resident CTAs per SM of a fused kernel instance (0: it cannot launch with this much shared memory)
template <typename K>
int resident_ctas(K kernel, int smem)
{
    int n = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kFirThreads + 256, (size_t)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int32_t plan_fused(Firpfbch2FastPlan& plan, bool synthesis)
{
    const int G = (int)plan.M / kFirThreads;
    plan.n_groups = plan.n_sm / G;
    int dev = 0, coop = 0;
    YG_CUDA(cudaGetDevice(&dev));
    YG_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    int fit = 0;                          // a cooperative grid of one CTA per SM must be resident all at once
    if (synthesis) {
        switch (plan.m) {
            case 1: fit = resident_ctas(k_large_synth_fused<4>, synth_fused_smem(4)); break;
            case 2: fit = resident_ctas(k_large_synth_fused<8>, synth_fused_smem(8)); break;
            case 3: fit = resident_ctas(k_large_synth_fused<12>, synth_fused_smem(12)); break;
            case 4: fit = resident_ctas(k_large_synth_fused<16>, synth_fused_smem(16)); break;
            case 5: fit = resident_ctas(k_large_synth_fused<20>, synth_fused_smem(20)); break;
            case 6: fit = resident_ctas(k_large_synth_fused<24>, synth_fused_smem(24)); break;
            case 7: fit = resident_ctas(k_large_synth_fused<28>, synth_fused_smem(28)); break;
            default: break;
        }
    } else {
        switch (plan.m) {
            case 1: fit = resident_ctas(k_large_fused<3>, kFusedSmem); break;
            case 2: fit = resident_ctas(k_large_fused<5>, kFusedSmem); break;
            case 3: fit = resident_ctas(k_large_fused<7>, kFusedSmem); break;
            case 4: fit = resident_ctas(k_large_fused<9>, kFusedSmem); break;
            case 5: fit = resident_ctas(k_large_fused<11>, kFusedSmem); break;
            case 6: fit = resident_ctas(k_large_fused<13>, kFusedSmem); break;
            case 7: fit = resident_ctas(k_large_fused<15>, kFusedSmem); break;
            case 8: fit = resident_ctas(k_large_fused<17>, kFusedSmem); break;
            default: break;
        }
    }
    if (!coop || fit < 1 || plan.n_groups < 1) { plan.n_groups = 0; return YG_OK; }
    YG_CUDA(cudaMalloc(&plan.d_scratch, (size_t)plan.n_groups * kSlots * 32 * plan.M * sizeof(float2)));
    YG_CUDA(cudaMalloc(&plan.d_flags, (size_t)plan.n_groups * kFlagStride * sizeof(unsigned)));
    return YG_OK;
}

// frames of intermediate kept per chunk: 96 MB (swept 16..96 MB at M = 1024: profiles/r01_large_chunk_sweep.log)
long long chunk_frames(uint32_t M) { return std::max<long long>(128, (((long long)96 << 20) / ((long long)M * 8)) & ~63LL); }

}  // namespace

int32_t firpfbch2_large_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if (!large_geometry_ok(M) || m < 1 || m > 8) return YG_OK;
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
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float2)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float2), cudaMemcpyHostToDevice));
    YG_TRY(plan_common(plan, M));
    if (plan.supported && ((M == 1024 && 2 * m + 1 <= (uint32_t)s1k::kMaxTaps) || (M == 512 && m <= (uint32_t)s5k::kMaxM))) {
        const char* e = getenv("YG_LARGE_SINGLE_SM");     // debugging knob: 0 keeps the group kernel
        plan.single_sm = !(e && e[0] == '0');
        if (plan.single_sm) YG_TRY(M == 1024 ? s1k::prepare(m) : s5k::prepare(m));
        const char* d = getenv("YG_PDL");
        plan.pdl = !(d && d[0] == '0');
    }
    if (plan.supported && !plan.single_sm) YG_TRY(plan_fused(plan, false));      // the group kernel's L2 ring: only where it runs
    return YG_OK;
}

int32_t firpfbch2_large_launch(const Firpfbch2FastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                               size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new, long long n_new, bool* hist_done)
{
    if (hist_done) *hist_done = false;
    if (!plan.supported) return fail(YG_EINTERNAL, "large-M path not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames & 1) return fail(YG_EINTERNAL, "large-M path needs an even number of frames");
    const int M = (int)plan.M;
    const long long n_pairs = (long long)(n_frames / 2);
    long long fused_pairs = 0;
    if (plan.single_sm && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {         // M = 1024, m <= 4: one CTA per SM, no exchange
        const int bp = M == 1024 ? s1k::kBP : s5k::kBP;
        fused_pairs = (n_pairs / bp) * bp;
        if (fused_pairs > 0) {
            LargeParams p;
            p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
            p.f0 = (long long)f0;
            p.pair_begin = 0;
            p.pair_end = fused_pairs;
            p.slabs = 0;
            p.M = M;
            p.taps = reinterpret_cast<const float2*>(plan.d_taps);
            p.twid = reinterpret_cast<const float2*>(plan.d_twid);
            if (hist_done && hist_new) { p.hist_new = hist_new; p.n_new = n_new; *hist_done = true; }
            if (M == 512) YG_TRY(s5k::for_m(plan.m, [&](auto tag) { return s5k::launch<decltype(tag)::value>(plan, p, st); }));
            else switch (plan.m) {
                case 1: YG_TRY(s1k::launch<3>(plan, p, st)); break;
                case 2: YG_TRY(s1k::launch<5>(plan, p, st)); break;
                case 3: YG_TRY(s1k::launch<7>(plan, p, st)); break;
                case 4: YG_TRY(s1k::launch<9>(plan, p, st)); break;
                default: return fail(YG_EINTERNAL, "single-SM large-M kernel not instantiated for m = %u", plan.m);
            }
        }
    } else if (plan.n_groups > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {      // the fused kernel stages 16-byte chunks
        // whole 16-pair batches go through the fused kernel; what is left (< 32 frames) takes the two-stage path below
        const long long n_batches = n_pairs / kPairsPerBatch;
        fused_pairs = n_batches * kPairsPerBatch;
        if (n_batches > 0) {
            FusedParams fp;
            LargeParams& p = fp.base;
            p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
            p.f0 = (long long)f0;
            p.pair_begin = 0;
            p.pair_end = fused_pairs;
            p.slabs = 0;
            p.M = M;
            p.taps = reinterpret_cast<const float2*>(plan.d_taps);
            p.twid = reinterpret_cast<const float2*>(plan.d_twid);
            fp.scratch = reinterpret_cast<float2*>(plan.d_scratch);
            fp.flags = reinterpret_cast<unsigned*>(plan.d_flags);
            fp.n_groups = (int)std::min<long long>(plan.n_groups, n_batches);
            YG_CUDA(cudaMemsetAsync(fp.flags, 0, (size_t)plan.n_groups * kFlagStride * sizeof(unsigned), st));
            switch (plan.m) {
                case 1: YG_TRY(launch_fused<3>(fp, st)); break;
                case 2: YG_TRY(launch_fused<5>(fp, st)); break;
                case 3: YG_TRY(launch_fused<7>(fp, st)); break;
                case 4: YG_TRY(launch_fused<9>(fp, st)); break;
                case 5: YG_TRY(launch_fused<11>(fp, st)); break;
                case 6: YG_TRY(launch_fused<13>(fp, st)); break;
                case 7: YG_TRY(launch_fused<15>(fp, st)); break;
                case 8: YG_TRY(launch_fused<17>(fp, st)); break;
                default: return fail(YG_EINTERNAL, "large-M path not instantiated for m = %u", plan.m);
            }
        }
    }
    // chunk so that a chunk's output (8 M bytes per frame) stays resident in L2 between the two stages
    const long long chunk_pairs = chunk_frames(plan.M) / 2;
    for (long long q = fused_pairs; q < n_pairs; q += chunk_pairs) {
        LargeParams p;
        p.hist = hist; p.Hlen = Hlen; p.x = x; p.y = y;
        p.f0 = (long long)f0;
        p.pair_begin = q;
        p.pair_end = std::min(n_pairs, q + chunk_pairs);
        p.M = M;
        const long long batches = (p.pair_end - p.pair_begin + kPairsPerBatch - 1) / kPairsPerBatch;
        p.slabs = (int)std::max<long long>(1, std::min<long long>(batches, (long long)plan.n_sm * 4 / (M / kFirThreads)));
        p.taps = reinterpret_cast<const float2*>(plan.d_taps);
        p.twid = reinterpret_cast<const float2*>(plan.d_twid);
        switch (plan.m) {
            case 1: YG_TRY(launch_fir<3>(p, st)); break;
            case 2: YG_TRY(launch_fir<5>(p, st)); break;
            case 3: YG_TRY(launch_fir<7>(p, st)); break;
            case 4: YG_TRY(launch_fir<9>(p, st)); break;
            case 5: YG_TRY(launch_fir<11>(p, st)); break;
            case 6: YG_TRY(launch_fir<13>(p, st)); break;
            case 7: YG_TRY(launch_fir<15>(p, st)); break;
            case 8: YG_TRY(launch_fir<17>(p, st)); break;
            default: return fail(YG_EINTERNAL, "large-M path not instantiated for m = %u", plan.m);
        }
        const long long nf = 2 * (p.pair_end - p.pair_begin);
        float2* yc = y + ((long long)f0 + 2 * p.pair_begin) * M;
        YG_TRY(launch_fft(plan, nullptr, yc, 0, yc, nf, 1, st));
    }
    return YG_OK;
}

int32_t firpfbch2_large_synth_plan(Firpfbch2FastPlan& plan, uint32_t M, uint32_t m, const float* h)
{
    plan.supported = false;
    plan.M = M;
    plan.m = m;
    if (!large_geometry_ok(M) || m < 1 || m > 7) return YG_OK;
    const int iM = (int)M, iM2 = iM / 2;
    const int kTaps = 4 * (int)m;
    std::vector<float> taps((size_t)iM * kTaps);
    for (int j = 0; j < iM; j++)
        for (int l = 0; l < kTaps; l++) taps[(size_t)j * kTaps + l] = 0.5f * h[(j & (iM2 - 1)) + l * iM2];
    YG_CUDA(cudaMalloc(&plan.d_taps, taps.size() * sizeof(float)));
    YG_CUDA(yg::memcpy_sync(plan.d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_TRY(plan_common(plan, M));
    if (plan.supported && ((M == 1024 && m <= (uint32_t)s1ks::kMaxM) || (M == 512 && m <= (uint32_t)s5ks::kMaxM))) {
        const char* e = getenv("YG_LARGE_SINGLE_SM");     // debugging knob: 0 keeps the group kernel
        plan.single_sm = !(e && e[0] == '0');
        if (plan.single_sm) YG_TRY(M == 1024 ? s1ks::prepare(m) : s5ks::prepare(m));
        const char* d = getenv("YG_PDL");
        plan.pdl = !(d && d[0] == '0');
    }
    if (plan.supported && !plan.single_sm) YG_TRY(plan_fused(plan, true));
    return YG_OK;
}

// `prefix` = the 32 input frames preceding x[0]; frames [f0, f0 + n_frames) of the call, f0 on even global
// parity, n_frames a multiple of 32; `scratch` holds firpfbch2_large_synth_scratch_frames(M) frames.
long long firpfbch2_large_synth_scratch_frames(uint32_t M) { return chunk_frames(M) + 32; }

namespace {
bool synth_fused_ok(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x)
{
    return (plan.n_groups > 0 || plan.single_sm) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(prefix)) & 15) == 0;
}
}  // namespace

bool firpfbch2_large_synth_needs_scratch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x)
{
    return !synth_fused_ok(plan, prefix, x);
}

int32_t firpfbch2_large_synth_launch(const Firpfbch2FastPlan& plan, const float2* prefix, const float2* x, float2* y,
                                     float2* scratch, size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new, long long n_new,
                                     bool* hist_done)
{
    if (hist_done) *hist_done = false;
    if (!plan.supported) return fail(YG_EINTERNAL, "large-M synthesis path not available for this geometry");
    if (n_frames == 0) return YG_OK;
    if (n_frames % 32) return fail(YG_EINTERNAL, "large-M synthesis path needs a multiple of 32 frames");
    const int M = (int)plan.M;
    if (plan.single_sm && synth_fused_ok(plan, prefix, x)) {       // M = 1024, m <= 4: one CTA per SM, no exchange
        s1ks::Params p;
        p.prefix = prefix; p.x = x; p.y = y;
        p.f0 = (long long)f0;
        p.n_batches = (long long)(n_frames / (M == 1024 ? s1ks::kFB : s5ks::kFB));
        p.taps = reinterpret_cast<const float*>(plan.d_taps);
        p.twid = reinterpret_cast<const float2*>(plan.d_twid);
        p.hist_new = nullptr; p.n_new = 0;
        if (hist_done && hist_new) { p.hist_new = hist_new; p.n_new = n_new; *hist_done = true; }     // x and the prefix are 16-byte aligned here
        if (M == 512) return s5ks::for_m(plan.m, [&](auto tag) { return s5ks::launch<decltype(tag)::value>(plan, p, st); });
        switch (plan.m) {
            case 1: return s1ks::launch<2>(plan, p, st);
            case 2: return s1ks::launch<4>(plan, p, st);
            case 3: return s1ks::launch<6>(plan, p, st);
            case 4: return s1ks::launch<8>(plan, p, st);
            default: return fail(YG_EINTERNAL, "single-SM large-M synthesis kernel not instantiated for m = %u", plan.m);
        }
    }
    if (synth_fused_ok(plan, prefix, x)) {                 // the fused kernel stages 16-byte chunks
        SynthFusedParams p;
        p.prefix = prefix; p.x = x; p.y = y;
        p.f0 = (long long)f0;
        p.n_batches = (long long)(n_frames / 32);
        p.M = M;
        p.taps = reinterpret_cast<const float*>(plan.d_taps);
        p.twid = reinterpret_cast<const float2*>(plan.d_twid);
        p.scratch = reinterpret_cast<float4*>(plan.d_scratch);
        p.flags = reinterpret_cast<unsigned*>(plan.d_flags);
        p.n_groups = (int)std::min<long long>(plan.n_groups, p.n_batches);
        YG_CUDA(cudaMemsetAsync(p.flags, 0, (size_t)plan.n_groups * kFlagStride * sizeof(unsigned), st));
        switch (plan.m) {
            case 1: return launch_synth_fused<4>(p, st);
            case 2: return launch_synth_fused<8>(p, st);
            case 3: return launch_synth_fused<12>(p, st);
            case 4: return launch_synth_fused<16>(p, st);
            case 5: return launch_synth_fused<20>(p, st);
            case 6: return launch_synth_fused<24>(p, st);
            case 7: return launch_synth_fused<28>(p, st);
            default: return fail(YG_EINTERNAL, "large-M synthesis path not instantiated for m = %u", plan.m);
        }
    }
    const long long chunk = chunk_frames(plan.M);                              // multiple of 32
    for (long long c0 = 0; c0 < (long long)n_frames; c0 += chunk) {
        const long long nf = std::min<long long>(chunk, (long long)n_frames - c0);
        // stage B': U[g] = IDFT_unnorm(X[f0 + c0 - 32 + g]), g = 0 .. nf + 31 (the 1/2 scale lives in the taps)
        YG_TRY(launch_fft(plan, prefix, x, (long long)f0 + c0 - 32, scratch, nf + 32, 0, st));
        WolaParams p;
        p.U = scratch;
        p.y = y + ((long long)f0 + c0) * (M / 2);
        p.n_frames = nf;
        p.M = M;
        p.slabs = (int)std::max<long long>(1, std::min<long long>(nf / 32, (long long)plan.n_sm * 4 / (M / kFirThreads)));
        p.taps = reinterpret_cast<const float*>(plan.d_taps);
        switch (plan.m) {
            case 1: YG_TRY(launch_wola<4>(p, st)); break;
            case 2: YG_TRY(launch_wola<8>(p, st)); break;
            case 3: YG_TRY(launch_wola<12>(p, st)); break;
            case 4: YG_TRY(launch_wola<16>(p, st)); break;
            case 5: YG_TRY(launch_wola<20>(p, st)); break;
            case 6: YG_TRY(launch_wola<24>(p, st)); break;
            case 7: YG_TRY(launch_wola<28>(p, st)); break;
            default: return fail(YG_EINTERNAL, "large-M synthesis path not instantiated for m = %u", plan.m);
        }
    }
    return YG_OK;
}

}  // namespace yg
