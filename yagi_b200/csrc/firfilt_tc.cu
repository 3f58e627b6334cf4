// firfilt_tc.cu -- batched firfilt_crcf (<= 161 taps) on the 5th-generation tensor cores: 3xTF32 banded-Toeplitz GEMM.
//
//   y[s][n] = scale * sum_k h[k] x[s][n-k]                 (src/filter/fir/firfilt.rs:241-245, :267-278)
//
// Why: at 63 taps the CUDA-core kernel (firfilt_fast.cu) is FP32-pipe bound -- 126 lane-FMAs per 16 bytes moved, a
// ceiling of ~0.70 of the HBM roofline -- so the only way up is to take the multiply-adds off the FMA pipe.
//
// Formulation (per CTA, M = 128 rows, N = 64 outputs, K = 128 inputs per tile; the numbers below are those of the <= 65-tap
// instance Cfg<64, 1> -- Cfg<32, 3> and Cfg<32, 5> use 32-sample blocks and 3 / 5 blocks of history for up to 97 / 161 taps):
//   rows    : 8 streams x 8 adjacent time segments of each (independent rows, each primed with the block before it)
//             x {re, im}                           -> the A operand, which lives in TENSOR MEMORY (lane = row);
//   columns : 64 consecutive output times of a tile, stored reversed (column n' <-> time 63 - n');
//   K       : the 128 input times [64 tau - 64, 64 tau + 64) = the previous and the current 64-sample block;
//   B[k][n']: h[127 - (n' + k)] (zero outside 0 .. h_len-1): a Toeplitz band.  Because it depends on n' + k only,
//             the 8 x 16-byte core matrices of the K-major no-swizzle shared-memory layout at (n'-group I, k-chunk J)
//             are all the same function of 2I + J: with SBO = 256 B and LBO = 128 B every descriptor of every K step
//             aliases ONE 5.9 KB table instead of a 32 KB operand per step -- the band never has to be materialised.
//   3xTF32  : x = x_hi + x_lo, h = h_hi + h_lo (hi = round-to-nearest TF32, lo = exact f32 remainder); the tile is
//             D = A_hi B_hi + A_lo B_hi + A_hi B_lo accumulated in f32 in TMEM (the dropped lo*lo term and the TF32
//             truncation of the lo parts are ~2^-22 relative): 48 tcgen05.mma (128 x 64 x 8, kind::tf32) per tile.
//
// Pipeline (one persistent CTA per SM, 320 threads, all 512 TMEM columns):
//   warp 8  TMA producer : cp.async.bulk.tensor 3-D boxes {16 samples, 8 segments, 8 streams} x 4 per 64-sample block
//                          into a 4-stage shared ring, 128-byte swizzle; out-of-range boxes (before the stream start,
//                          past the last segment or stream) are zero-filled / clipped by the hardware.
//   warps 0-3 converter  : thread = row: reads its 64 samples of the block (conflict-free LDS.128 through the swizzle),
//                          splits them into TF32 hi / lo and writes them to its TMEM lane with tcgen05.st, into a ring
//                          of three 64-column blocks (hi) + three (lo).
//   warp 9  MMA issuer   : one elected lane (warp-uniform control flow, so the operands are provably uniform) issues the 48
//                          MMAs of a tile (A from TMEM, B descriptors into the aliased table) into one of two 64-column
//                          accumulators; tcgen05.commit signals the epilogue and hands the oldest A block back to the converter.
//   warps 4-7 epilogue   : tcgen05.ld the accumulator, scale, write the tile to a swizzled staging buffer and store
//                          it with cp.async.bulk.tensor (shared -> global).
// A CTA walks a contiguous range of (tile, block) items; every run inside a tile starts with kHB priming blocks (the
// samples before it: earlier blocks, the tail of the previous segment, or the object's history at the start of a call).
#include "common.cuh"
#include "fused_common.cuh"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace yg {

namespace {

using namespace yg::dev;

constexpr int kRows = 128;                    // 64 (stream, segment) rows x {re, im} (M of the MMA)
constexpr int kTileStreams = 8, kTileSegs = 8;
constexpr int kSubBytes = 16 * 64 * 8;        // one TMA box: 16 samples x 64 rows = 8 KB
constexpr int kThreads = 320;

// Geometry of one instantiation: blocks of kBlk samples (= outputs per tile, N of the MMA) and kHB blocks of history, so the
// K window is (kHB + 1) kBlk samples and filters of up to kHB kBlk + 1 taps fit.  TMEM holds kHB + 2 blocks of A (hi) + as
// many (lo) + two accumulators: 2 (kHB + 2) kBlk + 2 kBlk <= 512 columns.
//   <64, 1>: up to 65 taps, K = 128 (BASELINE config #2)     <32, 3> / <32, 4> / <32, 5>: up to 97 / 129 / 161 taps, K = 128 / 160 / 192
template <int kBlk_, int kHB_>
struct Cfg {
    static constexpr int kBlk = kBlk_, kHB = kHB_;
    static constexpr int kSlots = kHB + 2;                               // ring of A blocks
    static constexpr int kKSteps = (kHB + 1) * kBlk / 8;                 // MMAs (x 3) per tile
    static constexpr int kSub = kBlk / 16;                               // TMA boxes per block
    static constexpr int kStageBytes = kBlk * 64 * 8;
    static constexpr int kStages = 131072 / kStageBytes;                 // 128 KB of input ring
    static constexpr int kToepBytes = (kBlk / 8 + kKSteps - 1) * 256;    // aliased Toeplitz table (one of hi / lo)
    static constexpr int kSmemIn = 0;
    static constexpr int kSmemOut = kSmemIn + kStages * kStageBytes;     // 2 staging buffers
    static constexpr int kSmemToep = kSmemOut + 2 * kStageBytes;         // hi table, lo table
    static constexpr int kSmemBar = kSmemToep + 2 * kToepBytes;
    // mbarrier slots
    static constexpr int kBarInFull = 0, kBarInEmpty = kStages, kBarAFull = 2 * kStages, kBarAFree = kBarAFull + kSlots,
                         kBarDFull = kBarAFree + kSlots, kBarDEmpty = kBarDFull + 2, kNumBars = kBarDEmpty + 2;
    static constexpr int kSmemBytes = kSmemBar + 8 * kNumBars + 64 + 1024;          // + TMEM slot + alignment slack
    // TMEM columns
    static constexpr uint32_t kColHi = 0, kColLo = kSlots * kBlk, kColD = 2 * kSlots * kBlk;
    static_assert(kColD + 2 * kBlk <= 512, "tensor memory holds 512 columns");
    static_assert(kBlk == 32 || kBlk == 64, "blocks of 32 or 64 samples");
    static constexpr int kMaxTaps = kHB * kBlk + 1;
};

struct TcParams {
    const float2* hist;       // [n_streams][Hlen], oldest first (the object's state), or null
    int Hlen;
    long long n;              // samples per stream in this call
    int n_streams;
    float scale;
    const float* toep;        // [2][kToepBytes / 4]: hi table, lo table
    long long n_blocks;       // blocks per segment: Q / kBlk
    int n_groups;             // tiles: ceil(n_streams / 8) stream groups x seg_groups segment groups
    int seg_groups;           // ceil((n / Q) / 8)
    int q_floats;             // 2 Q: floats per segment
};

// ---- tcgen05 / TMA wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], 128 x 64 x 8, TF32 inputs, f32 accumulate
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}

// One lane of a converged warp.  The tcgen05 / TMA instructions take their addresses from UNIFORM registers: issued under
// `if (lane == 0)` ptxas cannot prove the operands uniform and wraps every instruction in an ELECT / R2UR.BROADCAST loop
// (~60 clk per MMA, twice the instruction's own 32 clk); issued by an elected lane inside warp-uniform control flow, with
// operands computed from warp-uniform values, they are single instructions.
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

#define YG_R8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
#define YG_W8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
// 32 consecutive columns of this thread's TMEM lane  <->  32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), YG_R8(v, 0), YG_R8(v, 8), YG_R8(v, 16), YG_R8(v, 24) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : YG_W8(v, 0), YG_W8(v, 8), YG_W8(v, 16), YG_W8(v, 24) : "r"(taddr) : "memory");
}
#undef YG_R8
#undef YG_W8

// x = hi + lo with hi the nearest TF32 (10 explicit mantissa bits; ties away from zero) and lo = x - hi exact in f32.
// Two integer ops instead of cvt.rna.tf32.f32, which ptxas expands into an Inf/NaN test and selects; a non-finite x
// still poisons the output through lo = x - hi.
__device__ __forceinline__ uint32_t tf32_hi(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }

// K-major, no-swizzle shared-memory matrix descriptor: 8 x 16-byte core matrices, LBO = stride between the two
// 16-byte K chunks of an instruction, SBO = stride between 8-row groups (cute::UMMA::SmemDescriptor bit layout).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);                      // descriptor version 1 (sm_100); base offset 0, layout type 0 = no swizzle
}
// kind::tf32, f32 accumulate, A and B K-major, N = 64, M = 128 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24); }

// The walk every role performs: the CTA's items [e0, e1) of the (tile, block) grid, cut into runs inside one tile;
// a run starts with kHB priming steps for the blocks before it.  step(tile, block, prime) is called in the same order by
// every role, so running counters (blocks converted, tiles produced) agree across roles.
template <int kHB, typename Step>
__device__ __forceinline__ void walk(long long e0, long long e1, long long n_blocks, Step&& step)
{
    long long e = e0;
    while (e < e1) {
        const int g = (int)(e / n_blocks);
        long long b = e - (long long)g * n_blocks;
        const long long run_end = (e1 < (long long)(g + 1) * n_blocks) ? e1 : (long long)(g + 1) * n_blocks;
#pragma unroll
        for (int i = kHB; i >= 1; i--) step(g, b - i, true);
        for (; e < run_end; e++, b++) step(g, b, false);
    }
}

template <class C>
__global__ void __launch_bounds__(kThreads, 1) k_firfilt_tc(const __grid_constant__ CUtensorMap tm_in,
                                                             const __grid_constant__ CUtensorMap tm_out, const TcParams p)
{
    constexpr int kBlk = C::kBlk, kHB = C::kHB, kSlots = C::kSlots, kStages = C::kStages, kStageBytes = C::kStageBytes;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;             // the 128-byte swizzle wants 1024-byte alignment
    unsigned char* smem_gen = smem_raw + (smem - smem_u32(smem_raw));
    const uint32_t bar0 = smem + C::kSmemBar;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + C::kSmemBar + 8 * C::kNumBars);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;      // warp index, provably uniform

    const long long total = (long long)p.n_groups * p.n_blocks;
    const long long e0 = total * blockIdx.x / gridDim.x, e1 = total * (blockIdx.x + 1) / gridDim.x;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) { mbar_init(bar0 + 8 * (C::kBarInFull + i), 1); mbar_init(bar0 + 8 * (C::kBarInEmpty + i), 4); }
        for (int i = 0; i < kSlots; i++) { mbar_init(bar0 + 8 * (C::kBarAFull + i), 4); mbar_init(bar0 + 8 * (C::kBarAFree + i), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(bar0 + 8 * (C::kBarDFull + i), 1); mbar_init(bar0 + 8 * (C::kBarDEmpty + i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // Toeplitz tables (hi, lo) -> shared memory; the tensor core reads them through the async proxy
        float* dst = reinterpret_cast<float*>(smem_gen + C::kSmemToep);
        for (int i = threadIdx.x; i < 2 * C::kToepBytes / 4; i += kThreads) dst[i] = __ldg(&p.toep[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 8) {
        // ================================================================= TMA producer (warp-uniform loop, one elected lane issues)
        long long k = 0;
        walk<kHB>(e0, e1, p.n_blocks, [&](int g, long long b, bool) {
            const int st = (int)(k % kStages);
            const uint32_t full = bar0 + 8 * (C::kBarInFull + st);
            if (k >= kStages) mbar_wait(bar0 + 8 * (C::kBarInEmpty + st), (uint32_t)(((k / kStages) - 1) & 1));
            if (elect_one()) {
                mbar_expect_tx(full, kStageBytes);
                const uint32_t dst = smem + C::kSmemIn + st * kStageBytes;
                // coordinates: (float inside the segment, segment of the stream, stream).  The blocks before a segment's first
                // one are the last blocks of the segment before it (segment -1 does not exist: zero-filled by the hardware).
                const int c2 = (g / p.seg_groups) * kTileStreams;
                const int c1 = (g % p.seg_groups) * kTileSegs - (b < 0 ? 1 : 0);
                const int c0 = (b < 0 ? p.q_floats : 0) + (int)(2 * kBlk * b);
#pragma unroll
                for (int j = 0; j < C::kSub; j++) tma_load_3d(dst + j * kSubBytes, &tm_in, c0 + 32 * j, c1, c2, full);
            }
            __syncwarp();
            k++;
        });
    } else if (warp == 9) {
        // ================================================================= MMA issuer (warp-uniform loop, one elected lane issues)
        long long k = 0, u = 0;
        const uint32_t toep_hi = smem + C::kSmemToep, toep_lo = toep_hi + C::kToepBytes;
        walk<kHB>(e0, e1, p.n_blocks, [&](int, long long, bool prime) {
            const int slot = (int)(k % kSlots);
            mbar_wait(bar0 + 8 * (C::kBarAFull + slot), (uint32_t)((k / kSlots) & 1));
            tc_fence_after();
            if (!prime) {
                const int db = (int)(u & 1);
                if (u >= 2) { mbar_wait(bar0 + 8 * (C::kBarDEmpty + db), (uint32_t)(((u >> 1) - 1) & 1)); tc_fence_after(); }
                const uint32_t d = tmem + C::kColD + kBlk * db;
                const int oldest = (slot + 2) % kSlots;                    // slot of block k - kHB (kSlots = kHB + 2)
                if (elect_one()) {
                    if constexpr (kBlk == 64 && kHB == 1) {
                        // B[k][n'] is zero unless 63 <= n' + k <= 127: K steps 0-3 (k < 32) touch only the columns n' >= 32 and
                        // K steps 12-15 (k >= 96) only n' < 32, so those eight steps are issued at half width (N = 32: half the
                        // tensor-pipe time); the full-width steps 4-11 come first so that one of them initialises the accumulator.
                        const uint32_t col_prev = (uint32_t)(kBlk * oldest), col_cur = (uint32_t)(kBlk * slot);
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int s = (i < 8) ? i + 4 : (i < 12 ? i - 8 : i);              // 4..11, 0..3, 12..15
                            const uint32_t col = (s < 8) ? col_prev + 8 * s : col_cur + 8 * (s - 8);
                            const bool upper = s < 4, lower = s >= 12;
                            const uint32_t boff = 256u * s + (upper ? 4u * 256u : 0u);         // n' group 4 = +4 SBO
                            const uint64_t bh = smem_desc(toep_hi + boff, 128, 256), bl = smem_desc(toep_lo + boff, 128, 256);
                            const uint32_t dd = d + (upper ? 32u : 0u);
                            const uint32_t id = (upper || lower) ? idesc_n(kBlk / 2) : idesc_n(kBlk);
                            tc_mma_ts(dd, tmem + C::kColLo + col, bh, id, i > 0 ? 1u : 0u);    // small terms first
                            tc_mma_ts(dd, tmem + C::kColHi + col, bl, id, 1u);
                            tc_mma_ts(dd, tmem + C::kColHi + col, bh, id, 1u);
                        }
                    } else {
                        // (a run-time skip of the K steps whose slice of the band is all zero for the actual tap count was tried:
                        // the per-step branches push the operands out of the uniform datapath and the MMA warp, the pacemaker
                        // of these instances, gets slower -- 4.52 vs 4.03 ms at 127 taps; hence one instance per history depth)
#pragma unroll
                        for (int s = 0; s < C::kKSteps; s++) {
                            constexpr int kPerBlock = kBlk / 8;                                // K steps per block of A
                            int sl = oldest + s / kPerBlock;                                   // ring slot of the block this step reads
                            if (sl >= kSlots) sl -= kSlots;
                            const uint32_t col = (uint32_t)(kBlk * sl + 8 * (s % kPerBlock));
                            const uint64_t bh = smem_desc(toep_hi + 256u * s, 128, 256), bl = smem_desc(toep_lo + 256u * s, 128, 256);
                            tc_mma_ts(d, tmem + C::kColLo + col, bh, idesc_n(kBlk), s > 0 ? 1u : 0u);
                            tc_mma_ts(d, tmem + C::kColHi + col, bl, idesc_n(kBlk), 1u);
                            tc_mma_ts(d, tmem + C::kColHi + col, bh, idesc_n(kBlk), 1u);
                        }
                    }
                    tc_commit(bar0 + 8 * (C::kBarDFull + db));
                }
                __syncwarp();
                u++;
            }
            if (k >= kHB) {                                                // block k - kHB is no longer read by any later tile
                if (elect_one()) tc_commit(bar0 + 8 * (C::kBarAFree + (int)((k - kHB) % kSlots)));
                __syncwarp();
            }
            k++;
        });
    } else if (warp < 4) {
        // ================================================================= converter: thread = row (stream, segment, component)
        const int row = threadIdx.x;
        const int sl = row >> 1, c = row & 1;                   // (stream, segment) row inside the tile, component
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        // byte offset of this row's 128-byte line inside a box; chunk u of the line sits at ((u ^ (sl & 7)) << 4)
        const uint32_t row_off = (uint32_t)sl * 128u;
        long long k = 0;
        walk<kHB>(e0, e1, p.n_blocks, [&](int g, long long b, bool) {
            const int st = (int)(k % kStages);
            const int slot = (int)(k % kSlots);
            mbar_wait(bar0 + 8 * (C::kBarInFull + st), (uint32_t)((k / kStages) & 1));
            const uint32_t src = smem + C::kSmemIn + st * kStageBytes + row_off;
            // row sl of the tile = (stream sl / 8, segment sl % 8) of the tile's 8 x 8 patch
            const long long s_glob = (long long)(g / p.seg_groups) * kTileStreams + (sl >> 3);
            const int seg_glob = (g % p.seg_groups) * kTileSegs + (sl & 7);
            const bool from_hist = (b < 0) && seg_glob == 0 && p.hist != nullptr && s_glob < p.n_streams;
#pragma unroll
            for (int half = 0; half < kBlk / 32; half++) {
                uint32_t hi[32], lo[32];
                if (!from_hist) {
#pragma unroll
                    for (int q = 0; q < 16; q++) {               // 16-byte chunk = 2 samples (re0, im0, re1, im1)
                        const int t = 32 * half + 2 * q;         // first sample of the chunk inside the block
                        const uint32_t a = src + (uint32_t)(t >> 4) * kSubBytes + ((uint32_t)(((t & 15) >> 1) ^ (sl & 7)) << 4);
                        const float4 v = lds128(a);
                        const float v0 = c ? v.y : v.x, v1 = c ? v.w : v.z;
                        const uint32_t h0 = tf32_hi(v0), h1 = tf32_hi(v1);
                        hi[2 * q] = h0; hi[2 * q + 1] = h1;
                        lo[2 * q] = __float_as_uint(v0 - __uint_as_float(h0));
                        lo[2 * q + 1] = __float_as_uint(v1 - __uint_as_float(h1));
                    }
                } else {
                    // blocks before the first one of the call: the object's history (oldest first, Hlen <= kHB kBlk samples)
                    const float* hp = reinterpret_cast<const float*>(p.hist + s_glob * p.Hlen);
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const int idx = p.Hlen + (int)b * kBlk + 32 * half + i;          // b < 0
                        const float v = (idx >= 0) ? __ldg(hp + 2 * idx + c) : 0.0f;
                        const uint32_t h0 = tf32_hi(v);
                        hi[i] = h0;
                        lo[i] = __float_as_uint(v - __uint_as_float(h0));
                    }
                }
                if (half == 0 && k >= kSlots) {  // the MMAs that read this ring slot's previous block have completed
                    mbar_wait(bar0 + 8 * (C::kBarAFree + slot), (uint32_t)(((k / kSlots) - 1) & 1));
                    tc_fence_after();
                }
                __syncwarp();
                tmem_st32(lane_base + C::kColHi + kBlk * slot + 32 * half, hi);
                tmem_st32(lane_base + C::kColLo + kBlk * slot + 32 * half, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar0 + 8 * (C::kBarInEmpty + st));
                mbar_arrive(bar0 + 8 * (C::kBarAFull + slot));
            }
            k++;
        });
    } else {
        // ================================================================= epilogue: thread = row (stream, segment, component)
        const int row = threadIdx.x - 128;
        const int sl = row >> 1, c = row & 1;
        const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)sl * 128u + 4u * c;
        const bool issuer = lane == 0 && (warp & 3) < C::kSub;      // one box per epilogue warp, issued (and its bulk group owned) by lane 0
        long long u = 0;
        walk<kHB>(e0, e1, p.n_blocks, [&](int g, long long b, bool prime) {
            if (prime) return;
            const int db = (int)(u & 1);
            mbar_wait(bar0 + 8 * (C::kBarDFull + db), (uint32_t)((u >> 1) & 1));
            tc_fence_after();
            uint32_t dv[kBlk / 32][32];                             // column n' of the tile <-> time kBlk - 1 - n'
#pragma unroll
            for (int h = 0; h < kBlk / 32; h++) tmem_ld32(lane_base + C::kColD + kBlk * db + 32 * h, dv[h]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8 * (C::kBarDEmpty + db));
            // the staging buffer was last used two tiles ago: its bulk store must have finished reading it
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // 16 rows x 2 components per warp: rows v and v + 8 share a swizzle key, so in each store the upper eight rows
            // take the other sample of the 16-byte chunk (flip) -- all 32 lanes then hit distinct banks.
            const uint32_t dst = smem + C::kSmemOut + db * kStageBytes + row_off;
            const bool flip = (sl >> 3) & 1;
            auto tile_val = [&](int t) { return dv[(kBlk - 1 - t) >> 5][(kBlk - 1 - t) & 31]; };
#pragma unroll
            for (int tt = 0; tt < kBlk; tt += 2) {
                const uint32_t a = dst + (uint32_t)(tt >> 4) * kSubBytes + ((uint32_t)(((tt & 15) >> 1) ^ (sl & 7)) << 4);
                const uint32_t lo_bits = tile_val(tt), hi_bits = tile_val(tt + 1);
                const float first = __uint_as_float(flip ? hi_bits : lo_bits) * p.scale;       // sample tt + flip
                const float second = __uint_as_float(flip ? lo_bits : hi_bits) * p.scale;      // sample tt + 1 - flip
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + (flip ? 8u : 0u)), "f"(first) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + (flip ? 0u : 8u)), "f"(second) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            {
                const int j = warp & 3;
                const uint32_t src = smem + C::kSmemOut + db * kStageBytes + j * kSubBytes;
                const int c0 = (int)(2 * kBlk * b) + 32 * j, c1 = (g % p.seg_groups) * kTileSegs, c2 = (g / p.seg_groups) * kTileStreams;
                if (issuer) {
                    tma_store_3d(&tm_out, c0, c1, c2, src);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                __syncwarp();
            }
            u++;
        });
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// x[stream][n] cf32 viewed as f32 [n_streams][n / Q segments][2 Q]; box = 32 floats (16 samples, 128 bytes) x 8 segments
// x 8 streams, 128-byte swizzle.  Why segments: with rows = 64 different streams a CTA cycles through 64 + 64 pages that
// are a power-of-two row pitch apart, which thrashes the TLB (measured: 4.2 TB/s with 8 MiB rows against 5.3 TB/s with
// 2 MiB rows); 8 streams x 8 adjacent 64 KB segments touch 8 + 8 pages.
int32_t make_map(CUtensorMap* tm, const float2* base, long long n, long long pitch, long long n_streams, long long Q)
{
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(YG_EINTERNAL, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[3] = {(cuuint64_t)(2 * Q), (cuuint64_t)(n / Q), (cuuint64_t)n_streams};
    const cuuint64_t gstride[2] = {(cuuint64_t)(8 * Q), (cuuint64_t)(8 * pitch)};
    const cuuint32_t box[3] = {32, (cuuint32_t)kTileSegs, (cuuint32_t)kTileStreams};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float2*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);     // promotion none / 64 / 256 B measured: no difference
    if (r != CUDA_SUCCESS) return fail(YG_EINTERNAL, "cuTensorMapEncodeTiled failed (%d) for n = %lld, streams = %lld", (int)r, n, n_streams);
    return YG_OK;
}

// segment length: 8192 samples (64 KB) for long streams (1 K .. 128 K measured: no difference, profiles/r02_firfilt_tc_sweeps.log),
// an eighth of the stream for short ones
constexpr long long kLongSegment = 8192;
long long segment_len(long long n) { return n >= kTileSegs * kLongSegment ? kLongSegment : n / kTileSegs; }

float tf32_rna(float v)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    b = (b + 0x1000u) & 0xFFFFE000u;          // round to nearest (ties away), keep 10 mantissa bits
    float r;
    memcpy(&r, &b, 4);
    return r;
}

using Cfg65 = Cfg<64, 1>;       // up to 65 taps (BASELINE config #2)
using Cfg97 = Cfg<32, 3>;       // up to 97 taps
using Cfg129 = Cfg<32, 4>;      // up to 129 taps
using Cfg161 = Cfg<32, 5>;      // up to 161 taps

// shortest segment a variant accepts: whole blocks, and the kHB blocks of history must lie inside the previous segment
template <class C> constexpr long long min_segment() { return C::kBlk == 64 ? 64 : 256; }

template <class C>
long long prefix_t(long long n)
{
    const long long unit = (n >= kTileSegs * kLongSegment) ? kLongSegment : (long long)kTileSegs * min_segment<C>();
    const long long n_main = n / unit * unit;
    return n_main >= kTileSegs * min_segment<C>() ? n_main : 0;
}

// Aliased Toeplitz tables (hi, lo) for taps h: the entry at byte A of a table is B at n' + k = 8 (A >> 8) + 4 (A >> 7 & 1)
// + (A >> 4 & 7) + (A >> 2 & 3), i.e. tap (kHB + 1) kBlk - 1 - (n' + k).
template <class C>
int32_t plan_t(const float* h, size_t h_len, float** d_toep)
{
    std::vector<float> t(2 * C::kToepBytes / 4, 0.0f);
    for (int A = 0; A < C::kToepBytes; A += 4) {
        const int v = 8 * (A >> 8) + 4 * ((A >> 7) & 1) + ((A >> 4) & 7) + ((A >> 2) & 3);      // n' + k
        const int j = (C::kHB + 1) * C::kBlk - 1 - v;
        if (j >= 0 && j < (int)h_len) {
            const float hi = tf32_rna(h[j]);
            t[A / 4] = hi;
            t[C::kToepBytes / 4 + A / 4] = h[j] - hi;
        }
    }
    if (!*d_toep) YG_CUDA(cudaMalloc(d_toep, t.size() * sizeof(float)));
    YG_CUDA(memcpy_sync(*d_toep, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
    YG_CUDA(cudaFuncSetAttribute(k_firfilt_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    return YG_OK;
}

template <class C>
int32_t launch_t(const float* d_toep, float scale, const float2* hist, long long Hlen, const float2* x, float2* y,
                 long long n, long long pitch, long long n_streams, int n_sm, cudaStream_t st)
{
    CUtensorMap tm_in, tm_out;
    const long long Q = segment_len(n);
    if (Q < min_segment<C>() || Q % C::kBlk != 0 || n % Q != 0)
        return fail(YG_EINTERNAL, "tensor-core firfilt: %lld samples are not whole segments", n);
    YG_TRY(make_map(&tm_in, x, n, pitch, n_streams, Q));
    YG_TRY(make_map(&tm_out, y, n, pitch, n_streams, Q));
    TcParams p;
    p.hist = (Hlen > 0) ? hist : nullptr;
    p.Hlen = (int)Hlen;
    p.n = n;
    p.n_streams = (int)n_streams;
    p.scale = scale;
    p.toep = d_toep;
    p.n_blocks = Q / C::kBlk;
    p.seg_groups = (int)((n / Q + kTileSegs - 1) / kTileSegs);
    p.n_groups = (int)((n_streams + kTileStreams - 1) / kTileStreams) * p.seg_groups;
    p.q_floats = (int)(2 * Q);
    const long long total = p.n_blocks * p.n_groups;
    const int grid = (int)std::min<long long>(n_sm, total);
    k_firfilt_tc<C><<<grid, kThreads, C::kSmemBytes, st>>>(tm_in, tm_out, p);
    YG_LAUNCH_CHECK();
    return YG_OK;
}

}  // namespace

bool firfilt_tc_taps_ok(size_t h_len) { return h_len >= 1 && h_len <= (size_t)Cfg161::kMaxTaps; }

// The longest prefix of an n-sample stream the kernel takes: whole 8192-sample segments of long streams, an eighth of the
// prefix per segment (whole blocks, history inside the previous segment) for short ones; 0 if it takes nothing.
long long firfilt_tc_prefix(size_t h_len, long long n, long long n_streams, const void* x, const void* y)
{
    if (!firfilt_tc_taps_ok(h_len)) return 0;
    if (n & 1) return 0;                                                   // row pitch must be a multiple of 16 bytes
    if (n >= (1LL << 29) || n_streams > 0x7fffffffLL) return 0;            // int32 box coordinates
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return 0;
    const long long n_main = h_len <= (size_t)Cfg65::kMaxTaps ? prefix_t<Cfg65>(n) : prefix_t<Cfg161>(n);      // all 32-sample instances alike
    if (n_main * n_streams < (1LL << 16)) return 0;                        // tiny calls: not worth 148 persistent CTAs
    return encode_tiled() != nullptr ? n_main : 0;
}

int32_t firfilt_tc_plan(const float* h, size_t h_len, float** d_toep)
{
    if (h_len <= (size_t)Cfg65::kMaxTaps) return plan_t<Cfg65>(h, h_len, d_toep);
    if (h_len <= (size_t)Cfg97::kMaxTaps) return plan_t<Cfg97>(h, h_len, d_toep);
    if (h_len <= (size_t)Cfg129::kMaxTaps) return plan_t<Cfg129>(h, h_len, d_toep);
    return plan_t<Cfg161>(h, h_len, d_toep);
}

// Outputs [0, n) of every stream; rows of x and y are `pitch` samples apart (pitch >= n, even).
int32_t firfilt_tc_launch(const float* d_toep, size_t h_len, float scale, const float2* hist, long long Hlen, const float2* x, float2* y,
                          long long n, long long pitch, long long n_streams, int n_sm, cudaStream_t st)
{
    if (h_len <= (size_t)Cfg65::kMaxTaps) return launch_t<Cfg65>(d_toep, scale, hist, Hlen, x, y, n, pitch, n_streams, n_sm, st);
    if (h_len <= (size_t)Cfg97::kMaxTaps) return launch_t<Cfg97>(d_toep, scale, hist, Hlen, x, y, n, pitch, n_streams, n_sm, st);
    if (h_len <= (size_t)Cfg129::kMaxTaps) return launch_t<Cfg129>(d_toep, scale, hist, Hlen, x, y, n, pitch, n_streams, n_sm, st);
    return launch_t<Cfg161>(d_toep, scale, hist, Hlen, x, y, n, pitch, n_streams, n_sm, st);
}

}  // namespace yg
