// fused_common.cuh -- device helpers shared by the fused sm_100a kernels: mbarrier / TMA bulk-copy
// wrappers, explicit shared-space accesses, packed f32x2 arithmetic and the radix-4 x radix-4
// 16-point backward DFT on (even frame, odd frame) float2 lanes.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace yg {
namespace dev {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Spin on try_wait WITHOUT a suspend-time hint: with the hint (0x989680, as CUTLASS uses) the fused kernels
// got slower (firpfbch 0.38 -> 0.50 ms, sustained firpfbch2 analysis 1.27 -> 1.52 ms): wake-up latency
// matters more here than the issue slots the spin loop burns (~12 % of executed instructions).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still draining; pdl_wait() blocks until that predecessor has completed
// and its writes are visible (a no-op for a normally launched kernel), pdl_launch_dependents() lets the successor's
// CTAs be scheduled as soon as SMs free up.  Rule used here: nothing the caller could have produced with the
// previous kernel (x, the history) is read, and nothing is written, before pdl_wait(); only the prologue (barrier
// initialisation, plan-time tables -> registers) runs ahead.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// explicit shared-space accesses on 32-bit shared addresses (keeps them LDS/STS, never generic LD/ST)
__device__ __forceinline__ float2 lds64(uint32_t a)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v)
{
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// packed f32x2 helpers: a float2 holds the (even frame, odd frame) values of one real quantity
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 fnma2(float2 a, float2 b, float2 c) { return __ffma2_rn(make_float2(-a.x, -a.y), b, c); }

struct C2 { float2 re, im; };           // one complex value for each frame of the pair

__device__ __forceinline__ C2 cadd(C2 a, C2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ C2 csub(C2 a, C2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a + j b,  a - j b
__device__ __forceinline__ C2 caddj(C2 a, C2 b) { return {sub2(a.re, b.im), add2(a.im, b.re)}; }
__device__ __forceinline__ C2 csubj(C2 a, C2 b) { return {add2(a.re, b.im), sub2(a.im, b.re)}; }
// a * (wr + j wi), scalar twiddle shared by both frames
__device__ __forceinline__ C2 cmulw(C2 a, float wr, float wi)
{
    C2 r;
    r.re = fnma2(a.im, f2(wi), mul2(a.re, f2(wr)));
    r.im = fma2(a.im, f2(wr), mul2(a.re, f2(wi)));
    return r;
}

// 4-point backward DFT (W4 = +j), in place
__device__ __forceinline__ void dft4(C2& a0, C2& a1, C2& a2, C2& a3)
{
    const C2 s0 = cadd(a0, a2), d0 = csub(a0, a2);
    const C2 s1 = cadd(a1, a3), d1 = csub(a1, a3);
    a0 = cadd(s0, s1);
    a2 = csub(s0, s1);
    a1 = caddj(d0, d1);
    a3 = csubj(d0, d1);
}

// 16-point backward DFT: out[k] = sum_n v[n] e^{+j 2 pi n k / 16}.
// Input natural order; output left in v[] at index (k1 + 4 k2) -> stored at v[4 k1 + k2]
// (digit-reversed base 4); callers index through dr4().
__device__ __forceinline__ constexpr int dr4(int k) { return ((k & 3) << 2) | (k >> 2); }

// a + c p,  a - c p  (real scalar c shared by both frames)
__device__ __forceinline__ C2 cfma(C2 p, float c, C2 a) { return {fma2(p.re, f2(c), a.re), fma2(p.im, f2(c), a.im)}; }
// a + j c p,  a - j c p
__device__ __forceinline__ C2 cfmaj(C2 p, float c, C2 a) { return {fma2(p.im, f2(-c), a.re), fma2(p.re, f2(c), a.im)}; }
// t (1 + j) and t (-1 + j): the W16^2 and W16^6 twiddles without their 1/sqrt(2), which is folded
// into the FMAs of the butterfly that consumes them
__device__ __forceinline__ C2 w8u(C2 t) { return {sub2(t.re, t.im), add2(t.re, t.im)}; }
__device__ __forceinline__ C2 w8u3(C2 t)
{
    return {__fadd2_rn(make_float2(-t.re.x, -t.re.y), make_float2(-t.im.x, -t.im.y)), sub2(t.re, t.im)};
}

__device__ __forceinline__ void dft16(C2 (&v)[16])
{
    constexpr float c1 = 0.92387953251128674f;      // cos(pi/8)
    constexpr float s1 = 0.38268343236508977f;      // sin(pi/8)
    constexpr float r2 = 0.70710678118654752f;      // sqrt(1/2)
    // stage 1: for each b, DFT4 over a of v[4a + b]  -> T_b[k1] stored at v[4 k1 + b]
#pragma unroll
    for (int b = 0; b < 4; b++) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
    // stage 2: for each k1, DFT4 over b of W16^{b k1} T_b[k1] -> X[k1 + 4 k2] stored at v[4 k1 + k2]
    // k1 = 0: no twiddles
    dft4(v[0], v[1], v[2], v[3]);
    // k1 = 1: W^1, W^2 (folded), W^3
    {
        const C2 a0 = v[4], a1 = cmulw(v[5], c1, s1), p2 = w8u(v[6]), a3 = cmulw(v[7], s1, c1);
        const C2 s0 = cfma(p2, r2, a0), d0 = cfma(p2, -r2, a0);
        const C2 sA = cadd(a1, a3), dA = csub(a1, a3);
        v[4] = cadd(s0, sA); v[6] = csub(s0, sA); v[5] = caddj(d0, dA); v[7] = csubj(d0, dA);
    }
    // k1 = 2: W^2 (folded), W^4 = j, W^6 (folded)
    {
        const C2 a0 = v[8], p1 = w8u(v[9]), t2 = v[10], p3 = w8u3(v[11]);
        const C2 s0 = caddj(a0, t2), d0 = csubj(a0, t2);
        const C2 sU = cadd(p1, p3), dU = csub(p1, p3);       // both still to be scaled by r2
        v[8] = cfma(sU, r2, s0); v[10] = cfma(sU, -r2, s0); v[9] = cfmaj(dU, r2, d0); v[11] = cfmaj(dU, -r2, d0);
    }
    // k1 = 3: W^3, W^6 (folded), W^9
    {
        const C2 a0 = v[12], a1 = cmulw(v[13], s1, c1), p2 = w8u3(v[14]), a3 = cmulw(v[15], -c1, -s1);
        const C2 s0 = cfma(p2, r2, a0), d0 = cfma(p2, -r2, a0);
        const C2 sA = cadd(a1, a3), dA = csub(a1, a3);
        v[12] = cadd(s0, sA); v[14] = csub(s0, sA); v[13] = caddj(d0, dA); v[15] = csubj(d0, dA);
    }
}

// packed pair <-> one 16-byte shared-memory entry (re_e, re_o, im_e, im_o)
__device__ __forceinline__ C2 ldc2(uint32_t a)
{
    const float4 q = lds128(a);
    return {make_float2(q.x, q.y), make_float2(q.z, q.w)};
}
__device__ __forceinline__ void stc2(uint32_t a, C2 z) { sts128(a, make_float4(z.re.x, z.re.y, z.im.x, z.im.y)); }

// radix-R backward DFT on packed pairs, natural order in and out (R = 2, 4, 8, 16)
template <int R> __device__ __forceinline__ void dft_r(C2* u);
template <> __device__ __forceinline__ void dft_r<2>(C2* u)
{
    const C2 s = cadd(u[0], u[1]), d = csub(u[0], u[1]);
    u[0] = s; u[1] = d;
}
template <> __device__ __forceinline__ void dft_r<4>(C2* u) { dft4(u[0], u[1], u[2], u[3]); }
template <> __device__ __forceinline__ void dft_r<8>(C2* u)
{
    constexpr float r2 = 0.70710678118654752f;
    dft4(u[0], u[2], u[4], u[6]);                                        // E[k] at u[2k]
    dft4(u[1], u[3], u[5], u[7]);                                        // O[k] at u[2k + 1]
    const C2 e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6];
    const C2 o0 = u[1], p1 = w8u(u[3]), o2 = u[5], p3 = w8u3(u[7]);      // p1, p3 still to be scaled by r2
    u[0] = cadd(e0, o0);       u[4] = csub(e0, o0);
    u[1] = cfma(p1, r2, e1);   u[5] = cfma(p1, -r2, e1);
    u[2] = caddj(e2, o2);      u[6] = csubj(e2, o2);
    u[3] = cfma(p3, r2, e3);   u[7] = cfma(p3, -r2, e3);
}
template <> __device__ __forceinline__ void dft_r<16>(C2* u)
{
    C2(&v)[16] = *reinterpret_cast<C2(*)[16]>(u);
    dft16(v);
    C2 t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = v[dr4(k)];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = t[k];
}

}  // namespace dev
}  // namespace yg
