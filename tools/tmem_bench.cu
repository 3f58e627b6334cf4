// tools/tmem_bench.cu -- tensor-memory load / store throughput per SM when TMEM is used as a lane-private scratch
// (tcgen05.ld / tcgen05.st 32x32b.x16, 8 warps per CTA, one CTA per SM).  Question behind it: can the per-branch
// windows and taps of a large-M channelizer live in TMEM instead of registers / shared memory?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#define R16(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
#define W16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])

__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : W16(v) : "r"(a) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(a), R16(v) : "memory");
}

#define W32(v) W16(v), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
// 32 lanes x 32 columns in one instruction
__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : W32(v) : "r"(a) : "memory");
}
// 16 lanes x 256 bits, repeated 8 times along the columns: 32 registers per thread, 4 KB per warp instruction
__device__ __forceinline__ void ld16x256(uint32_t a, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : W32(v) : "r"(a) : "memory");
}

// mode 0: loads only (4 independent x16 loads per iteration, then wait::ld); mode 1: stores only; mode 2: ld + st + 16 FMAs
__global__ void __launch_bounds__(256, 1) k_tmem(int iters, int mode, unsigned* out, long long* clk)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 256u;    // 256 columns per warp half
    uint32_t v[4][16];
    for (int q = 0; q < 4; q++)
        for (int i = 0; i < 16; i++) v[q][i] = threadIdx.x * 64 + q * 16 + i;
    for (int q = 0; q < 4; q++) st16(base + 16 * q, v[q]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        const uint32_t a = base + (uint32_t)((it & 3) * 64);
        if (mode == 0 || mode == 2) {
            for (int q = 0; q < 4; q++) ld16(a + 16 * q, v[q]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (mode == 2)
            for (int q = 0; q < 4; q++)
                for (int i = 0; i < 16; i++) v[q][i] = v[q][i] * 3u + (unsigned)it;
        if (mode == 3) {                                   // 2 x (32x32b.x32): the same 64 columns per thread
            uint32_t w[32];
            ld32(a, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += w[it & 31];
            ld32(a + 32, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += w[(it + 7) & 31];
        }
        if (mode == 4) {                                   // 2 x (16x256b.x8): lanes 0-15 and 16-31 of the warp's quadrant, 64 columns
            uint32_t w[32];
            ld16x256(a, w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += w[it & 31];
            ld16x256(a + (16u << 16), w);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += w[(it + 7) & 31];
        }
        if (mode == 1 || mode == 2) {
            for (int q = 0; q < 4; q++) st16(a + 16 * q, v[q]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        acc += v[it & 3][it & 15];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main()
{
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int grid = prop.multiProcessorCount, iters = 20000;
    unsigned* out;
    long long* clk;
    CK(cudaMalloc(&out, grid * 256 * sizeof(unsigned)));
    CK(cudaMalloc(&clk, grid * sizeof(long long)));
    const char* names[5] = {"tcgen05.ld 4 x (32x32b.x16) per iteration", "tcgen05.st 4 x (32x32b.x16) per iteration", "ld + 64 IMAD + st per iteration",
                            "tcgen05.ld 2 x (32x32b.x32) per iteration", "tcgen05.ld 2 x (16x256b.x8) per iteration"};
    for (int mode = 0; mode < 5; mode++) {
        k_tmem<<<grid, 256>>>(iters, mode, out, clk);
        CK(cudaDeviceSynchronize());
        long long h[512];
        CK(cudaMemcpy(h, clk, grid * sizeof(long long), cudaMemcpyDeviceToHost));
        double c = 0;
        for (int i = 0; i < grid; i++) c += (double)h[i];
        c /= grid;
        const double bytes = (double)iters * 256 * 64 * 4 * (mode == 2 ? 2 : 1);      // per CTA: 256 threads x 64 columns x 4 B
        printf("%-48s %8.1f clk per iteration, %7.1f B/clk/SM (8 warps, 64 columns per thread per iteration)\n", names[mode], c / iters, bytes / c);
    }
    return 0;
}
