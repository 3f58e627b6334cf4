// tools/microbench.cu -- B200 micro-measurements that size the fused channelizer kernel:
// FP32 pipe throughput (FFMA vs packed FFMA2/FADD2/FMUL2), shared-memory bandwidth, and the
// achievable HBM bandwidth of a 1:2 read:write streaming pattern (8 B in, 16 B out per sample).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

constexpr int ILP = 8;
constexpr int ITERS = 2048;

// mode 0: scalar FFMA, 1: FFMA2, 2: FADD2, 3: FMUL2, 4: FFMA2 with broadcast scalar b, 5: scalar FADD
template <int MODE>
__global__ void k_fp(float* out, long long* cycles, float seed)
{
    float a[ILP * 2];
#pragma unroll
    for (int i = 0; i < ILP * 2; i++) a[i] = seed + i + threadIdx.x;
    const float b = seed * 0.5f, c = seed * 0.25f;
    unsigned long long p[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
    const unsigned long long pb = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(b);
    const unsigned long long pbb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) { a[2 * i] = fmaf(a[2 * i], b, c); a[2 * i + 1] = fmaf(a[2 * i + 1], b, c); }
            else if (MODE == 1) p[i] = f2_fma(p[i], pb, pb);
            else if (MODE == 2) p[i] = f2_add(p[i], pb);
            else if (MODE == 3) p[i] = f2_mul(p[i], pb);
            else if (MODE == 4) p[i] = f2_fma(p[i], pbb, pb);
            else if (MODE == 5) { a[2 * i] = a[2 * i] + b; a[2 * i + 1] = a[2 * i + 1] + c; }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) {
        s += a[2 * i] + a[2 * i + 1];
        s += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory bandwidth: each thread reads W-byte words, conflict-free
template <int W>
__global__ void k_lds(float* out, long long* cycles)
{
    extern __shared__ float4 sm4[];
    float* sm = reinterpret_cast<float*>(sm4);
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < 1024; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int idx = ((it * 8 + u) * 64 + threadIdx.x * (W / 4)) & 8191 & ~(W / 4 - 1);
            if (W == 16) { float4 v = *reinterpret_cast<float4*>(sm + idx); acc += v.x + v.y + v.z + v.w; }
            else if (W == 8) { float2 v = *reinterpret_cast<float2*>(sm + idx); acc += v.x + v.y; }
            else acc += sm[idx];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// streaming pattern of the channelizer: read n float2, write 2n float2 (float4 stores)
template <int HINT>
__global__ void k_stream12(const float2* __restrict__ x, float4* __restrict__ y, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float2 v;
        if (HINT) v = __ldcs(&x[i]); else v = x[i];
        float4 o = make_float4(v.x, v.y, v.x + 1.f, v.y + 1.f);
        if (HINT) __stcs(&y[i], o); else y[i] = o;
    }
}

__global__ void k_copy(const float4* __restrict__ x, float4* __restrict__ y, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = x[i];
}

template <typename F>
float time_ms(F&& f, int reps = 10)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

template <int MODE>
void run_fp(const char* name, int warps, int nsm, float* d_out, long long* d_cyc, double flop_per_op)
{
    k_fp<MODE><<<nsm, warps * 32>>>(d_out, d_cyc, 1.0f);
    CK(cudaDeviceSynchronize());
    float ms = time_ms([&] { k_fp<MODE><<<nsm, warps * 32>>>(d_out, d_cyc, 1.0f); }, 5);
    std::vector<long long> cyc(nsm);
    CK(cudaMemcpy(cyc.data(), d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : cyc) avg += c; avg /= nsm;
    const double instr = (double)ITERS * ILP * (MODE == 0 || MODE == 5 ? 2 : 1) * warps;   // warp-instr per SM
    printf("%-28s warps/SM=%2d  warp-instr/clk/SM=%.3f  lane-flop/clk/SM=%.1f  (%.3f ms, %.0f cyc, ~%.0f MHz)\n", name, warps,
           instr / avg, instr * 32 * flop_per_op / avg, ms, avg, avg / (ms * 1e3));
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    printf("device: %s, %d SMs, smem/SM %zu KB, L2 %d MB\n", prop.name, nsm, prop.sharedMemPerMultiprocessor / 1024, prop.l2CacheSize >> 20);
    float* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, nsm * 1024 * sizeof(float)));
    CK(cudaMalloc(&d_cyc, nsm * sizeof(long long)));

    for (int warps : {4, 8, 16, 32}) {
        run_fp<0>("FFMA (scalar)", warps, nsm, d_out, d_cyc, 2);
        run_fp<1>("FFMA2 (f32x2)", warps, nsm, d_out, d_cyc, 4);
        run_fp<4>("FFMA2 bcast-b", warps, nsm, d_out, d_cyc, 4);
        run_fp<2>("FADD2", warps, nsm, d_out, d_cyc, 2);
        run_fp<3>("FMUL2", warps, nsm, d_out, d_cyc, 2);
        run_fp<5>("FADD (scalar)", warps, nsm, d_out, d_cyc, 1);
    }

    // shared memory
    for (int warps : {8, 16, 32}) {
        auto run = [&](auto kern, int W, const char* name) {
            kern<<<nsm, warps * 32, 32768>>>(d_out, d_cyc);
            CK(cudaDeviceSynchronize());
            std::vector<long long> cyc(nsm);
            CK(cudaMemcpy(cyc.data(), d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0; for (auto c : cyc) avg += c; avg /= nsm;
            printf("%-10s warps/SM=%2d  bytes/clk/SM=%.1f\n", name, warps, 1024.0 * 8 * warps * 32 * W / avg);
        };
        run(k_lds<4>, 4, "LDS.32");
        run(k_lds<8>, 8, "LDS.64");
        run(k_lds<16>, 16, "LDS.128");
    }

    // HBM streaming, 1:2 read:write, 2^28 samples (2 GiB in, 4 GiB out)
    const long long n = 1LL << 28;
    float2* x; float4* y;
    CK(cudaMalloc(&x, n * sizeof(float2)));
    CK(cudaMalloc(&y, n * sizeof(float4)));
    CK(cudaMemset(x, 0, n * sizeof(float2)));
    for (int bpsm : {4, 8, 16, 32}) {
        for (int threads : {256, 512}) {
            float ms0 = time_ms([&] { k_stream12<0><<<nsm * bpsm, threads>>>(x, y, n); });
            float ms1 = time_ms([&] { k_stream12<1><<<nsm * bpsm, threads>>>(x, y, n); });
            printf("stream 1:2  grid=%dx%d thr=%d  default: %.3f ms %.0f GB/s | .cs hints: %.3f ms %.0f GB/s\n", nsm, bpsm, threads,
                   ms0, 24.0 * n / ms0 * 1e-6, ms1, 24.0 * n / ms1 * 1e-6);
        }
    }
    {
        const long long n4 = (1LL << 30) / 16 * 2;    // 2 GiB copy
        float ms = time_ms([&] { k_copy<<<nsm * 16, 512>>>((const float4*)y, (float4*)y + n4, n4); });
        printf("copy 1:1 (2 GiB): %.3f ms %.0f GB/s\n", ms, 32.0 * n4 / ms * 1e-6);
        float msm = time_ms([&] { cudaMemcpyAsync((float4*)y + n4, y, n4 * 16, cudaMemcpyDeviceToDevice); });
        printf("cudaMemcpy D2D (2 GiB): %.3f ms %.0f GB/s\n", msm, 32.0 * n4 / msm * 1e-6);
    }
    return 0;
}
