// tools/tma_bench.cu -- how fast can one SM pull HBM through cp.async.bulk (1-D TMA) as a function of
// copy size and the number of copies kept in flight?  One CTA per SM, one issuing thread.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_tma(const char* __restrict__ src, long long bytes_per_cta, int copy_bytes, int depth, int issuers, float* sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // depth barriers
    const uint32_t data = s32(smem) + 1024;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < depth; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char* base = src + (long long)blockIdx.x * bytes_per_cta;
    const long long n = bytes_per_cta / copy_bytes;
    // thread t < issuers owns copies i with i % issuers == t; stage = i % depth
    if (tid < issuers) {
        for (long long i = tid; i < n + depth; i += issuers) {
            const int st = (int)(i % depth);
            const uint32_t bar = s32(&bars[st]);
            if (i >= depth) {
                const uint32_t par = (uint32_t)(((i / depth) - 1) & 1);
                asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(par) : "memory");
            }
            if (i < n) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(copy_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(data + st * copy_bytes), "l"(base + i * copy_bytes), "r"(copy_bytes), "r"(bar) : "memory");
            }
        }
    }
    __syncthreads();
    if (tid == 0) sink[blockIdx.x] = (float)smem[1024];
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nsm = prop.multiProcessorCount;
    const long long per = 16LL << 20;          // 16 MiB per CTA -> 2.3 GiB total
    char* src; float* sink;
    CK(cudaMalloc(&src, per * nsm)); CK(cudaMemset(src, 1, per * nsm)); CK(cudaMalloc(&sink, nsm * 4));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int issuers : {1, 4}) for (int copy : {2048, 4096, 16384, 32768}) for (int inflight_kb : {32, 64, 128, 192}) {
        int depth = inflight_kb * 1024 / copy;
        if (depth < 1 || depth > 96) continue;
        if (depth % issuers) continue;
        const size_t smem = 1024 + (size_t)depth * copy;
        float best = 1e30f;
        for (int r = 0; r < 4; r++) {
            cudaEventRecord(e0);
            k_tma<<<nsm, 128, smem>>>(src, per, copy, depth, issuers, sink);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("issuers=%d copy=%6d B  in-flight=%3d KB (depth %2d): %.3f ms  %.0f GB/s  (%.1f B/ns per SM)\n", issuers, copy, inflight_kb, depth,
               best, per * nsm / best * 1e-6, per / best * 1e-6);
    }
    return 0;
}
