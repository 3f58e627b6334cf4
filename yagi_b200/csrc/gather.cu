// gather.cu -- channel-major view of analysis output (the stream front-end).
//
// The analysers write frame-major y[frame][M] (what consecutive reference `execute` calls produce).  A consumer
// that wants one contiguous time series per channel gets it with this tiled transpose: 32 x 32 tiles of cf32
// through padded shared memory, 256-byte coalesced rows on both sides.  HBM-bound: 16 B per sample.
#include "common.cuh"

namespace {

constexpr int kTile = 32;

__global__ void __launch_bounds__(256) k_channel_major(const float2* __restrict__ y, float2* __restrict__ out,
                                                       long long n_frames, int M)
{
    __shared__ float2 tile[kTile][kTile + 1];
    const long long f0 = (long long)blockIdx.x * kTile;
    const int c0 = blockIdx.y * kTile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const long long f = f0 + r;
        const int c = c0 + tx;
        if (f < n_frames && c < M) tile[r][tx] = __ldcs(&y[f * M + c]);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < kTile; r += 8) {
        const int c = c0 + r;
        const long long f = f0 + tx;
        if (c < M && f < n_frames) __stcs(&out[(long long)c * n_frames + f], tile[tx][r]);
    }
}

}  // namespace

extern "C" int32_t yg_channel_major_dev(const yg_cf32* d_frames, size_t n_frames, uint32_t M, yg_cf32* d_out, void* cuda_stream)
{
    using namespace yg;
    if (n_frames == 0 || M == 0) return YG_OK;
    if (!d_frames || !d_out) return fail(YG_EVALUE, "null buffer");
    if (d_frames == d_out) return fail(YG_EVALUE, "the transpose is out of place");
    const unsigned gx = (unsigned)((n_frames + kTile - 1) / kTile), gy = (M + kTile - 1) / kTile;
    if (gy > 65535u) return fail(YG_ECONFIG, "too many channels (%u)", M);
    k_channel_major<<<dim3(gx, gy), 256, 0, (cudaStream_t)cuda_stream>>>(reinterpret_cast<const float2*>(d_frames),
                                                                          reinterpret_cast<float2*>(d_out), (long long)n_frames, (int)M);
    YG_LAUNCH_CHECK();
    return YG_OK;
}
