// firpfbch_fast.cuh -- interface of the fused firpfbch (critically sampled) analysis kernel (sm_100a).
#pragma once
#include "common.cuh"

namespace yg {

struct FirpfbchFastPlan {
    bool supported = false;
    uint32_t p = 0;
    uint32_t M = 0;
    int32_t type = 0;
    void* d_taps = nullptr;
    void* d_twid = nullptr;
    int n_sm = 0;
};

int32_t firpfbch_fast_plan(FirpfbchFastPlan& plan, int32_t type, uint32_t M, uint32_t p, const float* h);
void firpfbch_fast_release(FirpfbchFastPlan& plan);
// n_streams must be a multiple of 4; layouts as the generic kernel: x[stream][n_frames*64], hist[stream][Hlen]
int32_t firpfbch_fast_launch(const FirpfbchFastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                             long long n_frames, long long n_streams, cudaStream_t st);

int32_t firpfbch_fast_synth_launch(const FirpfbchFastPlan& plan, const float2* hist, long long hist_frames, const float2* x,
                                   float2* y, long long n_frames, long long n_streams, cudaStream_t st);

// Tiny-M kernels (firpfbch_tiny.cu, M = 8 / 16 / 32, analysis and synthesis by plan.type): n_streams must be a multiple
// of 32 / M; hist as above (analysis: Hlen = (p-1) M samples; synthesis: Hlen = hist_frames M, hist_frames >= 16);
// x, y and hist 16-byte aligned.
int32_t firpfbch_tiny_plan(FirpfbchFastPlan& plan, int32_t type, uint32_t M, uint32_t p, const float* h);
int32_t firpfbch_tiny_launch(const FirpfbchFastPlan& plan, const float2* hist, long long Hlen, const float2* x, float2* y,
                             long long n_frames, long long n_streams, cudaStream_t st);

}  // namespace yg
