// firpfbch2_fast.cu -- fused firpfbch2 analysis kernel (placeholder until the kernel lands).
#include "firpfbch2_fast.cuh"

namespace yg {

int32_t firpfbch2_fast_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float*)
{
    p.supported = false;
    p.M = M;
    p.m = m;
    return YG_OK;
}

void firpfbch2_fast_release(Firpfbch2FastPlan& p)
{
    if (p.d_taps) cudaFree(p.d_taps);
    if (p.d_twid) cudaFree(p.d_twid);
    p.d_taps = p.d_twid = nullptr;
    p.supported = false;
}

int32_t firpfbch2_fast_launch(const Firpfbch2FastPlan&, const float2*, long long, const float2*, float2*, size_t,
                              size_t, cudaStream_t)
{
    return fail(YG_EINTERNAL, "fused kernel not available");
}

}  // namespace yg
