// firpfbch2_fast.cuh -- interface of the fused firpfbch2 analysis kernel (sm_100a).
#pragma once
#include "common.cuh"

namespace yg {

struct Firpfbch2FastPlan {
    bool supported = false;
    uint32_t M = 0, m = 0;
    size_t min_frames = 0;        // below this the generic kernel is used
    void* d_taps = nullptr;       // kernel-specific tap layout (device)
    void* d_twid = nullptr;       // kernel-specific twiddle layout (device)
    int n_sm = 0;
    bool pdl = true;              // launch with programmatic stream serialization (fused M=256 analysis)
    void* d_scratch = nullptr;    // large-M fused analysis: per-group V ring (device)
    void* d_flags = nullptr;      //   and its counters
    int n_groups = 0;             //   0: fused large-M kernel not available
    bool single_sm = false;       // M = 1024, m <= 4: the one-CTA-per-SM analysis / synthesis kernel takes the call
};

// Decide whether (M, m) has a fused kernel and upload its tap / twiddle tables.
int32_t firpfbch2_fast_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
void firpfbch2_fast_release(Firpfbch2FastPlan& p);

// Frames [f0, f0 + n_frames) of the call (f0 even-parity in the global frame count, n_frames
// even).  `x` points at the first sample of the call, `hist` holds the Hlen samples before it.
// With `hist_new` non-null the kernel also writes the object's next state (the last Hlen samples of
// hist ++ x[0 .. n_new)) there, which saves the separate k_update_hist launch.
int32_t firpfbch2_fast_launch(const Firpfbch2FastPlan& p, const float2* hist, long long Hlen, const float2* x,
                              float2* y, size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new = nullptr,
                              long long n_new = 0);

// Fused synthesis (firpfbch2_synth_fast.cu).  `prefix` = the 32 input frames preceding x[0] of the
// call; frames [f0, f0 + n_frames) of the call, f0 on even global parity, n_frames a multiple of 32.
int32_t firpfbch2_synth_fast_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
int32_t firpfbch2_synth_fast_launch(const Firpfbch2FastPlan& p, const float2* prefix, const float2* x, float2* y,
                                    size_t f0, size_t n_frames, cudaStream_t st);

// Small-M fused synthesis (firpfbch2_small_synth.cu, M = 64 / 128): 256 / M time slabs per CTA; same contract as
// firpfbch2_synth_fast_launch.
int32_t firpfbch2_small_synth_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
int32_t firpfbch2_small_synth_launch(const Firpfbch2FastPlan& p, const float2* prefix, const float2* x, float2* y,
                                     size_t f0, size_t n_frames, cudaStream_t st);

// Small-M fused analysis (firpfbch2_small.cu, M = 64): four time slabs per CTA.
int32_t firpfbch2_small_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
int32_t firpfbch2_small_launch(const Firpfbch2FastPlan& p, const float2* hist, long long Hlen, const float2* x, float2* y,
                               size_t f0, size_t n_frames, cudaStream_t st);

// Tiny-M fused analysis (firpfbch2_tiny.cu, M = 8 / 16 / 32): eight (FIR warp, DFT warp) units per CTA; needs a
// 16-byte aligned input and output pointers.
int32_t firpfbch2_tiny_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
int32_t firpfbch2_tiny_launch(const Firpfbch2FastPlan& p, const float2* hist, long long Hlen, const float2* x, float2* y,
                              size_t f0, size_t n_frames, cudaStream_t st);

// Tiny-M fused synthesis (firpfbch2_tiny_synth.cu, M = 8 / 16 / 32); same contract as firpfbch2_synth_fast_launch.
int32_t firpfbch2_tiny_synth_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
int32_t firpfbch2_tiny_synth_launch(const Firpfbch2FastPlan& p, const float2* prefix, const float2* x, float2* y,
                                    size_t f0, size_t n_frames, cudaStream_t st);

// Large-M analysis (firpfbch2_large.cu, M = 512 / 1024 / 2048 / 4096): one fused cooperative kernel (FIR role -> L2 ring
// -> DFT teams) for whole 32-frame batches; FIR stage + in-place FFT stage per L2-sized chunk for the rest.
int32_t firpfbch2_large_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
// With `hist_new` and `hist_done` non-null the kernel that takes the call may also write the object's next state (see
// firpfbch2_fast_launch); *hist_done tells whether it did.
int32_t firpfbch2_large_launch(const Firpfbch2FastPlan& p, const float2* hist, long long Hlen, const float2* x, float2* y,
                               size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new = nullptr, long long n_new = 0,
                               bool* hist_done = nullptr);

// Large-M synthesis (same sizes): one fused cooperative kernel (DFT teams -> L2 ring -> overlap-add role), or, when
// that cannot launch, IFFT stage into an L2-resident scratch + overlap-add stage per chunk.
int32_t firpfbch2_large_synth_plan(Firpfbch2FastPlan& p, uint32_t M, uint32_t m, const float* h);
long long firpfbch2_large_synth_scratch_frames(uint32_t M);
// false when the fused kernel will take the call (its ring lives in the plan) and `scratch` may be null
bool firpfbch2_large_synth_needs_scratch(const Firpfbch2FastPlan& p, const float2* prefix, const float2* x);
// `hist_new` / `hist_done` as in firpfbch2_large_launch (n_new counts input samples, M per frame).
int32_t firpfbch2_large_synth_launch(const Firpfbch2FastPlan& p, const float2* prefix, const float2* x, float2* y,
                                     float2* scratch, size_t f0, size_t n_frames, cudaStream_t st, float2* hist_new = nullptr,
                                     long long n_new = 0, bool* hist_done = nullptr);

}  // namespace yg
