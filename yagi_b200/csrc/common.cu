// common.cu -- error plumbing, host-side Kaiser design, pinned-memory helpers.
#include "common.cuh"

#include <atomic>
#include <cmath>

namespace yg {

std::string& last_error_ref()
{
    thread_local std::string e;
    return e;
}

int32_t fail(int32_t code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

int32_t require_device(int* dev_out)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(YG_EINTERNAL, "no CUDA device available (%s); yagi_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    int dev = 0;
    YG_CUDA(cudaGetDevice(&dev));
    *dev_out = dev;
    return YG_OK;
}

int sm_count(int dev)
{
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = 1;
    }
    return n;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

cudaError_t memcpy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, cudaStreamPerThread);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamPerThread);
}

cudaError_t memset_sync(void* dst, int value, size_t bytes)
{
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dst, value, bytes, cudaStreamPerThread);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamPerThread);
}

int32_t HostPipe::init()
{
    if (inited) return YG_OK;
    YG_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    YG_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
        YG_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
        YG_CUDA(cudaEventCreateWithFlags(&ev_comp[b], cudaEventDisableTiming));
        YG_CUDA(cudaEventCreateWithFlags(&ev_out[b], cudaEventDisableTiming));
    }
    inited = true;
    return YG_OK;
}

void HostPipe::destroy()
{
    if (!inited) return;
    cudaStreamDestroy(s_in);
    cudaStreamDestroy(s_out);
    for (int b = 0; b < 2; b++) {
        cudaEventDestroy(ev_in[b]);
        cudaEventDestroy(ev_comp[b]);
        cudaEventDestroy(ev_out[b]);
        dx[b].release();
        dy[b].release();
    }
    inited = false;
}

// ------------------------------------------------------------------ Kaiser design
// Host-side, cold path.  Follows the reference's f32 formulae so that taps designed here
// equal the taps a yagi caller would have designed:
//   lngammaf   src/math/gamma.rs:7-22        lnbesselif / besseli0f  src/math/bessel.rs:9-67
//   sincf      src/math/mod.rs:63-69         kaiser window           src/math/windows.rs:76-90
//   beta(As)   src/filter/fir/design/kaiser.rs:62-72     design loop  kaiser.rs:16-51
static float lngammaf_(float z)
{
    if (z < 10.0f) return lngammaf_(z + 1.0f) - logf(z);        // gamma.rs:10-15 (z > 0 guaranteed by callers)
    const float pi = 3.14159265358979323846f;
    float g = 0.5f * (logf(2.0f * pi) - logf(z));
    g += z * (logf(z + (1.0f / (12.0f * z - 0.1f / z))) - 1.0f);
    return g;
}

static float besseli0f_(float z)
{
    if (z == 0.0f) return 1.0f;
    if (z < 1e-3f) return 1.0f / expf(lngammaf_(1.0f));
    const float lz = logf(0.5f * z);
    float y = 0.0f;
    for (int k = 0; k < 64; k++) {
        const float t1 = 2.0f * (float)k * lz;
        const float t2 = lngammaf_((float)k + 1.0f);
        y += expf(t1 - t2 - t2);
    }
    return expf(logf(y));
}

static float sincf_(float x)
{
    const float pi = 3.14159265358979323846f;
    if (fabsf(x) < 0.01f) return cosf(pi * x / 2.0f) * cosf(pi * x / 4.0f) * cosf(pi * x / 8.0f);
    return sinf(pi * x) / (pi * x);
}

int32_t fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h)
{
    if (mu <= -0.5f || mu > 0.5f)
        return fail(YG_ECONFIG, "fractional sample offset (%g) out of range (-0.5, 0.5)", (double)mu);
    if (fc <= 0.0f || fc > 0.5f) return fail(YG_ECONFIG, "cutoff frequency (%g) out of range (0, 0.5)", (double)fc);
    if (n == 0) return fail(YG_ECONFIG, "filter length must be greater than zero");
    if (as <= 0.0f) return fail(YG_ECONFIG, "stop-band attenuation must be greater than zero");
    if (h == nullptr) return fail(YG_EVALUE, "null output pointer");

    const float a = fabsf(as);
    float beta = 0.0f;
    if (a > 50.0f) beta = 0.1102f * (a - 8.7f);
    else if (a > 21.0f) beta = 0.5842f * powf(a - 21.0f, 0.4f) + 0.07886f * (a - 21.0f);
    const float i0b = besseli0f_(beta);

    for (uint32_t i = 0; i < n; i++) {
        const float t = (float)i - ((float)n - 1.0f) / 2.0f + mu;
        const float h1 = sincf_(2.0f * fc * t);
        const float tw = (float)i - (float)(n - 1) / 2.0f;
        const float r = 2.0f * tw / (float)(n - 1);
        const float h2 = besseli0f_(beta * sqrtf(1.0f - r * r)) / i0b;
        h[i] = h1 * h2;
    }
    return YG_OK;
}

bool plan_radices(uint32_t M, TiledPass& tp)
{
    tp = TiledPass{};
    for (uint32_t rem = M; rem > 1;) {
        uint32_t r = rem;
        if ((rem & 3) == 0) r = 4;
        else if ((rem & 1) == 0) r = 2;
        else
            for (uint32_t f = 3; f * f <= rem; f += 2)
                if (rem % f == 0) { r = f; break; }
        if (r > 255 || tp.n_pass >= 24) return false;
        tp.radix[tp.n_pass++] = (unsigned char)r;
        rem /= r;
    }
    return true;
}

void make_twiddles(uint32_t M, std::vector<float2>& tw)
{
    tw.resize(M);
    for (uint32_t k = 0; k < M; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)M;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
}

}  // namespace yg

// ------------------------------------------------------------------ C ABI: library
extern "C" {

int32_t yg_version(void) { return 0x000100; }

const char* yg_last_error(void) { return yg::last_error_ref().c_str(); }

int32_t yg_device_count(int32_t* n)
{
    if (!n) return yg::fail(YG_EVALUE, "null pointer");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { cudaGetLastError(); c = 0; }
    *n = c;
    return YG_OK;
}

int32_t yg_launch_count(uint64_t* n)
{
    if (!n) return yg::fail(YG_EVALUE, "null pointer");
    *n = (uint64_t)yg::launch_count();
    return YG_OK;
}

int32_t yg_host_alloc(void** p, size_t bytes)
{
    if (!p) return yg::fail(YG_EVALUE, "null pointer");
    YG_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return YG_OK;
}

int32_t yg_host_free(void* p)
{
    if (!p) return YG_OK;
    YG_CUDA(cudaFreeHost(p));
    return YG_OK;
}

int32_t yg_fir_design_kaiser(uint32_t n, float fc, float as, float mu, float* h)
{
    return yg::fir_design_kaiser(n, fc, as, mu, h);
}

}  // extern "C"
