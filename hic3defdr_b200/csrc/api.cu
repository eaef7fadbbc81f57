// Library-level entry points of libh3d: version, error text, launch counter.
#include <atomic>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace h3d {

static thread_local char g_error[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n); }

}  // namespace h3d

namespace h3d {

// register-resident FP64 FMA chains: 8 independent accumulators per thread
__global__ void __launch_bounds__(256)
fp64_peak_kernel(double* out, int iters, double a, double b) {
    double v0 = threadIdx.x * 1e-3, v1 = v0 + 1.0, v2 = v0 + 2.0, v3 = v0 + 3.0;
    double v4 = v0 + 4.0, v5 = v0 + 5.0, v6 = v0 + 6.0, v7 = v0 + 7.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
            v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
        }
    }
    const double s = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (s == 123.456) out[0] = s;      // keeps the chains alive
}

// small device -> host-visible copy done by the SMs (see h3d_publish)
__global__ void publish_kernel(const unsigned long long* __restrict__ src,
                               volatile unsigned long long* __restrict__ dst, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
}

}  // namespace h3d

extern "C" {

// Copies ``nbytes`` (a multiple of 8, at most 64 KiB) from device memory to
// pinned, device-mapped host memory with a kernel instead of the copy engine:
// small control read-backs (pixel counts, convergence counters) then never
// queue behind the bulk device->host output copies that share the DMA engine.
// The data are valid on the host after the stream is synchronised.
int h3d_publish(const void* dev_src, void* host_mapped_dst, size_t nbytes, h3d_stream_t stream) {
    H3D_REQUIRE(nbytes % 8 == 0 && nbytes <= 65536, "publish size must be a multiple of 8, <= 64 KiB");
    if (nbytes == 0) return H3D_OK;
    h3d::publish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)dev_src, (volatile unsigned long long*)host_mapped_dst,
        (int)(nbytes / 8));
    H3D_LAUNCHED("publish_kernel");
    return H3D_OK;
}

int h3d_version(void) { return 100; }

// Measures the FP64 FMA issue rate of the current device (the roofline
// denominator of the FP64-bound kernels; MEASURED_PEAKS.json has no FP64
// entry).  Returns TFLOP/s (FMA = 2 flops) in *tflops_out.  Synchronises.
int h3d_fp64_peak(double* scratch, double* tflops_out, h3d_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    H3D_CHECK(cudaEventCreate(&e0));
    H3D_CHECK(cudaEventCreate(&e1));
    const int blocks = h3d::kNumSMs * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        H3D_CHECK(cudaEventRecord(e0, st));
        h3d::fp64_peak_kernel<<<blocks, 256, 0, st>>>(scratch, iters, 0.999999, 1e-6);
        H3D_LAUNCHED("fp64_peak_kernel");
        H3D_CHECK(cudaEventRecord(e1, st));
        H3D_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        H3D_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)blocks;
        const double tf = flops / ((double)ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_out = best;
    return H3D_OK;
}


const char* h3d_last_error(void) { return h3d::g_error; }

unsigned long long h3d_launch_count(void) { return h3d::g_launches.load(); }

void h3d_reset_launch_count(void) { h3d::g_launches.store(0); }

}  // extern "C"
