// Host-side unpacking of outputs that travel over PCIe in a compact form.
//
// Three of the reference's per-pixel output arrays are redundant on the wire:
// ``size_factors`` (N, R) and ``disp`` (N_d, C) are pure functions of the
// pixel's distance (hic3defdr/util/scaling.py:92-104 interpolates a per-distance
// curve; analysis/analysis.py:218 evaluates disp_fn(dist)), and ``raw`` (N, R)
// is int64 (analysis/analysis.py:92-96) holding counts that fit 32 bits.  The
// end-to-end path therefore ships the (D + 1) x R / (D + 1) x C tables and an
// int32 copy of ``raw`` (2.6 of 7.6 GB less for the mouse genome) and
// rebuilds the arrays the drop-in contract promises -- same names, shapes,
// dtypes, values -- in the caller's host buffers with a few host threads, while
// the GPU is busy with the dispersion estimate.  Plain C++ threads; no device
// code in this file.
#include <stdint.h>
#include <string.h>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

template <typename F>
void parallel_ranges(long long n, int n_threads, F fn) {
    if (n_threads < 1) n_threads = 1;
    if (n < (1 << 16)) n_threads = 1;
    if (n_threads == 1) { fn(0, n, 0); return; }
    std::vector<std::thread> pool;
    const long long per = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        const long long lo = t * per, hi = (lo + per < n) ? lo + per : n;
        if (lo >= hi) break;
        pool.emplace_back(fn, lo, hi, t);
    }
    for (auto& th : pool) th.join();
}

}  // namespace

// out[k, :] = table[col[i] - row[i], :] for the pixels i with mask[i] != 0 (all
// pixels when mask is NULL), k = rank of i among them.  All pointers HOST.
extern "C" int h3d_host_expand_by_distance(const double* table, int n_rows, int n_cols, const int* row,
                                           const int* col, const unsigned char* mask, long long n,
                                           double* out, int n_threads) {
    H3D_REQUIRE(n_rows >= 1 && n_cols >= 1 && n >= 0, "bad arguments");
    if (n == 0) return H3D_OK;
    if (n_threads < 1) n_threads = 1;
    std::vector<long long> start(n_threads + 1, 0);
    if (mask) {
        std::vector<long long> cnt(n_threads, 0);
        parallel_ranges(n, n_threads, [&](long long lo, long long hi, int t) {
            long long c = 0;
            for (long long i = lo; i < hi; ++i) c += mask[i] != 0;
            cnt[t] = c;
        });
        for (int t = 0; t < n_threads; ++t) start[t + 1] = start[t] + cnt[t];
    }
    int bad = 0;
    parallel_ranges(n, n_threads, [&](long long lo, long long hi, int t) {
        long long k = mask ? start[t] : lo;
        for (long long i = lo; i < hi; ++i) {
            if (mask && !mask[i]) continue;
            const int d = col[i] - row[i];
            if (d < 0 || d >= n_rows) { bad = 1; ++k; continue; }
            const double* src = table + (long long)d * n_cols;
            double* dst = out + k * n_cols;
            for (int c = 0; c < n_cols; ++c) dst[c] = src[c];
            ++k;
        }
    });
    if (bad) { h3d::set_error("pixel distance outside the table"); return H3D_ERR_ARG; }
    return H3D_OK;
}

extern "C" int h3d_host_widen_i32(const int* in, long long* out, long long n, int n_threads) {
    if (n <= 0) return H3D_OK;
    parallel_ranges(n, n_threads, [&](long long lo, long long hi, int) {
        for (long long i = lo; i < hi; ++i) out[i] = (long long)in[i];
    });
    return H3D_OK;
}

namespace h3d {
__global__ void __launch_bounds__(256)
narrow_i64_kernel(const long long* __restrict__ in, long long n, int* __restrict__ out,
                  int* __restrict__ overflow) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long v = in[i];
    out[i] = (int)v;
    if (v != (long long)(int)v) *overflow = 1;
}
}  // namespace h3d

// int64 -> int32 on the device; *overflow (device int, zeroed by the caller) is
// set when a value does not fit
extern "C" int h3d_narrow_i64(const long long* in, long long n, int* out, int* overflow,
                              h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    h3d::narrow_i64_kernel<<<h3d::div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, out, overflow);
    H3D_LAUNCHED("narrow_i64_kernel");
    return H3D_OK;
}
