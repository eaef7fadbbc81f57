// Robust locally weighted regression (lowess) of the distance -> dispersion
// trend, one thread block.
//
// Replaces lib5c.util.lowess.lowess as called from
// hic3defdr/util/lowess.py:72 (lib5c 0.6.0 is a port of statsmodels'
// _smoothers_lowess.pyx; neither is vendored, the algorithm is Cleveland's:
// k nearest neighbours, tricube kernel, local linear fit, ``it`` bisquare
// robustifying passes, ``delta`` skipping with linear interpolation).
// The problem is tiny (at most a few thousand duplicated points), the anchor
// sequence is data dependent and serial; the block parallelises the sums
// inside each local regression and the median of the residuals.
#include "common.cuh"

namespace h3d {

constexpr int kLwThreads = 512;

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
    // deterministic: warp shuffle tree then thread 0 adds the warp sums in order
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kLwThreads / 32; ++w) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(kLwThreads)
lowess_kernel(const double* __restrict__ x, const double* __restrict__ y, int n, int k_nn, int n_pass,
              double delta, double* __restrict__ y_fit, double* __restrict__ resid_w,
              double* __restrict__ resid, double* __restrict__ wbuf) {
    __shared__ double sh[kLwThreads / 32];
    __shared__ int s_i, s_last, s_left, s_right, s_done;
    __shared__ double s_med[2];
    const int t = threadIdx.x;
    for (int j = t; j < n; j += kLwThreads) resid_w[j] = 1.0;
    __syncthreads();
    for (int pass = 0; pass < n_pass; ++pass) {
        for (int j = t; j < n; j += kLwThreads) y_fit[j] = 0.0;
        if (t == 0) { s_i = 0; s_last = -1; s_left = 0; s_right = k_nn; s_done = 0; }
        __syncthreads();
        while (!s_done) {
            if (t == 0) {
                int l = s_left, r = s_right;
                const double xi = x[s_i];
                while (r < n && (xi - x[l]) > (x[r] - xi)) { ++l; ++r; }
                s_left = l; s_right = r;
            }
            __syncthreads();
            const int i = s_i, left = s_left, right = s_right;
            const double xi = x[i];
            const double radius = fmax(xi - x[left], x[right - 1] - xi);
            // tricube weights times robustness weights
            double part = 0.0;
            for (int j = left + t; j < right; j += kLwThreads) {
                const double u = fabs(x[j] - xi) / radius;
                const double c = 1.0 - u * u * u;
                const double w = c * c * c * resid_w[j];
                wbuf[j] = w;
                part += w;
            }
            const double sw = block_reduce_sum(part, sh);
            double fit;
            if (!(sw > 0.0)) {
                fit = y[i];
            } else {
                part = 0.0;
                for (int j = left + t; j < right; j += kLwThreads) {
                    const double w = wbuf[j] / sw;
                    wbuf[j] = w;
                    part += w * x[j];
                }
                const double xbar = block_reduce_sum(part, sh);
                part = 0.0;
                for (int j = left + t; j < right; j += kLwThreads) {
                    const double dx = x[j] - xbar;
                    part += wbuf[j] * dx * dx;
                }
                const double sq = block_reduce_sum(part, sh);
                part = 0.0;
                for (int j = left + t; j < right; j += kLwThreads) {
                    const double p = wbuf[j] * (1.0 + (xi - xbar) * (x[j] - xbar) / sq);
                    part += p * y[j];
                }
                fit = block_reduce_sum(part, sh);
            }
            __syncthreads();
            if (t == 0) {
                y_fit[i] = fit;
                int last = s_last;
                if (last < i - 1) {                      // anchors skipped because of delta
                    const double a = xi - x[last];
                    for (int j = last + 1; j < i; ++j) {
                        const double al = (x[j] - x[last]) / a;
                        y_fit[j] = al * fit + (1.0 - al) * y_fit[last];
                    }
                }
                last = i;
                const double cut = x[last] + delta;
                int kk = last + 1;
                bool stopped = false;
                for (kk = last + 1; kk < n; ++kk) {
                    if (x[kk] > cut) { stopped = true; break; }
                    if (x[kk] == x[last]) { y_fit[kk] = y_fit[last]; last = kk; }
                }
                if (!stopped) kk = n - 1;
                const int nxt = (kk - 1 > last + 1) ? kk - 1 : last + 1;
                s_last = last;
                s_i = nxt;
                if (last >= n - 1) s_done = 1;
            }
            __syncthreads();
        }
        if (pass < n_pass - 1) {
            // bisquare robustness weights from |resid| / (6 median|resid|)
            for (int j = t; j < n; j += kLwThreads) resid[j] = fabs(y[j] - y_fit[j]);
            __syncthreads();
            const int k1 = (n - 1) / 2, k2 = n / 2;
            for (int j = t; j < n; j += kLwThreads) {
                const double v = resid[j];
                int lt = 0, le = 0;
                for (int m = 0; m < n; ++m) { const double u = resid[m]; lt += (u < v); le += (u <= v); }
                if (lt <= k1 && k1 < le) s_med[0] = v;
                if (lt <= k2 && k2 < le) s_med[1] = v;
            }
            __syncthreads();
            const double med = (s_med[0] + s_med[1]) / 2.0;
            for (int j = t; j < n; j += kLwThreads) {
                double r = resid[j];
                if (med == 0.0) r = (r > 0.0) ? 1.0 : 0.0;
                else r = r / (6.0 * med);
                if (r >= 1.0) r = 1.0;
                const double b = 1.0 - r * r;
                resid_w[j] = b * b;
            }
            __syncthreads();
        }
    }
}

}  // namespace h3d

using namespace h3d;

extern "C" size_t h3d_lowess_ws_bytes(int n) { return 3 * ws_pad((size_t)n * 8); }

extern "C" int h3d_lowess(const double* x, const double* y, int n, double frac, int it, double delta,
                          double* y_fit, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n >= 2, "lowess needs at least two points");
    H3D_REQUIRE(it >= 0, "negative number of robustifying iterations");
    Workspace w(ws, ws_bytes);
    double* resid_w = w.take<double>(n);
    double* resid = w.take<double>(n);
    double* wbuf = w.take<double>(n);
    if (!resid_w || !resid || !wbuf) { set_error("lowess workspace too small"); return H3D_ERR_WORKSPACE; }
    int k = (int)(frac * (double)n + 1e-10);
    if (k < 2) k = 2;
    if (k > n) k = n;
    lowess_kernel<<<1, kLwThreads, 0, (cudaStream_t)stream>>>(x, y, n, k, it + 1, delta, y_fit, resid_w,
                                                             resid, wbuf);
    H3D_LAUNCHED("lowess_kernel");
    return H3D_OK;
}
