// Robust locally weighted regression (lowess) of the distance -> dispersion
// trend, one thread block.
//
// Replaces lib5c.util.lowess.lowess as called from
// hic3defdr/util/lowess.py:72 (lib5c 0.6.0 is a port of statsmodels'
// _smoothers_lowess.pyx; neither is vendored, the algorithm is Cleveland's:
// k nearest neighbours, tricube kernel, local linear fit, ``it`` bisquare
// robustifying passes, ``delta`` skipping with linear interpolation).
// The problem is tiny (at most a few thousand duplicated points): one block;
// the anchor sequence is planned once, the local regressions of a pass run
// one warp per anchor, the median of the residuals by rank counting.
#include "common.cuh"

namespace h3d {

constexpr int kLwThreads = 512;

// Which points get their own local regression ("anchors") depends on x alone
// (the delta rule and the ties), not on y or on the robustness weights.  The
// kernel therefore walks the anchor sequence once (serially, thread 0), and
// every robustifying pass then fits all anchors in parallel, one warp per
// anchor, and fills the skipped / tied points from the recorded plan.
struct LowessPlan {
    int* anchor;        // [3 * n]: i, left, right of every anchor
    int* a0;            // [n] anchor supplying the value (or the left end of the interpolation)
    int* a1;            // [n] anchor at the right end of the interpolation, -1: copy a0
    double* al;         // [n] interpolation weight of a1
};

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one smoothing problem of a batch: its slice [off, off + n) of the packed
// arrays (x, y in; y_fit out; scratch of the same layout)
struct LowessJob { int off, n, k_nn, n_pass; double delta; };

__global__ void __launch_bounds__(kLwThreads)
lowess_kernel(const double* __restrict__ x_all, const double* __restrict__ y_all,
              const LowessJob* __restrict__ jobs, double* __restrict__ y_fit_all,
              double* __restrict__ resid_w_all, double* __restrict__ resid_all,
              double* __restrict__ fitv_all, LowessPlan plan_all) {
    const LowessJob job = jobs[blockIdx.x];
    const int n = job.n, k_nn = job.k_nn, n_pass = job.n_pass;
    const double delta = job.delta;
    const double* __restrict__ x = x_all + job.off;
    const double* __restrict__ y = y_all + job.off;
    double* __restrict__ y_fit = y_fit_all + job.off;
    double* __restrict__ resid_w = resid_w_all + job.off;
    double* __restrict__ resid = resid_all + job.off;
    double* __restrict__ fitv = fitv_all + job.off;
    LowessPlan plan;
    plan.anchor = plan_all.anchor + 3 * job.off;
    plan.a0 = plan_all.a0 + job.off;
    plan.a1 = plan_all.a1 + job.off;
    plan.al = plan_all.al + job.off;
    __shared__ int s_n_anchor;
    __shared__ double s_med[2];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    for (int j = t; j < n; j += kLwThreads) resid_w[j] = 1.0;
    if (t == 0) {
        // the anchor walk of Cleveland's lowess with delta skipping
        int i = 0, last = -1, left = 0, right = k_nn, na = 0;
        while (true) {
            const double xi = x[i];
            while (right < n && (xi - x[left]) > (x[right] - xi)) { ++left; ++right; }
            plan.anchor[3 * na] = i; plan.anchor[3 * na + 1] = left; plan.anchor[3 * na + 2] = right;
            if (last < i - 1) {                      // points skipped because of delta: interpolate
                const double a = xi - x[last];
                const int prev = plan.a0[last];      // the anchor whose value y_fit[last] holds
                for (int j = last + 1; j < i; ++j) {
                    plan.a0[j] = prev; plan.a1[j] = na;
                    plan.al[j] = (x[j] - x[last]) / a;
                }
            }
            plan.a0[i] = na; plan.a1[i] = -1;
            last = i;
            const double cut = x[last] + delta;
            int kk;
            bool stopped = false;
            for (kk = last + 1; kk < n; ++kk) {
                if (x[kk] > cut) { stopped = true; break; }
                if (x[kk] == x[last]) { plan.a0[kk] = na; plan.a1[kk] = -1; last = kk; }   // ties copy the fit
            }
            if (!stopped) kk = n - 1;
            ++na;
            if (last >= n - 1) break;
            i = (kk - 1 > last + 1) ? kk - 1 : last + 1;
        }
        s_n_anchor = na;
    }
    __syncthreads();
    const int n_anchor = s_n_anchor;
    for (int pass = 0; pass < n_pass; ++pass) {
        for (int a = wid; a < n_anchor; a += kLwThreads / 32) {
            const int i = plan.anchor[3 * a], left = plan.anchor[3 * a + 1], right = plan.anchor[3 * a + 2];
            const double xi = x[i];
            const double radius = fmax(xi - x[left], x[right - 1] - xi);
            // tricube weights times robustness weights
            double part = 0.0;
            for (int j = left + lane; j < right; j += 32) {
                const double u = fabs(x[j] - xi) / radius;
                const double c = 1.0 - u * u * u;
                part += c * c * c * resid_w[j];
            }
            const double sw = warp_sum_all(part);
            double fit;
            if (!(sw > 0.0)) {
                fit = y[i];
            } else {
                double p1 = 0.0;
                for (int j = left + lane; j < right; j += 32) {
                    const double u = fabs(x[j] - xi) / radius;
                    const double c = 1.0 - u * u * u;
                    p1 += (c * c * c * resid_w[j] / sw) * x[j];
                }
                const double xbar = warp_sum_all(p1);
                double p2 = 0.0;
                for (int j = left + lane; j < right; j += 32) {
                    const double u = fabs(x[j] - xi) / radius;
                    const double c = 1.0 - u * u * u;
                    const double dx = x[j] - xbar;
                    p2 += (c * c * c * resid_w[j] / sw) * dx * dx;
                }
                const double sq = warp_sum_all(p2);
                double p3 = 0.0;
                for (int j = left + lane; j < right; j += 32) {
                    const double u = fabs(x[j] - xi) / radius;
                    const double c = 1.0 - u * u * u;
                    const double w = c * c * c * resid_w[j] / sw;
                    p3 += w * (1.0 + (xi - xbar) * (x[j] - xbar) / sq) * y[j];
                }
                fit = warp_sum_all(p3);
            }
            if (lane == 0) fitv[a] = fit;
        }
        __syncthreads();
        for (int j = t; j < n; j += kLwThreads) {
            const int b = plan.a1[j];
            const double v0 = fitv[plan.a0[j]];
            y_fit[j] = (b < 0) ? v0 : plan.al[j] * fitv[b] + (1.0 - plan.al[j]) * v0;
        }
        __syncthreads();
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

extern "C" size_t h3d_lowess_ws_bytes(int n) { return 4 * ws_pad((size_t)n * 8) + ws_pad((size_t)n * 12) + 2 * ws_pad((size_t)n * 4) + 256; }

static int lowess_launch(const double* x, const double* y, int n_total, const LowessJob* jobs_dev,
                         int n_jobs, double* y_fit, void* ws, size_t ws_bytes, cudaStream_t st) {
    Workspace w(ws, ws_bytes);
    double* resid_w = w.take<double>(n_total);
    double* resid = w.take<double>(n_total);
    double* fitv = w.take<double>(n_total);
    LowessPlan plan;
    plan.al = w.take<double>(n_total);
    plan.anchor = w.take<int>((size_t)3 * n_total);
    plan.a0 = w.take<int>(n_total);
    plan.a1 = w.take<int>(n_total);
    if (!resid_w || !resid || !fitv || !plan.al || !plan.anchor || !plan.a0 || !plan.a1) {
        set_error("lowess workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    lowess_kernel<<<n_jobs, kLwThreads, 0, st>>>(x, y, jobs_dev, y_fit, resid_w, resid, fitv, plan);
    H3D_LAUNCHED("lowess_kernel");
    return H3D_OK;
}

static int lowess_k(double frac, int n) {
    int k = (int)(frac * (double)n + 1e-10);
    if (k < 2) k = 2;
    if (k > n) k = n;
    return k;
}

extern "C" int h3d_lowess(const double* x, const double* y, int n, double frac, int it, double delta,
                          double* y_fit, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n >= 2, "lowess needs at least two points");
    H3D_REQUIRE(it >= 0, "negative number of robustifying iterations");
    cudaStream_t st = (cudaStream_t)stream;
    // the job descriptor lives at the end of the workspace
    H3D_REQUIRE(ws_bytes >= h3d_lowess_ws_bytes(n), "lowess workspace too small");
    LowessJob job = {0, n, lowess_k(frac, n), it + 1, delta};
    LowessJob* job_dev = (LowessJob*)((char*)ws + h3d_lowess_ws_bytes(n) - 256);
    H3D_CHECK(cudaMemcpyAsync(job_dev, &job, sizeof(job), cudaMemcpyHostToDevice, st));
    H3D_CHECK(cudaStreamSynchronize(st));      // the descriptor is on the stack
    return lowess_launch(x, y, n, job_dev, 1, y_fit, ws, ws_bytes - 256, st);
}

// n_jobs independent smoothing problems in one launch (one thread block each):
// problem j is the slice [off[j], off[j] + n[j]) of the packed x / y / y_fit.
// off_host, n_host, frac_host, delta_host: HOST arrays.
extern "C" size_t h3d_lowess_batch_ws_bytes(int n_total, int n_jobs) {
    return h3d_lowess_ws_bytes(n_total) + ws_pad((size_t)n_jobs * sizeof(LowessJob));
}

extern "C" int h3d_lowess_batch(const double* x, const double* y, const int* off_host,
                                const int* n_host, const double* frac_host, int it,
                                const double* delta_host, int n_jobs, double* y_fit, void* ws,
                                size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n_jobs >= 1 && n_jobs <= 64 && it >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    LowessJob jobs[64];
    int n_total = 0;
    for (int j = 0; j < n_jobs; ++j) {
        H3D_REQUIRE(n_host[j] >= 2, "lowess needs at least two points");
        jobs[j] = {off_host[j], n_host[j], lowess_k(frac_host[j], n_host[j]), it + 1, delta_host[j]};
        if (off_host[j] + n_host[j] > n_total) n_total = off_host[j] + n_host[j];
    }
    H3D_REQUIRE(ws_bytes >= h3d_lowess_batch_ws_bytes(n_total, n_jobs), "lowess workspace too small");
    const size_t main_bytes = h3d_lowess_ws_bytes(n_total) - 256;
    LowessJob* jobs_dev = (LowessJob*)((char*)ws + main_bytes);
    H3D_CHECK(cudaMemcpyAsync(jobs_dev, jobs, (size_t)n_jobs * sizeof(LowessJob), cudaMemcpyHostToDevice, st));
    H3D_CHECK(cudaStreamSynchronize(st));      // the descriptors are on the stack
    return lowess_launch(x, y, n_total, jobs_dev, n_jobs, y_fit, ws, main_bytes, st);
}
