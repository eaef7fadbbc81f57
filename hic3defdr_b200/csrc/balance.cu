// Knight-Ruiz matrix balancing and the sparse-bin filter that precedes it
// (the README's simulate -> balance -> re-run
// loop).
//
// Replaces the iteration of hic3defdr/util/balancing.py:86-174 (kr_balance:
// an inexact Newton method whose inner solver is a conjugate-gradient loop
// around sparse matrix-vector products) and the band counts of
// hic3defdr/util/filtering.py:50-53 (filter_sparse_rows_count).  The symmetric
// matrix stays resident as CSR; every inner step is one SpMV (a warp per row:
// ~2 (k + 1) stored entries per row of a banded contact matrix) plus fused
// vector updates whose dot products / extrema are reduced in the same pass; the
// host reads two or three scalars per step to take the reference's branches.
// HBM-bound: 12 B per stored entry per SpMV.
#include <vector>

#include "common.cuh"

namespace h3d {

// out[i] = sum_j A[i, j] * a[j] * (b ? b[j] : 1)
__global__ void __launch_bounds__(256)
kr_spmv_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
               const double* __restrict__ data, int n, const double* __restrict__ a,
               const double* __restrict__ b, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int rowi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (rowi >= n) return;
    double acc = 0.0;
    for (int k = indptr[rowi] + lane; k < indptr[rowi + 1]; k += 32) {
        const int j = indices[k];
        acc += data[k] * (b ? a[j] * b[j] : a[j]);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[rowi] = acc;
}

struct KrVecs {
    double *x, *v, *rk, *y, *z, *p, *w, *ap, *ynew, *t;
};

enum : int { KR_INIT, KR_FIRST, KR_PUPDATE, KR_W, KR_STEP, KR_GAMMA_LO, KR_GAMMA_HI, KR_YAXPY,
             KR_ACCEPT, KR_OUTER, KR_ONES };

// elementwise phase ``op`` over the vectors; up to two block-reduced results
// (sum in r0 for the dot products; min in r0 / max in r1 for KR_STEP; min in r0
// for the gamma searches) go to partial[2 * block + {0, 1}]
__global__ void __launch_bounds__(256)
kr_vec_kernel(int op, int n, KrVecs q, double s0, double s1, double* __restrict__ partial) {
    __shared__ double sh0[8], sh1[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < n;
    double r0 = 0.0, r1 = 0.0;
    const bool is_min = (op == KR_STEP || op == KR_GAMMA_LO || op == KR_GAMMA_HI);
    if (is_min) { r0 = INFINITY; r1 = -INFINITY; }
    if (on) {
        switch (op) {
            case KR_ONES: q.y[i] = 1.0; break;
            case KR_INIT: {                    // v = x * (A x); rk = 1 - v; sum rk^2
                const double v = q.x[i] * q.t[i];
                q.v[i] = v; q.rk[i] = 1.0 - v; r0 = (1.0 - v) * (1.0 - v);
                break;
            }
            case KR_FIRST: {                   // z = rk / v; p = z; sum rk z
                const double z = q.rk[i] / q.v[i];
                q.z[i] = z; q.p[i] = z; r0 = q.rk[i] * z;
                break;
            }
            case KR_PUPDATE: q.p[i] = q.z[i] + s0 * q.p[i]; break;          // s0 = beta
            case KR_W: {                       // w = x * A(x p) + v p; sum p w
                const double w = q.x[i] * q.t[i] + q.v[i] * q.p[i];
                q.w[i] = w; r0 = q.p[i] * w;
                break;
            }
            case KR_STEP: {                    // ap = alpha p; ynew = y + ap; min, max ynew
                const double ap = s0 * q.p[i];
                const double yn = q.y[i] + ap;
                q.ap[i] = ap; q.ynew[i] = yn; r0 = yn; r1 = yn;
                break;
            }
            case KR_GAMMA_LO:                  // min over ap < 0 of (delta - y) / ap
                if (q.ap[i] < 0.0) r0 = (s0 - q.y[i]) / q.ap[i];
                break;
            case KR_GAMMA_HI:                  // min over ynew > Delta of (Delta - y) / ap
                if (q.ynew[i] > s0) r0 = (s0 - q.y[i]) / q.ap[i];
                break;
            case KR_YAXPY: q.y[i] += s0 * q.ap[i]; break;                   // s0 = gamma
            case KR_ACCEPT: {                  // y = ynew; rk -= alpha w; z = rk / v; sum rk z
                q.y[i] = q.ynew[i];
                const double rk = q.rk[i] - s0 * q.w[i];
                const double z = rk / q.v[i];
                q.rk[i] = rk; q.z[i] = z; r0 = rk * z;
                break;
            }
            case KR_OUTER: q.x[i] *= q.y[i]; break;
        }
    }
    if (partial == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_down_sync(0xffffffffu, r0, o), b = __shfl_down_sync(0xffffffffu, r1, o);
        if (is_min) { r0 = fmin(r0, a); r1 = fmax(r1, b); } else { r0 += a; r1 += b; }
    }
    if ((threadIdx.x & 31) == 0) { sh0[threadIdx.x >> 5] = r0; sh1[threadIdx.x >> 5] = r1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) {
            if (is_min) { r0 = fmin(r0, sh0[k]); r1 = fmax(r1, sh1[k]); } else { r0 += sh0[k]; r1 += sh1[k]; }
        }
        partial[2 * blockIdx.x] = r0; partial[2 * blockIdx.x + 1] = r1;
    }
}

__global__ void __launch_bounds__(256)
kr_final_kernel(const double* __restrict__ partial, int n_blocks, int is_min, double* __restrict__ result) {
    __shared__ double sh0[256], sh1[256];
    double r0 = is_min ? INFINITY : 0.0, r1 = is_min ? -INFINITY : 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) {
        if (is_min) { r0 = fmin(r0, partial[2 * b]); r1 = fmax(r1, partial[2 * b + 1]); }
        else { r0 += partial[2 * b]; r1 += partial[2 * b + 1]; }
    }
    sh0[threadIdx.x] = r0; sh1[threadIdx.x] = r1;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            if (is_min) { sh0[threadIdx.x] = fmin(sh0[threadIdx.x], sh0[threadIdx.x + o]);
                          sh1[threadIdx.x] = fmax(sh1[threadIdx.x], sh1[threadIdx.x + o]); }
            else { sh0[threadIdx.x] += sh0[threadIdx.x + o]; sh1[threadIdx.x] += sh1[threadIdx.x + o]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { result[0] = sh0[0]; result[1] = sh1[0]; }
}

// entries > 0 within k bins upstream (column i, rows i-k .. i-1) and downstream
// (row i, columns i+1 .. i+k) of every bin of an upper-triangular CSR matrix
__global__ void __launch_bounds__(256)
band_nnz_kernel(const long long* __restrict__ indptr, const int* __restrict__ indices,
                const double* __restrict__ data, int n, int k, int* __restrict__ up, int* __restrict__ down) {
    const int lane = threadIdx.x & 31;
    const int rowi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (rowi >= n) return;
    int cnt = 0;
    for (long long e = indptr[rowi] + lane; e < indptr[rowi + 1]; e += 32) {
        const int j = indices[e];
        const int d = j - rowi;
        if (d > 0 && d <= k && data[e] > 0.0) { ++cnt; atomicAdd(&up[j], 1); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0) down[rowi] = cnt;
}

}  // namespace h3d

using namespace h3d;

extern "C" size_t h3d_kr_balance_ws_bytes(int n) {
    return 10 * ws_pad((size_t)n * 8) + ws_pad((size_t)(n / 256 + 2) * 16) + ws_pad(64);
}

extern "C" int h3d_kr_balance(const int* indptr, const int* indices, const double* data, int n,
                              double tol, const double* x0, double delta, double ddelta, int max_iter,
                              double* x_out, double* res_host, int res_cap, int* n_res_host,
                              int* n_matvec_host, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n >= 1 && tol > 0.0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w(ws, ws_bytes);
    KrVecs q;
    q.x = x_out;
    q.v = w.take<double>(n); q.rk = w.take<double>(n); q.y = w.take<double>(n); q.z = w.take<double>(n);
    q.p = w.take<double>(n); q.w = w.take<double>(n); q.ap = w.take<double>(n); q.ynew = w.take<double>(n);
    q.t = w.take<double>(n);
    const int blocks = div_up(n, 256);
    double* partial = w.take<double>((size_t)blocks * 2);
    double* result = w.take<double>(2);
    if (!q.v || !q.rk || !q.y || !q.z || !q.p || !q.w || !q.ap || !q.ynew || !q.t || !partial || !result) {
        set_error("kr_balance workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    const int sgrid = div_up((long long)n * 32, 256);
    double h[2];
    int matvecs = 0;
    // elementwise phase (+ reduction read back into h[0], h[1])
    auto vec = [&](int op, double s0, bool reduce, bool is_min) -> int {
        kr_vec_kernel<<<blocks, 256, 0, st>>>(op, n, q, s0, 0.0, reduce ? partial : nullptr);
        H3D_LAUNCHED("kr_vec_kernel");
        if (reduce) {
            kr_final_kernel<<<1, 256, 0, st>>>(partial, blocks, is_min ? 1 : 0, result);
            H3D_LAUNCHED("kr_final_kernel");
            H3D_CHECK(cudaMemcpyAsync(h, result, 16, cudaMemcpyDeviceToHost, st));
            H3D_CHECK(cudaStreamSynchronize(st));
        }
        return H3D_OK;
    };
    auto spmv = [&](const double* a, const double* b) -> int {
        kr_spmv_kernel<<<sgrid, 256, 0, st>>>(indptr, indices, data, n, a, b, q.t);
        H3D_LAUNCHED("kr_spmv_kernel");
        ++matvecs;
        return H3D_OK;
    };
#define KR(call) do { int rc_ = (call); if (rc_) return rc_; } while (0)
    if (x0) H3D_CHECK(cudaMemcpyAsync(q.x, x0, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    else { KrVecs o = q; o.y = q.x; kr_vec_kernel<<<blocks, 256, 0, st>>>(KR_ONES, n, o, 0.0, 0.0, nullptr); H3D_LAUNCHED("kr_vec_kernel"); }
    // balancing.py:100-118
    const double g = 0.9, eta_max = 0.1, stop_tol = tol * 0.5, rt = tol * tol;
    double eta = eta_max;
    KR(spmv(q.x, nullptr));
    KR(vec(KR_INIT, 0.0, true, false));
    double rho_km1 = h[0], rout = rho_km1, rold = rout;
    int i = 0, n_res = 0;
    while (rout > rt) {
        if (max_iter >= 0 && i > max_iter) break;
        ++i;
        int k = 0;
        KR(vec(KR_ONES, 0.0, false, false));
        const double innertol = fmax(eta * eta * rout, rt);
        double rho_km2 = 0.0;
        while (rho_km1 > innertol) {
            ++k;
            if (k == 1) {
                KR(vec(KR_FIRST, 0.0, true, false));
                rho_km1 = h[0];
            } else {
                KR(vec(KR_PUPDATE, rho_km1 / rho_km2, false, false));
            }
            KR(spmv(q.x, q.p));
            KR(vec(KR_W, 0.0, true, false));
            const double alpha = rho_km1 / h[0];
            KR(vec(KR_STEP, alpha, true, true));
            const double ymin = h[0], ymax = h[1];
            if (ymin <= delta) {
                if (delta == 0.0) break;
                KR(vec(KR_GAMMA_LO, delta, true, true));
                KR(vec(KR_YAXPY, h[0], false, false));
                break;
            }
            if (ymax >= ddelta) {
                KR(vec(KR_GAMMA_HI, ddelta, true, true));
                KR(vec(KR_YAXPY, h[0], false, false));
                break;
            }
            KR(vec(KR_ACCEPT, alpha, true, false));
            rho_km2 = rho_km1;
            rho_km1 = h[0];
        }
        KR(vec(KR_OUTER, 0.0, false, false));
        KR(spmv(q.x, nullptr));
        KR(vec(KR_INIT, 0.0, true, false));
        rho_km1 = h[0];
        rout = rho_km1;
        const double rat = rout / rold;
        rold = rout;
        const double res_norm = sqrt(rout);
        const double eta_0 = eta;
        eta = g * rat;
        if (g * eta_0 * eta_0 > 0.1) eta = fmax(eta, g * eta_0 * eta_0);
        eta = fmax(fmin(eta, eta_max), stop_tol / res_norm);
        if (res_host && n_res < res_cap) res_host[n_res] = res_norm;
        ++n_res;
    }
#undef KR
    if (n_res_host) *n_res_host = n_res;
    if (n_matvec_host) *n_matvec_host = matvecs;
    H3D_CHECK(cudaStreamSynchronize(st));
    return H3D_OK;
}

extern "C" int h3d_band_nnz(const long long* indptr, const int* indices, const double* data, int n,
                            int k, int* upstream, int* downstream, h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    H3D_CHECK(cudaMemsetAsync(upstream, 0, (size_t)n * 4, st));
    band_nnz_kernel<<<div_up((long long)n * 32, 256), 256, 0, st>>>(indptr, indices, data, n, k, upstream,
                                                                   downstream);
    H3D_LAUNCHED("band_nnz_kernel");
    return H3D_OK;
}
