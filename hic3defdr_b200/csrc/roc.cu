// ROC / FDR-control curves of q-values against ground-truth labels.
//
// Replaces sklearn.metrics.roc_curve as called by
// hic3defdr/util/evaluation.py:44-79 (evaluate) from
// hic3defdr/analysis/simulation.py:146-239: the q-values are sorted once
// (descending score 1 - q, the one-sweep radix sort of sort.cu), thresholds sit
// at the distinct scores, true / false positive counts are prefix counts of
// the labels in sorted order, and only the corners of the curve (points where
// a second difference of either count is non-zero) are kept.  All counts are
// integers: the curve is exactly sklearn's.  HBM-bound; 38.7 M q-values in a
// few milliseconds against seconds for the argsort + cumsum on the host.
#include "common.cuh"

namespace h3d {

size_t sort_pairs_ws(long long n);
int sort_pairs_u64(unsigned long long* keys_a, int* idx_a, unsigned long long* keys_b, int* idx_b,
                   long long n, void* ws, size_t ws_bytes, cudaStream_t st);

// ascending key <=> descending score 1 - q
__global__ void __launch_bounds__(256)
roc_keys_kernel(const double* __restrict__ q, long long n, unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = 1.0 - q[i];
    unsigned long long b = (unsigned long long)__double_as_longlong(s);
    b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);      // order-preserving map
    keys[i] = ~b;
}

__device__ __forceinline__ double roc_key_to_score(unsigned long long k) {
    unsigned long long b = ~k;
    b = (b >> 63) ? (b & 0x7FFFFFFFFFFFFFFFull) : ~b;
    return __longlong_as_double((long long)b);
}

// boundary[i]: last element of a run of equal scores; ys[i]: label in sorted order
__global__ void __launch_bounds__(256)
roc_flags_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                 const unsigned char* __restrict__ y_true, long long n,
                 unsigned char* __restrict__ boundary, unsigned char* __restrict__ ys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    boundary[i] = (i == n - 1 || keys[i] != keys[i + 1]) ? 1 : 0;
    ys[i] = y_true[idx[i]] ? 1 : 0;
}

// per threshold t (sorted position thr_idx[t]): tps = positives at positions
// <= thr_idx[t] (pos_idx: sorted positions of the positives), fps = the rest
__global__ void __launch_bounds__(256)
roc_counts_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ thr_idx,
                  long long m, const int* __restrict__ pos_idx, long long n_pos,
                  long long* __restrict__ tps, long long* __restrict__ fps,
                  double* __restrict__ thresholds) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int at = thr_idx[t];
    long long lo = 0, hi = n_pos;
    while (lo < hi) {                              // first positive beyond ``at``
        const long long mid = (lo + hi) >> 1;
        if (pos_idx[mid] <= at) lo = mid + 1; else hi = mid;
    }
    tps[t] = lo;
    fps[t] = 1 + (long long)at - lo;
    thresholds[t] = roc_key_to_score(keys[at]);
}

// sklearn's drop_intermediate: keep the end points and every point where the
// second difference of fps or of tps is non-zero
__global__ void __launch_bounds__(256)
roc_corners_kernel(const long long* __restrict__ tps, const long long* __restrict__ fps, long long m,
                   unsigned char* __restrict__ keep) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    if (m <= 2 || t == 0 || t == m - 1) { keep[t] = 1; return; }
    const long long d2f = fps[t + 1] - 2 * fps[t] + fps[t - 1];
    const long long d2t = tps[t + 1] - 2 * tps[t] + tps[t - 1];
    keep[t] = (d2f != 0 || d2t != 0) ? 1 : 0;
}

}  // namespace h3d

using namespace h3d;

extern "C" size_t h3d_roc_sort_ws_bytes(long long n) {
    if (n < 1) n = 1;
    return ws_pad((size_t)n * 8) + 2 * ws_pad((size_t)n * 4) + ws_pad(sort_pairs_ws(n));
}

extern "C" int h3d_roc_sort(const double* qvalues, const unsigned char* y_true, long long n,
                            unsigned long long* keys_sorted, unsigned char* boundary,
                            unsigned char* y_sorted, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return H3D_OK;
    Workspace w(ws, ws_bytes);
    unsigned long long* keys_b = w.take<unsigned long long>(n);
    int* idx_a = w.take<int>(n);
    int* idx_b = w.take<int>(n);
    const size_t sws = sort_pairs_ws(n);
    void* sort_ws = w.take<char>(sws);
    if (!keys_b || !idx_a || !idx_b || !sort_ws) { set_error("roc workspace too small"); return H3D_ERR_WORKSPACE; }
    const int grid = div_up(n, 256);
    roc_keys_kernel<<<grid, 256, 0, st>>>(qvalues, n, keys_sorted);
    H3D_LAUNCHED("roc_keys_kernel");
    { int rc = sort_pairs_u64(keys_sorted, idx_a, keys_b, idx_b, n, sort_ws, sws, st); if (rc) return rc; }
    roc_flags_kernel<<<grid, 256, 0, st>>>(keys_sorted, idx_a, y_true, n, boundary, y_sorted);
    H3D_LAUNCHED("roc_flags_kernel");
    return H3D_OK;
}

extern "C" int h3d_roc_points(const unsigned long long* keys_sorted, const int* thr_idx, long long m,
                              const int* pos_idx, long long n_pos, long long* tps, long long* fps,
                              double* thresholds, unsigned char* keep, h3d_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 0) return H3D_OK;
    const int grid = div_up(m, 256);
    roc_counts_kernel<<<grid, 256, 0, st>>>(keys_sorted, thr_idx, m, pos_idx, n_pos, tps, fps, thresholds);
    H3D_LAUNCHED("roc_counts_kernel");
    roc_corners_kernel<<<grid, 256, 0, st>>>(tps, fps, m, keep);
    H3D_LAUNCHED("roc_corners_kernel");
    return H3D_OK;
}
