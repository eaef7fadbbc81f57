// Connected components of a sparse set of matrix pixels (the step right after
// bh(): hic3defdr/analysis/analysis.py:366-430 threshold(), :432-496
// classify()).
//
// Replaces hic3defdr/util/clusters.py:15-96 (DirectedDisjointSet +
// find_clusters with connectivity 1): the reference walks every pixel in a
// Python loop and adds an edge to each of its four neighbours through a
// dict-of-sets union-find; the groups it ends with are the 4-connected
// components of the pixel set.  Here: pixels arrive sorted by (row, col) (they
// are subsets of the union pixel order), the right neighbour of a pixel is the
// next element if it exists, the lower neighbour is found by a binary search
// for (row + 1, col), and components are merged with a lock-free union-find
// (hook the larger root under the smaller by compare-and-swap, path halving).
// The representative of a component is its first pixel in (row, col) order.
#include "common.cuh"

namespace h3d {

__device__ __forceinline__ int cc_find(int* __restrict__ parent, int x) {
    // path halving; concurrent hooks only ever replace a root by a smaller
    // index, so an ancestor stays an ancestor and the racy writes are benign
    int p = __ldcg(parent + x);
    while (p != x) {
        const int gp = __ldcg(parent + p);
        if (gp != p) __stcg(parent + x, gp);
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void cc_unite(int* __restrict__ parent, int a, int b) {
    for (;;) {
        a = cc_find(parent, a);
        b = cc_find(parent, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }       // a < b: hook b under a
        const int old = atomicCAS(parent + b, b, a);
        if (old == b) return;
        b = old;                                            // b was hooked meanwhile: retry
    }
}

__global__ void __launch_bounds__(256)
cc_init_kernel(int* __restrict__ parent, int* __restrict__ size, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = (int)i; size[i] = 0; }
}

__device__ __forceinline__ long long cc_key(const int* __restrict__ row, const int* __restrict__ col,
                                            long long i) {
    return ((long long)row[i] << 32) | (long long)(unsigned)col[i];
}

__global__ void __launch_bounds__(256)
cc_link_kernel(const int* __restrict__ row, const int* __restrict__ col, long long n,
               int* __restrict__ parent, int* __restrict__ unsorted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = row[i], c = col[i];
    if (i + 1 < n) {
        const long long k0 = cc_key(row, col, i), k1 = cc_key(row, col, i + 1);
        if (k1 <= k0) *unsorted = 1;                        // contract: sorted, unique
        if (row[i + 1] == r && col[i + 1] == c + 1) cc_unite(parent, (int)i, (int)(i + 1));
    }
    // lower neighbour (r + 1, c): first position with key >= target in (i, n)
    // galloping from the pixel's own position: the neighbour is at most a row's
    // worth of pixels away, so the probes stay in nearby cache lines (~16
    // instead of ~25 across the whole array)
    const long long target = ((long long)(r + 1) << 32) | (long long)(unsigned)c;
    long long lo = i + 1, hi = n;
    for (long long step = 1; lo + step < n; step <<= 1) {
        if (cc_key(row, col, lo + step) >= target) { hi = lo + step; break; }
        lo += step;
    }
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (cc_key(row, col, mid) < target) lo = mid + 1; else hi = mid;
    }
    if (lo < n && cc_key(row, col, lo) == target) cc_unite(parent, (int)i, (int)lo);
}

__global__ void __launch_bounds__(256)
cc_label_kernel(int* __restrict__ parent, long long n, int* __restrict__ label,
                int* __restrict__ size) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int root = cc_find(parent, (int)i);
    label[i] = root;
    atomicAdd(size + root, 1);
}

}  // namespace h3d

using namespace h3d;

extern "C" size_t h3d_connected_components_ws_bytes(long long n) {
    return ws_pad((size_t)(n > 0 ? n : 1) * 4) + ws_pad(4);
}

extern "C" int h3d_connected_components(const int* row, const int* col, long long n, int* label,
                                        int* size, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    H3D_REQUIRE(n < 2147483647LL, "more than 2^31 pixels");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w(ws, ws_bytes);
    int* parent = w.take<int>(n);
    int* unsorted = w.take<int>(1);
    if (!parent || !unsorted) { set_error("connected_components workspace too small"); return H3D_ERR_WORKSPACE; }
    H3D_CHECK(cudaMemsetAsync(unsorted, 0, sizeof(int), st));
    const int grid = div_up(n, 256);
    cc_init_kernel<<<grid, 256, 0, st>>>(parent, size, n);
    H3D_LAUNCHED("cc_init_kernel");
    cc_link_kernel<<<grid, 256, 0, st>>>(row, col, n, parent, unsorted);
    H3D_LAUNCHED("cc_link_kernel");
    cc_label_kernel<<<grid, 256, 0, st>>>(parent, n, label, size);
    H3D_LAUNCHED("cc_label_kernel");
    int flag = 0;
    H3D_CHECK(cudaMemcpyAsync(&flag, unsorted, sizeof(int), cudaMemcpyDeviceToHost, st));
    H3D_CHECK(cudaStreamSynchronize(st));
    if (flag) {
        set_error("connected_components: pixels must be sorted by (row, col) and unique");
        return H3D_ERR_ARG;
    }
    return H3D_OK;
}
