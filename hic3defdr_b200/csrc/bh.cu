// Benjamini-Hochberg q-values on the device.
//
// Replaces lib5c.util.statistics.adjust_pvalues (-> statsmodels multipletests
// 'fdr_bh') as called at hic3defdr/analysis/analysis.py:300: over the finite
// p-values, sort ascending, q_(i) = p_(i) / (i / n), running minimum from the
// largest p downwards, clip at 1, undo the sort; non-finite entries stay NaN.
//
// HBM-bound.  The sort is an LSD radix sort (6 passes of 11 bits) built on the
// library's own stable counting-rank primitive (rank.cu); the suffix minimum
// is a three-phase tiled scan.
#include "common.cuh"

namespace h3d {

constexpr unsigned long long kNonFinite = 0xFFFFFFFFFFFFFFFFull;
constexpr int kBhTile = 2048;

// order-preserving map double -> uint64; non-finite values go to the very end
__global__ void __launch_bounds__(256)
bh_keys_kernel(const double* __restrict__ p, long long n, unsigned long long* __restrict__ keys,
               int* __restrict__ idx, unsigned long long* __restrict__ n_finite) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool fin = false;
    if (i < n) {
        const double v = p[i];
        fin = isfinite(v);
        unsigned long long b = (unsigned long long)__double_as_longlong(v);
        b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        keys[i] = fin ? b : kNonFinite;
        idx[i] = (int)i;
    }
    const unsigned m = __ballot_sync(0xffffffffu, fin);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_finite, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
bh_digit_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int mask,
                int* __restrict__ digit) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) digit[i] = (int)((keys[i] >> shift) & (unsigned long long)mask);
}

__global__ void __launch_bounds__(256)
bh_scatter_kernel(const unsigned long long* __restrict__ keys_in, const int* __restrict__ idx_in,
                  const int* __restrict__ rank, long long n, unsigned long long* __restrict__ keys_out,
                  int* __restrict__ idx_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int r = rank[i];
        keys_out[r] = keys_in[i];
        idx_out[r] = idx_in[i];
    }
}

__device__ __forceinline__ double key_to_double(unsigned long long b) {
    b = (b >> 63) ? (b & 0x7FFFFFFFFFFFFFFFull) : ~b;
    return __longlong_as_double((long long)b);
}

// raw BH ratio of the i-th smallest finite p-value: p / ((i + 1) / n)
__device__ __forceinline__ double bh_raw(unsigned long long key, long long i, double n) {
    return key_to_double(key) / ((double)(i + 1) / n);
}

__global__ void __launch_bounds__(256)
bh_tile_min_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ n_finite,
                   double* __restrict__ tile_min) {
    __shared__ double sh[8];
    const long long nf = (long long)*n_finite;
    const long long base = (long long)blockIdx.x * kBhTile;
    if (base >= nf) return;
    double m = INFINITY;
    for (int k = threadIdx.x; k < kBhTile; k += 256) {
        const long long i = base + k;
        if (i < nf) m = fmin(m, bh_raw(keys[i], i, (double)nf));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmin(m, sh[w]);
        tile_min[blockIdx.x] = m;
    }
}

// suffix minimum over tiles, exclusive: carry[t] = min over tiles > t
__global__ void bh_tile_scan_kernel(const double* __restrict__ tile_min, const unsigned long long* __restrict__ n_finite,
                                    double* __restrict__ carry) {
    const long long nf = (long long)*n_finite;
    const long long tiles = (nf + kBhTile - 1) / kBhTile;
    double m = INFINITY;
    for (long long t = tiles - 1; t >= 0; --t) {
        carry[t] = m;
        m = fmin(m, tile_min[t]);
    }
}

__global__ void __launch_bounds__(256)
bh_finish_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                 const unsigned long long* __restrict__ n_finite, const double* __restrict__ carry,
                 long long n, double* __restrict__ q) {
    __shared__ double sh[256];
    const long long nf = (long long)*n_finite;
    const long long base = (long long)blockIdx.x * kBhTile;
    constexpr int per = kBhTile / 256;
    const long long first = base + (long long)threadIdx.x * per;
    if (base >= nf) {
        // non-finite inputs keep NaN
        for (int k = 0; k < per; ++k) {
            const long long i = first + k;
            if (i < n) q[idx[i]] = NAN;
        }
        return;
    }
    double v[per];
    double m = INFINITY;
    for (int k = per - 1; k >= 0; --k) {
        const long long i = first + k;
        if (i < nf) m = fmin(m, bh_raw(keys[i], i, (double)nf));
        v[k] = m;                      // suffix min inside the thread's run
    }
    sh[threadIdx.x] = m;
    __syncthreads();
    // exclusive suffix min over the threads of the tile
    double after = carry[blockIdx.x];
    for (int t = 255; t > (int)threadIdx.x; --t) after = fmin(after, sh[t]);
    for (int k = 0; k < per; ++k) {
        const long long i = first + k;
        if (i < nf) {
            double out = fmin(v[k], after);
            if (out > 1.0) out = 1.0;
            q[idx[i]] = out;
        } else if (i < n) {
            q[idx[i]] = NAN;
        }
    }
}

}  // namespace h3d

using namespace h3d;

static const int kDigitBits[6] = {11, 11, 11, 11, 11, 9};

extern "C" size_t h3d_bh_ws_bytes(long long n) {
    if (n < 1) n = 1;
    const long long tiles = (n + kBhTile - 1) / kBhTile;
    return 2 * ws_pad((size_t)n * 8) + 4 * ws_pad((size_t)n * 4) + stable_rank_ws(n, 2048) +
           ws_pad(2049 * 8) + 2 * ws_pad((size_t)tiles * 8) + ws_pad(64);
}

extern "C" int h3d_bh(const double* p, long long n, double* q, void* ws, size_t ws_bytes,
                      h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    H3D_REQUIRE(n < 2147483647LL, "more than 2^31 p-values");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w(ws, ws_bytes);
    const long long tiles = (n + kBhTile - 1) / kBhTile;
    unsigned long long* keys_a = w.take<unsigned long long>(n);
    unsigned long long* keys_b = w.take<unsigned long long>(n);
    int* idx_a = w.take<int>(n);
    int* idx_b = w.take<int>(n);
    int* digit = w.take<int>(n);
    int* rank = w.take<int>(n);
    const size_t rws = stable_rank_ws(n, 2048);
    void* rank_ws = w.take<char>(rws);
    long long* key_start = w.take<long long>(2049);
    double* tile_min = w.take<double>(tiles);
    double* carry = w.take<double>(tiles);
    unsigned long long* n_finite = w.take<unsigned long long>(1);
    if (!keys_a || !keys_b || !idx_a || !idx_b || !digit || !rank || !rank_ws || !key_start ||
        !tile_min || !carry || !n_finite) {
        set_error("bh workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    const int grid = div_up(n, 256);
    H3D_CHECK(cudaMemsetAsync(n_finite, 0, 8, st));
    bh_keys_kernel<<<grid, 256, 0, st>>>(p, n, keys_a, idx_a, n_finite);
    H3D_LAUNCHED("bh_keys_kernel");
    int shift = 0;
    for (int pass = 0; pass < 6; ++pass) {
        const int bits = kDigitBits[pass];
        bh_digit_kernel<<<grid, 256, 0, st>>>(keys_a, n, shift, (1 << bits) - 1, digit);
        H3D_LAUNCHED("bh_digit_kernel");
        int rc = stable_rank_impl(digit, n, 1 << bits, rank, key_start, rank_ws, rws, st);
        if (rc) return rc;
        bh_scatter_kernel<<<grid, 256, 0, st>>>(keys_a, idx_a, rank, n, keys_b, idx_b);
        H3D_LAUNCHED("bh_scatter_kernel");
        unsigned long long* tk = keys_a; keys_a = keys_b; keys_b = tk;
        int* ti = idx_a; idx_a = idx_b; idx_b = ti;
        shift += bits;
    }
    bh_tile_min_kernel<<<(int)tiles, 256, 0, st>>>(keys_a, n_finite, tile_min);
    H3D_LAUNCHED("bh_tile_min_kernel");
    bh_tile_scan_kernel<<<1, 1, 0, st>>>(tile_min, n_finite, carry);
    H3D_LAUNCHED("bh_tile_scan_kernel");
    bh_finish_kernel<<<(int)tiles, 256, 0, st>>>(keys_a, idx_a, n_finite, carry, n, q);
    H3D_LAUNCHED("bh_finish_kernel");
    return H3D_OK;
}
