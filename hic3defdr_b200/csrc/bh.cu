// Benjamini-Hochberg q-values on the device.
//
// Replaces lib5c.util.statistics.adjust_pvalues (-> statsmodels multipletests
// 'fdr_bh') as called at hic3defdr/analysis/analysis.py:300: over the finite
// p-values, sort ascending, q_(i) = p_(i) / (i / n), running minimum from the
// largest p downwards, clip at 1, undo the sort; non-finite entries stay NaN.
//
// HBM-bound.  The sort is the one-sweep-per-digit LSD radix sort of sort.cu
// (8 sweeps of 8 bits, 200 B moved per p-value); the suffix minimum is a
// three-phase tiled scan.
#include "common.cuh"

namespace h3d {

constexpr unsigned long long kNonFinite = 0xFFFFFFFFFFFFFFFFull;
constexpr int kBhTile = 2048;

// order-preserving map double -> uint64; non-finite values go to the very end.
// Each block converts kKeysPerBlock values and adds its finite count with ONE
// atomic (a per-warp atomic on the single counter serialised 1.2 M updates:
// 1.1 ms of a 0.15 ms kernel).
constexpr int kKeysPerBlock = 4096;
__global__ void __launch_bounds__(256)
bh_keys_kernel(const double* __restrict__ p, long long n, unsigned long long* __restrict__ keys,
               unsigned long long* __restrict__ n_finite) {
    __shared__ int sh[8];
    const long long base = (long long)blockIdx.x * kKeysPerBlock;
    int cnt = 0;
#pragma unroll 4
    for (int k = threadIdx.x; k < kKeysPerBlock; k += 256) {
        const long long i = base + k;
        if (i < n) {
            const double v = p[i];
            const bool fin = isfinite(v);
            unsigned long long b = (unsigned long long)__double_as_longlong(v);
            b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
            keys[i] = fin ? b : kNonFinite;
            cnt += fin ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        if (t) atomicAdd(n_finite, (unsigned long long)t);
    }
}

__device__ __forceinline__ double key_to_double(unsigned long long b) {
    b = (b >> 63) ? (b & 0x7FFFFFFFFFFFFFFFull) : ~b;
    return __longlong_as_double((long long)b);
}

// raw BH ratio of the p-value of (1-based) rank i + 1 among n: p / ((i + 1) / n)
__device__ __forceinline__ double bh_raw(unsigned long long key, long long i, double n) {
    return key_to_double(key) / ((double)(i + 1) / n);
}

// rank of the first local value minus one, and the number of tested hypotheses:
// the local finite count for a single-device correction, host-supplied values
// for one bucket of a distributed one
struct BhScope { long long rank_offset, n_total; };
__device__ __forceinline__ double bh_total(const BhScope& sc, long long nf) {
    return (double)(sc.n_total > 0 ? sc.n_total : nf);
}

__global__ void __launch_bounds__(256)
bh_tile_min_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ n_finite,
                   BhScope sc, double* __restrict__ tile_min) {
    __shared__ double sh[8];
    const long long nf = (long long)*n_finite;
    const long long base = (long long)blockIdx.x * kBhTile;
    if (base >= nf) return;
    const double nt = bh_total(sc, nf);
    double m = INFINITY;
    for (int k = threadIdx.x; k < kBhTile; k += 256) {
        const long long i = base + k;
        if (i < nf) m = fmin(m, bh_raw(keys[i], sc.rank_offset + i, nt));
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

// suffix minimum over tiles, exclusive: carry[t] = min over tiles > t; one
// block, every thread owns a contiguous run of tiles.  Also reports the minimum
// over all tiles (the bucket minimum of a distributed correction).
__global__ void __launch_bounds__(1024)
bh_tile_scan_kernel(const double* __restrict__ tile_min, const unsigned long long* __restrict__ n_finite,
                    double* __restrict__ carry, double* __restrict__ min_out) {
    __shared__ double sh[1024];
    const long long nf = (long long)*n_finite;
    const long long tiles = (nf + kBhTile - 1) / kBhTile;
    const long long per = (tiles + 1023) / 1024;
    const long long t0 = (long long)threadIdx.x * per;
    const long long t1 = (t0 + per < tiles) ? t0 + per : tiles;
    double m = INFINITY;
    for (long long t = t0; t < t1; ++t) m = fmin(m, tile_min[t]);
    sh[threadIdx.x] = m;
    __syncthreads();
    // exclusive suffix min over threads (Hillis-Steele on the reversed order)
    double incl = m;
    for (int o = 1; o < 1024; o <<= 1) {
        const double other = (threadIdx.x + o < 1024) ? sh[threadIdx.x + o] : INFINITY;
        __syncthreads();
        incl = fmin(incl, other);
        sh[threadIdx.x] = incl;
        __syncthreads();
    }
    double after = (threadIdx.x + 1 < 1024) ? sh[threadIdx.x + 1] : INFINITY;
    if (threadIdx.x == 0 && min_out) *min_out = sh[0];
    for (long long t = t1 - 1; t >= t0; --t) {
        carry[t] = after;
        after = fmin(after, tile_min[t]);
    }
}

__global__ void __launch_bounds__(256)
bh_finish_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ idx,
                 const unsigned long long* __restrict__ n_finite, const double* __restrict__ carry,
                 BhScope sc, long long n, double* __restrict__ q) {
    __shared__ double sh[256];
    const long long nf = (long long)*n_finite;
    const double nt = bh_total(sc, nf);
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
        if (i < nf) m = fmin(m, bh_raw(keys[i], sc.rank_offset + i, nt));
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

// q <- min(q, carry), NaN kept (a distributed correction's contribution of
// the buckets holding larger p-values)
__global__ void __launch_bounds__(256)
bh_apply_carry_kernel(double* __restrict__ q, long long n, double carry,
                      const double* __restrict__ carry_dev) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (carry_dev) carry = *carry_dev;
    if (i < n) {
        const double v = q[i];
        if (v > carry) q[i] = carry;
    }
}

}  // namespace h3d

using namespace h3d;

namespace h3d {
// sort.cu
size_t sort_pairs_ws(long long n);
int sort_pairs_u64(unsigned long long* keys_a, int* idx_a, unsigned long long* keys_b, int* idx_b,
                   long long n, void* ws, size_t ws_bytes, cudaStream_t st);
}

extern "C" size_t h3d_bh_ws_bytes(long long n) {
    if (n < 1) n = 1;
    const long long tiles = (n + kBhTile - 1) / kBhTile;
    return 2 * ws_pad((size_t)n * 8) + 2 * ws_pad((size_t)n * 4) + ws_pad(sort_pairs_ws(n)) +
           2 * ws_pad((size_t)tiles * 8) + ws_pad(64);
}

extern "C" int h3d_bh(const double* p, long long n, double* q, void* ws, size_t ws_bytes,
                      h3d_stream_t stream) {
    return h3d_bh_ranked(p, n, 0, 0, q, nullptr, ws, ws_bytes, stream);
}

extern "C" int h3d_bh_apply_carry(double* q, long long n, double carry, h3d_stream_t stream) {
    if (n <= 0 || !(carry < 1.0)) return H3D_OK;      // q is already clipped at 1
    bh_apply_carry_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(q, n, carry, nullptr);
    H3D_LAUNCHED("bh_apply_carry_kernel");
    return H3D_OK;
}

extern "C" int h3d_bh_apply_carry_dev(double* q, long long n, const double* carry_dev,
                                      h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    bh_apply_carry_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(q, n, 0.0, carry_dev);
    H3D_LAUNCHED("bh_apply_carry_kernel");
    return H3D_OK;
}

extern "C" int h3d_bh_ranked(const double* p, long long n, long long rank_offset, long long n_total,
                             double* q, double* min_out, void* ws, size_t ws_bytes,
                             h3d_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) {
        if (min_out) {
            const double inf = INFINITY;
            H3D_CHECK(cudaMemcpyAsync(min_out, &inf, 8, cudaMemcpyHostToDevice, st));
            H3D_CHECK(cudaStreamSynchronize(st));
        }
        return H3D_OK;
    }
    H3D_REQUIRE(n < (1LL << 30), "more than 2^30 p-values on one device");
    H3D_REQUIRE(rank_offset >= 0 && n_total >= 0, "negative rank offset / total");
    const BhScope sc = {rank_offset, n_total};
    Workspace w(ws, ws_bytes);
    const long long tiles = (n + kBhTile - 1) / kBhTile;
    unsigned long long* keys_a = w.take<unsigned long long>(n);
    unsigned long long* keys_b = w.take<unsigned long long>(n);
    int* idx_a = w.take<int>(n);
    int* idx_b = w.take<int>(n);
    const size_t sws = sort_pairs_ws(n);
    void* sort_ws = w.take<char>(sws);
    double* tile_min = w.take<double>(tiles);
    double* carry = w.take<double>(tiles);
    unsigned long long* n_finite = w.take<unsigned long long>(1);
    if (!keys_a || !keys_b || !idx_a || !idx_b || !sort_ws || !tile_min || !carry || !n_finite) {
        set_error("bh workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    H3D_CHECK(cudaMemsetAsync(n_finite, 0, 8, st));
    bh_keys_kernel<<<div_up(n, kKeysPerBlock), 256, 0, st>>>(p, n, keys_a, n_finite);
    H3D_LAUNCHED("bh_keys_kernel");
    // ascending by p, payload = original position; non-finite keys sort last
    { int rc = sort_pairs_u64(keys_a, idx_a, keys_b, idx_b, n, sort_ws, sws, st); if (rc) return rc; }
    bh_tile_min_kernel<<<(int)tiles, 256, 0, st>>>(keys_a, n_finite, sc, tile_min);
    H3D_LAUNCHED("bh_tile_min_kernel");
    bh_tile_scan_kernel<<<1, 1024, 0, st>>>(tile_min, n_finite, carry, min_out);
    H3D_LAUNCHED("bh_tile_scan_kernel");
    bh_finish_kernel<<<(int)tiles, 256, 0, st>>>(keys_a, idx_a, n_finite, carry, sc, n, q);
    H3D_LAUNCHED("bh_finish_kernel");
    return H3D_OK;
}
