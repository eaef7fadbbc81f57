// prepare_data kernels: bias filter, union of per-replicate CSR matrices with
// raw / balanced gather, scaling + disp_idx, mask compaction, small gathers.
//
// Replaces hic3defdr/analysis/core.py:56-60 (load_bias filter),
// hic3defdr/util/matrices.py:8-62,92-129 (deconvolute, wipe_distances,
// sparse_union), hic3defdr/analysis/analysis.py:92-101 (raw / balanced),
// :109-115 (scaled, disp_idx), :117-125 (loop_idx), :181-183 (combined factor).
//
// All of these are HBM-bound streaming kernels.  The union works on the band
// structure directly: one warp owns one matrix row, builds the row's
// (dist_max + 1)-bit occupancy bitmap in shared memory from the R replicate
// rows (already sorted by column), and a second pass with the same bitmap
// writes row/col/raw/balanced for the row's contiguous output range -- no
// global sort, no global atomics, output in (row, col) order by construction.
#include "common.cuh"

namespace h3d {

// ---------------------------------------------------------------------------
// single-block exclusive scan of int32 counts (n up to a few million).
// out[i] = sum_{j<i} in[j], out[n] = total; optional int64 total.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scan_exclusive_kernel(const int* __restrict__ in, int* __restrict__ out, long long n,
                      long long* __restrict__ total64) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = (i < n) ? (long long)in[i] : 0;
        long long s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[wid] = s;
        __syncthreads();
        if (wid == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long prev_warps = (wid > 0) ? warp_tot[wid - 1] : 0;
        if (i < n) out[i] = (int)(carry + prev_warps + s - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[n] = (int)carry_s;
        if (total64) *total64 = carry_s;
    }
}

// ---------------------------------------------------------------------------
// load_bias filter
// ---------------------------------------------------------------------------
__global__ void bias_filter_kernel(double* __restrict__ bias, int n_bins, int n_reps,
                                   double thresh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bins) return;
    const double hi = 1.0 / thresh;
    bool bad = false;
    for (int r = 0; r < n_reps; ++r) {
        const double b = bias[(long long)i * n_reps + r];
        bad = bad || (b < thresh) || (b > hi);      // NaN compares false (core.py:58-59)
    }
    if (bad)
        for (int r = 0; r < n_reps; ++r) bias[(long long)i * n_reps + r] = 0.0;
}

// ---------------------------------------------------------------------------
// union
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long lower_bound_col(const int* __restrict__ idx,
                                                     long long lo, long long hi, int key) {
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// normalised value used for union membership: (inv_i * v) * inv_j with 1/0 -> 0
// (matrices.py:29-38)
__device__ __forceinline__ double normalised(double v, const double* __restrict__ bias,
                                             int i, int j, int r, int n_reps) {
    if (bias == nullptr) return v;
    const double bi = bias[(long long)i * n_reps + r], bj = bias[(long long)j * n_reps + r];
    const double ii = (bi == 0.0) ? 0.0 : 1.0 / bi;
    const double ij = (bj == 0.0) ? 0.0 : 1.0 / bj;
    return (ii * v) * ij;
}

// sum over replicates of the normalised values stored at (i, j), replicate order
__device__ double cell_sum(const CsrReps& reps, int n_reps, int is64, int dtype,
                           const double* __restrict__ bias, int i, int j) {
    double s = 0.0;
    bool first = true;
    for (int r = 0; r < n_reps; ++r) {
        const long long lo = load_indptr(reps.indptr[r], i, is64);
        const long long hi = load_indptr(reps.indptr[r], i + 1, is64);
        const long long k = lower_bound_col(reps.indices[r], lo, hi, j);
        if (k < hi && reps.indices[r][k] == j) {
            const double vn = normalised(load_value(reps.data[r], k, dtype), bias, i, j, r, n_reps);
            if (vn != 0.0) { s = first ? vn : s + vn; first = false; }
        }
    }
    return s;
}

// One warp per matrix row.  Shared memory per warp: 4 * W words
// (present, bad, negative, prefix), W = ceil((dist_max + 1) / 32).
template <bool EMIT>
__global__ void __launch_bounds__(256)
union_kernel(CsrReps reps, int n_reps, int is64, int dtype, const double* __restrict__ bias,
             int n_bins, int dist_max, int W, int* __restrict__ row_count,
             const int* __restrict__ row_offset, int* __restrict__ row_out,
             int* __restrict__ col_out, int* __restrict__ dist_out,
             long long* __restrict__ raw_out, double* __restrict__ bal_out) {
    extern __shared__ unsigned smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    unsigned* present = smem + (size_t)wid * 4 * W;
    unsigned* bad = present + W;
    unsigned* neg = bad + W;
    unsigned* prefix = neg + W;
    for (int i = blockIdx.x * warps_per_block + wid; i < n_bins; i += gridDim.x * warps_per_block) {
        for (int w = lane; w < W; w += 32) { present[w] = 0; bad[w] = 0; neg[w] = 0; }
        __syncwarp();
        const int jmax = (i + dist_max < n_bins - 1) ? i + dist_max : n_bins - 1;
        for (int r = 0; r < n_reps; ++r) {
            const int* __restrict__ idx = reps.indices[r];
            const long long lo = load_indptr(reps.indptr[r], i, is64);
            const long long hi = load_indptr(reps.indptr[r], i + 1, is64);
            const long long k0 = lower_bound_col(idx, lo, hi, i);    // drops dist < 0
            for (long long k = k0 + lane; k < hi; k += 32) {
                const int j = idx[k];
                if (j > jmax) break;
                const double vn = normalised(load_value(reps.data[r], k, dtype), bias, i, j, r, n_reps);
                if (vn != 0.0) {          // NaN != 0: kept here, rejected by the finite test
                    const int d = j - i;
                    const unsigned bit = 1u << (d & 31);
                    atomicOr(&present[d >> 5], bit);
                    if (!isfinite(vn)) atomicOr(&bad[d >> 5], bit);
                    else if (vn < 0.0) atomicOr(&neg[d >> 5], bit);
                }
            }
        }
        __syncwarp();
        // finalise validity: sum of normalised values finite and >= 0 (and,
        // as scipy's CSR addition drops exact zeros, != 0); only cells that
        // hold a negative value need the explicit sum.
        for (int w = lane; w < W; w += 32) {
            unsigned v = present[w] & ~bad[w];
            unsigned ng = neg[w] & v;
            while (ng) {
                const int b = __ffs(ng) - 1;
                ng &= ng - 1;
                const double s = cell_sum(reps, n_reps, is64, dtype, bias, i, i + w * 32 + b);
                if (!(s > 0.0)) v &= ~(1u << b);
            }
            present[w] = v;
        }
        __syncwarp();
        // exclusive prefix of popcounts over the row's words
        int run = 0;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int w = w0 + lane;
            const int c = (w < W) ? __popc(present[w]) : 0;
            int s = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            if (w < W) prefix[w] = run + s - c;
            run += __shfl_sync(0xffffffffu, s, 31);
        }
        if (!EMIT) {
            if (lane == 0) row_count[i] = run;
            __syncwarp();
            continue;
        }
        __syncwarp();
        const long long base = row_offset[i];
        // row / col and the zero fill of raw / balanced, coalesced over d
        for (int d = lane; d <= dist_max; d += 32) {
            const unsigned word = present[d >> 5];
            if ((word >> (d & 31)) & 1u) {
                const long long pos = base + prefix[d >> 5] + __popc(word & ((1u << (d & 31)) - 1u));
                const int j = i + d;
                row_out[pos] = i;
                col_out[pos] = j;
                if (dist_out) dist_out[pos] = d;
                for (int r = 0; r < n_reps; ++r) {
                    raw_out[pos * n_reps + r] = 0;
                    double z = 0.0;
                    if (bias != nullptr)       // absent entry: 0 / (b_i b_j) (analysis.py:100-101)
                        z = 0.0 / (bias[(long long)i * n_reps + r] * bias[(long long)j * n_reps + r]);
                    bal_out[pos * n_reps + r] = z;
                }
            }
        }
        __syncwarp();
        for (int r = 0; r < n_reps; ++r) {
            const int* __restrict__ idx = reps.indices[r];
            const long long lo = load_indptr(reps.indptr[r], i, is64);
            const long long hi = load_indptr(reps.indptr[r], i + 1, is64);
            const long long k0 = lower_bound_col(idx, lo, hi, i);
            for (long long k = k0 + lane; k < hi; k += 32) {
                const int j = idx[k];
                if (j > jmax) break;
                const int d = j - i;
                const unsigned word = present[d >> 5];
                if ((word >> (d & 31)) & 1u) {
                    const long long pos = base + prefix[d >> 5] + __popc(word & ((1u << (d & 31)) - 1u));
                    const double v = load_value(reps.data[r], k, dtype);
                    raw_out[pos * n_reps + r] = (long long)v;          // int64, truncating
                    double bb = 1.0;
                    if (bias != nullptr)
                        bb = bias[(long long)i * n_reps + r] * bias[(long long)j * n_reps + r];
                    bal_out[pos * n_reps + r] = v / bb;
                }
            }
        }
        __syncwarp();
    }
}

// Emit pass with the row staged in shared memory (the fast path; the kernel
// above stays as the generic one for bands too wide to stage).  One warp per
// matrix row: the replicate values of the row are scattered into a
// (dist_max + 1) x R tile of shared memory while the occupancy bitmap is
// built, then every present pixel is written once, its R raw values and R
// balanced values as full 32-byte sectors -- no zero-fill pass, no second walk
// over the CSR rows, every output byte written exactly once (ncu r01d: the
// two-pass kernel moved 2.7x the algorithmic bytes).
// Shared memory per warp: 4 W words of bitmaps + (dist_max + 1) * R doubles.
__global__ void __launch_bounds__(256)
union_emit_staged_kernel(CsrReps reps, int n_reps, int is64, int dtype,
                         const double* __restrict__ bias, int n_bins, int dist_max, int W,
                         int warp_stride_words, const int* __restrict__ row_offset,
                         int* __restrict__ row_out, int* __restrict__ col_out,
                         int* __restrict__ dist_out, long long* __restrict__ raw_out,
                         double* __restrict__ bal_out) {
    extern __shared__ unsigned smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    unsigned* present = smem + (size_t)wid * warp_stride_words;
    unsigned* bad = present + W;
    unsigned* neg = bad + W;
    unsigned* prefix = neg + W;
    double* val = (double*)(present + ((4 * W + 1) & ~1));      // 8-byte aligned
    const int n_val = (dist_max + 1) * n_reps;
    for (int i = blockIdx.x * warps_per_block + wid; i < n_bins; i += gridDim.x * warps_per_block) {
        for (int w = lane; w < W; w += 32) { present[w] = 0; bad[w] = 0; neg[w] = 0; }
        for (int t = lane; t < n_val; t += 32) val[t] = 0.0;
        __syncwarp();
        const int jmax = (i + dist_max < n_bins - 1) ? i + dist_max : n_bins - 1;
        for (int r = 0; r < n_reps; ++r) {
            const int* __restrict__ idx = reps.indices[r];
            const long long lo = load_indptr(reps.indptr[r], i, is64);
            const long long hi = load_indptr(reps.indptr[r], i + 1, is64);
            // upper-triangular inputs start at or right of the diagonal
            const long long k0 = (lo < hi && idx[lo] >= i) ? lo : lower_bound_col(idx, lo, hi, i);
            for (long long k = k0 + lane; k < hi; k += 32) {
                const int j = idx[k];
                if (j > jmax) break;
                const double v = load_value(reps.data[r], k, dtype);
                const double vn = normalised(v, bias, i, j, r, n_reps);
                const int d = j - i;
                val[r * (dist_max + 1) + d] = v;        // replicate-major: conflict-free
                if (vn != 0.0) {
                    const unsigned bit = 1u << (d & 31);
                    atomicOr(&present[d >> 5], bit);
                    if (!isfinite(vn)) atomicOr(&bad[d >> 5], bit);
                    else if (vn < 0.0) atomicOr(&neg[d >> 5], bit);
                }
            }
        }
        __syncwarp();
        for (int w = lane; w < W; w += 32) {
            unsigned v = present[w] & ~bad[w];
            unsigned ng = neg[w] & v;
            while (ng) {
                const int b = __ffs(ng) - 1;
                ng &= ng - 1;
                const double s = cell_sum(reps, n_reps, is64, dtype, bias, i, i + w * 32 + b);
                if (!(s > 0.0)) v &= ~(1u << b);
            }
            present[w] = v;
        }
        __syncwarp();
        int run = 0;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int w = w0 + lane;
            const int c = (w < W) ? __popc(present[w]) : 0;
            int s = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += t;
            }
            if (w < W) prefix[w] = run + s - c;
            run += __shfl_sync(0xffffffffu, s, 31);
        }
        __syncwarp();
        const long long base = row_offset[i];
        for (int d = lane; d <= dist_max; d += 32) {
            const unsigned word = present[d >> 5];
            if (!((word >> (d & 31)) & 1u)) continue;
            const long long pos = base + prefix[d >> 5] + __popc(word & ((1u << (d & 31)) - 1u));
            const int j = i + d;
            row_out[pos] = i;
            col_out[pos] = j;
            if (dist_out) dist_out[pos] = d;
            const double* __restrict__ vrow = val + d;
            const int vs = dist_max + 1;
            long long* __restrict__ raw_p = raw_out + pos * n_reps;
            double* __restrict__ bal_p = bal_out + pos * n_reps;
            if ((n_reps & 1) == 0) {
                for (int r = 0; r < n_reps; r += 2) {
                    const double v0 = vrow[r * vs], v1 = vrow[(r + 1) * vs];
                    double b0 = 1.0, b1 = 1.0;
                    if (bias != nullptr) {      // absent entry: 0 / (b_i b_j) (analysis.py:100-101)
                        b0 = bias[(long long)i * n_reps + r] * bias[(long long)j * n_reps + r];
                        b1 = bias[(long long)i * n_reps + r + 1] * bias[(long long)j * n_reps + r + 1];
                    }
                    *(longlong2*)(raw_p + r) = make_longlong2((long long)v0, (long long)v1);
                    *(double2*)(bal_p + r) = make_double2(v0 / b0, v1 / b1);
                }
            } else {
                for (int r = 0; r < n_reps; ++r) {
                    const double v = vrow[r * vs];
                    double bb = 1.0;
                    if (bias != nullptr)
                        bb = bias[(long long)i * n_reps + r] * bias[(long long)j * n_reps + r];
                    raw_p[r] = (long long)v;
                    bal_p[r] = v / bb;
                }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// scaled / size_factors / disp_idx
// ---------------------------------------------------------------------------
struct DesignF {
    double w[H3D_MAX_REPS][H3D_MAX_CONDS];   // design as 0.0 / 1.0
    double n_in_cond[H3D_MAX_CONDS];
};

__global__ void __launch_bounds__(256)
scale_filter_kernel(const int* __restrict__ row, const int* __restrict__ col,
                    double* __restrict__ data, const double* __restrict__ sf_table,
                    int sf_per_dist, DesignF dz, long long n_px, int n_reps, int n_conds,
                    int dist_max, double mean_thresh, int dist_min,
                    double* __restrict__ sf_out, unsigned char* __restrict__ disp_idx) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_px) return;
    const int d = col[p] - row[p];
    double acc[H3D_MAX_CONDS];
    for (int c = 0; c < n_conds; ++c) acc[c] = 0.0;
    for (int r = 0; r < n_reps; ++r) {
        const double s = sf_per_dist ? sf_table[(long long)d * n_reps + r] : sf_table[r];
        const double v = data[p * n_reps + r] / s;
        data[p * n_reps + r] = v;
        if (sf_out) sf_out[p * n_reps + r] = s;
        for (int c = 0; c < n_conds; ++c) acc[c] = __dadd_rn(acc[c], __dmul_rn(v, dz.w[r][c]));
    }
    bool ok = d >= dist_min;
    for (int c = 0; c < n_conds; ++c) ok = ok && ((acc[c] / dz.n_in_cond[c]) >= mean_thresh);
    disp_idx[p] = ok ? 1 : 0;
}

// ---------------------------------------------------------------------------
// mask -> index list (ordered stream compaction), 3 launches
// ---------------------------------------------------------------------------
constexpr int kCompactTile = 4096;

__global__ void __launch_bounds__(256)
mask_count_kernel(const unsigned char* __restrict__ mask, long long n, int* __restrict__ tile_count) {
    const long long base = (long long)blockIdx.x * kCompactTile;
    int c = 0;
    for (int k = threadIdx.x; k < kCompactTile; k += 256) {
        const long long i = base + k;
        if (i < n && mask[i]) ++c;
    }
    __shared__ int sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        tile_count[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
mask_emit_kernel(const unsigned char* __restrict__ mask, long long n,
                 const int* __restrict__ tile_offset, int* __restrict__ index_out) {
    __shared__ int warp_base[8];
    __shared__ int running;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * kCompactTile;
    if (threadIdx.x == 0) running = tile_offset[blockIdx.x];
    __syncthreads();
    for (int k0 = 0; k0 < kCompactTile; k0 += 256) {
        const long long i = base + k0 + threadIdx.x;
        const bool set = (i < n) && mask[i];
        const unsigned ballot = __ballot_sync(0xffffffffu, set);
        if (lane == 0) warp_base[wid] = __popc(ballot);
        __syncthreads();
        int before = running;
        for (int w = 0; w < wid; ++w) before += warp_base[w];
        if (set) index_out[before + __popc(ballot & ((1u << lane) - 1u))] = (int)i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_base[w];
            running += t;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// small gathers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loop_membership_kernel(const int* __restrict__ row, const int* __restrict__ col,
                       const int* __restrict__ index, long long n,
                       const long long* __restrict__ keys, long long n_keys,
                       unsigned char* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long u = index ? (long long)index[i] : i;
    const long long key = ((long long)row[u] << 32) | (unsigned)col[u];
    long long lo = 0, hi = n_keys;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    out[i] = (lo < n_keys && keys[lo] == key) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
gather_counts_factors_kernel(const int* __restrict__ row, const int* __restrict__ col,
                             const int* __restrict__ index, long long n_sel,
                             const long long* __restrict__ raw, const double* __restrict__ sf,
                             int sf_per_pixel, const double* __restrict__ bias, int n_reps,
                             const int* __restrict__ dest, long long ld,
                             double* __restrict__ x_out, double* __restrict__ f_out,
                             int* __restrict__ dist_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel) return;
    const long long u = index ? (long long)index[i] : i;
    const long long pos = dest ? (long long)dest[i] : i;
    const int ri = row[u], ci = col[u];
    if (dist_out) dist_out[pos] = ci - ri;
    if (x_out == nullptr) return;
    for (int r = 0; r < n_reps; ++r) {
        const double s = sf_per_pixel ? sf[u * n_reps + r] : sf[r];
        x_out[(long long)r * ld + pos] = (double)raw[u * n_reps + r];
        f_out[(long long)r * ld + pos] =
            bias[(long long)ri * n_reps + r] * bias[(long long)ci * n_reps + r] * s;
    }
}

__global__ void __launch_bounds__(256)
gather_table_kernel(const int* __restrict__ dist, long long n, const double* __restrict__ table,
                    int n_cols, int n_rows, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = dist[i];
    for (int c = 0; c < n_cols; ++c)
        out[i * n_cols + c] = (d >= 0 && d < n_rows) ? table[(long long)d * n_cols + c] : NAN;
}

static int fill_reps(CsrReps* reps, int n_reps, const void* const* indptr, const int* const* indices,
                     const void* const* data) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    for (int r = 0; r < kMaxReps; ++r) {
        reps->indptr[r] = (r < n_reps) ? indptr[r] : nullptr;
        reps->indices[r] = (r < n_reps) ? indices[r] : nullptr;
        reps->data[r] = (r < n_reps) ? data[r] : nullptr;
    }
    return H3D_OK;
}

}  // namespace h3d

using namespace h3d;

extern "C" int h3d_bias_filter(double* bias, int n_bins, int n_reps, double bias_thresh,
                               h3d_stream_t stream) {
    if (n_bins <= 0) return H3D_OK;
    bias_filter_kernel<<<div_up(n_bins, 256), 256, 0, (cudaStream_t)stream>>>(bias, n_bins, n_reps,
                                                                            bias_thresh);
    H3D_LAUNCHED("bias_filter_kernel");
    return H3D_OK;
}

extern "C" size_t h3d_union_ws_bytes(int n_bins) { return ws_pad((size_t)(n_bins + 1) * sizeof(int)); }

static int union_launch_shape(int n_bins, int dist_max, int* W, int* grid, size_t* smem) {
    *W = (dist_max + 1 + 31) / 32;
    *smem = (size_t)8 * 4 * (*W) * sizeof(unsigned);
    H3D_REQUIRE(*smem <= 200 * 1024, "dist_max too large for the per-row bitmap");
    int g = div_up(n_bins, 8);
    const int cap = kNumSMs * 8;
    *grid = g < cap ? g : cap;
    return H3D_OK;
}

extern "C" int h3d_union_count(int n_reps, const void* const* indptr_host, int indptr_is64,
                               const int* const* indices_host, const void* const* data_host,
                               int data_dtype, const double* bias, int n_bins, int dist_max,
                               int* row_offset, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    CsrReps reps;
    int rc = fill_reps(&reps, n_reps, indptr_host, indices_host, data_host);
    if (rc) return rc;
    H3D_REQUIRE(n_bins >= 1 && dist_max >= 0, "empty chromosome");
    Workspace w(ws, ws_bytes);
    int* counts = w.take<int>(n_bins + 1);
    if (!counts) { set_error("union workspace too small"); return H3D_ERR_WORKSPACE; }
    int W, grid; size_t smem;
    rc = union_launch_shape(n_bins, dist_max, &W, &grid, &smem);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (smem > 48 * 1024)
        H3D_CHECK(cudaFuncSetAttribute(union_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    union_kernel<false><<<grid, 256, smem, st>>>(reps, n_reps, indptr_is64, data_dtype, bias, n_bins,
                                                 dist_max, W, counts, nullptr, nullptr, nullptr,
                                                 nullptr, nullptr, nullptr);
    H3D_LAUNCHED("union_kernel<count>");
    scan_exclusive_kernel<<<1, 1024, 0, st>>>(counts, row_offset, n_bins, nullptr);
    H3D_LAUNCHED("scan_exclusive_kernel");
    return H3D_OK;
}

extern "C" int h3d_union_emit(int n_reps, const void* const* indptr_host, int indptr_is64,
                              const int* const* indices_host, const void* const* data_host,
                              int data_dtype, const double* bias, int n_bins, int dist_max,
                              const int* row_offset, int* row, int* col, int* dist,
                              long long* raw, double* balanced, h3d_stream_t stream) {
    CsrReps reps;
    int rc = fill_reps(&reps, n_reps, indptr_host, indices_host, data_host);
    if (rc) return rc;
    int W, grid; size_t smem;
    rc = union_launch_shape(n_bins, dist_max, &W, &grid, &smem);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // staged fast path: the row's (dist_max + 1) x R values fit in shared memory
    const size_t warp_words = (size_t)((4 * W + 1) & ~1) + 2 * (size_t)(dist_max + 1) * n_reps;
    const size_t warp_bytes = warp_words * sizeof(unsigned);
    if (warp_bytes <= 56 * 1024) {
        int wpb = (int)((56 * 1024) / warp_bytes);
        if (wpb > 8) wpb = 8;
        const size_t sm = warp_bytes * wpb;
        if (sm > 48 * 1024)
            H3D_CHECK(cudaFuncSetAttribute(union_emit_staged_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int g = div_up(n_bins, wpb);
        const int cap = kNumSMs * 16;
        if (g > cap) g = cap;
        union_emit_staged_kernel<<<g, 32 * wpb, sm, st>>>(
            reps, n_reps, indptr_is64, data_dtype, bias, n_bins, dist_max, W, (int)warp_words,
            row_offset, row, col, dist, raw, balanced);
        H3D_LAUNCHED("union_emit_staged_kernel");
        return H3D_OK;
    }
    if (smem > 48 * 1024)
        H3D_CHECK(cudaFuncSetAttribute(union_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    union_kernel<true><<<grid, 256, smem, st>>>(reps, n_reps, indptr_is64, data_dtype, bias, n_bins,
                                                dist_max, W, nullptr, row_offset, row, col, dist,
                                                raw, balanced);
    H3D_LAUNCHED("union_kernel<emit>");
    return H3D_OK;
}

extern "C" int h3d_scale_filter(const int* row, const int* col, double* data,
                                const double* sf_table, int sf_per_dist,
                                const unsigned char* design_host, long long n_px, int n_reps,
                                int n_conds, int dist_max, double mean_thresh, int dist_min,
                                double* size_factors_out, unsigned char* disp_idx,
                                h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_conds >= 1 && n_conds <= H3D_MAX_CONDS, "n_conds out of range");
    if (n_px <= 0) return H3D_OK;
    DesignF dz;
    for (int c = 0; c < H3D_MAX_CONDS; ++c) dz.n_in_cond[c] = 0.0;
    for (int r = 0; r < H3D_MAX_REPS; ++r)
        for (int c = 0; c < H3D_MAX_CONDS; ++c) {
            const bool on = (r < n_reps && c < n_conds) && design_host[r * n_conds + c];
            dz.w[r][c] = on ? 1.0 : 0.0;
            if (on) dz.n_in_cond[c] += 1.0;
        }
    scale_filter_kernel<<<div_up(n_px, 256), 256, 0, (cudaStream_t)stream>>>(
        row, col, data, sf_table, sf_per_dist, dz, n_px, n_reps, n_conds, dist_max, mean_thresh,
        dist_min, size_factors_out, disp_idx);
    H3D_LAUNCHED("scale_filter_kernel");
    return H3D_OK;
}

extern "C" size_t h3d_mask_to_index_ws_bytes(long long n) {
    const long long tiles = (n + kCompactTile - 1) / kCompactTile;
    return 2 * ws_pad((size_t)(tiles + 1) * sizeof(int));
}

extern "C" int h3d_mask_to_index(const unsigned char* mask, long long n, int* index_out,
                                 long long* n_set_out, void* ws, size_t ws_bytes,
                                 h3d_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) {
        H3D_CHECK(cudaMemsetAsync(n_set_out, 0, sizeof(long long), st));
        return H3D_OK;
    }
    H3D_REQUIRE(n < 2147483647LL, "mask longer than int32 indexing");
    const int tiles = div_up(n, kCompactTile);
    Workspace w(ws, ws_bytes);
    int* counts = w.take<int>(tiles + 1);
    int* offsets = w.take<int>(tiles + 1);
    if (!counts || !offsets) { set_error("mask_to_index workspace too small"); return H3D_ERR_WORKSPACE; }
    mask_count_kernel<<<tiles, 256, 0, st>>>(mask, n, counts);
    H3D_LAUNCHED("mask_count_kernel");
    scan_exclusive_kernel<<<1, 1024, 0, st>>>(counts, offsets, tiles, n_set_out);
    H3D_LAUNCHED("scan_exclusive_kernel");
    mask_emit_kernel<<<tiles, 256, 0, st>>>(mask, n, offsets, index_out);
    H3D_LAUNCHED("mask_emit_kernel");
    return H3D_OK;
}

extern "C" int h3d_loop_membership(const int* row, const int* col, const int* index, long long n,
                                   const long long* sorted_keys, long long n_keys,
                                   unsigned char* out, h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    loop_membership_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(row, col, index, n,
                                                                            sorted_keys, n_keys, out);
    H3D_LAUNCHED("loop_membership_kernel");
    return H3D_OK;
}

extern "C" int h3d_gather_counts_factors(const int* row, const int* col, const int* index,
                                         long long n_sel, const long long* raw,
                                         const double* size_factors, int sf_per_pixel,
                                         const double* bias, int n_reps, const int* dest,
                                         long long ld, double* x_out, double* f_out,
                                         int* dist_out, h3d_stream_t stream) {
    if (n_sel <= 0) return H3D_OK;
    gather_counts_factors_kernel<<<div_up(n_sel, 256), 256, 0, (cudaStream_t)stream>>>(
        row, col, index, n_sel, raw, size_factors, sf_per_pixel, bias, n_reps, dest, ld, x_out,
        f_out, dist_out);
    H3D_LAUNCHED("gather_counts_factors_kernel");
    return H3D_OK;
}

extern "C" int h3d_gather_table(const int* dist, long long n, const double* table, int n_cols,
                                int n_rows, double* out, h3d_stream_t stream) {
    if (n <= 0) return H3D_OK;
    gather_table_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(dist, n, table, n_cols,
                                                                         n_rows, out);
    H3D_LAUNCHED("gather_table_kernel");
    return H3D_OK;
}
