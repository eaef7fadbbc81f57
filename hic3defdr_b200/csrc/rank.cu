// Stable rank by a small integer key, and the size factors built on it.
//
// Replaces hic3defdr/util/binning.py:4-25 (equal_bin; canonical stable
// tie-break), hic3defdr/util/scaling.py:27-149 (median_of_ratios,
// simple_scaling, conditional*), np.median / interp1d underneath them, and
// the ``raw[dist == d]`` pooling of hic3defdr/analysis/analysis.py:196-204.
//
// stable_rank: the classic three-phase counting sort restricted to what is
// needed (positions, not a permuted copy).  Each warp owns a contiguous
// segment of the input; phase A builds one key histogram per segment in
// shared memory, phase B scans the (segment x key) table down the segments
// and across the keys, phase C replays each segment in order and hands out
// positions with __match_any_sync so that equal keys keep their input order.
// Pixels arrive in (row, col) order, hence rank order == (distance, row)
// order: every distance (and every equal-count bin) becomes one contiguous
// range, which is what both the medians and the dispersion pooling want.
//
// medians: the ratios of one (bin, replicate) are a contiguous slice of the
// rank-ordered SoA buffer; one CTA runs an exact 5-digit (13 bit) MSD radix
// select for both middle order statistics over that slice (L2 resident after
// the first sweep).  HBM traffic: dist twice, balanced once, ratios written
// once and swept from L2.
#include "common.cuh"

namespace h3d {

// --------------------------------------------------------------------------
// stable rank
// --------------------------------------------------------------------------
struct RankPlan {
    long long n;
    int n_keys, seg_len, n_segs, warps_per_block;
    size_t table_bytes, smem_bytes;
};

static RankPlan make_rank_plan(long long n, int n_keys) {
    RankPlan p;
    p.n = n; p.n_keys = n_keys;
    long long t = 2048;
    const long long budget = 96LL << 20;     // bytes for the segment x key table
    // one warp per segment: keep >= 32 warps per SM in flight for small inputs
    // (a chromosome is 2-5 M pixels), as long as the table stays small
    while (t > 256 && (n + t - 1) / t < (long long)kNumSMs * 32 &&
           (n + t / 2 - 1) / (t / 2) * (long long)n_keys * 4 <= (8LL << 20)) t /= 2;
    while ((n + t - 1) / t * (long long)n_keys * 4 > budget) t *= 2;
    p.seg_len = (int)t;
    p.n_segs = (int)((n + t - 1) / t);
    if (p.n_segs < 1) p.n_segs = 1;
    // one counter array per warp in shared memory
    int wpb = 8;
    while (wpb > 1 && (size_t)wpb * n_keys * 4 > 160 * 1024) wpb /= 2;
    p.warps_per_block = wpb;
    p.smem_bytes = (size_t)wpb * n_keys * 4;
    p.table_bytes = (size_t)p.n_segs * n_keys * 4;
    return p;
}

// where a key comes from
struct IntKeys {
    const int* __restrict__ keys;
    __device__ __forceinline__ int operator()(long long i) const { return keys[i]; }
};
template <typename KeyFn>
__global__ void rank_hist_kernel(KeyFn keys, long long n, int n_keys,
                                 int seg_len, int n_segs, int* __restrict__ table,
                                 int* __restrict__ bad_key) {
    extern __shared__ int sh_cnt[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int* cnt = sh_cnt + (size_t)wid * n_keys;
    for (int seg = blockIdx.x * wpb + wid; seg < n_segs; seg += gridDim.x * wpb) {
        for (int k = lane; k < n_keys; k += 32) cnt[k] = 0;
        __syncwarp();
        const long long lo = (long long)seg * seg_len;
        const long long hi = (lo + seg_len < n) ? lo + seg_len : n;
        for (long long i = lo + lane; i < hi; i += 32) {
            const int k = keys(i);
            if (k < 0 || k >= n_keys) { *bad_key = 1; continue; }
            atomicAdd(&cnt[k], 1);
        }
        __syncwarp();
        for (int k = lane; k < n_keys; k += 32) table[(size_t)seg * n_keys + k] = cnt[k];
        __syncwarp();
    }
}

// Exclusive scan of the (segment x key) table down the segments, per key.
// Three small kernels so that every SM takes part whatever the shape (201 keys x
// 9 k segments for a chromosome, 2048 keys x 9 k segments for a BH pass):
//   A  grid (key tiles of 32, segment chunks): per-chunk column sums
//   B  one warp-lane per key: exclusive scan of the chunk sums, key totals
//   C  grid as A: exclusive scan inside the chunk, offset by the chunk's start
constexpr int kScanChunks = 64;

__global__ void __launch_bounds__(256)
rank_scan_sums_kernel(const int* __restrict__ table, int n_keys, int n_segs, int chunk_len,
                      long long* __restrict__ chunk_sum) {
    __shared__ long long part[8][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + lane;
    const int c0 = blockIdx.y * chunk_len;
    const int c1 = (c0 + chunk_len < n_segs) ? c0 + chunk_len : n_segs;
    long long sum = 0;
    if (k < n_keys)
        for (int s = c0 + wid; s < c1; s += 8) sum += table[(size_t)s * n_keys + k];
    part[wid][lane] = sum;
    __syncthreads();
    if (wid == 0 && k < n_keys) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += part[w][lane];
        chunk_sum[(size_t)blockIdx.y * n_keys + k] = t;
    }
}

__global__ void rank_scan_chunks_kernel(long long* __restrict__ chunk_sum, int n_keys, int n_chunks,
                                        long long* __restrict__ key_total) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_keys) return;
    long long run = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const long long t = chunk_sum[(size_t)c * n_keys + k];
        chunk_sum[(size_t)c * n_keys + k] = run;
        run += t;
    }
    key_total[k] = run;
}

__global__ void __launch_bounds__(256)
rank_scan_apply_kernel(int* __restrict__ table, int n_keys, int n_segs, int chunk_len,
                       const long long* __restrict__ chunk_sum) {
    __shared__ long long part[8][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + lane;
    const int c0 = blockIdx.y * chunk_len;
    const int c1 = (c0 + chunk_len < n_segs) ? c0 + chunk_len : n_segs;
    // warp w owns the contiguous slice [s0, s1) of the chunk
    const int per = (chunk_len + 7) / 8;
    const int s0 = c0 + wid * per;
    const int s1 = (s0 + per < c1) ? s0 + per : c1;
    long long sum = 0;
    if (k < n_keys)
        for (int s = s0; s < s1; ++s) sum += table[(size_t)s * n_keys + k];
    part[wid][lane] = sum;
    __syncthreads();
    long long run = (k < n_keys) ? chunk_sum[(size_t)blockIdx.y * n_keys + k] : 0;
    for (int w = 0; w < wid; ++w) run += part[w][lane];
    if (k < n_keys)
        for (int s = s0; s < s1; ++s) {
            const int t = table[(size_t)s * n_keys + k];
            table[(size_t)s * n_keys + k] = (int)run;
            run += t;
        }
}

// exclusive scan of per-key totals -> key_start[0..n_keys]
__global__ void __launch_bounds__(1024)
key_start_kernel(const long long* __restrict__ key_total, int n_keys, long long* __restrict__ key_start) {
    __shared__ long long sh[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_keys; base += 1024) {
        const int k = base + threadIdx.x;
        const long long v = (k < n_keys) ? key_total[k] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const long long t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (k < n_keys) key_start[k] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) key_start[n_keys] = carry;
}

// what happens with an element's position
struct StoreRank {
    int* __restrict__ rank_out;
    __device__ __forceinline__ void operator()(long long i, int pos) const { rank_out[i] = pos; }
};
template <typename KeyFn, typename Sink>
__global__ void rank_emit_kernel(KeyFn keys, long long n, int n_keys,
                                 int seg_len, int n_segs, const int* __restrict__ table,
                                 const long long* __restrict__ key_start, Sink sink) {
    extern __shared__ int sh_cnt[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int* cnt = sh_cnt + (size_t)wid * n_keys;
    const unsigned lt = (1u << lane) - 1u;
    for (int seg = blockIdx.x * wpb + wid; seg < n_segs; seg += gridDim.x * wpb) {
        for (int k = lane; k < n_keys; k += 32)
            cnt[k] = (int)key_start[k] + table[(size_t)seg * n_keys + k];
        __syncwarp();
        const long long lo = (long long)seg * seg_len;
        const long long hi = (lo + seg_len < n) ? lo + seg_len : n;
        // the key of the next round is loaded before this round's (serial)
        // match / count / store chain, so that the chain does not start with a
        // global-memory latency every 32 elements
        int k_next = (lo + lane < hi) ? keys(lo + lane) : -1;
        for (long long base = lo; base < hi; base += 32) {
            const long long i = base + lane;
            int k = k_next;
            k_next = (i + 32 < hi) ? keys(i + 32) : -1;
            if (k < 0 || k >= n_keys) k = -1;
            const unsigned act = __ballot_sync(0xffffffffu, k >= 0);
            if (k >= 0) {
                const unsigned peers = __match_any_sync(act, k);
                const int mine = cnt[k] + __popc(peers & lt);
                sink(i, mine);
                __syncwarp(act);
                if ((peers >> lane) == 1u) cnt[k] += __popc(peers);   // highest lane of the group
                __syncwarp(act);
            }
        }
        __syncwarp();
    }
}

template <typename KeyFn, typename Sink>
static int counting_rank(KeyFn keys, Sink sink, long long n, int n_keys, long long* key_start,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
    H3D_REQUIRE(n >= 0 && n < 2147483647LL, "n out of int32 range");
    H3D_REQUIRE(n_keys >= 1 && n_keys <= 40000, "n_keys out of range");
    const RankPlan p = make_rank_plan(n, n_keys);
    Workspace w(ws, ws_bytes);
    int* table = w.take<int>((size_t)p.n_segs * n_keys);
    long long* key_total = w.take<long long>(n_keys + 1);
    long long* chunk_sum = w.take<long long>((size_t)kScanChunks * n_keys);
    int* bad = w.take<int>(1);
    if (!table || !key_total || !chunk_sum || !bad) { set_error("stable_rank workspace too small"); return H3D_ERR_WORKSPACE; }
    H3D_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    const int threads = p.warps_per_block * 32;
    int grid = div_up(p.n_segs, p.warps_per_block);
    if (p.smem_bytes > 48 * 1024) {
        H3D_CHECK(cudaFuncSetAttribute(rank_hist_kernel<KeyFn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
        H3D_CHECK(cudaFuncSetAttribute(rank_emit_kernel<KeyFn, Sink>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    }
    rank_hist_kernel<KeyFn><<<grid, threads, p.smem_bytes, st>>>(keys, n, n_keys, p.seg_len, p.n_segs, table, bad);
    H3D_LAUNCHED("rank_hist_kernel");
    {
        int n_chunks = kScanChunks;
        if (n_chunks > p.n_segs) n_chunks = p.n_segs;
        const int chunk_len = div_up(p.n_segs, n_chunks);
        n_chunks = div_up(p.n_segs, chunk_len);
        dim3 sgrid(div_up(n_keys, 32), n_chunks);
        rank_scan_sums_kernel<<<sgrid, 256, 0, st>>>(table, n_keys, p.n_segs, chunk_len, chunk_sum);
        H3D_LAUNCHED("rank_scan_sums_kernel");
        rank_scan_chunks_kernel<<<div_up(n_keys, 128), 128, 0, st>>>(chunk_sum, n_keys, n_chunks, key_total);
        H3D_LAUNCHED("rank_scan_chunks_kernel");
        rank_scan_apply_kernel<<<sgrid, 256, 0, st>>>(table, n_keys, p.n_segs, chunk_len, chunk_sum);
        H3D_LAUNCHED("rank_scan_apply_kernel");
    }
    key_start_kernel<<<1, 1024, 0, st>>>(key_total, n_keys, key_start);
    H3D_LAUNCHED("key_start_kernel");
    rank_emit_kernel<KeyFn, Sink><<<grid, threads, p.smem_bytes, st>>>(keys, n, n_keys, p.seg_len, p.n_segs, table, key_start, sink);
    H3D_LAUNCHED("rank_emit_kernel");
    return H3D_OK;
}

int stable_rank_impl(const int* keys, long long n, int n_keys, int* rank_out,
                     long long* key_start, void* ws, size_t ws_bytes, cudaStream_t st) {
    return counting_rank(IntKeys{keys}, StoreRank{rank_out}, n, n_keys, key_start, ws, ws_bytes, st);
}

size_t stable_rank_ws(long long n, int n_keys) {
    const RankPlan p = make_rank_plan(n, n_keys);
    return ws_pad(p.table_bytes) + ws_pad((size_t)(n_keys + 1) * 8) +
           ws_pad((size_t)kScanChunks * n_keys * 8) + ws_pad(4);
}

// --------------------------------------------------------------------------
// size factors
// --------------------------------------------------------------------------
constexpr unsigned long long kInvalidKey = 0xFFFFFFFFFFFFFFFFull;

// group g covers ranks [gstart[g], gstart[g+1])
__global__ void group_bounds_kernel(long long n, int n_groups, int n_bins, int mode_exact,
                                    const long long* __restrict__ key_start,
                                    long long* __restrict__ gstart) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (n_bins <= 0) {                    // exact distance groups, or a single global group
        gstart[g] = mode_exact ? key_start[g] : (g == 0 ? 0 : n);
        return;
    }
    // equal_bin: bin(rank) = floor(rank * (n_bins / n)) as numpy.linspace(0, n_bins, n,
    // endpoint=False, dtype=int) evaluates it; first rank whose bin is >= g
    const double step = (double)n_bins / (double)n;
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)floor((double)mid * step) < (long long)g) lo = mid + 1; else hi = mid;
    }
    gstart[g] = lo;
}

// ratio (or plain value) of every pixel and replicate, scattered to rank order
__global__ void __launch_bounds__(256)
ratio_scatter_kernel(const double* __restrict__ balanced, const int* __restrict__ rank,
                     long long n, int n_reps, int want_ratio,
                     unsigned long long* __restrict__ sorted) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long pos = rank ? (long long)rank[p] : p;
    double v[H3D_MAX_REPS];
    bool ok = true;
    double slog = 0.0;
    for (int r = 0; r < n_reps; ++r) {
        v[r] = balanced[p * n_reps + r];
        ok = ok && (v[r] > 0.0);
        slog += log(v[r] + 1.0);
    }
    if (want_ratio) {
        const double gm = exp(slog / (double)n_reps) - 1.0;      // gmean, pseudocount 1
        for (int r = 0; r < n_reps; ++r)
            sorted[(long long)r * n + pos] =
                ok ? (unsigned long long)__double_as_longlong(v[r] / gm) : kInvalidKey;
    } else {
        for (int r = 0; r < n_reps; ++r)
            sorted[(long long)r * n + pos] = (unsigned long long)__double_as_longlong(v[r]);
    }
}

// Exact median of one (group, replicate) slice by MSD radix select.
// grid = (n_groups, n_reps), NT threads, 64 KB dynamic shared memory.  NT = 512
// lets two CTAs share an SM: the 160 CTAs of a chromosome (40 bins x 4
// replicates) then run as one wave on 148 SMs instead of one and a tail.
template <int NT>
__global__ void __launch_bounds__(NT)
median_select_kernel(const unsigned long long* __restrict__ sorted, long long n,
                     const long long* __restrict__ gstart, double* __restrict__ med_out,
                     long long* __restrict__ valid_out, int n_reps) {
    extern __shared__ unsigned hist[];          // [2][8192]
    __shared__ unsigned scan_part[NT];
    __shared__ unsigned long long prefix[2];
    __shared__ long long kth[2];
    __shared__ int done_digit[2];
    const int g = blockIdx.x, r = blockIdx.y, t = threadIdx.x;
    const long long lo = gstart[g], hi = gstart[g + 1];
    const unsigned long long* __restrict__ seg = sorted + (long long)r * n;
    const int shifts[5] = {51, 38, 25, 12, 0};
    const int widths[5] = {13, 13, 13, 13, 12};
    if (t < 2) { prefix[t] = 0; kth[t] = 0; }
    long long m_valid = 0;
    for (int pass = 0; pass < 5; ++pass) {
        for (int b = t; b < 2 * 8192; b += NT) hist[b] = 0;
        __syncthreads();
        const int sh = shifts[pass], wd = widths[pass];
        const unsigned long long p0 = prefix[0], p1 = prefix[1];
        const unsigned dmask = (1u << wd) - 1u;
        const bool same = (p0 == p1);
        for (long long i = lo + t; i < hi; i += NT) {
            const unsigned long long key = seg[i];
            const unsigned long long hi_bits = (pass == 0) ? 0ull : (key >> (sh + wd));
            const unsigned dg = (unsigned)(key >> sh) & dmask;
            if (pass == 0 || hi_bits == p0) atomicAdd(&hist[dg], 1u);
            if (!same && hi_bits == p1) atomicAdd(&hist[8192 + dg], 1u);
        }
        __syncthreads();
        if (pass == 0) {
            // invalid keys (all ones) sit in the top bucket of digit 0
            m_valid = (hi - lo) - (long long)hist[8191];
            if (t == 0) { kth[0] = (m_valid - 1) / 2; kth[1] = m_valid / 2; }
            __syncthreads();
            if (m_valid <= 0) break;
        }
        // locate the digit holding the k-th element for both problems
        for (int q = 0; q < 2; ++q) {
            const unsigned* h = hist + ((q == 1 && !same) ? 8192 : 0);
            const int nb = 1 << wd, per = nb / NT > 0 ? nb / NT : 1;
            unsigned s = 0;
            for (int b = 0; b < per; ++b) { const int idx = t * per + b; if (idx < nb) s += h[idx]; }
            scan_part[t] = s;
            __syncthreads();
            for (int o = 1; o < NT; o <<= 1) {
                const unsigned v = (t >= o) ? scan_part[t - o] : 0;
                __syncthreads();
                scan_part[t] += v;
                __syncthreads();
            }
            const long long before = (long long)scan_part[t] - (long long)s;
            const long long k = kth[q];
            __syncthreads();                      // everyone has read kth[q] before the owner rewrites it
            if (k >= before && k < before + (long long)s) {
                long long run = before;
                for (int b = 0; b < per; ++b) {
                    const int idx = t * per + b;
                    const long long c = (idx < nb) ? (long long)h[idx] : 0;
                    if (k < run + c) { done_digit[q] = idx; kth[q] = k - run; break; }
                    run += c;
                }
            }
            __syncthreads();
        }
        if (t < 2) prefix[t] = (prefix[t] << wd) | (unsigned long long)done_digit[t];
        __syncthreads();
    }
    if (t == 0) {
        double med = NAN;
        if (m_valid > 0) {
            const double a = __longlong_as_double((long long)prefix[0]);
            const double b = __longlong_as_double((long long)prefix[1]);
            med = (m_valid & 1) ? a : (a + b) / 2.0;      // np.median: mean of the two middles
        }
        med_out[(long long)g * n_reps + r] = med;
        if (r == 0) valid_out[g] = m_valid;
    }
}

// deterministic sum of one (group, replicate) slice (simple_scaling reducer)
__global__ void __launch_bounds__(1024)
group_sum_kernel(const unsigned long long* __restrict__ sorted, long long n,
                 const long long* __restrict__ gstart, double* __restrict__ sum_out, int n_reps) {
    __shared__ double sh[1024];
    const int g = blockIdx.x, r = blockIdx.y, t = threadIdx.x;
    const long long lo = gstart[g], hi = gstart[g + 1];
    const unsigned long long* __restrict__ seg = sorted + (long long)r * n;
    double s = 0.0;
    for (long long i = lo + t; i < hi; i += 1024) s += __longlong_as_double((long long)seg[i]);
    sh[t] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (t < o) sh[t] += sh[t + o];
        __syncthreads();
    }
    if (t == 0) sum_out[(long long)g * n_reps + r] = sh[0];
}

// builds the (dist_max + 1, n_reps) table (conditional modes) or the (n_reps)
// vector; single block.
__global__ void __launch_bounds__(256)
sf_table_kernel(const double* __restrict__ red, const long long* __restrict__ gstart,
                const long long* __restrict__ key_start, int n_groups, int n_reps, int dist_max,
                int n_bins, int norm, double* __restrict__ d_b, double* __restrict__ s_b,
                int* __restrict__ n_occ_out, double* __restrict__ table) {
    __shared__ int n_occ;
    const int t = threadIdx.x;
    const bool scaling = (norm == H3D_NORM_CONDITIONAL_SCALING || norm == H3D_NORM_SIMPLE_SCALING);
    // mean distance of every group (scaling.py:94): exact integer sum / count;
    // one thread per group, written to the (still uncompacted) slot g of d_b
    for (int g = t; g < n_groups; g += blockDim.x) {
        const long long lo = gstart[g], hi = gstart[g + 1];
        long long sumd = 0;
        if (key_start && hi > lo) {
            for (int d = 0; d <= dist_max; ++d) {
                const long long a = key_start[d] > lo ? key_start[d] : lo;
                const long long b = key_start[d + 1] < hi ? key_start[d + 1] : hi;
                if (b > a) sumd += (long long)d * (b - a);
            }
        }
        d_b[g] = (hi > lo) ? (double)sumd / (double)(hi - lo) : 0.0;     // compacted in place below (k <= g)
    }
    __syncthreads();
    if (t == 0) {
        int k = 0;
        for (int g = 0; g < n_groups; ++g) {
            const long long lo = gstart[g], hi = gstart[g + 1];
            if (hi <= lo) continue;                       // np.unique(bins): occupied bins only
            d_b[k] = d_b[g];
            if (scaling) {
                // simple_scaling: s / gmean(s) with pseudocount 1 (scaling.py:64-65)
                double sl = 0.0;
                for (int r = 0; r < n_reps; ++r) sl += log(red[(long long)g * n_reps + r] + 1.0);
                const double gm = exp(sl / (double)n_reps) - 1.0;
                for (int r = 0; r < n_reps; ++r) s_b[(long long)k * n_reps + r] = red[(long long)g * n_reps + r] / gm;
            } else {
                for (int r = 0; r < n_reps; ++r) s_b[(long long)k * n_reps + r] = red[(long long)g * n_reps + r];
            }
            ++k;
        }
        n_occ = k;
        *n_occ_out = k;
    }
    __syncthreads();
    const int k = n_occ;
    if (norm == H3D_NORM_MEDIAN_OF_RATIOS || norm == H3D_NORM_SIMPLE_SCALING) {
        for (int r = t; r < n_reps; r += blockDim.x) table[r] = (k > 0) ? s_b[r] : NAN;
        return;
    }
    for (int e = t; e < (dist_max + 1) * n_reps; e += blockDim.x) {
        const int d = e / n_reps, r = e % n_reps;
        double out = NAN;
        if (n_bins <= 0) {
            // exact-distance mode (scaling.py:101-104): the group of distance d itself
            int idx = -1, seen = 0;
            for (int g = 0; g < n_groups; ++g) {
                if (gstart[g + 1] > gstart[g]) { if (g == d) idx = seen; ++seen; }
            }
            if (idx >= 0) out = s_b[(long long)idx * n_reps + r];
        } else if (k == 1) {
            out = s_b[r];
        } else if (k >= 2) {
            // interp1d(kind='linear', fill_value='extrapolate') as scipy evaluates it
            const double x = (double)d;
            int hi = 0;                                   // searchsorted(d_b, x), side='left'
            while (hi < k && d_b[hi] < x) ++hi;
            if (hi < 1) hi = 1;
            if (hi > k - 1) hi = k - 1;
            const int lo = hi - 1;
            const double xl = d_b[lo], xh = d_b[hi];
            const double yl = s_b[(long long)lo * n_reps + r], yh = s_b[(long long)hi * n_reps + r];
            out = __dadd_rn(__dmul_rn((x - xl) / (xh - xl), yh), __dmul_rn((xh - x) / (xh - xl), yl));
        }
        table[e] = out;
    }
}

}  // namespace h3d

using namespace h3d;

extern "C" size_t h3d_stable_rank_ws_bytes(long long n, int n_keys) { return stable_rank_ws(n, n_keys); }

extern "C" int h3d_stable_rank(const int* keys, long long n, int n_keys, int* rank_out,
                               long long* key_start, void* ws, size_t ws_bytes,
                               h3d_stream_t stream) {
    return stable_rank_impl(keys, n, n_keys, rank_out, key_start, ws, ws_bytes, (cudaStream_t)stream);
}

static int sf_groups(int dist_max, int n_bins, int norm) {
    if (norm == H3D_NORM_MEDIAN_OF_RATIOS || norm == H3D_NORM_SIMPLE_SCALING) return 1;
    return n_bins > 0 ? n_bins : dist_max + 1;
}

extern "C" size_t h3d_size_factors_ws_bytes(long long n_px, int n_reps, int dist_max) {
    const int gmax = dist_max + 2;
    return stable_rank_ws(n_px, dist_max + 1) + ws_pad((size_t)n_px * 4) +
           ws_pad((size_t)n_px * n_reps * 8) + ws_pad((size_t)(dist_max + 2) * 8) +
           4 * ws_pad((size_t)(gmax + 1) * 8) + 3 * ws_pad((size_t)gmax * n_reps * 8) + ws_pad(64);
}

// the four stages of the size-factor computation; h3d_size_factors chains them on
// one device, hic3defdr_b200/dist.py puts collectives between them when the pixels
// of one chromosome are sharded over several GPUs by row range
static int sf_group_bounds(long long n_total, int dist_max, int n_bins, int norm,
                           const long long* key_start, long long* gstart, cudaStream_t st) {
    const bool conditional = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_CONDITIONAL_SCALING);
    if (!conditional) n_bins = 0;
    const int n_groups = sf_groups(dist_max, n_bins, norm);
    group_bounds_kernel<<<div_up(n_groups + 1, 128), 128, 0, st>>>(
        n_total, n_groups, n_bins, conditional ? 1 : 0, conditional ? key_start : nullptr, gstart);
    H3D_LAUNCHED("group_bounds_kernel");
    return H3D_OK;
}

static int sf_values(const double* balanced, const int* rank, long long n_px, int n_reps, int norm,
                     unsigned long long* values, cudaStream_t st) {
    const bool want_ratio = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_MEDIAN_OF_RATIOS);
    if (n_px <= 0) return H3D_OK;
    ratio_scatter_kernel<<<div_up(n_px, 256), 256, 0, st>>>(balanced, rank, n_px, n_reps,
                                                           want_ratio ? 1 : 0, values);
    H3D_LAUNCHED("ratio_scatter_kernel");
    return H3D_OK;
}

static int sf_group_reduce(const unsigned long long* values, long long ld, const long long* gstart,
                           int n_groups, int n_reps, int norm, double* red, long long* valid,
                           cudaStream_t st) {
    const bool want_ratio = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_MEDIAN_OF_RATIOS);
    if (n_groups <= 0) return H3D_OK;
    dim3 grid(n_groups, n_reps);
    if (want_ratio) {
        const size_t smem = 2 * 8192 * sizeof(unsigned);
        H3D_CHECK(cudaFuncSetAttribute(median_select_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        median_select_kernel<512><<<grid, 512, smem, st>>>(values, ld, gstart, red, valid, n_reps);
        H3D_LAUNCHED("median_select_kernel");
    } else {
        group_sum_kernel<<<grid, 1024, 0, st>>>(values, ld, gstart, red, n_reps);
        H3D_LAUNCHED("group_sum_kernel");
    }
    return H3D_OK;
}

extern "C" size_t h3d_sf_table_ws_bytes(int n_groups, int n_reps) {
    return ws_pad((size_t)(n_groups + 1) * 8) + ws_pad((size_t)n_groups * n_reps * 8) + ws_pad(64);
}

static int sf_table(const double* red, const long long* gstart, const long long* key_start,
                    int n_groups, int n_reps, int dist_max, int n_bins, int norm, double* table,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
    const bool conditional = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_CONDITIONAL_SCALING);
    if (!conditional) n_bins = 0;
    Workspace w(ws, ws_bytes);
    double* d_b = w.take<double>(n_groups + 1);
    double* s_b = w.take<double>((size_t)n_groups * n_reps);
    int* n_occ = w.take<int>(1);
    if (!d_b || !s_b || !n_occ) {
        set_error("sf_table workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    sf_table_kernel<<<1, 256, 0, st>>>(red, gstart, conditional ? key_start : nullptr, n_groups, n_reps,
                                       dist_max, n_bins, norm, d_b, s_b, n_occ, table);
    H3D_LAUNCHED("sf_table_kernel");
    return H3D_OK;
}

extern "C" int h3d_size_factors(const int* dist, const double* balanced, long long n_px,
                                int n_reps, int dist_max, int n_bins, int norm, double* sf_table_out,
                                void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(norm >= 0 && norm <= 3, "unknown norm");
    H3D_REQUIRE(n_px >= 1, "no pixels");
    cudaStream_t st = (cudaStream_t)stream;
    const bool conditional = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_CONDITIONAL_SCALING);
    if (!conditional) n_bins = 0;
    const int n_groups = sf_groups(dist_max, n_bins, norm);
    Workspace w(ws, ws_bytes);
    const size_t rank_ws_bytes = stable_rank_ws(n_px, dist_max + 1);
    void* rank_ws = w.take<char>(rank_ws_bytes);
    int* rank = w.take<int>(n_px);
    unsigned long long* sorted = w.take<unsigned long long>((size_t)n_px * n_reps);
    long long* key_start = w.take<long long>(dist_max + 2);
    long long* gstart = w.take<long long>(n_groups + 1);
    long long* valid = w.take<long long>(n_groups + 1);
    double* red = w.take<double>((size_t)n_groups * n_reps);
    const size_t table_ws_bytes = h3d_sf_table_ws_bytes(n_groups, n_reps);
    void* table_ws = w.take<char>(table_ws_bytes);
    if (!rank_ws || !rank || !sorted || !key_start || !gstart || !valid || !red || !table_ws) {
        set_error("size_factors workspace too small");
        return H3D_ERR_WORKSPACE;
    }
    int rc;
    if (conditional) {
        rc = stable_rank_impl(dist, n_px, dist_max + 1, rank, key_start, rank_ws, rank_ws_bytes, st);
        if (rc) return rc;
    }
    if ((rc = sf_group_bounds(n_px, dist_max, n_bins, norm, key_start, gstart, st))) return rc;
    if ((rc = sf_values(balanced, conditional ? rank : nullptr, n_px, n_reps, norm, sorted, st))) return rc;
    if ((rc = sf_group_reduce(sorted, n_px, gstart, n_groups, n_reps, norm, red, valid, st))) return rc;
    return sf_table(red, gstart, key_start, n_groups, n_reps, dist_max, n_bins, norm, sf_table_out,
                    table_ws, table_ws_bytes, st);
}

// ---- the same stages, one entry point each (row-sharded chromosomes) -------

extern "C" int h3d_sf_num_groups(int dist_max, int n_bins, int norm) {
    const bool conditional = (norm == H3D_NORM_CONDITIONAL_MOR || norm == H3D_NORM_CONDITIONAL_SCALING);
    return sf_groups(dist_max, conditional ? n_bins : 0, norm);
}

extern "C" int h3d_sf_group_bounds(long long n_total, int dist_max, int n_bins, int norm,
                                   const long long* key_start, long long* gstart,
                                   h3d_stream_t stream) {
    H3D_REQUIRE(norm >= 0 && norm <= 3, "unknown norm");
    H3D_REQUIRE(n_total >= 1, "no pixels");
    return sf_group_bounds(n_total, dist_max, n_bins, norm, key_start, gstart, (cudaStream_t)stream);
}

extern "C" int h3d_sf_values(const double* balanced, const int* rank, long long n_px, int n_reps,
                             int norm, double* values, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(norm >= 0 && norm <= 3, "unknown norm");
    return sf_values(balanced, rank, n_px, n_reps, norm, (unsigned long long*)values,
                     (cudaStream_t)stream);
}

extern "C" int h3d_sf_group_reduce(const double* values, long long ld, const long long* gstart,
                                   int n_groups, int n_reps, int norm, double* red,
                                   long long* valid, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(norm >= 0 && norm <= 3, "unknown norm");
    return sf_group_reduce((const unsigned long long*)values, ld, gstart, n_groups, n_reps, norm,
                           red, valid, (cudaStream_t)stream);
}

extern "C" int h3d_sf_table(const double* red, const long long* gstart, const long long* key_start,
                            int n_groups, int n_reps, int dist_max, int n_bins, int norm,
                            double* sf_table_out, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(norm >= 0 && norm <= 3, "unknown norm");
    H3D_REQUIRE(n_groups == h3d_sf_num_groups(dist_max, n_bins, norm), "n_groups does not match the mode");
    return sf_table(red, gstart, key_start, n_groups, n_reps, dist_max, n_bins, norm, sf_table_out,
                    ws, ws_bytes, (cudaStream_t)stream);
}
