// Stable LSD radix sort of (64-bit key, 32-bit payload) pairs, one sweep per
// 8-bit digit ("onesweep" organisation), used by the Benjamini-Hochberg step
// (bh.cu; hic3defdr/analysis/analysis.py:300 -> lib5c adjust_pvalues, whose
// argsort this replaces).
//
// HBM-bound: per digit every pair is read once and written once (24 B), plus
// one up-front pass that builds the histograms of all eight digits (8 B):
// 200 B per pair in total, against 444 B for the rank-table formulation it
// replaces (ncu, profiles/r01h).  Per sweep a CTA
//   1. takes the next tile (atomic ticket: tiles start in order, which makes
//      the look-back below deadlock free),
//   2. ranks its 4096 keys by digit, stably: each warp owns a contiguous run,
//      __match_any_sync groups equal digits inside a round of 32, per-warp
//      counters in shared memory carry the order across rounds,
//   3. publishes the tile's digit counts and obtains, by decoupled look-back
//      over the preceding tiles' status words, the number of equal digits in
//      all earlier tiles (one thread per digit),
//   4. reorders the tile in shared memory and writes it out digit by digit:
//      consecutive threads write consecutive addresses inside a digit's run
//      (~128-byte runs instead of single 12-byte scatters).
#include "common.cuh"

namespace h3d {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;     // 4096 pairs
constexpr int kSortPasses = 8;
constexpr unsigned kFlagAgg = 1u << 30, kFlagIncl = 2u << 30, kValMask = (1u << 30) - 1u;

// histograms of all eight digits in one pass over the keys
__global__ void __launch_bounds__(256)
sort_hist_kernel(const unsigned long long* __restrict__ keys, long long n,
                 unsigned* __restrict__ hist /* [8][256] */) {
    __shared__ unsigned sh[kSortPasses * 256];
    for (int k = threadIdx.x; k < kSortPasses * 256; k += 256) sh[k] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * 256;
    const long long n_round = (n + 31) / 32 * 32;           // whole warps take part in the votes
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        const unsigned long long key = valid ? keys[i] : 0ull;
#pragma unroll
        for (int p = 0; p < kSortPasses; ++p) {
            const unsigned d = (unsigned)(key >> (8 * p)) & 255u;
            if (p < 5) {
                // low mantissa bytes of p-values: spread out, plain atomics
                if (valid) atomicAdd(&sh[p * 256 + d], 1u);
            } else {
                // exponent bytes: nearly constant -> one atomic per group of equal digits
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256u);
                if (valid && (peers >> lane) == 1u) atomicAdd(&sh[p * 256 + d], (unsigned)__popc(peers));
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kSortPasses * 256; k += 256)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

// exclusive scan of each digit's histogram: first output position per value
__global__ void __launch_bounds__(256)
sort_scan_kernel(const unsigned* __restrict__ hist, unsigned* __restrict__ digit_start) {
    __shared__ unsigned sh[256];
    const int p = blockIdx.x;
    const unsigned v = hist[p * 256 + threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const unsigned t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    digit_start[p * 256 + threadIdx.x] = sh[threadIdx.x] - v;
}

__device__ __forceinline__ unsigned ld_status(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// one sweep: stable partition of the pairs by the digit (key >> shift) & 255.
// first != 0: the payload is the element's own position (no payload array yet).
__global__ void __launch_bounds__(kSortThreads, 3)
sort_pass_kernel(const unsigned long long* __restrict__ keys_in, const int* __restrict__ idx_in,
                 unsigned long long* __restrict__ keys_out, int* __restrict__ idx_out, long long n,
                 int shift, int first, const unsigned* __restrict__ digit_start /* [256] of this pass */,
                 unsigned* __restrict__ status /* [tiles][256] of this pass, zeroed */,
                 unsigned* __restrict__ ticket) {
    extern __shared__ unsigned char sort_smem[];
    unsigned long long* s_keys = (unsigned long long*)sort_smem;
    int* s_idx = (int*)(sort_smem + (size_t)kSortTile * 8);
    __shared__ unsigned warp_cnt[8][256];
    __shared__ unsigned tile_start[256];
    __shared__ long long glob_base[256];
    __shared__ unsigned warp_tot[8];
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    if (tid == 0) s_tile = (int)atomicAdd(ticket, 1u);
#pragma unroll
    for (int k = 0; k < 8; ++k) warp_cnt[k][tid] = 0;
    __syncthreads();
    const int tile = s_tile;
    const long long base = (long long)tile * kSortTile;
    const int cnt = (int)((n - base < kSortTile) ? n - base : kSortTile);

    unsigned long long key[kSortItems];
    int idx[kSortItems];
    unsigned lrank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int j = w * (kSortItems * 32) + r * 32 + lane;
        const bool valid = j < cnt;
        key[r] = valid ? keys_in[base + j] : 0xFFFFFFFFFFFFFFFFull;
        idx[r] = valid ? (first ? (int)(base + j) : idx_in[base + j]) : 0;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const unsigned d = (unsigned)(key[r] >> shift) & 255u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        lrank[r] = warp_cnt[w][d] + (unsigned)__popc(peers & lt);
        __syncwarp();
        if ((peers >> lane) == 1u) warp_cnt[w][d] += (unsigned)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // thread d: digit d's counts per warp -> exclusive over the warps; tile total
    unsigned tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned c = warp_cnt[k][tid];
        warp_cnt[k][tid] = tot;
        tot += c;
    }
    // exclusive scan of the tile's digit totals over the 256 digits
    unsigned incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    unsigned before = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) before += (k < w) ? warp_tot[k] : 0u;
    const unsigned my_start = before + incl - tot;
    tile_start[tid] = my_start;
    // decoupled look-back: equal digits in all earlier tiles
    {
        unsigned* mine = status + (size_t)tile * 256 + tid;
        unsigned prefix = 0;
        if (tile == 0) {
            st_status(mine, kFlagIncl | tot);
        } else {
            st_status(mine, kFlagAgg | tot);
            int t = tile - 1;
            while (true) {
                const unsigned v = ld_status(status + (size_t)t * 256 + tid);
                if ((v >> 30) == 0u) continue;
                prefix += v & kValMask;
                if ((v >> 30) == 2u) break;
                --t;
            }
            st_status(mine, kFlagIncl | (prefix + tot));
        }
        glob_base[tid] = (long long)digit_start[tid] + (long long)prefix - (long long)my_start;
    }
    __syncthreads();
    // reorder inside the tile
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const unsigned d = (unsigned)(key[r] >> shift) & 255u;
        const unsigned pos = tile_start[d] + warp_cnt[w][d] + lrank[r];
        s_keys[pos] = key[r];
        s_idx[pos] = idx[r];
    }
    __syncthreads();
    for (int j = tid; j < cnt; j += kSortThreads) {
        const unsigned long long k = s_keys[j];
        const unsigned d = (unsigned)(k >> shift) & 255u;
        const long long g = glob_base[d] + j;
        keys_out[g] = k;
        idx_out[g] = s_idx[j];
    }
}

size_t sort_pairs_ws(long long n) {
    const long long tiles = (n + kSortTile - 1) / kSortTile;
    return ws_pad((size_t)kSortPasses * 256 * 4) * 2 + ws_pad((size_t)kSortPasses * 4) +
           ws_pad((size_t)kSortPasses * tiles * 256 * 4);
}

// Sorts n pairs by key (ascending, stable).  keys_a holds the keys on entry and
// the sorted keys on return; idx_a receives the payload = original position;
// keys_b / idx_b are scratch of the same size.  n < 2^30.
int sort_pairs_u64(unsigned long long* keys_a, int* idx_a, unsigned long long* keys_b, int* idx_b,
                   long long n, void* ws, size_t ws_bytes, cudaStream_t st) {
    H3D_REQUIRE(n >= 1 && n < (1LL << 30), "sort size out of range");
    const long long tiles = (n + kSortTile - 1) / kSortTile;
    Workspace w(ws, ws_bytes);
    unsigned* hist = w.take<unsigned>(kSortPasses * 256);
    unsigned* digit_start = w.take<unsigned>(kSortPasses * 256);
    unsigned* ticket = w.take<unsigned>(kSortPasses);
    unsigned* status = w.take<unsigned>((size_t)kSortPasses * tiles * 256);
    if (!hist || !digit_start || !ticket || !status) { set_error("sort workspace too small"); return H3D_ERR_WORKSPACE; }
    // hist .. status are adjacent carvings: one memset clears them all
    H3D_CHECK(cudaMemsetAsync(hist, 0, (size_t)((char*)(status + (size_t)kSortPasses * tiles * 256) - (char*)hist), st));
    int hgrid = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (hgrid > kNumSMs * 8) hgrid = kNumSMs * 8;
    sort_hist_kernel<<<hgrid, 256, 0, st>>>(keys_a, n, hist);
    H3D_LAUNCHED("sort_hist_kernel");
    sort_scan_kernel<<<kSortPasses, 256, 0, st>>>(hist, digit_start);
    H3D_LAUNCHED("sort_scan_kernel");
    const size_t smem = (size_t)kSortTile * 12;
    H3D_CHECK(cudaFuncSetAttribute(sort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long *ki = keys_a, *ko = keys_b;
    int *ii = idx_a, *io = idx_b;
    for (int p = 0; p < kSortPasses; ++p) {
        sort_pass_kernel<<<(int)tiles, kSortThreads, smem, st>>>(
            ki, ii, ko, io, n, 8 * p, p == 0 ? 1 : 0, digit_start + p * 256,
            status + (size_t)p * tiles * 256, ticket + p);
        H3D_LAUNCHED("sort_pass_kernel");
        unsigned long long* tk = ki; ki = ko; ko = tk;
        int* ti = ii; ii = io; io = ti;
    }
    return H3D_OK;       // an even number of sweeps: the result is back in keys_a / idx_a
}

}  // namespace h3d
