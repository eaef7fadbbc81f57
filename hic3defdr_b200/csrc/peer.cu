// Pooling of the tested pixels by distance, fused with the multi-GPU exchange.
//
// hic3defdr/analysis/analysis.py:169-206 pools the tested pixels genome-wide
// per distance (raw[dist == d], f[dist == d]).  Here the pooled arrays are SoA
// (2 R, ld) matrices in (owner rank, distance, chromosome, row) order, built in
// two passes:
//   1. h3d_pool_index (per chromosome): each tested pixel drops an 8-byte
//      record {union index, chromosome << 24 | row} at its pooled position
//      (the stable rank by distance) -- the only scattered writes, 8 B/pixel;
//   2. h3d_pool_pull (all chromosomes): one thread per POOLED position reads
//      the record, gathers the pixel's R counts (one 8 R-byte run) and its
//      factors bias[row] bias[col] size_factor[d], and writes the 2 R values
//      coalesced (consecutive threads -> consecutive addresses in every row).
// With one process per GPU every distance is owned by one rank, and pass 2
// writes each pixel INTO THE OWNER'S BUFFER over NVLink: the receive buffers
// are cudaMalloc allocations shared through CUDA IPC, the position of a pixel
// in its owner's buffer follows from the all-gathered (rank x distance) counts,
// and the consumer reads a distance as one run per source rank
// (h3d_estimate_dispersion_runs).  Gather + exchange in one kernel, with
// coalesced remote stores (8-byte scattered stores over NVLink reached 83 GB/s
// on this box, the coalesced form runs at link speed).
#include "common.cuh"

namespace h3d {

constexpr int kMaxPeers = 16;
struct PeerBases { double* base[kMaxPeers]; };

__global__ void __launch_bounds__(256)
pool_index_kernel(const int* __restrict__ row, const int* __restrict__ index, long long n_sel,
                  int chrom_id, const int* __restrict__ dest, int2* __restrict__ rec) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel) return;
    const int u = index ? index[i] : (int)i;
    rec[dest[i]] = make_int2(u, (chrom_id << 24) | row[u]);
}

// per chromosome: {raw (int64 (N, R)), bias ((n_bins, R)), size factors, their form}
struct ChromSrc { const long long* raw; const double* bias; const double* sf; long long sf_mode; };
// sf_mode 0: (R,) per replicate; 1: (N, R) per union pixel; 2: (D + 1, R) per distance

__global__ void __launch_bounds__(256)
pool_pull_kernel(const int2* __restrict__ rec, long long n_local,
                 const long long* __restrict__ key_start, int n_keys,
                 const int* __restrict__ dist_of_key, const int* __restrict__ owner_of_key,
                 const long long* __restrict__ shift_of_key, const ChromSrc* __restrict__ chroms,
                 int n_reps, PeerBases peers, long long ld) {
    __shared__ int s_key;
    const long long p0 = (long long)blockIdx.x * blockDim.x;
    if (threadIdx.x == 0) {
        // last key whose start is <= p0 (keys may be empty)
        int lo = 0, hi = n_keys;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (key_start[mid] <= p0) lo = mid; else hi = mid;
        }
        s_key = lo;
    }
    __syncthreads();
    const long long p = p0 + threadIdx.x;
    if (p >= n_local) return;
    int key = s_key;
    while (key + 1 < n_keys && key_start[key + 1] <= p) ++key;
    const int d = dist_of_key[key];
    const int2 rc = rec[p];
    const long long u = rc.x;
    const int c = (int)((unsigned)rc.y >> 24), ri = rc.y & 0xFFFFFF, ci = ri + d;
    const ChromSrc src = chroms[c];
    double* __restrict__ out = peers.base[owner_of_key[key]];
    const long long pos = p + shift_of_key[key];
    for (int r = 0; r < n_reps; ++r) {
        const double s = src.sf_mode == 1 ? src.sf[u * n_reps + r]
                       : (src.sf_mode == 2 ? src.sf[(long long)d * n_reps + r] : src.sf[r]);
        out[(long long)r * ld + pos] = (double)src.raw[u * n_reps + r];
        out[(long long)(n_reps + r) * ld + pos] =
            src.bias[(long long)ri * n_reps + r] * src.bias[(long long)ci * n_reps + r] * s;
    }
}

// contiguous slices from local memory into (peer) buffers, 8-byte words,
// consecutive threads -> consecutive addresses (full-width NVLink stores)
constexpr int kMaxSlices = 16;
struct CopySlices {
    const unsigned long long* src[kMaxSlices];
    unsigned long long* dst[kMaxSlices];
    long long words[kMaxSlices];
};

__global__ void __launch_bounds__(256)
peer_copy_kernel(CopySlices sl) {
    const int k = blockIdx.y;
    const unsigned long long* __restrict__ src = sl.src[k];
    unsigned long long* __restrict__ dst = sl.dst[k];
    const long long n = sl.words[k];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = src[i];
}

}  // namespace h3d

using namespace h3d;

extern "C" int h3d_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out) {
    H3D_REQUIRE(bytes > 0 && ptr_out && handle_out, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == H3D_PEER_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    H3D_CHECK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return H3D_ERR_CUDA;
    }
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return H3D_OK;
}

extern "C" int h3d_peer_free(void* ptr) {
    if (ptr) H3D_CHECK(cudaFree(ptr));
    return H3D_OK;
}

extern "C" int h3d_peer_open(const unsigned char* handle, void** ptr_out) {
    H3D_REQUIRE(handle && ptr_out, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    H3D_CHECK(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return H3D_OK;
}

extern "C" int h3d_peer_close(void* ptr) {
    if (ptr) H3D_CHECK(cudaIpcCloseMemHandle(ptr));
    return H3D_OK;
}

extern "C" int h3d_pool_index(const int* row, const int* index, long long n_sel, int chrom_id,
                              const int* dest, int* rec, h3d_stream_t stream) {
    H3D_REQUIRE(chrom_id >= 0 && chrom_id < 256, "at most 256 chromosomes per rank");
    if (n_sel <= 0) return H3D_OK;
    pool_index_kernel<<<div_up(n_sel, 256), 256, 0, (cudaStream_t)stream>>>(row, index, n_sel, chrom_id,
                                                                           dest, (int2*)rec);
    H3D_LAUNCHED("pool_index_kernel");
    return H3D_OK;
}

extern "C" int h3d_pool_pull(const int* rec, long long n_local, const long long* key_start, int n_keys,
                             const int* dist_of_key, const int* owner_of_key,
                             const long long* shift_of_key, const void* chrom_table, int n_chroms,
                             int n_reps, void* const* peer_base_host, int n_ranks, long long ld,
                             h3d_stream_t stream) {
    H3D_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeers, "at most 16 ranks");
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_chroms >= 0 && n_chroms <= 256 && n_keys >= 1, "bad table sizes");
    static_assert(sizeof(ChromSrc) == 32, "chromosome table rows are four 8-byte words");
    if (n_local <= 0) return H3D_OK;
    PeerBases pb;
    for (int k = 0; k < kMaxPeers; ++k) pb.base[k] = (k < n_ranks) ? (double*)peer_base_host[k] : nullptr;
    pool_pull_kernel<<<div_up(n_local, 256), 256, 0, (cudaStream_t)stream>>>(
        (const int2*)rec, n_local, key_start, n_keys, dist_of_key, owner_of_key, shift_of_key,
        (const ChromSrc*)chrom_table, n_reps, pb, ld);
    H3D_LAUNCHED("pool_pull_kernel");
    return H3D_OK;
}

extern "C" int h3d_peer_copy(const void* src_base, const long long* src_off_host,
                             void* const* dst_base_host, const long long* dst_off_host,
                             const long long* bytes_host, int n_slices, h3d_stream_t stream) {
    H3D_REQUIRE(n_slices >= 0 && n_slices <= kMaxSlices, "at most 16 slices per call");
    CopySlices sl;
    long long longest = 0;
    for (int k = 0; k < kMaxSlices; ++k) {
        sl.src[k] = nullptr; sl.dst[k] = nullptr; sl.words[k] = 0;
        if (k >= n_slices) continue;
        H3D_REQUIRE(src_off_host[k] % 8 == 0 && dst_off_host[k] % 8 == 0 && bytes_host[k] % 8 == 0 &&
                    bytes_host[k] >= 0, "slices are multiples of 8 bytes");
        sl.src[k] = (const unsigned long long*)((const char*)src_base + src_off_host[k]);
        sl.dst[k] = (unsigned long long*)((char*)dst_base_host[k] + dst_off_host[k]);
        sl.words[k] = bytes_host[k] / 8;
        if (sl.words[k] > longest) longest = sl.words[k];
    }
    if (n_slices == 0 || longest == 0) return H3D_OK;
    int gx = div_up(longest, 256 * 8);
    const int cap = kNumSMs * 8 / n_slices + 1;
    if (gx > cap) gx = cap;
    peer_copy_kernel<<<dim3(gx, n_slices), 256, 0, (cudaStream_t)stream>>>(sl);
    H3D_LAUNCHED("peer_copy_kernel");
    return H3D_OK;
}
