// Shared plumbing of libh3d: error reporting across the C ABI, the launch
// counter, small device helpers.  No torch types anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/h3d.h"
#include "h3d_math.cuh"

namespace h3d {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

constexpr int kNumSMs = 148;   // B200

#define H3D_CHECK(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            h3d::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, \
                           cudaGetErrorString(_e));                            \
            return H3D_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

// call right after a kernel launch
#define H3D_LAUNCHED(name)                                                     \
    do {                                                                       \
        h3d::count_launch();                                                   \
        cudaError_t _e = cudaGetLastError();                                   \
        if (_e != cudaSuccess) {                                               \
            h3d::set_error("launch of %s failed at %s:%d: %s", name, __FILE__, \
                           __LINE__, cudaGetErrorString(_e));                  \
            return H3D_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

#define H3D_REQUIRE(cond, msg)                                                 \
    do {                                                                       \
        if (!(cond)) {                                                         \
            h3d::set_error("invalid argument: %s (%s)", msg, #cond);           \
            return H3D_ERR_ARG;                                                \
        }                                                                      \
    } while (0)

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// bump allocator over a caller-owned workspace
struct Workspace {
    char* base;
    size_t size, used;
    __host__ Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {}
    template <typename T>
    __host__ T* take(size_t count) {
        size_t off = (used + 255) & ~(size_t)255;
        size_t bytes = count * sizeof(T);
        used = off + bytes;
        if (base == nullptr || used > size) return nullptr;
        return (T*)(base + off);
    }
};
static inline size_t ws_pad(size_t bytes) { return ((bytes + 255) & ~(size_t)255) + 256; }

// replicate-indexed pointer bundles passed to kernels by value
struct CsrReps {
    const void* indptr[kMaxReps];
    const int* indices[kMaxReps];
    const void* data[kMaxReps];
};

__device__ __forceinline__ long long load_indptr(const void* p, long long i, int is64) {
    return is64 ? ((const long long*)p)[i] : (long long)((const int*)p)[i];
}

__device__ __forceinline__ double load_value(const void* p, long long i, int dtype) {
    switch (dtype) {
        case H3D_DTYPE_I64: return (double)((const long long*)p)[i];
        case H3D_DTYPE_F64: return ((const double*)p)[i];
        case H3D_DTYPE_I32: return (double)((const int*)p)[i];
        default: return (double)((const float*)p)[i];
    }
}

// rank.cu: positions of the elements in a stable sort by an integer key
int stable_rank_impl(const int* keys, long long n, int n_keys, int* rank_out,
                     long long* key_start, void* ws, size_t ws_bytes, cudaStream_t st);
size_t stable_rank_ws(long long n, int n_keys);
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace h3d
