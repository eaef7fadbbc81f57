// Simulation of raw contact matrices.
//
// Replaces the per-replicate body of hic3defdr/util/simulation.py:177-202
// (simulate -> gen): f = bias[row] bias[col] size_factor, biased mean bm = m f,
// counts ~ NB(mean bm, variance bm + disp bm^2) (scaled_nb.mvr, util/
// scaled_nb.py:36-48), and the cluster perturbation of :12-67 once its
// per-pixel factors are known.  The reference draws with scipy's frozen nbinom
// from numpy's global generator; here every pixel owns a counter-based
// Philox4x32-10 stream keyed by (seed, replicate), and a negative binomial is
// drawn as a gamma-Poisson mixture: lambda ~ Gamma(1 / disp, disp bm)
// (Marsaglia-Tsang), count ~ Poisson(lambda) (multiplication method below 10,
// Hoermann's transformed rejection PTRS above).  The draws are a different
// realisation of the same distribution; parity is checked on moments and on
// goodness of fit (tests/test_gpu_simulate.py), the deterministic parts
// (classes, perturbed means, file layout) exactly.
#include "common.cuh"

namespace h3d {

struct Philox {
    unsigned c0, c1, c2, c3, k0, k1;
    unsigned out[4];
    int have;
    __device__ __forceinline__ void init(unsigned long long seed, unsigned long long index, unsigned stream) {
        k0 = (unsigned)seed; k1 = (unsigned)(seed >> 32);
        c0 = 0; c1 = stream; c2 = (unsigned)index; c3 = (unsigned)(index >> 32);
        have = 0;
    }
    __device__ __forceinline__ void block() {
        unsigned x0 = c0, x1 = c1, x2 = c2, x3 = c3, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const unsigned hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const unsigned hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const unsigned y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
            x0 = y0; x1 = y1; x2 = y2; x3 = y3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
        c0 += 1;                       // 2^32 blocks per (pixel, stream): never exhausted
        have = 4;
    }
    // uniform in (0, 1), 53 bits
    __device__ __forceinline__ double uniform() {
        if (have < 2) block();
        const unsigned hi = out[have - 1], lo = out[have - 2];
        have -= 2;
        const unsigned long long bits = (((unsigned long long)hi << 32) | lo) >> 11;
        return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double normal() {
        const double u1 = uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
};

__device__ double sample_gamma(Philox& g, double shape, double scale) {
    double boost = 1.0;
    if (shape < 1.0) {                              // Gamma(a) = Gamma(a + 1) U^(1/a)
        boost = pow(g.uniform(), 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 1000; ++it) {
        const double x = g.normal();
        double v = 1.0 + c * x;
        if (v <= 0.0) continue;
        v = v * v * v;
        const double u = g.uniform();
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v)))
            return d * v * scale * boost;
    }
    return d * scale * boost;
}

__device__ long long sample_poisson(Philox& g, double lam) {
    if (!(lam > 0.0)) return 0;
    if (lam < 10.0) {
        const double limit = exp(-lam);
        long long k = 0;
        double p = g.uniform();
        while (p > limit) { ++k; p *= g.uniform(); }
        return k;
    }
    // PTRS, W. Hoermann, Insurance: Mathematics and Economics 12 (1993)
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
        const double U = g.uniform() - 0.5, V = g.uniform();
        const double us = 0.5 - fabs(U);
        const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return (long long)kf;
        if (kf < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0))
            return (long long)kf;
    }
    return (long long)floor(lam);
}

// sf_mode 0: size_factors (n_sim,); 1: (n_dist, n_sim) by distance.
// disp_mode 0: disp_table (n_dist,) by distance; 1: per-pixel disp array.
__global__ void __launch_bounds__(256)
nb_simulate_kernel(const int* __restrict__ row, const int* __restrict__ col,
                   const double* __restrict__ mean, long long n, const double* __restrict__ bias,
                   int n_sim, const double* __restrict__ sf, int sf_mode, int n_dist,
                   const double* __restrict__ disp, int disp_mode, int rep,
                   unsigned long long seed, long long* __restrict__ out,
                   double* __restrict__ biased_mean_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = row[i], c = col[i];
    int d = c - r;
    if (d < 0) d = -d;
    const double s = sf_mode ? sf[(long long)(d < n_dist ? d : n_dist - 1) * n_sim + rep] : sf[rep];
    const double f = bias[(long long)r * n_sim + rep] * bias[(long long)c * n_sim + rep] * s;
    const double bm = mean[i] * f;
    if (biased_mean_out) biased_mean_out[i] = bm;
    if (out == nullptr) return;                    // biased means only
    const double phi = disp_mode ? disp[i] : disp[d < n_dist ? d : n_dist - 1];
    Philox g;
    g.init(seed, (unsigned long long)i, (unsigned)rep);
    double lam = bm;
    if (phi > 0.0) lam = sample_gamma(g, 1.0 / phi, phi * bm);   // mean bm, variance phi bm^2
    out[i] = sample_poisson(g, lam);
}

// mean[idx(key)] *= factor for a list of (pixel key, factor); pixel keys sorted
__global__ void __launch_bounds__(256)
perturb_kernel(const long long* __restrict__ pixel_keys, long long n_px,
               const long long* __restrict__ keys, const double* __restrict__ factor, long long n,
               double* __restrict__ mean) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long key = keys[i];
    long long lo = 0, hi = n_px;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (pixel_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    if (lo < n_px && pixel_keys[lo] == key) mean[lo] *= factor[i];   // respect_zeros: absent pixels stay absent
}

}  // namespace h3d

using namespace h3d;

extern "C" int h3d_nb_simulate(const int* row, const int* col, const double* mean, long long n,
                               const double* bias, int n_sim, const double* size_factors,
                               int sf_by_distance, int n_dist, const double* disp, int disp_per_pixel,
                               int rep, unsigned long long seed, long long* counts_out,
                               double* biased_mean_out, h3d_stream_t stream) {
    H3D_REQUIRE(n_sim >= 1 && rep >= 0 && rep < n_sim, "replicate out of range");
    H3D_REQUIRE(n_dist >= 1, "n_dist must be positive");
    if (n <= 0) return H3D_OK;
    nb_simulate_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(
        row, col, mean, n, bias, n_sim, size_factors, sf_by_distance, n_dist, disp, disp_per_pixel,
        rep, seed, counts_out, biased_mean_out);
    H3D_LAUNCHED("nb_simulate_kernel");
    return H3D_OK;
}

extern "C" int h3d_perturb(const long long* pixel_keys, long long n_px, const long long* keys,
                           const double* factor, long long n, double* mean, h3d_stream_t stream) {
    if (n <= 0 || n_px <= 0) return H3D_OK;
    perturb_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(pixel_keys, n_px, keys, factor, n,
                                                                     mean);
    H3D_LAUNCHED("perturb_kernel");
    return H3D_OK;
}
