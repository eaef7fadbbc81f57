// fit_mu_hat and the likelihood-ratio test, one thread per pixel.
//
// Replaces hic3defdr/util/scaled_nb.py:71-183 (fit_mu_hat),
// hic3defdr/util/lrt.py:7-50 (lrt) and the gathers that feed it at
// hic3defdr/analysis/analysis.py:261-278.
//
// FP64-pipe bound: ~2.5 k FP64 instruction-equivalents per pixel against
// 96 algorithmic bytes, so the layout goal is only
// that loads/stores are coalesced (each thread reads its n_reps contiguous
// values; a warp covers one contiguous 32 * n_reps * 8 byte span) and that the
// per-replicate arrays stay in registers (static indexing under a replicate
// mask, one template instance per replicate-count class).
#include "common.cuh"

namespace h3d {

struct DesignMasks {
    unsigned cond_mask[H3D_MAX_CONDS];   // bit r set: replicate r belongs to condition c
    unsigned all_mask;
    int n_reps, n_conds;
};

static int make_masks(const unsigned char* design_host, int n_reps, int n_conds, DesignMasks* m) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_conds >= 1 && n_conds <= H3D_MAX_CONDS, "n_conds out of range");
    m->n_reps = n_reps; m->n_conds = n_conds;
    m->all_mask = (n_reps == 32) ? 0xffffffffu : ((1u << n_reps) - 1u);
    for (int c = 0; c < H3D_MAX_CONDS; ++c) m->cond_mask[c] = 0;
    for (int r = 0; r < n_reps; ++r)
        for (int c = 0; c < n_conds; ++c)
            if (design_host[r * n_conds + c]) m->cond_mask[c] |= (1u << r);
    return H3D_OK;
}

template <int MAXR>
__global__ void __launch_bounds__(128)
fit_mu_kernel(const double* __restrict__ x, const double* __restrict__ b,
              const double* __restrict__ alpha, long long a_spx, long long a_srep,
              long long n, int n_reps, double* __restrict__ mu_out, int* __restrict__ n_failed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xr[MAXR], br[MAXR], ar[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        if (r < n_reps) {
            xr[r] = x[i * n_reps + r];
            br[r] = b[i * n_reps + r];
            ar[r] = alpha[i * a_spx + r * a_srep];
        }
    }
    int st = 0;
    const unsigned mask = (1u << n_reps) - 1u;
    mu_out[i] = fit_mu<MAXR>(xr, br, ar, mask, &st);
    if (st) atomicAdd(n_failed, 1);
}

// Shared body of the two LRT kernels: everything after the per-replicate
// inputs (counts x, factors f, per-condition dispersions) are in registers.
// one out-of-line copy of the Newton fit per kernel instead of one per call
// site (null model + the loop over conditions): the straight-line code of the
// fused kernel otherwise outgrows the instruction cache (profiles/r01g:
// no_instruction 2.5 stalls per issue)
#ifndef H3D_LRT_NOINLINE_FIT
#define H3D_LRT_NOINLINE_FIT 0
#endif
template <int MAXR>
#if H3D_LRT_NOINLINE_FIT
__device__ __noinline__
#else
__device__ __forceinline__
#endif
double lrt_fit_mu(const double* xr, const double* fr, const double* ar, unsigned mask, int* st) {
    return fit_mu<MAXR>(xr, fr, ar, mask, st);
}

template <int MAXR>
__device__ __forceinline__ void lrt_pixel(const double* xr, const double* fr,
                                          const double* dc, const DesignMasks& dm,
                                          int refit_mu, double* p_out, double* llr_out,
                                          double* mu0_out, double* mu1_out, int* fail) {
    // widen dispersions to replicates: (disp @ design.T)[r]
    double ar[MAXR];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        double a = 0.0;
        for (int c = 0; c < dm.n_conds; ++c)
            if ((dm.cond_mask[c] >> r) & 1u) a += dc[c];
        ar[r] = a;
    }
    double mu0, mu1[H3D_MAX_CONDS];
    int st = 0;
    if (refit_mu) {
        mu0 = lrt_fit_mu<MAXR>(xr, fr, ar, dm.all_mask, &st);
        for (int c = 0; c < dm.n_conds; ++c) {
            int s2 = 0;
            mu1[c] = lrt_fit_mu<MAXR>(xr, fr, ar, dm.cond_mask[c], &s2);
            st |= s2;
        }
    } else {
        // plain means of raw / f (lrt.py:41-44)
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < MAXR; ++r)
            if ((dm.all_mask >> r) & 1u) s += xr[r] / fr[r];
        mu0 = s / (double)dm.n_reps;
        for (int c = 0; c < dm.n_conds; ++c) {
            double sc = 0.0; int k = 0;
#pragma unroll
            for (int r = 0; r < MAXR; ++r)
                if ((dm.cond_mask[c] >> r) & 1u) { sc += xr[r] / fr[r]; ++k; }
            mu1[c] = sc / (double)k;
        }
    }
    // llr = sum_r logpmf(x; mu0 f, phi) - logpmf(x; mu1_wide f, phi)
    double llr = 0.0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        if ((dm.all_mask >> r) & 1u) {
            double m1w = 0.0;
            for (int c = 0; c < dm.n_conds; ++c)
                if ((dm.cond_mask[c] >> r) & 1u) m1w += mu1[c];
            llr += nb_llr_term(xr[r], mu0 * fr[r], m1w * fr[r], ar[r]);
        }
    }
    *llr_out = llr;
    *p_out = chi2_sf(-2.0 * llr, dm.n_conds - 1);
    *mu0_out = mu0;
    for (int c = 0; c < dm.n_conds; ++c) mu1_out[c] = mu1[c];
    *fail = st;
}

template <int MAXR>
__global__ void __launch_bounds__(128)
lrt_kernel(const double* __restrict__ raw, const double* __restrict__ f,
           const double* __restrict__ disp, DesignMasks dm, long long n, int refit_mu,
           double* __restrict__ pvalues, double* __restrict__ llr,
           double* __restrict__ mu_null, double* __restrict__ mu_alt,
           int* __restrict__ n_failed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xr[MAXR], fr[MAXR], dc[H3D_MAX_CONDS];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        xr[r] = 0.0; fr[r] = 1.0;
        if (r < dm.n_reps) {
            xr[r] = raw[i * dm.n_reps + r];
            fr[r] = f[i * dm.n_reps + r];
        }
    }
    for (int c = 0; c < dm.n_conds; ++c) dc[c] = disp[i * dm.n_conds + c];
    double p, l, m0, m1[H3D_MAX_CONDS];
    int st;
    lrt_pixel<MAXR>(xr, fr, dc, dm, refit_mu, &p, &l, &m0, m1, &st);
    pvalues[i] = p; llr[i] = l; mu_null[i] = m0;
    for (int c = 0; c < dm.n_conds; ++c) mu_alt[i * dm.n_conds + c] = m1[c];
    if (st) atomicAdd(n_failed, 1);
}

// 8 resident CTAs per SM (64 registers, ~200 bytes of spills) against 5 at 88
// registers: 8.15 -> 7.46 ms per step (profiles/r01g: 30 % of the warp slots were
// in use, FP64 pipe active 32 %)
#ifndef H3D_LRT_MIN_BLOCKS
#define H3D_LRT_MIN_BLOCKS 8
#endif
template <int MAXR>
__global__ void __launch_bounds__(128, H3D_LRT_MIN_BLOCKS)
lrt_fused_kernel(const int* __restrict__ row, const int* __restrict__ col,
                 const int* __restrict__ index, long long n_sel,
                 const long long* __restrict__ raw, const double* __restrict__ sf,
                 int sf_per_pixel, const double* __restrict__ bias,
                 const double* __restrict__ disp, DesignMasks dm, int refit_mu,
                 double* __restrict__ pvalues, double* __restrict__ llr,
                 double* __restrict__ mu_null, double* __restrict__ mu_alt,
                 int* __restrict__ n_failed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel) return;
    const long long u = index ? (long long)index[i] : i;
    const long long ro = (long long)row[u] * dm.n_reps, co = (long long)col[u] * dm.n_reps;
    double xr[MAXR], fr[MAXR], dc[H3D_MAX_CONDS];
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        xr[r] = 0.0; fr[r] = 1.0;
        if (r < dm.n_reps) {
            xr[r] = (double)raw[u * dm.n_reps + r];
            const double s = sf_per_pixel ? sf[u * dm.n_reps + r] : sf[r];
            fr[r] = bias[ro + r] * bias[co + r] * s;      // analysis.py:272-275
        }
    }
    for (int c = 0; c < dm.n_conds; ++c) dc[c] = disp[i * dm.n_conds + c];
    double p, l, m0, m1[H3D_MAX_CONDS];
    int st;
    lrt_pixel<MAXR>(xr, fr, dc, dm, refit_mu, &p, &l, &m0, m1, &st);
    pvalues[i] = p; llr[i] = l; mu_null[i] = m0;
    for (int c = 0; c < dm.n_conds; ++c) mu_alt[i * dm.n_conds + c] = m1[c];
    if (st) atomicAdd(n_failed, 1);
}

}  // namespace h3d

using namespace h3d;

#define H3D_DISPATCH_REPS(n_reps, CALL)            \
    if ((n_reps) <= 2) { CALL(2); }                \
    else if ((n_reps) <= 4) { CALL(4); }           \
    else if ((n_reps) <= 8) { CALL(8); }           \
    else { CALL(16); }

extern "C" int h3d_fit_mu_hat(const double* x, const double* b, const double* alpha,
                              long long alpha_stride_px, long long alpha_stride_rep,
                              long long n, int n_reps, double* mu_out, int* n_failed,
                              h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    if (n <= 0) return H3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = div_up(n, 128);
#define CALL(M) fit_mu_kernel<M><<<grid, 128, 0, st>>>(x, b, alpha, alpha_stride_px, \
        alpha_stride_rep, n, n_reps, mu_out, n_failed)
    H3D_DISPATCH_REPS(n_reps, CALL)
#undef CALL
    H3D_LAUNCHED("fit_mu_kernel");
    return H3D_OK;
}

extern "C" int h3d_lrt(const double* raw, const double* f, const double* disp,
                       const unsigned char* design_host, long long n, int n_reps,
                       int n_conds, int refit_mu, double* pvalues, double* llr,
                       double* mu_hat_null, double* mu_hat_alt, int* n_failed,
                       h3d_stream_t stream) {
    DesignMasks dm;
    int rc = make_masks(design_host, n_reps, n_conds, &dm);
    if (rc) return rc;
    if (n <= 0) return H3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = div_up(n, 128);
#define CALL(M) lrt_kernel<M><<<grid, 128, 0, st>>>(raw, f, disp, dm, n, refit_mu, pvalues, \
        llr, mu_hat_null, mu_hat_alt, n_failed)
    H3D_DISPATCH_REPS(n_reps, CALL)
#undef CALL
    H3D_LAUNCHED("lrt_kernel");
    return H3D_OK;
}

extern "C" int h3d_lrt_fused(const int* row, const int* col, const int* index,
                             long long n_sel, const long long* raw,
                             const double* size_factors, int sf_per_pixel,
                             const double* bias, const double* disp,
                             const unsigned char* design_host, int n_reps, int n_conds,
                             int refit_mu, double* pvalues, double* llr,
                             double* mu_hat_null, double* mu_hat_alt, int* n_failed,
                             h3d_stream_t stream) {
    DesignMasks dm;
    int rc = make_masks(design_host, n_reps, n_conds, &dm);
    if (rc) return rc;
    if (n_sel <= 0) return H3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = div_up(n_sel, 128);
#define CALL(M) lrt_fused_kernel<M><<<grid, 128, 0, st>>>(row, col, index, n_sel, raw, \
        size_factors, sf_per_pixel, bias, disp, dm, refit_mu, pvalues, llr, mu_hat_null, \
        mu_hat_alt, n_failed)
    H3D_DISPATCH_REPS(n_reps, CALL)
#undef CALL
    H3D_LAUNCHED("lrt_fused_kernel");
    return H3D_OK;
}
