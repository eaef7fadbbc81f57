// TEST TOOL: compiles the scalar FP64 building blocks of
// hic3defdr_b200/csrc/h3d_math.cuh as HOST code so that their algorithms can be
// checked against scipy on a machine without a GPU.  Never used by the product.
#define H3D_HOST_STATS
#include "../../hic3defdr_b200/csrc/h3d_math.cuh"
H3dStats g_h3d_stats = {};
extern "C" {
void hc_fit_mu(const double* x, const double* b, const double* alpha, int n, int R,
               double* out, int* status) {
    for (int i = 0; i < n; ++i) {
        int st = 0;
        out[i] = h3d::fit_mu<h3d::kMaxReps>(x + (long)i * R, b + (long)i * R,
                                            alpha + (long)i * R, (1u << R) - 1u, &st);
        status[i] = st;
    }
}
void hc_gamma_pq(const double* a, const double* x, int n, double* p, double* q) {
    for (int i = 0; i < n; ++i) {
        double lga = lgamma(a[i]);
        p[i] = h3d::gamma_p(a[i], x[i], lga);
        q[i] = h3d::gamma_q(a[i], x[i], lga);
    }
}
void hc_gamma_inv(const double* a, const double* t, int n, int upper, double* y) {
    for (int i = 0; i < n; ++i)
        y[i] = h3d::gamma_tail_inv(a[i], t[i], lgamma(a[i]), upper != 0, a[i]);
}
void hc_q2q(const double* x, const double* mu_in, const double* mu_out, double alpha,
            int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = h3d::q2q_one(x[i], mu_in[i], mu_out[i], alpha);
}
void hc_lgamma_pos(const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = h3d::lgamma_pos(x[i]);
}
void hc_fast_log(const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = h3d::m_log(x[i]);
}
// lgamma through the likelihood's Stirling core (shifted): core - x + .5 ln 2pi
void hc_lgamma_core(const double* x, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        // the kernel shifts by ceil(10 - r) with r <= x the smallest argument of
        // the evaluation; here: the smallest shift that brings x itself to >= 10
        const int sh = (x[i] < 10.0) ? (int)ceil(10.0 - x[i]) : 0;
        const double c = (sh > 0) ? h3d::stirling_core_shifted(x[i], sh)
                                  : h3d::stirling_core(x[i]);
        out[i] = (c - (double)sh) - x[i] + 0.9189385332046727;
    }
}
void hc_chi2_sf(const double* x, int n, int df, double* out) {
    for (int i = 0; i < n; ++i) out[i] = h3d::chi2_sf(x[i], df);
}
void hc_stats(long long* out, int reset) {
    out[0] = g_h3d_stats.n_tail_eval; out[1] = g_h3d_stats.n_series_it;
    out[2] = g_h3d_stats.n_cf_it; out[3] = g_h3d_stats.n_q2q;
    for (int k = 0; k < 16; ++k) out[4 + k] = g_h3d_stats.step_hist[k];
    for (int k = 0; k < 8; ++k) out[20 + k] = g_h3d_stats.evals_hist[k];
    if (reset) g_h3d_stats = H3dStats{};
}
// drives the Brent state machine with a callback
typedef double (*hc_fn)(double);
double hc_brent(hc_fn f, double lo, double hi, double xatol, int maxfun, int* nfev, int* flag) {
    h3d::BrentState s;
    h3d::brent_begin(s, lo, hi);
    bool more = true;
    while (more) more = h3d::brent_advance(s, f(s.x_eval), xatol, maxfun);
    *nfev = s.num; *flag = s.flag;
    return s.xf;
}
}
