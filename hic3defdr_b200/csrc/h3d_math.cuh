// FP64 scalar building blocks of the scaled-NB pipeline, usable from device
// code (the product) and from host code (tests/hostcheck only).
//
// Reference behaviour being reproduced (paths relative to /root/reference):
//   fit_mu          hic3defdr/util/scaled_nb.py:139-183  (root of the score)
//   q2q_*           hic3defdr/util/scaled_nb.py:239-275  (edgeR q2qnbinom)
//   nb_llr          hic3defdr/util/lrt.py:46-48 + scaled_nb.py:31-33
//   chi2_sf         hic3defdr/util/lrt.py:49 (scipy chi2.sf = gammaincc)
//   gamma_p/gamma_q scipy.special.gammainc/gammaincc (regularised)
//   gamma_*_inv     scipy.special.gammaincinv/gammainccinv
// None of this is a translation of scipy/cephes: the incomplete gamma pair is
// the textbook series + Lentz continued fraction, and the inverses are a
// bracketed Newton iteration on the log of the tail probability.
#pragma once
#include <math.h>
#include <float.h>
#include <string.h>

#ifdef __CUDACC__
#define H3D_HD __host__ __device__ __forceinline__
#define H3D_HDN static __host__ __device__ __noinline__
#else
#define H3D_HD inline
#define H3D_HDN static inline
#endif

// optional host-side instrumentation (tests/hostcheck only)
#ifdef H3D_HOST_STATS
struct H3dStats { long long n_tail_eval, n_series_it, n_cf_it, n_q2q; long long step_hist[16]; long long evals_hist[8]; };
extern H3dStats g_h3d_stats;
#define H3D_STAT(field) (++g_h3d_stats.field)
#define H3D_STAT_STEP(it, rel) do { if ((it) == 0) { int b_ = (int)(-log10((rel) + 1e-300)); \
    ++g_h3d_stats.step_hist[b_ < 0 ? 0 : (b_ > 15 ? 15 : b_)]; } } while (0)
#define H3D_STAT_EVALS(n) (++g_h3d_stats.evals_hist[(n) > 7 ? 7 : (n)])
#else
#define H3D_STAT(field) ((void)0)
#define H3D_STAT_STEP(it, rel) ((void)0)
#define H3D_STAT_EVALS(n) ((void)0)
#endif

// Halley iteration of the incomplete-gamma inverse: a step smaller than this
// (in units of the distribution's width, y / sqrt(a)) ends the iteration -- the
// error left after a step of size e is ~ C e^3 (cubic convergence).  Worst
// relative error of the whole quantile map against scipy over 3e5 random cases
// (tests/test_hostcheck_math.py) and equalize_kernel time per step on B200:
//   2e-4: 1.4e-12, 81.2 ms   6e-4: 9.5e-12, 78.1 ms   1e-3: 4.5e-11
//   2e-3: 4.7e-10, 73.1 ms   4e-3: 3.8e-9
// (scipy's own gammaincinv is good to ~1e-11 for a ~ 1e3).  1e-3 keeps the
// pseudo-data 20x inside the 1e-9 budget of the outputs they feed.
#ifndef H3D_HALLEY_TOL
#define H3D_HALLEY_TOL 1e-3
#endif

namespace h3d {

constexpr int kMaxReps = 16;
constexpr double kEps = 2.220446049250313e-16;

// One out-of-line copy per kernel of the library transcendentals that the
// special-function code calls from many places: inlined, they made up 40 % of
// equalize_kernel's 77 KB of SASS and the warps starved on instruction fetch
// (ncu: stall "no instruction" dominant).
H3D_HDN double m_exp(double x) { return exp(x); }
H3D_HDN double m_log1p(double x) { return log1p(x); }

// ---------------------------------------------------------------------------
// Table-driven natural logarithm for positive, finite, normal arguments that
// are not close to 1 (the log-gamma arguments y + r >= r of the conditional
// likelihood, evaluated ~10^10 times per genome): x = 2^e m, m in [1, 2);
// c_j ~ 1 / (centre of the j-th of 128 mantissa intervals);
// ln x = e ln2 + (-ln c_j) + log1p(m c_j - 1), |m c_j - 1| <= 2^-8, so a
// degree-7 polynomial suffices.  ~12 FP64 instructions instead of ~30 for the
// general routine; error <= ~1.5 ulp for |ln x| >= 1 (the cancellation-free
// range); checked against libm in tests/test_hostcheck_math.py.
// ---------------------------------------------------------------------------
struct LogTabEntry { double c, neg_log_c; };
constexpr int kLogTabSize = 128;

// Polynomial coefficients of the likelihood kernel's hot loop.  In device code
// they are read as constant-bank operands of the FP64 instructions themselves;
// as literals the compiler materialises each one with a pair of uniform-register
// moves in front of its use (66 of the 213 instructions of the loop, ncu r01e),
// which costs issue slots the FP64 pipe is waiting for.
#ifdef __CUDACC__
static __constant__ double kDevLogPoly[6] = {1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5};
static __constant__ double kDevStirling[7] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0,
                                              1.0 / 1188.0, -691.0 / 360360.0, 1.0 / 156.0};
static __constant__ double kDevLn2[2] = {0.693147180559945286, 2.319046813846299616e-17};
// 1/27, 1/25, ..., 1/3: the atanh series of log1p_minus_x
static __constant__ double kDevAtanh[13] = {1.0 / 27.0, 1.0 / 25.0, 1.0 / 23.0, 1.0 / 21.0, 1.0 / 19.0,
                                            1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0,
                                            1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};
// 1/10!, 1/9!, ..., 1/2!: exp(u) - 1 - u for small u
static __constant__ double kDevExpSmall[9] = {1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                                              1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
// the table itself: 2 KB, read through L1 (a divergent index would serialise
// in the constant cache)
static __device__ const LogTabEntry kDevLogTab[kLogTabSize] = {
#include "h3d_logtab.inc"
};
#endif
static const LogTabEntry kHostLogTab[kLogTabSize] = {
#include "h3d_logtab.inc"
};

H3D_HD double fast_log_pos(double x) {
#ifdef __CUDA_ARCH__
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
#else
    long long b;
    memcpy(&b, &x, 8);
    const int hi = (int)(b >> 32);
    const long long mb = (b & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL;
    double m;
    memcpy(&m, &mb, 8);
#endif
    const int ex = (hi >> 20) - 1023;
#ifdef __CUDA_ARCH__
    const double2 tt = __ldg((const double2*)&kDevLogTab[(hi >> 13) & (kLogTabSize - 1)]);
    LogTabEntry t;
    t.c = tt.x; t.neg_log_c = tt.y;
#else
    const LogTabEntry t = kHostLogTab[(hi >> 13) & (kLogTabSize - 1)];
#endif
    const double r = fma(m, t.c, -1.0);
    const double e = (double)ex;
#ifdef __CUDA_ARCH__
    double p = fma(r, kDevLogPoly[0], kDevLogPoly[1]);
    p = fma(p, r, kDevLogPoly[2]);
    p = fma(p, r, kDevLogPoly[3]);
    p = fma(p, r, kDevLogPoly[4]);
    p = fma(p, r, kDevLogPoly[5]);
    const double tail = fma(r * r, p, e * kDevLn2[1]);                    // + e ln2_lo
    return fma(e, kDevLn2[0], t.neg_log_c) + (r + tail);
#else
    double p = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma(p, r, 1.0 / 5.0);
    p = fma(p, r, -1.0 / 4.0);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double tail = fma(r * r, p, e * 2.319046813846299616e-17);      // + e ln2_lo
    return fma(e, 0.693147180559945286, t.neg_log_c) + (r + tail);
#endif
}

// Natural logarithm as used by the special-function code below: absolute error
// <= ~1.5e-16 max(1, |ln x|) (what log-probabilities and (a - .5) ln a terms
// need; NOT a relative-accuracy log near x = 1), table-driven for normal
// positive arguments, the library routine otherwise (zero, subnormal, inf, NaN,
// negative).  One out-of-line copy per kernel (instruction-cache footprint).
H3D_HDN double m_log(double x) {
#ifdef __CUDA_ARCH__
    const unsigned hi = (unsigned)__double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    const unsigned hi = (unsigned)(b >> 32);
#endif
#ifndef H3D_MLOG_LIB
    if (hi - 0x00100000u < 0x7fe00000u) return fast_log_pos(x);   // normal and positive
#else
    (void)hi;
#endif
    return log(x);
}

// reciprocal of a positive normal number to ~1 ulp (not correctly rounded):
// hardware seed (2^-23) refined by one cubic step
H3D_HD double fast_rcp_pos(double x) {
#ifdef __CUDA_ARCH__
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e = fma(-x, y0, 1.0);
    return fma(y0, fma(e, e, e), y0);
#else
    return 1.0 / x;
#endif
}

// ---------------------------------------------------------------------------
// fit_mu: the unique positive root of
//     g(mu) = sum_r (x_r - mu b_r) / (mu + alpha_r mu^2 b_r)
// which is also the root of h(mu) = sum_r (x_r - mu b_r) / (1 + alpha_r b_r mu).
// h is strictly decreasing and convex on mu > 0 (h' = -sum b(1+alpha x)/(1+a mu)^2,
// h'' > 0), so Newton started left of the root increases monotonically to it.
// The reference reaches the same root (to ~1e-15) with a secant sweep plus a
// brentq fallback.  Returns the root; *status = 1 when sum(x) == 0 (the
// reference raises there), 2 when the iteration budget ran out.
// ---------------------------------------------------------------------------
template <int MAXR>
H3D_HD double fit_mu(const double* x, const double* b, const double* alpha,
                     unsigned mask, int* status) {
    // ``mask`` selects the replicates that take part (bit r = replicate r);
    // arrays are indexed statically so that they stay in registers.
    double a[MAXR], c[MAXR];
    double sx = 0.0, sc = 0.0, mu = 0.0;
    int nrep = 0;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
        if ((mask >> r) & 1u) {
            a[r] = alpha[r] * b[r];
            c[r] = b[r] * (1.0 + alpha[r] * x[r]);
            sx += x[r];
            sc += c[r];
            mu += x[r] / b[r];
            nrep += 1;
        }
    }
    if (!(sx > 0.0)) { *status = 1; return 0.0; }
    mu /= (double)nrep;                      // the reference's start, mean(x/b)
    const double safe = sx / sc;             // Newton step from mu = 0: left of the root
    bool left = false;                       // true once we know mu <= root
    for (int it = 0; it < 100; ++it) {
        double h = 0.0, dh = 0.0;
#pragma unroll
        for (int r = 0; r < MAXR; ++r) {
            if ((mask >> r) & 1u) {
                const double t = 1.0 / (1.0 + a[r] * mu);
                h += (x[r] - mu * b[r]) * t;
                dh -= c[r] * t * t;
            }
        }
        double nxt = mu - h / dh;
        if (!left) {
            // first step may come from the right of the root: the tangent then
            // lands left of the root, possibly at or below zero.
            if (!(nxt > 0.0) || !isfinite(nxt)) nxt = safe;
            left = true;
            mu = nxt;
            continue;
        }
        if (!(nxt > mu)) return mu;          // monotone sequence hit round-off
        // quadratic convergence from the left: the step just taken is (to first
        // order) the error of mu, the error of nxt is ~ K step^2 / nxt with
        // K = a mu / (1 + a mu) < 1 -- a relative step below 1e-8 leaves nxt
        // converged to round-off, and one Newton sweep of five is saved
        const bool done = (nxt - mu) <= 1e-8 * nxt;
        mu = nxt;
        if (done) return mu;
    }
    *status = 2;
    return mu;
}

// ---------------------------------------------------------------------------
// Regularised incomplete gamma functions, organised around
//   k(a, x) = x^a e^-x / Gamma(a)          (gamma_logk returns log k)
//   P(a, x) = k * S(a, x),  S = (1/a) sum_n x^n / ((a+1)...(a+n))   (series)
//   Q(a, x) = k * H(a, x),  H = 1/(x+1-a- 1(1-a)/(x+3-a- 2(2-a)/...)) (cont. fr.)
// The series is used for x < max(a, 1), the continued fraction otherwise.
// Both loops are arranged so that the FP64 divider is touched once per four
// terms: the series shares one reciprocal among four consecutive terms, the
// continued fraction is evaluated by the forward (Wallis) recurrence of its
// convergents and only divides when it tests convergence.
// ---------------------------------------------------------------------------
// Stirling's correction sum 1/(12x) - 1/(360x^3) + ... given ix = 1/x (x >= 10:
// truncation < 2e-17).  Device: coefficients as constant-bank operands.
H3D_HD double stirling_sum(double ix) {
    const double ix2 = ix * ix;
#ifdef __CUDA_ARCH__
    double c = fma(ix2, kDevStirling[6], kDevStirling[5]);
    c = fma(ix2, c, kDevStirling[4]);
    c = fma(ix2, c, kDevStirling[3]);
    c = fma(ix2, c, kDevStirling[2]);
    c = fma(ix2, c, kDevStirling[1]);
    c = fma(ix2, c, kDevStirling[0]);
    return ix * c;
#else
    return ix * (1.0 / 12.0 + ix2 * (-1.0 / 360.0 + ix2 * (1.0 / 1260.0 +
        ix2 * (-1.0 / 1680.0 + ix2 * (1.0 / 1188.0 + ix2 * (-691.0 / 360360.0 +
        ix2 * (1.0 / 156.0)))))));
#endif
}

// exp(u) for the small arguments of a converging iteration (|u| < 1/16: the
// degree-10 Taylor polynomial is exact to 2e-21), the library routine otherwise
H3D_HD double exp_step(double u) {
    if (fabs(u) < 0.0625) {
#ifdef __CUDA_ARCH__
        double p = kDevExpSmall[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) p = fma(p, u, kDevExpSmall[k]);
#else
        const double cf[9] = {1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                              1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
        double p = cf[0];
        for (int k = 1; k < 9; ++k) p = fma(p, u, cf[k]);
#endif
        return fma(u * u, p, u) + 1.0;
    }
    return m_exp(u);
}

H3D_HD double log1p_minus_x(double u) {
    if (fabs(u) < 0.25) {
        // ln(1+u) - u via the atanh series in w = u / (2 + u):
        // ln(1+u) = 2 (w + w^3/3 + w^5/5 + ...), and 2w - u = -u w
        const double w = u / (2.0 + u);
        const double w2 = w * w;
#ifdef __CUDA_ARCH__
        double s = kDevAtanh[0];
#pragma unroll
        for (int k = 1; k < 13; ++k) s = fma(s, w2, kDevAtanh[k]);
#else
        double s = 0.0;
        for (int k = 27; k >= 3; k -= 2) s = s * w2 + 1.0 / (double)k;
#endif
        return 2.0 * w * w2 * s - u * w;
    }
    return m_log1p(u) - u;
}

struct GammaShape {
    double a;
    double lead;   // a >= 10: 0.5 ln(a / 2 pi) - stirling(a);  else: -lgamma(a)
    double inv_a;
    bool big;
};

H3D_HD double lgamma_pos(double x);

H3D_HDN GammaShape gamma_shape(double a) {
    GammaShape s;
    s.a = a;
    s.inv_a = 1.0 / a;
    s.big = a >= 10.0;
    if (s.big) {
        // direct a ln x - x - lgamma(a) cancels catastrophically for large a;
        // Stirling: ln Gamma(a) = (a - .5) ln a - a + .5 ln 2pi + corr(a)
        s.lead = 0.5 * (m_log(a) - 1.8378770664093453) - stirling_sum(s.inv_a);
    } else {
        s.lead = -lgamma_pos(a);
    }
    return s;
}

// log of x^a e^-x / Gamma(a)
H3D_HD double gamma_logk(const GammaShape& s, double x) {
    if (s.big) {
        const double u = (x - s.a) * s.inv_a;
        // away from u = 0 take the log of x / a itself: log1p(u) would inherit
        // the rounding of u, amplified by 1 / (1 + u) in the far left tail
        const double l = (fabs(u) < 0.25) ? log1p_minus_x(u) : m_log(x * s.inv_a) - u;
        return s.lead + s.a * l;
    }
    return s.a * m_log(x) - x + s.lead;
}

// S(a, x) = (1/a) * sum_{n>=0} x^n / ((a+1)...(a+n)); terms decrease
// monotonically because the series is only used for x < a + 1.
H3D_HD double gamma_series_factor(double a, double x) {
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    double r = a, c = 1.0, ans = 1.0;
    for (int i = 0; i < 5000; ++i) {
        H3D_STAT(n_series_it);
        const double p1 = r + 1.0, p2 = r + 2.0, p3 = r + 3.0, p4 = r + 4.0;
        const double p34 = p3 * p4, p234 = p2 * p34;
        const double q = c / (p1 * p234);            // one division per four terms
        const double c1 = q * x * p234, c2 = q * x2 * p34, c3 = q * x3 * p4, c4 = q * x4;
        ans += (c1 + c2) + (c3 + c4);
        c = c4;
        r = p4;
        if (c4 <= ans * (0.25 * kEps)) break;
    }
    return ans / a;
}

// H(a, x): forward recurrence of the convergents A_n / B_n of
// b0 + a1/(b1 + a2/(b2 + ...)), b_n = x + 2n + 1 - a, a_n = -n (n - a);
// H = B_n / A_n.  Used for x >= max(a, 1), where it converges.
H3D_HD double gamma_cf_factor(double a, double x) {
    double bn = x + 1.0 - a;
    double Am = 1.0, A = bn, Bm = 0.0, B = 1.0;       // A_{-1}, A_0, B_{-1}, B_0
    double h_prev = 1.0 / bn;
    double n = 0.0;
    for (int i = 0; i < 5000; ++i) {
        H3D_STAT(n_cf_it);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            n += 1.0;
            bn += 2.0;
            const double an = -n * (n - a);
            const double An = bn * A + an * Am, Bn = bn * B + an * Bm;
            Am = A; A = An; Bm = B; B = Bn;
        }
        const double h = B / A;
        if (fabs(h - h_prev) <= fabs(h) * kEps) return h;
        h_prev = h;
        if (fabs(A) > 1e150) { A *= 1e-150; Am *= 1e-150; B *= 1e-150; Bm *= 1e-150; }
    }
    return h_prev;
}

// The power series converges for every x; the continued fraction only pays for
// itself when x is well above a.  For small x (a handful of counts, the bulk
// of the far-distance pixels) the fraction needs ~50 steps where the series
// needs ~30 cheaper terms, so the series also serves the upper tail there as
// long as the complement 1 - P does not cancel (z-score below kSeriesZMax,
// i.e. Q > ~0.02: at most ~50 eps relative error in Q).
#ifndef H3D_SERIES_XMAX
#define H3D_SERIES_XMAX 6.0
#endif
#ifndef H3D_SERIES_ZMAX
#define H3D_SERIES_ZMAX 1.8
#endif
H3D_HD bool gamma_use_series(double a, double x) {
    if ((x < 1.0) || (x < a)) return true;
    const double dz = x - a;
    return (x < H3D_SERIES_XMAX) && (dz * dz < (H3D_SERIES_ZMAX * H3D_SERIES_ZMAX) * a);
}

// log of the lower (upper = false) or upper tail at y > 0,
// ratio = T / (y * pdf(y)) (so that d log T / d log y = -/+ 1 / ratio), and the
// log of the kernel k(a, y).
// (not inlined on purpose: one copy of the series / continued-fraction code per
// kernel keeps the hot loop inside the instruction cache)
H3D_HDN void gamma_log_tail(const GammaShape& s, double y, bool upper, double* log_t,
                            double* ratio, double* log_k) {
    H3D_STAT(n_tail_eval);
    const double lk = gamma_logk(s, y);
    *log_k = lk;
    const bool series = gamma_use_series(s.a, y);
    const double F = series ? gamma_series_factor(s.a, y) : gamma_cf_factor(s.a, y);
    if (series != upper) {
        *log_t = lk + m_log(F);                     // log P (series) or log Q (fraction)
        *ratio = F;
        return;
    }
    // the complementary tail was asked for: T = 1 - k F, ratio = T / k
    // (one exponential; no logarithm of F is needed on this branch)
    const double k = m_exp(lk);
    const double c = k * F;
    *log_t = m_log1p(-c);
    *ratio = (1.0 - c) / k;
}

H3D_HD double gamma_p(double a, double x, double lga) {
    (void)lga;
    if (!(x > 0.0)) return (x == 0.0) ? 0.0 : NAN;
    if (isinf(x)) return 1.0;
    const GammaShape s = gamma_shape(a);
    double lt, ratio, lk;
    gamma_log_tail(s, x, false, &lt, &ratio, &lk);
    return exp(lt);
}

H3D_HD double gamma_q(double a, double x, double lga) {
    (void)lga;
    if (!(x > 0.0)) return (x == 0.0) ? 1.0 : NAN;
    if (isinf(x)) return 0.0;
    const GammaShape s = gamma_shape(a);
    double lt, ratio, lk;
    gamma_log_tail(s, x, true, &lt, &ratio, &lk);
    return exp(lt);
}

// ---------------------------------------------------------------------------
// Inverse in log space: solve log T(a, y) = lt for y by Newton with a
// maintained bracket.  log Q is close to linear in y and log P close to linear
// in log y, and in both cases the slope is 1 / ratio -- no extra
// transcendental is needed per step beyond the tail evaluation itself.
// ``guess`` seeds the iteration (any positive finite value is acceptable).
// ---------------------------------------------------------------------------
H3D_HD double gamma_log_tail_inv(const GammaShape& s, double lt, bool upper, double guess) {
    if (!(lt < 0.0)) return (lt == 0.0) ? (upper ? 0.0 : INFINITY) : NAN;
    const double a = s.a;
    double lo = 0.0, hi = INFINITY;
    // the natural unit of a step is the distribution's width, y / sqrt(a)
    const double step_tol = (a > 1.0) ? H3D_HALLEY_TOL / sqrt(a) : H3D_HALLEY_TOL;
    double y = (guess > 0.0 && isfinite(guess)) ? guess : a;
    if (!upper) {
        // far lower tail: P(a,y) ~ y^a / Gamma(a+1).  Only a seed: single
        // precision (the exponent is bounded: lt > -745, a > 0)
        const float af = (float)a;
        const float ys = expf(((float)lt + lgammaf(af + 1.0f)) / af);
        if (ys < 0.2f * (af + 1.0f) && ys > 0.0f) y = (double)ys;
    }
    if (!(y > 0.0)) y = 1.0;
    for (int it = 0; it < 100; ++it) {
        double lT, ratio, lk;
        gamma_log_tail(s, y, upper, &lT, &ratio, &lk);
        if (lT == lt) return y;
        // bracket update: upper tail decreases in y, lower tail increases
        const bool y_too_small = upper ? (lT > lt) : (lT < lt);
        if (y_too_small) lo = y; else hi = y;
        double nxt = NAN;
        if (isfinite(lT) && isfinite(ratio) && ratio > 0.0) {
            // Halley step on g = log T - lt (cubic convergence); the second
            // derivative is analytic: with lambda = pdf / T = 1 / (y ratio),
            //   upper: g' = -lambda, g'' = -lambda ((a-1)/y - 1 + lambda)   (in y)
            //   lower: g' = kappa = y lambda, g'' = kappa (a - y - kappa)    (in log y)
            const double g = lT - lt;
            if (upper) {
                const double lam = 1.0 / (y * ratio);
                double den = 2.0 * lam + g * ((a - 1.0) / y - 1.0 + lam);
                if (!(den > lam)) den = 2.0 * lam;            // fall back to Newton
                nxt = y + 2.0 * g / den;
            } else {
                const double kap = 1.0 / ratio;
                double den = 2.0 * kap - g * (a - y - kap);
                if (!(den > kap)) den = 2.0 * kap;
                nxt = y * exp_step(-2.0 * g / den);
            }
            // a step this small means nxt is converged to round-off (the next
            // correction would be below 1e-18 relative)
            H3D_STAT_STEP(it, fabs(nxt - y) / fabs(nxt) * (a > 1.0 ? sqrt(a) : 1.0));
            if (fabs(nxt - y) <= step_tol * fabs(nxt)) { H3D_STAT_EVALS(it + 1); return nxt; }
        }
        if (!((nxt > lo) && (nxt < hi))) {
            // Newton left the bracket (or was not available): bisect
            // (geometrically when possible), or expand an open side
            if (isinf(hi)) nxt = (lo > 0.0) ? lo * 2.0 : 1.0;
            else if (lo == 0.0) nxt = hi * 0.25;
            else nxt = sqrt(lo) * sqrt(hi);
            if ((hi - lo) <= 4.0 * kEps * hi) return nxt;
        }
        y = nxt;
    }
    return y;
}

// probability-space wrapper (tests): T(a, y) = t
H3D_HD double gamma_tail_inv(double a, double t, double lga, bool upper, double guess) {
    (void)lga;
    if (!(t > 0.0)) return (t == 0.0) ? (upper ? INFINITY : 0.0) : NAN;
    if (t >= 1.0) return (t == 1.0) ? (upper ? 0.0 : INFINITY) : NAN;
    return gamma_log_tail_inv(gamma_shape(a), m_log(t), upper, guess);
}

// Wilson-Hilferty: Gamma(a,1) variate as a cube of a normal one.
// Residual of the Wilson-Hilferty approximation on a grid of (shape, normal
// score): z_wh(a, y(a, z)) = z + e(a, z) with y the exact quantile
// (tools/gen_whtab.py).  Bilinear interpolation in (log2 a, z); false outside
// the grid (a in [0.25, 64], z in [-6, 8]).  Seeds only.
constexpr int kWhNA = 49, kWhNZ = 57;
#ifdef __CUDACC__
static __device__ const float kDevWhTab[kWhNA * kWhNZ] = {
#include "h3d_whtab.inc"
};
#endif
static const float kHostWhTab[kWhNA * kWhNZ] = {
#include "h3d_whtab.inc"
};

H3D_HD bool wh_residual(float a, float z, float* e) {
    const float fa = 6.0f * log2f(4.0f * a), fz = 4.0f * (z + 6.0f);
    if (!(fa >= 0.0f && fa < (float)(kWhNA - 1) && fz >= 0.0f && fz < (float)(kWhNZ - 1))) return false;
    const int ja = (int)fa, jz = (int)fz;
    const float ta = fa - (float)ja, tz = fz - (float)jz;
#ifdef __CUDA_ARCH__
    const float* t = kDevWhTab + ja * kWhNZ + jz;
    const float v00 = __ldg(t), v01 = __ldg(t + 1), v10 = __ldg(t + kWhNZ), v11 = __ldg(t + kWhNZ + 1);
#else
    const float* t = kHostWhTab + ja * kWhNZ + jz;
    const float v00 = t[0], v01 = t[1], v10 = t[kWhNZ], v11 = t[kWhNZ + 1];
#endif
    const float v = (1.0f - ta) * ((1.0f - tz) * v00 + tz * v01) + ta * ((1.0f - tz) * v10 + tz * v11);
    *e = v;
    return isfinite(v);
}

// (only seeds an iteration: single precision)
H3D_HD float wh_to_normal(float a, float x) {
    return (cbrtf(x / a) - (1.0f - 1.0f / (9.0f * a))) * 3.0f * sqrtf(a);
}
H3D_HD float wh_from_normal(float a, float z) {
    const float t = 1.0f - 1.0f / (9.0f * a) + z / (3.0f * sqrtf(a));
    return a * t * t * t;
}

// ---------------------------------------------------------------------------
// q2q for one replicate of one pixel (scaled_nb.py:239-275).  mu_in/mu_out are
// the already clamped means.  Only the tail the reference selects is
// evaluated.  The normal-to-normal quantile map is the affine map of the
// z-score (ndtri(ndtr(z)) == z up to round-off).  The gamma-to-gamma map is
// carried out on the LOG of the tail probability.  Where the reference's tail
// probability underflows to zero it returns +inf (right tail) or 0 (left
// tail); that is mimicked at the same thresholds.
// ---------------------------------------------------------------------------
// the x = 0, left-tail case: the gamma quantile map is exactly 0 (cdf(0) = 0
// -> ppf(0) = 0), only the normal map contributes
H3D_HD double q2q_zero(double mu_in, double mu_out, double alpha) {
    const double v_in = mu_in * (1.0 + alpha * mu_in);
    const double v_out = mu_out * (1.0 + alpha * mu_out);
    const double z = (0.0 - mu_in) / sqrt(v_in);
    double q_norm = mu_out + sqrt(v_out) * z;
    if (0.5 * z * z > 709.782712893384) q_norm = -INFINITY;
    double out = (q_norm + 0.0) / 2.0;
    if (!(out >= 0.0)) out = 0.0;
    return out;
}

H3D_HD double q2q_one(double x, double mu_in, double mu_out, double alpha) {
    H3D_STAT(n_q2q);
    const double r_in = 1.0 + alpha * mu_in;
    const double r_out = 1.0 + alpha * mu_out;
    const double v_in = mu_in * r_in;
    const double v_out = mu_out * r_out;
    const bool right = x >= mu_in;
    // normal part
    const double z = (x - mu_in) / sqrt(v_in);
    double q_norm;
    // scipy's ndtr goes through cephes erfc, which flushes to zero once
    // z^2 / 2 exceeds MAXLOG = 709.78...; isf(0) / ppf(0) are then +/-inf
    if (0.5 * z * z > 709.782712893384) q_norm = right ? INFINITY : -INFINITY;
    else q_norm = mu_out + sqrt(v_out) * z;
    // gamma part
    const double a_in = mu_in / r_in, a_out = mu_out / r_out;
    const double xs = x / r_in;
    double q_gamma;
    if (!right && !(xs > 0.0)) {
        q_gamma = 0.0;                       // cdf(0) = 0 -> ppf(0) = 0
    } else {
        const GammaShape s_in = gamma_shape(a_in);
        double lt, ratio, lk;
        gamma_log_tail(s_in, xs, right, &lt, &ratio, &lk);
        // scipy's igam / igamc (cephes igam_fac) return 0 when the kernel
        // x^a e^-x / Gamma(a) is below exp(-MAXLOG) -- but only on their
        // "far" branch |a - x| > 0.4 a; closer to the mode the kernel is formed
        // as a product that keeps denormal values, and the tail probability
        // only vanishes where it underflows the double range itself
        const bool far = fabs(a_in - xs) > 0.4 * a_in;
        if ((far && lk < -709.782712893384) || lk < -744.44 || lt < -744.44) {
            q_gamma = right ? INFINITY : 0.0;            // sf / cdf underflow in the reference
        } else {
            double guess = xs * (a_out / a_in);
            {
                // Wilson-Hilferty score of xs under a_in, corrected by the
                // tabulated residual to the exact normal score z of its tail
                // probability (two fixed-point steps), mapped back under a_out
                // with that shape's residual: seed errors of ~1e-4 distribution
                // widths instead of ~1e-2, so that most inversions need a single
                // tail evaluation
                const float ai = (float)a_in, ao = (float)a_out;
                const float zi = wh_to_normal(ai, (float)xs);
                float e1, e2, eo, g2 = -1.0f;
                if (wh_residual(ai, zi, &e1) && wh_residual(ai, zi - e1, &e2) &&
                    wh_residual(ao, zi - e2, &eo))
                    g2 = wh_from_normal(ao, zi - e2 + eo);
                else if (a_in > 2.0 && a_out > 2.0)
                    g2 = wh_from_normal(ao, zi);
                if (g2 > 0.0f && isfinite(g2)) guess = (double)g2;
            }
            q_gamma = r_out * gamma_log_tail_inv(gamma_shape(a_out), lt, right, guess);
        }
    }
    double out = (q_norm + q_gamma) / 2.0;
    if (!(out >= 0.0)) out = 0.0;
    return out;
}

// ---------------------------------------------------------------------------
// log Gamma(x) for x > 0 (the only case the conditional NB likelihood needs,
// dispersion.py:72-75): Stirling's series for x >= 10, the recurrence
// Gamma(x) = Gamma(x + n) / (x (x+1) ... (x+n-1)) below.  One log (two below
// 10) and one reciprocal instead of the general-purpose library routine.
// ---------------------------------------------------------------------------
// correction term of Stirling's series for x >= 10 (truncation < 2e-17)
H3D_HD double stirling_corr(double x) {
#ifdef __CUDA_ARCH__
    const double ix = __drcp_rn(x);
#else
    const double ix = 1.0 / x;
#endif
    const double ix2 = ix * ix;
    return ix * (1.0 / 12.0 + ix2 * (-1.0 / 360.0 + ix2 * (1.0 / 1260.0 +
        ix2 * (-1.0 / 1680.0 + ix2 * (1.0 / 1188.0 + ix2 * (-691.0 / 360360.0 +
        ix2 * (1.0 / 156.0)))))));
}

// Stirling's series term of one log-gamma argument x >= 10 without the
// "- x + .5 ln 2pi" part (which cancels or is constant in the conditional
// likelihood): (x - .5) ln x + corr(x), table-driven log, refined-seed
// reciprocal.  NT = number of terms of the correction series that the smallest
// argument of the evaluation calls for (truncation below 1e-17): 4 for
// x >= 40 (next term 1 / (1188 x^9) = 3e-18), 5 for x >= 20, 7 for x >= 10.
template <int NT>
H3D_HD double stirling_core_nt(double x) {
    static_assert(NT == 4 || NT == 5 || NT == 7, "4, 5 or 7 terms");
    const double ix = fast_rcp_pos(x);
    const double ix2 = ix * ix;
#ifdef __CUDA_ARCH__
    double c = kDevStirling[NT - 1];
#pragma unroll
    for (int k = NT - 2; k >= 0; --k) c = fma(ix2, c, kDevStirling[k]);
    const double corr = ix * c;
#else
    const double cf[7] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0, 1.0 / 1188.0,
                          -691.0 / 360360.0, 1.0 / 156.0};
    double c = cf[NT - 1];
    for (int k = NT - 2; k >= 0; --k) c = fma(ix2, c, cf[k]);
    const double corr = ix * c;
#endif
    return fma(x - 0.5, fast_log_pos(x), corr);
}

H3D_HD double stirling_core(double x) { return stirling_core_nt<7>(x); }

// the same for arguments that may be below 10: ALL arguments of a likelihood
// evaluation are shifted up by the same n = ceil(10 - r) >= 1 unit steps (r > 0
// is the smallest possible argument, so x + n >= 10 for every pixel):
//   log Gamma(x) = log Gamma(x + n) - ln(x (x+1) ... (x+n-1)).
// A uniform n keeps the loop trip count identical across a thread block (no
// predication) and makes the "- n" terms a per-evaluation constant.  The
// product is >= r (r+1) ... (r+n-1) >= 3.7e3 for r >= 0.0101 (delta <= 100/101),
// i.e. far from 1, where the table logarithm is accurate.
H3D_HD double stirling_core_shifted(double x, int n) {
    double p = x, xs = x + 1.0;
    for (int k = 1; k < n; ++k) { p *= xs; xs += 1.0; }
    return stirling_core(xs) - fast_log_pos(p);
}

H3D_HD double lgamma_pos(double x) {
    double shift = 0.0;
    if (x < 10.0) {
        double p = 1.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (x < 10.0) { p *= x; x += 1.0; }
        }
        shift = m_log(p);
    }
    const double corr = stirling_sum(1.0 / x);
    return ((x - 0.5) * m_log(x) - x + 0.9189385332046727 + corr) - shift;
}

// ---------------------------------------------------------------------------
// LRT pieces.
// ---------------------------------------------------------------------------
// log NB pmf ratio summed over one replicate: logpmf(k; m0 f, phi) - logpmf(k; m1 f, phi)
// with the lgamma / r ln r terms cancelled analytically (scaled_nb.py:31-33):
//   (r + k) * log((r + m1) / (r + m0)) + k * log(m0 / m1)
H3D_HD double nb_llr_term(double k, double m0, double m1, double phi) {
    const double r = 1.0 / phi;
    double t = (r + k) * log1p((m1 - m0) / (r + m0));
    if (k > 0.0) t += k * log(m0 / m1);
    return t;
}

// chi-square survival function, df degrees of freedom (scipy chi2.sf =
// gammaincc(df/2, x/2); x <= 0 -> 1).
H3D_HD double chi2_sf(double x, int df) {
    if (isnan(x)) return NAN;
    if (!(x > 0.0)) return 1.0;
    if (df == 1) return erfc(sqrt(0.5 * x));
    if (df == 2) return exp(-0.5 * x);
    const double a = 0.5 * (double)df;
    return gamma_q(a, 0.5 * x, lgamma(a));
}

// ---------------------------------------------------------------------------
// One evaluation step of scipy's bounded Brent minimiser
// (scipy/optimize/_optimize.py:2289-2436, method='bounded', xatol=1e-5,
// maxiter=500), restructured as a resumable state machine: ``begin`` yields
// the first abscissa, ``advance`` consumes f(x_eval) and either yields the next
// abscissa (returns true) or finishes (returns false; xf is the minimiser).
// ---------------------------------------------------------------------------
struct BrentState {
    double a, b, xf, fx, nfc, fnfc, fulc, ffulc, e, rat, x_eval;
    int num;      // function evaluations so far
    int flag;     // 0 ok, 1 maxfun reached, 2 NaN
};

H3D_HD bool brent_propose(BrentState& s, double xatol) {
    // loop head of the reference: test convergence, otherwise choose x_eval
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    const double xm = 0.5 * (s.a + s.b);
    const double tol1 = sqrt_eps * fabs(s.xf) + xatol / 3.0;
    const double tol2 = 2.0 * tol1;
    if (!(fabs(s.xf - xm) > (tol2 - 0.5 * (s.b - s.a)))) return false;
    bool golden = true;
    if (fabs(s.e) > tol1) {
        golden = false;
        double r = (s.xf - s.nfc) * (s.fx - s.ffulc);
        double q = (s.xf - s.fulc) * (s.fx - s.fnfc);
        double p = (s.xf - s.fulc) * q - (s.xf - s.nfc) * r;
        q = 2.0 * (q - r);
        if (q > 0.0) p = -p;
        q = fabs(q);
        r = s.e;
        s.e = s.rat;
        if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (s.a - s.xf)) &&
            (p < q * (s.b - s.xf))) {
            s.rat = (p + 0.0) / q;
            const double x = s.xf + s.rat;
            if (((x - s.a) < tol2) || ((s.b - x) < tol2)) {
                const double dm = xm - s.xf;
                const double si = (dm > 0.0) - (dm < 0.0) + (dm == 0.0);
                s.rat = tol1 * si;
            }
        } else {
            golden = true;
        }
    }
    if (golden) {
        s.e = (s.xf >= xm) ? (s.a - s.xf) : (s.b - s.xf);
        s.rat = golden_mean * s.e;
    }
    const double si = (s.rat > 0.0) - (s.rat < 0.0) + (s.rat == 0.0);
    s.x_eval = s.xf + si * fmax(fabs(s.rat), tol1);
    return true;
}

H3D_HD void brent_begin(BrentState& s, double lo, double hi) {
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    s.a = lo; s.b = hi;
    s.fulc = lo + golden_mean * (hi - lo);
    s.nfc = s.fulc; s.xf = s.fulc;
    s.rat = 0.0; s.e = 0.0;
    s.x_eval = s.xf;
    s.num = 0; s.flag = 0;
    s.fx = s.fnfc = s.ffulc = 0.0;
}

// returns true when another evaluation at s.x_eval is wanted
H3D_HD bool brent_advance(BrentState& s, double fu, double xatol, int maxfun) {
    if (s.num == 0) {
        s.fx = fu; s.num = 1;
        s.ffulc = s.fnfc = fu;
        if (isnan(fu)) { s.flag = 2; return false; }
        return brent_propose(s, xatol);
    }
    const double x = s.x_eval;
    s.num += 1;
    if (fu <= s.fx) {
        if (x >= s.xf) s.a = s.xf; else s.b = s.xf;
        s.fulc = s.nfc; s.ffulc = s.fnfc;
        s.nfc = s.xf; s.fnfc = s.fx;
        s.xf = x; s.fx = fu;
    } else {
        if (x < s.xf) s.a = x; else s.b = x;
        if ((fu <= s.fnfc) || (s.nfc == s.xf)) {
            s.fulc = s.nfc; s.ffulc = s.fnfc;
            s.nfc = x; s.fnfc = fu;
        } else if ((fu <= s.ffulc) || (s.fulc == s.xf) || (s.fulc == s.nfc)) {
            s.fulc = x; s.ffulc = fu;
        }
    }
    if (isnan(fu) || isnan(s.fx) || isnan(s.xf)) { s.flag = 2; return false; }
    if (s.num >= maxfun) { s.flag = 1; return false; }
    return brent_propose(s, xatol);
}

}  // namespace h3d
