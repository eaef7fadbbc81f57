// Per-(distance, condition) dispersion estimation: qCML, CML, MME.
//
// Replaces hic3defdr/util/dispersion.py:10-131 (qcml, cml, mme,
// mme_per_pixel) and hic3defdr/util/scaled_nb.py:186-275 (equalize,
// q2qnbinom), i.e. the body of the distance loop at
// hic3defdr/analysis/analysis.py:191-206.
//
// The reference runs one scipy optimisation per (distance, condition) bin in
// a process pool.  Here ALL bins advance in lock step: the pooled pixels sit
// distance-major in SoA arrays, every launch works on every bin that is in the
// corresponding phase, and the tiny per-bin control state (qCML fixed point,
// Brent bracket) lives in device arrays updated by a one-thread-per-bin kernel.
//   equalize_kernel : per pixel fit_mu + q2q pseudo-data at the bin's current
//                     dispersion (FP64 special functions; the dominant cost)
//   nll_kernel      : the conditional NB log-likelihood of every searching bin
//                     at its current Brent abscissa, summed in 128-bit fixed
//                     point (exact, hence independent of the pixel order and
//                     of how many GPUs pooled the pixels)
//   step_kernel     : advances scipy's bounded Brent state machine, applies
//                     the qCML stopping rule |delta disp| <= 1e-4
// The host queues rounds of these ahead of the device and reads the per-round
// convergence counters from mapped memory without synchronising the stream.
// FP64-pipe bound: ~44 k FP64 instruction-
// equivalents per pixel against 16 R_c bytes read + 8 R_c written per sweep.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"

namespace h3d {

constexpr int kChunk = 1024;        // pixels per partial sum
constexpr double kQcmlTol = 1e-4;   // dispersion.py:10
constexpr double kBrentXatol = 1e-5;
constexpr int kBrentMaxfun = 500;
constexpr double kDeltaLo = 1e-4, kDeltaHi = 100.0 / 101.0;   // dispersion.py:77

enum : int { ST_EMPTY = 0, ST_NEED_EQ = 1, ST_IN_BRENT = 2, ST_DONE = 3, ST_FAILED = 4, ST_DONE_CAPPED = 5 };

struct CondReps {
    int rep[H3D_MAX_CONDS][H3D_MAX_REPS];   // replicate indices of each condition
    int n_in[H3D_MAX_CONDS];
    int pseudo_row[H3D_MAX_CONDS];          // first row of the condition in the pseudo buffer
    int n_conds;
};

struct Problem {                 // one (segment, condition)
    BrentState brent;
    double disp;
    long long n_px;
    int status, outer_iters, nfev_total, pad;
};

struct Counters { int n_need_eq, n_in_brent, n_failed, n_fit_failed; };

__global__ void init_problems_kernel(Problem* __restrict__ prob, const long long* __restrict__ seg_npx,
                                     int n_seg, int n_conds, int estimator, Counters* cnt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { cnt->n_need_eq = 0; cnt->n_in_brent = 0; cnt->n_failed = 0; cnt->n_fit_failed = 0; }
    if (p >= n_seg * n_conds) return;
    const int s = p / n_conds;
    Problem& q = prob[p];
    q.n_px = seg_npx[s];
    q.outer_iters = 0; q.nfev_total = 0;
    q.disp = (q.n_px > 0) ? 0.01 : NAN;                 // dispersion.py:33 / analysis.py:205
    q.status = (q.n_px > 0) ? ST_NEED_EQ : ST_EMPTY;
    // a bin that waits for pseudo-data has its Brent search already set up: the
    // likelihood kernel of the same round evaluates the first abscissa
    brent_begin(q.brent, kDeltaLo, kDeltaHi);
    (void)estimator;
}

// pseudo-data for every pixel of every bin that waits for it.
// grid = (n_chunks, n_conds); blockDim = 256; chunk -> (segment, first pixel)
//
// Two phases per batch of pixels, staged through shared memory:
//   1. per pixel: fit_mu, the clamped (mu_in, mu_out) of every replicate;
//      replicates whose quantile map needs no incomplete gamma function
//      (x = 0 in the left tail) are finished here, the others become tasks
//      (x, mu_in, mu_out, destination), tasks whose incomplete gamma function
//      is evaluated by the continued fraction filling the task array from
//      the front, power-series tasks from the back;
//   2. the tasks are processed in array order, so that the lanes of a warp
//      run the same branch of the incomplete gamma code instead of diverging
//      per replicate.
#ifndef H3D_EQ_TASKS
#define H3D_EQ_TASKS 512
#endif
constexpr int kEqTasks = H3D_EQ_TASKS;      // task slots per batch (32 B of shared memory each)

struct EqTask { double x, mu_in, mu_out; };

// Task order inside a batch: continued-fraction tasks first, series tasks
// after them, and inside each of the two lists by the predicted trip count of
// the incomplete-gamma loop, in cost buckets of a quarter octave each (counting
// sort through shared-memory counters), so that the 32 tasks a warp runs
// together take the same branch and have similar lengths (A/B on B200, mouse
// genome: 84.9 -> 81.5 ms per step; 2 or 6 buckets per octave are within 1 %,
// 8 are slower).  Warps then take groups of 32 consecutive tasks from the
// expensive end of the order through a shared counter, so that no warp waits
// at the batch's closing barrier for longer than one cheap group (barrier
// stalls were 13.6 % of the warp samples with a static assignment, ncu r01g).
// Also measured and rejected: splitting the quantile map after the first
// evaluation of its inverse and running the ~40 % of the tasks that need
// another one as a compacted second pass -- the resumable form of the Halley
// loop costs more (spills) than the idle lanes it saves.
#ifndef H3D_EQ_BUCKETS
#define H3D_EQ_BUCKETS 48
#endif
constexpr int kEqNB = H3D_EQ_BUCKETS;
#ifndef H3D_EQ_BUCKET_SCALE
#define H3D_EQ_BUCKET_SCALE 4.0f            // buckets per octave of predicted trip count
#endif

// predicted iterations of the series / continued fraction at (a, x), FP32, only
// a scheduling hint.  Both grow like sqrt(x) and shrink with the distance from
// the mode in units of sqrt(x): the series needs n terms with
// n^2 + 2 n (a - x) = 72 x (terms below 2^-52 of the sum), four per iteration.
__device__ __forceinline__ int eq_cost_bucket(float a, float x, bool series) {
    const float sx = sqrtf(fmaxf(x, 0.25f));
    const float z = (a - x) / sx;
    float pred;
    if (series) pred = 1.0f + 0.25f * sx * (sqrtf(z * z + 72.0f) - z);
    else pred = 1.0f + sx * (1.0f - 0.14f * fminf(fmaxf(-z, 0.0f), 4.0f));
    int b = (int)(__log2f(pred) * H3D_EQ_BUCKET_SCALE);
    return b < 0 ? 0 : (b >= kEqNB ? kEqNB - 1 : b);
}

// 4 resident CTAs per SM (<= 64 registers, a few spilled words) beat 3 without
// spills by 18 %: the kernel waits on dependent FP64 latency, and 32 warps hide
// more of it than 24 (A/B on B200: 105 -> 87 ms per step; 5 CTAs: 92 ms).
#ifndef H3D_EQ_MIN_BLOCKS
#define H3D_EQ_MIN_BLOCKS 4
#endif
template <int MAXRC>
__global__ void __launch_bounds__(256, H3D_EQ_MIN_BLOCKS)
equalize_kernel(const double* __restrict__ x, const double* __restrict__ f, long long ld,
                const int* __restrict__ chunk_seg, const long long* __restrict__ chunk_lo,
                const long long* __restrict__ chunk_hi, CondReps cr, int estimator,
                const Problem* __restrict__ prob, double* __restrict__ pseudo, Counters* cnt) {
    extern __shared__ unsigned char eq_smem[];
    EqTask* task = (EqTask*)eq_smem;                                    // [slot]
    int* task_code = (int*)(eq_smem + (size_t)kEqTasks * sizeof(EqTask));   // [slot]: (list << 16) | position, -1: none
    int* order = task_code + kEqTasks;                                  // [position in processing order] -> slot
    __shared__ int bucket_cnt[2 * kEqNB];       // [continued fraction | series] x cost bucket
    __shared__ int bucket_base[2 * kEqNB + 1];
    __shared__ int next_group;
    const int c = blockIdx.y;
    const int s = chunk_seg[blockIdx.x];
    const Problem& q = prob[s * cr.n_conds + c];
    if (q.status != ST_NEED_EQ) return;
    const double alpha = q.disp;
    const long long lo = chunk_lo[blockIdx.x];
    const long long hi = chunk_hi[blockIdx.x];
    const int nr = cr.n_in[c];
    const int lane = threadIdx.x & 31;
    double* __restrict__ out_base = pseudo + (long long)cr.pseudo_row[c] * ld;
    constexpr int kBatchPx = kEqTasks / MAXRC;
    if (estimator == H3D_EST_CML) {
        // cml(raw, f): data / f (dispersion.py:67-68, intended semantics)
        for (long long i = lo + threadIdx.x; i < hi; i += 256)
#pragma unroll
            for (int k = 0; k < MAXRC; ++k)
                if (k < nr) {
                    const int r = cr.rep[c][k];
                    out_base[(long long)k * ld + i] = x[(long long)r * ld + i] / f[(long long)r * ld + i];
                }
        return;
    }
    for (long long b0 = lo; b0 < hi; b0 += kBatchPx) {
        for (int b = threadIdx.x; b < 2 * kEqNB; b += 256) bucket_cnt[b] = 0;
        if (threadIdx.x == 0) next_group = 0;
        __syncthreads();
        const long long b1 = (b0 + kBatchPx < hi) ? b0 + kBatchPx : hi;
        // ---- per pixel: fit_mu, clamped means, task records -------------
        for (long long i0 = b0; i0 < b1; i0 += 256) {
            const long long i = i0 + threadIdx.x;
            const bool valid = i < b1;
            double xr[MAXRC], fr[MAXRC], ar[MAXRC];
            double slog = 0.0;
#pragma unroll
            for (int k = 0; k < MAXRC; ++k) {
                xr[k] = 0.0; fr[k] = 1.0; ar[k] = alpha;
                if (valid && k < nr) {
                    const int r = cr.rep[c][k];
                    xr[k] = x[(long long)r * ld + i];
                    fr[k] = f[(long long)r * ld + i];
                    slog += m_log(fr[k]);
                }
            }
            // equalize (scaled_nb.py:207-214)
            double mu_hat = 1.0, mu_out = 1.0;
            if (valid) {
                const double f_mean = m_exp(slog / (double)nr);     // gmean, pseudocount 0
                int st = 0;
                mu_hat = fit_mu<MAXRC>(xr, fr, ar, (1u << nr) - 1u, &st);
                if (st) atomicAdd(&cnt->n_fit_failed, 1);
                mu_out = mu_hat * f_mean;
            }
#pragma unroll
            for (int k = 0; k < MAXRC; ++k) {
                if (k < nr) {                                   // uniform across the block
                    double mu_in = mu_hat * fr[k];
                    // order-dependent clamp shared across replicates (scaled_nb.py:240-242)
                    if (!((mu_in >= 0.25) && (mu_out >= 0.25))) { mu_in = 0.25; mu_out = 0.25; }
                    const bool right = xr[k] >= mu_in;
                    const bool cheap = !right && !(xr[k] > 0.0);     // no gamma evaluation needed
                    if (valid && cheap)
                        out_base[(long long)k * ld + i] = q2q_zero(mu_in, mu_out, alpha);
                    const int slot = k * kBatchPx + (int)(i - b0);
                    int code = -1;
                    if (valid && !cheap) {
                        // which algorithm the tail evaluations of this task will use
                        // (FP32 copy of gamma_use_series: only a scheduling hint)
                        const float rin_f = 1.0f + (float)alpha * (float)mu_in;
                        const float a_f = (float)mu_in / rin_f, x_f = (float)xr[k] / rin_f;
                        const bool series = gamma_use_series((double)a_f, (double)x_f);
                        const int list = (series ? kEqNB : 0) + eq_cost_bucket(a_f, x_f, series);
                        // the task stays in its own (replicate, pixel) slot; its place in
                        // the processing order is settled once the batch's counters are complete
                        code = (list << 16) | atomicAdd(&bucket_cnt[list], 1);
                        task[slot].x = xr[k];
                        task[slot].mu_in = mu_in;
                        task[slot].mu_out = mu_out;
                    }
                    if (i - b0 < kBatchPx) task_code[slot] = code;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // exclusive scan of the 2 * kEqNB counters by warp 0: each lane owns a
            // contiguous run of them
            constexpr int kPer = (2 * kEqNB + 31) / 32;
            int own[kPer];
            int sum = 0;
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int b = (int)threadIdx.x * kPer + j;
                own[j] = (b < 2 * kEqNB) ? bucket_cnt[b] : 0;
                sum += own[j];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)threadIdx.x >= o) incl += v;
            }
            int run = incl - sum;
#pragma unroll
            for (int j = 0; j < kPer; ++j) {
                const int b = (int)threadIdx.x * kPer + j;
                if (b < 2 * kEqNB) bucket_base[b] = run;
                run += own[j];
            }
            if (threadIdx.x == 31) bucket_base[2 * kEqNB] = incl;
        }
        __syncthreads();
        {
            const int n_px = (int)(b1 - b0);
            for (int k = 0; k < nr; ++k)
                for (int j = threadIdx.x; j < n_px; j += 256) {
                    const int slot = k * kBatchPx + j;
                    const int code = task_code[slot];
                    if (code >= 0) order[bucket_base[code >> 16] + (code & 0xffff)] = slot;
                }
        }
        __syncthreads();
        {
            // groups of 32 consecutive tasks, from the expensive end of the order
            const int n_tasks = bucket_base[2 * kEqNB];
            const int n_groups = (n_tasks + 31) >> 5;
            while (true) {
                int g = 0;
                if (lane == 0) g = atomicAdd(&next_group, 1);
                g = __shfl_sync(0xffffffffu, g, 0);
                if (g >= n_groups) break;
                const int t = ((n_groups - 1 - g) << 5) + lane;
                if (t < n_tasks) {
                    const int slot = order[t];
                    const EqTask tk = task[slot];
                    const int k = slot / kBatchPx;
                    const long long i = b0 + (slot - k * kBatchPx);
                    out_base[(long long)k * ld + i] = q2q_one(tk.x, tk.mu_in, tk.mu_out, alpha);
                }
            }
        }
        __syncthreads();
    }
}

// Order-independent sums.  Every per-pixel term v is split, exactly, into a
// multiple of 2^-10 and a multiple of 2^-58 below it (two additions of a
// "magic" constant put the integer into the low mantissa bits: no conversion
// instruction), and the two integers are added in 64-bit integer arithmetic,
// which is associative and commutative: the sum of a bin does not depend on how
// its pixels are ordered, chunked, or spread over thread blocks -- and
// therefore not on how many GPUs pooled them.  What is
// dropped of a term is below 2^-59 (4e-18) whatever its size.  Terms that are
// not finite or reach 2^40 are counted in ``bad`` instead.
struct Fix128 {
    unsigned long long lo;   // units of 2^-58, kept in [0, 2^48) between blocks
    long long hi;            // units of 2^-10
    long long bad;           // number of non-representable terms
    long long cnt;           // number of terms (used by the MME mean)
};

// per-thread accumulator (registers): count-free, 32-bit flag
struct FixAcc { long long lo, hi; int bad; };

__device__ __forceinline__ void fix_add(FixAcc& a, double v) {
    constexpr double kM1 = 6597069766656.0;          // 1.5 * 2^42: ulp 2^-10
    constexpr double kM2 = 0.0234375;                // 1.5 * 2^-6:  ulp 2^-58
    if (!(fabs(v) < 1099511627776.0)) { a.bad += 1; return; }       // 2^40
    const double t1 = v + kM1;                       // rounds v to a multiple of 2^-10
    const double e = v - (t1 - kM1);                 // exact, |e| <= 2^-11
    const double t2 = e + kM2;                       // rounds e to a multiple of 2^-58
    a.hi += __double_as_longlong(t1) - __double_as_longlong(kM1);
    a.lo += __double_as_longlong(t2) - __double_as_longlong(kM2);
}

__host__ __device__ __forceinline__ double fix_value(const Fix128& a) {
    return (double)a.hi * 9.765625e-4 + (double)a.lo * 3.469446951953614189e-18;   // 2^-10, 2^-58
}

// block-wide sum of the threads' accumulators (and of a per-thread term
// count), normalised (lo in [0, 2^48)); the result is valid in thread 0
__device__ __forceinline__ Fix128 block_sum_fix(FixAcc v, int cnt, Fix128* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.lo += __shfl_down_sync(0xffffffffu, v.lo, o);
        v.hi += __shfl_down_sync(0xffffffffu, v.hi, o);
        v.bad += __shfl_down_sync(0xffffffffu, v.bad, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = Fix128{(unsigned long long)v.lo, v.hi, (long long)v.bad, (long long)cnt};
    __syncthreads();
    Fix128 t = {0ull, 0ll, 0ll, 0ll};
    if (threadIdx.x == 0) {
        long long lo = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            lo += (long long)sh[w].lo;
            t.hi += sh[w].hi;
            t.bad += sh[w].bad;
            t.cnt += sh[w].cnt;
        }
        const long long carry = lo >> 48;            // floor: 2^48 units of 2^-58 = one of 2^-10
        t.hi += carry;
        t.lo = (unsigned long long)(lo - (carry << 48));
    }
    return t;
}

// conditional NB negative log-likelihood (dispersion.py:72-75):
//   sum_px [ sum_k lgamma(y_k + r) + lgamma(n r) - lgamma(z + n r) - n lgamma(r) ]
// Every log-gamma goes through Stirling's series at an argument >= 10
// (shifted up by the recurrence when smaller); the "- x" terms of the
// (R_c + 1) log-gammas of a pixel cancel exactly (sum_k (y_k + r) == z + n r)
// and are never computed, so the per-pixel term is
//   sum_k [core(x_k) - n_k] - [core(X) - N] + (R_c - 1) .5 ln 2pi,
// core(x) = (x - .5) ln x + corr(x) at the shifted argument minus the log of
// the shift product, n = number of unit shifts.
// The per-pixel terms are added in 128-bit fixed point (order independent, see
// Fix128) into the bin's accumulator.
// 8 resident CTAs per SM (32 registers, 44 bytes of spills) against 6 without
// spills: 35.4 -> 34.5 ms per step -- the kernel waits on dependent FP64 latency
#ifndef H3D_NLL_MIN_BLOCKS
#define H3D_NLL_MIN_BLOCKS 8
#endif

// adds a block's total to the (segment, condition) accumulator; integer
// atomics commute, so the result is exact whatever the arrival order (every
// block contributes lo < 2^48: no overflow below 2^15 chunks = 32 M pixels of
// one bin, and hi has 2^53 units of headroom)
__device__ __forceinline__ void fix_atomic_add(Fix128* dst, const Fix128& t) {
    if (t.lo) atomicAdd(&dst->lo, t.lo);
    if (t.hi) atomicAdd((unsigned long long*)&dst->hi, (unsigned long long)t.hi);
    if (t.bad) atomicAdd((unsigned long long*)&dst->bad, (unsigned long long)t.bad);
    if (t.cnt) atomicAdd((unsigned long long*)&dst->cnt, (unsigned long long)t.cnt);
}

template <int MAXRC>
__global__ void __launch_bounds__(256, H3D_NLL_MIN_BLOCKS)
nll_kernel(const double* __restrict__ pseudo, long long ld, const int* __restrict__ chunk_seg,
           const long long* __restrict__ chunk_lo, const long long* __restrict__ chunk_hi,
           CondReps cr, const Problem* __restrict__ prob, Fix128* __restrict__ acc) {
    __shared__ Fix128 sh[8];
    const int c = blockIdx.y;
    const int s = chunk_seg[blockIdx.x];
    const Problem& q = prob[s * cr.n_conds + c];
    if (q.status != ST_IN_BRENT && q.status != ST_NEED_EQ) return;
    const double delta = q.brent.x_eval;
    const double r = 1.0 / delta - 1.0;
    const int nr = cr.n_in[c];
    const double nrr = (double)nr * r;
    const double cst = lgamma_pos(nrr) - (double)nr * lgamma_pos(r) +
                       (double)(nr - 1) * 0.9189385332046727;
    const long long lo = chunk_lo[blockIdx.x];
    const long long hi = chunk_hi[blockIdx.x];
    const double* __restrict__ base = pseudo + (long long)cr.pseudo_row[c] * ld;
    FixAcc a = {0ll, 0ll, 0};
    if (r >= 10.0) {
        // every argument is >= r: the number of Stirling correction terms is
        // uniform over the block (h3d_math.cuh, stirling_core_nt)
#define H3D_NLL_LOOP(NT)                                                        \
        for (long long i = lo + threadIdx.x; i < hi; i += 256) {                \
            double z = 0.0, t = 0.0;                                            \
            _Pragma("unroll")                                                   \
            for (int k = 0; k < MAXRC; ++k) {                                   \
                if (k < nr) {                                                   \
                    const double y = base[(long long)k * ld + i];               \
                    z += y;                                                     \
                    t += stirling_core_nt<NT>(y + r);                           \
                }                                                               \
            }                                                                   \
            fix_add(a, (t + cst) - stirling_core_nt<NT>(z + nrr));              \
        }
        if (r >= 40.0) { H3D_NLL_LOOP(4) }
        else if (r >= 20.0) { H3D_NLL_LOOP(5) }
        else { H3D_NLL_LOOP(7) }
#undef H3D_NLL_LOOP
    } else {
        // uniform shift of every argument by n_shift units (see stirling_core_shifted);
        // the "- n" terms of the (R_c + 1) shifted log-gammas leave (R_c - 1) n_shift
        const int n_shift = (int)ceil(10.0 - r);
        const double cst_s = cst - (double)(nr - 1) * (double)n_shift;
        for (long long i = lo + threadIdx.x; i < hi; i += 256) {
            double z = 0.0, t = 0.0;
#pragma unroll
            for (int k = 0; k < MAXRC; ++k) {
                if (k < nr) {
                    const double y = base[(long long)k * ld + i];
                    z += y;
                    t += stirling_core_shifted(y + r, n_shift);
                }
            }
            fix_add(a, (t + cst_s) - stirling_core_shifted(z + nrr, n_shift));
        }
    }
    const Fix128 tot = block_sum_fix(a, 0, sh);
    if (threadIdx.x == 0) fix_atomic_add(&acc[s * cr.n_conds + c], tot);
}

// MME: sum and count of the non-NaN per-pixel estimates
// (dispersion.py:101-105, 129-131)
template <int MAXRC>
__global__ void __launch_bounds__(256)
mme_kernel(const double* __restrict__ x, const double* __restrict__ f, long long ld,
           const int* __restrict__ chunk_seg, const long long* __restrict__ chunk_lo,
           const long long* __restrict__ chunk_hi, CondReps cr, Fix128* __restrict__ acc) {
    __shared__ Fix128 sh[8];
    const int c = blockIdx.y;
    const int s = chunk_seg[blockIdx.x];
    const long long lo = chunk_lo[blockIdx.x];
    const long long hi = chunk_hi[blockIdx.x];
    const int nr = cr.n_in[c];
    FixAcc a = {0ll, 0ll, 0};
    int n_est = 0;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        double v[MAXRC];
        double m = 0.0;
#pragma unroll
        for (int k = 0; k < MAXRC; ++k) {
            v[k] = 0.0;
            if (k < nr) {
                const int r = cr.rep[c][k];
                v[k] = x[(long long)r * ld + i] / f[(long long)r * ld + i];
                m += v[k];
            }
        }
        m /= (double)nr;
        double ss = 0.0;
#pragma unroll
        for (int k = 0; k < MAXRC; ++k)
            if (k < nr) { const double dlt = v[k] - m; ss += dlt * dlt; }
        const double var = ss / (double)(nr - 1);
        const double est = (var - m) / (m * m);             // inverse_mvr (scaled_nb.py:68)
        if (!isnan(est)) { fix_add(a, est); n_est += 1; }   // nanmean: infinities stay in
    }
    const Fix128 tot = block_sum_fix(a, n_est, sh);
    if (threadIdx.x == 0) fix_atomic_add(&acc[s * cr.n_conds + c], tot);
}

// one thread per problem.  mode 1: consume the bin's NLL accumulator, advance
// the Brent search and the qCML fixed point; mode 2: finish MME.
__global__ void step_kernel(Problem* __restrict__ prob, Fix128* __restrict__ acc,
                            int n_prob, int estimator, int mode, Counters* cnt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prob) return;
    Problem& q = prob[p];
    if (mode == 2) {
        if (q.status == ST_EMPTY) return;
        const Fix128 a = acc[p];
        // nanmean (dispersion.py:131); an infinite per-pixel estimate makes it
        // infinite (or NaN) there, NaN here
        q.disp = a.bad ? NAN : fix_value(a) / (double)a.cnt;
        q.status = ST_DONE;
        return;
    }
    if (q.status != ST_IN_BRENT && q.status != ST_NEED_EQ) return;
    const Fix128 a = acc[p];
    acc[p] = Fix128{0ull, 0ll, 0ll, 0ll};
    const double fu = a.bad ? NAN : -fix_value(a);
    const bool more = brent_advance(q.brent, fu, kBrentXatol, kBrentMaxfun);
    if (more) { q.status = ST_IN_BRENT; atomicAdd(&cnt->n_in_brent, 1); return; }
    q.nfev_total += q.brent.num;
    if (q.brent.flag != 0) {                                // assert res.success (dispersion.py:78)
        q.status = ST_FAILED; q.disp = NAN;
        atomicAdd(&cnt->n_failed, 1);
        return;
    }
    const double dh = q.brent.xf;
    const double nd = dh / (1.0 - dh);
    q.outer_iters += 1;
    if (estimator == H3D_EST_CML) { q.disp = nd; q.status = ST_DONE; return; }
    const double dl = fabs(q.disp - nd);                    // dispersion.py:36-42
    q.disp = nd;
    if (dl > kQcmlTol && q.outer_iters < H3D_QCML_MAX_OUTER) {
        // next outer iteration: pseudo-data at the new dispersion, new search
        brent_begin(q.brent, kDeltaLo, kDeltaHi);
        q.status = ST_NEED_EQ; atomicAdd(&cnt->n_need_eq, 1);
    } else {
        q.status = (dl > kQcmlTol) ? ST_DONE_CAPPED : ST_DONE;     // see include/h3d.h
    }
}

// End of a round: the convergence counters go to the host through pinned,
// device-mapped memory (written by an SM: must not queue on the copy engine
// behind bulk output copies of other streams), sequence number last; the
// per-round counters are cleared for the next round.
struct RoundSlot { unsigned long long seq; int n_need_eq, n_in_brent, n_failed, n_fit_failed; unsigned long long pad; };
static_assert(sizeof(RoundSlot) == 32, "RoundSlot layout");

__global__ void publish_round_kernel(Counters* cnt, volatile RoundSlot* slot, unsigned long long seq) {
    slot->n_need_eq = cnt->n_need_eq;
    slot->n_in_brent = cnt->n_in_brent;
    slot->n_failed = cnt->n_failed;
    slot->n_fit_failed = cnt->n_fit_failed;
    __threadfence_system();
    slot->seq = seq;
    cnt->n_need_eq = 0; cnt->n_in_brent = 0;
}

__global__ void publish_words_kernel(const unsigned long long* __restrict__ src,
                                     volatile unsigned long long* __restrict__ dst, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
}

__global__ void collect_kernel(const Problem* __restrict__ prob, int n, double* __restrict__ disp_out,
                               long long* __restrict__ stats) {
    // single block
    __shared__ long long it_sum, fev_sum, eq_sum, cap_sum;
    if (threadIdx.x == 0) { it_sum = 0; fev_sum = 0; eq_sum = 0; cap_sum = 0; }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        disp_out[p] = prob[p].disp;
        atomicAdd((unsigned long long*)&it_sum, (unsigned long long)prob[p].outer_iters);
        atomicAdd((unsigned long long*)&fev_sum, (unsigned long long)prob[p].nfev_total);
        atomicAdd((unsigned long long*)&eq_sum,
                  (unsigned long long)((long long)prob[p].outer_iters * prob[p].n_px));
        if (prob[p].status == ST_DONE_CAPPED) atomicAdd((unsigned long long*)&cap_sum, 1ull);
    }
    __syncthreads();
    if (threadIdx.x == 0) { stats[0] = it_sum; stats[1] = fev_sum; stats[2] = eq_sum; stats[3] = cap_sum; }
}

// one problem per condition over a single segment, waiting for pseudo-data at a
// given dispersion (h3d_equalize)
__global__ void single_bin_problems_kernel(Problem* prob, int n_conds, long long n_px, double alpha,
                                           double delta, int status, Counters* cnt) {
    const int p = threadIdx.x;
    if (p == 0) { cnt->n_need_eq = 0; cnt->n_in_brent = 0; cnt->n_failed = 0; cnt->n_fit_failed = 0; }
    if (p >= n_conds) return;
    prob[p].n_px = n_px; prob[p].outer_iters = 0; prob[p].nfev_total = 0;
    prob[p].disp = alpha; prob[p].status = status;
    brent_begin(prob[p].brent, kDeltaLo, kDeltaHi);
    prob[p].brent.x_eval = delta;
}

__global__ void fix_value_kernel(const Fix128* acc, double* out) {
    out[0] = acc->bad ? NAN : -fix_value(*acc);
}

}  // namespace h3d

using namespace h3d;

// per-thread pinned scratch for control read-backs (grown on demand): entry
// points may be driven from several host threads, each with its own stream
struct PinnedScratch {
    void* p = nullptr;
    size_t bytes = 0;
    ~PinnedScratch() { if (p) cudaFreeHost(p); }
};
static thread_local PinnedScratch g_pinned;
static int pinned_scratch(size_t bytes, void** host, void** dev) {
    if (bytes > g_pinned.bytes) {
        if (g_pinned.p) cudaFreeHost(g_pinned.p);
        g_pinned.p = nullptr; g_pinned.bytes = 0;
        H3D_CHECK(cudaHostAlloc(&g_pinned.p, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
        g_pinned.bytes = bytes;
    }
    *host = g_pinned.p;
    H3D_CHECK(cudaHostGetDevicePointer(dev, g_pinned.p, 0));
    return H3D_OK;
}

// CUDA events that are destroyed on every return path
struct EventPool {
    std::vector<cudaEvent_t> ev;
    int create(int n) {
        ev.assign(n, nullptr);
        for (int k = 0; k < n; ++k) H3D_CHECK(cudaEventCreate(&ev[k]));
        return H3D_OK;
    }
    ~EventPool() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
};

static int make_cond_reps(const unsigned char* design_host, int n_reps, int n_conds, CondReps* cr,
                          int* rows_out, int* max_rc_out) {
    cr->n_conds = n_conds;
    int rows = 0, max_rc = 0;
    for (int c = 0; c < H3D_MAX_CONDS; ++c) {
        cr->n_in[c] = 0; cr->pseudo_row[c] = rows;
        for (int k = 0; k < H3D_MAX_REPS; ++k) cr->rep[c][k] = 0;
        if (c >= n_conds) continue;
        for (int r = 0; r < n_reps; ++r)
            if (design_host[r * n_conds + c]) cr->rep[c][cr->n_in[c]++] = r;
        H3D_REQUIRE(cr->n_in[c] >= 1, "condition without replicates");
        rows += cr->n_in[c];
        if (cr->n_in[c] > max_rc) max_rc = cr->n_in[c];
    }
    *rows_out = rows; *max_rc_out = max_rc;
    return H3D_OK;
}

// Chunk tables, built on the host and copied.  A segment (one distance) is a
// list of runs [lo, hi) of the pooled arrays -- one run in a one-process
// estimate, one run per source rank after a multi-GPU exchange (the likelihood
// sums do not depend on the pixel order, so the runs need not be regrouped) --
// and every run is cut into chunks of at most kChunk pixels.
struct ChunkTables { int n_chunks; int* chunk_seg; long long *chunk_lo, *chunk_hi, *seg_npx; };
static int make_chunk_tables(const int* run_seg, const long long* run_lo, const long long* run_hi,
                             int n_runs, int n_seg, Workspace& w, cudaStream_t st, ChunkTables* t) {
    long long n_chunks_ll = 0;
    for (int r = 0; r < n_runs; ++r) {
        H3D_REQUIRE(run_seg[r] >= 0 && run_seg[r] < n_seg && run_hi[r] >= run_lo[r] && run_lo[r] >= 0,
                    "malformed run");
        n_chunks_ll += (run_hi[r] - run_lo[r] + kChunk - 1) / kChunk;
    }
    H3D_REQUIRE(n_chunks_ll < 2147483647LL, "too many chunks");
    const int n_chunks = (int)n_chunks_ll;
    std::vector<int> h_chunk_seg(n_chunks);
    std::vector<long long> h_chunk_lo(n_chunks), h_chunk_hi(n_chunks), h_seg_npx(n_seg, 0);
    int k = 0;
    for (int r = 0; r < n_runs; ++r) {
        h_seg_npx[run_seg[r]] += run_hi[r] - run_lo[r];
        for (long long lo = run_lo[r]; lo < run_hi[r]; lo += kChunk) {
            h_chunk_seg[k] = run_seg[r]; h_chunk_lo[k] = lo;
            h_chunk_hi[k] = (lo + kChunk < run_hi[r]) ? lo + kChunk : run_hi[r];
            ++k;
        }
    }
    t->n_chunks = n_chunks;
    t->chunk_seg = w.take<int>(n_chunks);
    t->chunk_lo = w.take<long long>(n_chunks);
    t->chunk_hi = w.take<long long>(n_chunks);
    t->seg_npx = w.take<long long>(n_seg);
    if (!t->chunk_seg || !t->chunk_lo || !t->chunk_hi || !t->seg_npx) {
        set_error("dispersion workspace too small (%zu bytes given)", w.size);
        return H3D_ERR_WORKSPACE;
    }
    H3D_CHECK(cudaMemcpyAsync(t->chunk_seg, h_chunk_seg.data(), (size_t)n_chunks * 4, cudaMemcpyHostToDevice, st));
    H3D_CHECK(cudaMemcpyAsync(t->chunk_lo, h_chunk_lo.data(), (size_t)n_chunks * 8, cudaMemcpyHostToDevice, st));
    H3D_CHECK(cudaMemcpyAsync(t->chunk_hi, h_chunk_hi.data(), (size_t)n_chunks * 8, cudaMemcpyHostToDevice, st));
    H3D_CHECK(cudaMemcpyAsync(t->seg_npx, h_seg_npx.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, st));
    // the host vectors must outlive the async copies
    H3D_CHECK(cudaStreamSynchronize(st));
    return H3D_OK;
}

static int set_equalize_smem(size_t eq_smem) {
    H3D_CHECK(cudaFuncSetAttribute(equalize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eq_smem));
    H3D_CHECK(cudaFuncSetAttribute(equalize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eq_smem));
    H3D_CHECK(cudaFuncSetAttribute(equalize_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eq_smem));
    H3D_CHECK(cudaFuncSetAttribute(equalize_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eq_smem));
    return H3D_OK;
}

#define DISPATCH_RC(CALL)                       \
    if (max_rc <= 2) { CALL(2); }               \
    else if (max_rc <= 4) { CALL(4); }          \
    else if (max_rc <= 8) { CALL(8); }          \
    else { CALL(16); }

static size_t disp_ws_bytes(long long n_px, int n_runs, int n_seg, int n_reps, int n_conds) {
    const long long max_chunks = n_px / kChunk + n_runs + 1;
    size_t b = 0;
    b += ws_pad((size_t)n_reps * n_conds * n_px * 8);               // pseudo (worst case)
    b += ws_pad((size_t)max_chunks * 4) + 2 * ws_pad((size_t)max_chunks * 8);
    b += ws_pad((size_t)(n_seg + 1) * 8);
    b += ws_pad((size_t)n_seg * n_conds * sizeof(Problem));
    b += ws_pad((size_t)n_seg * n_conds * sizeof(Fix128));
    b += ws_pad((size_t)n_seg * n_conds * 8) + ws_pad(64) + ws_pad(64);
    return b;
}

extern "C" size_t h3d_estimate_dispersion_ws_bytes(long long n_px, int n_seg, int n_reps, int n_conds) {
    return disp_ws_bytes(n_px, n_seg, n_seg, n_reps, n_conds);
}

extern "C" size_t h3d_estimate_dispersion_runs_ws_bytes(long long n_px, int n_runs, int n_seg, int n_reps,
                                                        int n_conds) {
    return disp_ws_bytes(n_px, n_runs, n_seg, n_reps, n_conds);
}

// rounds the host keeps queued ahead of the last round whose counters it has seen
static int qcml_run_ahead() {
    const char* e = getenv("H3D_QCML_AHEAD");
    int v = e ? atoi(e) : 6;
    return v < 1 ? 1 : (v > 64 ? 64 : v);
}

extern "C" int h3d_estimate_dispersion(const double* x, const double* f, long long ld,
                                       const long long* seg_start_host, int n_seg,
                                       const unsigned char* design_host, int n_reps, int n_conds,
                                       int estimator, double* disp_per_dist_host,
                                       long long* stats_host, void* ws, size_t ws_bytes,
                                       h3d_stream_t stream) {
    H3D_REQUIRE(n_seg >= 1, "no segments");
    H3D_REQUIRE(seg_start_host[0] == 0, "segments must start at 0");
    std::vector<int> run_seg(n_seg);
    for (int s = 0; s < n_seg; ++s) run_seg[s] = s;
    return h3d_estimate_dispersion_runs(x, f, ld, run_seg.data(), seg_start_host, seg_start_host + 1,
                                        n_seg, n_seg, design_host, n_reps, n_conds, estimator,
                                        disp_per_dist_host, stats_host, ws, ws_bytes, stream);
}

extern "C" int h3d_estimate_dispersion_runs(const double* x, const double* f, long long ld,
                                            const int* run_seg_host, const long long* run_lo_host,
                                            const long long* run_hi_host, int n_runs, int n_seg,
                                            const unsigned char* design_host, int n_reps, int n_conds,
                                            int estimator, double* disp_per_dist_host,
                                            long long* stats_host, void* ws, size_t ws_bytes,
                                            h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_conds >= 1 && n_conds <= H3D_MAX_CONDS, "n_conds out of range");
    H3D_REQUIRE(estimator >= 0 && estimator <= 2, "unknown estimator");
    H3D_REQUIRE(n_seg >= 1, "no segments");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long launches_before = h3d_launch_count();
    CondReps cr;
    int rows = 0, max_rc = 0;
    { int rc = make_cond_reps(design_host, n_reps, n_conds, &cr, &rows, &max_rc); if (rc) return rc; }
    long long n_px = 0;
    for (int r = 0; r < n_runs; ++r) {
        H3D_REQUIRE(run_hi_host[r] <= ld, "runs must fit in ld");
        n_px += run_hi_host[r] - run_lo_host[r];
    }
    const int n_prob = n_seg * n_conds;
    for (int p = 0; p < n_prob; ++p) disp_per_dist_host[p] = NAN;
    if (stats_host) for (int k = 0; k < 9; ++k) stats_host[k] = 0;
    if (n_px == 0) return H3D_OK;

    Workspace w(ws, ws_bytes);
    ChunkTables ct;
    { int rc = make_chunk_tables(run_seg_host, run_lo_host, run_hi_host, n_runs, n_seg, w, st, &ct); if (rc) return rc; }
    const int n_chunks = ct.n_chunks;
    double* pseudo = w.take<double>((size_t)rows * ld);
    Problem* prob = w.take<Problem>(n_prob);
    Fix128* acc = w.take<Fix128>(n_prob);
    double* disp_dev = w.take<double>(n_prob);
    Counters* cnt = w.take<Counters>(1);
    long long* stats_dev = w.take<long long>(4);
    if (!pseudo || !prob || !acc || !disp_dev || !cnt || !stats_dev) {
        set_error("estimate_dispersion workspace too small (%zu bytes given)", ws_bytes);
        return H3D_ERR_WORKSPACE;
    }
    H3D_CHECK(cudaMemsetAsync(acc, 0, (size_t)n_prob * sizeof(Fix128), st));

    const int pgrid = div_up(n_prob, 128);
    init_problems_kernel<<<pgrid, 128, 0, st>>>(prob, ct.seg_npx, n_seg, n_conds, estimator, cnt);
    H3D_LAUNCHED("init_problems_kernel");
    const dim3 cgrid(n_chunks, n_conds);
    const int ahead = qcml_run_ahead();
    const int ring = 2 * ahead;
    void *pin_host = nullptr, *pin_dev = nullptr;
    const size_t ring_bytes = (size_t)ring * sizeof(RoundSlot);
    const size_t off_stats = (ring_bytes + 255) & ~(size_t)255, off_disp = off_stats + 256;
    {
        int rc = pinned_scratch(off_disp + (size_t)n_prob * 8 + 64, &pin_host, &pin_dev);
        if (rc) return rc;
    }
    volatile RoundSlot* ring_host = (volatile RoundSlot*)pin_host;
    RoundSlot* ring_dev = (RoundSlot*)pin_dev;
    for (int k = 0; k < ring; ++k) ring_host[k].seq = 0;     // nothing of an earlier call is in flight
    const size_t eq_smem = (size_t)kEqTasks * (sizeof(EqTask) + 2 * sizeof(int));
    { int rc = set_equalize_smem(eq_smem); if (rc) return rc; }
    // per-kernel device time of the two heavy kernels (CUDA events on the
    // launching stream, read once the round's counters have arrived)
    EventPool evp;
    { int rc = evp.create(4 * ring); if (rc) return rc; }
    double eq_us = 0.0, nll_us = 0.0;
    long long eq_launches = 0, nll_launches = 0;

    if (estimator == H3D_EST_MME) {
#define CALL(M) mme_kernel<M><<<cgrid, 256, 0, st>>>(x, f, ld, ct.chunk_seg, ct.chunk_lo, ct.chunk_hi, cr, acc)
        DISPATCH_RC(CALL)
#undef CALL
        H3D_LAUNCHED("mme_kernel");
        step_kernel<<<pgrid, 128, 0, st>>>(prob, acc, n_prob, estimator, 2, cnt);
        H3D_LAUNCHED("step_kernel");
    } else {
        // Rounds of [pseudo-data for the bins that wait for it | likelihood of
        // every searching bin at its current abscissa | Brent / fixed-point
        // update | counters to the host].  Every kernel skips the bins that are
        // not in the matching state, so a round queued after the last useful
        // one is a no-op: the host keeps ``ahead`` rounds queued and reads the
        // counters of finished rounds from mapped memory WITHOUT synchronising
        // the stream (one host-device round trip per round was 2.4 ms of a
        // 149 ms step, and 13 % of an 8-GPU step).
        int launched = 0, seen = 0;
        bool done = false;
        int prev_need_eq = 1;                 // the first round computes pseudo-data
        int err = H3D_OK;
        while (!done) {
            while (launched - seen < ahead) {
                cudaEvent_t* ev = &evp.ev[4 * (launched % ring)];
                H3D_CHECK(cudaEventRecord(ev[0], st));
#define CALL(M) equalize_kernel<M><<<cgrid, 256, eq_smem, st>>>(x, f, ld, ct.chunk_seg, ct.chunk_lo, ct.chunk_hi, cr, \
        estimator, prob, pseudo, cnt)
                DISPATCH_RC(CALL)
#undef CALL
                H3D_LAUNCHED("equalize_kernel");
                H3D_CHECK(cudaEventRecord(ev[1], st));
#define CALL(M) nll_kernel<M><<<cgrid, 256, 0, st>>>(pseudo, ld, ct.chunk_seg, ct.chunk_lo, ct.chunk_hi, cr, prob, acc)
                DISPATCH_RC(CALL)
#undef CALL
                H3D_LAUNCHED("nll_kernel");
                H3D_CHECK(cudaEventRecord(ev[2], st));
                step_kernel<<<pgrid, 128, 0, st>>>(prob, acc, n_prob, estimator, 1, cnt);
                H3D_LAUNCHED("step_kernel");
                publish_round_kernel<<<1, 1, 0, st>>>(cnt, ring_dev + (launched % ring),
                                                     (unsigned long long)launched + 1ull);
                H3D_LAUNCHED("publish_round_kernel");
                ++launched;
                if (launched > 200000) break;
            }
            // counters of round ``seen``
            volatile RoundSlot* slot = ring_host + (seen % ring);
            long long spins = 0;
            while (slot->seq != (unsigned long long)seen + 1ull) {
                if ((++spins & 0xffff) == 0) {
                    const cudaError_t qe = cudaStreamQuery(st);
                    if (qe != cudaSuccess && qe != cudaErrorNotReady) {
                        set_error("qCML round %d: %s", seen, cudaGetErrorString(qe));
                        return H3D_ERR_CUDA;
                    }
                    if (qe == cudaSuccess && slot->seq != (unsigned long long)seen + 1ull) {
                        set_error("qCML round %d finished without publishing its counters", seen);
                        return H3D_ERR_CUDA;
                    }
                }
            }
            __sync_synchronize();
            const int n_need_eq = slot->n_need_eq, n_in_brent = slot->n_in_brent;
            const int n_failed = slot->n_failed, n_fit_failed = slot->n_fit_failed;
            {
                cudaEvent_t* ev = &evp.ev[4 * (seen % ring)];
                float ms = 0.f;
                if (prev_need_eq) {
                    H3D_CHECK(cudaEventElapsedTime(&ms, ev[0], ev[1]));
                    eq_us += 1e3 * (double)ms; ++eq_launches;
                }
                H3D_CHECK(cudaEventElapsedTime(&ms, ev[1], ev[2]));
                nll_us += 1e3 * (double)ms; ++nll_launches;
            }
            ++seen;
            prev_need_eq = n_need_eq > 0;
            if (n_failed > 0) {
                set_error("bounded Brent search failed for %d (distance, condition) bins "
                          "(NaN likelihood or evaluation budget exhausted)", n_failed);
                err = H3D_ERR_NUMERIC; break;
            }
            if (n_fit_failed > 0) {
                set_error("fit_mu_hat: %d pixels with all-zero counts inside a condition", n_fit_failed);
                err = H3D_ERR_NUMERIC; break;
            }
            if (n_need_eq == 0 && n_in_brent == 0) done = true;
            else if (seen > 200000) { set_error("qCML did not terminate"); err = H3D_ERR_NUMERIC; break; }
        }
        if (err != H3D_OK) { cudaStreamSynchronize(st); return err; }
    }
    collect_kernel<<<1, 256, 0, st>>>(prob, n_prob, disp_dev, stats_dev);
    H3D_LAUNCHED("collect_kernel");
    publish_words_kernel<<<1, 256, 0, st>>>((const unsigned long long*)disp_dev,
                                            (volatile unsigned long long*)((char*)pin_dev + off_disp), n_prob);
    H3D_LAUNCHED("publish_words_kernel");
    publish_words_kernel<<<1, 32, 0, st>>>((const unsigned long long*)stats_dev,
                                           (volatile unsigned long long*)((char*)pin_dev + off_stats), 4);
    H3D_LAUNCHED("publish_words_kernel");
    H3D_CHECK(cudaStreamSynchronize(st));       // also drains the queued no-op rounds
    memcpy(disp_per_dist_host, (char*)pin_host + off_disp, (size_t)n_prob * 8);
    long long h_stats[4] = {0, 0, 0, 0};
    memcpy(h_stats, (char*)pin_host + off_stats, 4 * sizeof(long long));
    if (stats_host) {
        stats_host[0] = h_stats[0]; stats_host[1] = h_stats[1]; stats_host[2] = h_stats[2];
        stats_host[3] = (long long)(h3d_launch_count() - launches_before);
        stats_host[4] = eq_launches; stats_host[5] = (long long)eq_us;
        stats_host[6] = nll_launches; stats_host[7] = (long long)nll_us;
        stats_host[8] = h_stats[3];
    }
    return H3D_OK;
}

// Pseudo-data of ONE bin at a given dispersion: hic3defdr/util/scaled_nb.py:
// 186-214 (equalize).  x, f: SoA (n_reps, ld), every replicate belongs to the
// one condition; pseudo_out: SoA (n_reps, ld).  The same kernel the qCML
// driver launches; exposed so that the device pseudo-data can be compared
// element by element with the reference's.
extern "C" size_t h3d_equalize_ws_bytes(long long n_px) {
    const long long max_chunks = n_px / kChunk + 2;
    return ws_pad((size_t)max_chunks * 4) + 2 * ws_pad((size_t)max_chunks * 8) + ws_pad(16) +
           ws_pad(sizeof(Problem)) + ws_pad(64);
}

extern "C" int h3d_equalize(const double* x, const double* f, long long ld, long long n_px, int n_reps,
                            double alpha, double* pseudo_out, int* n_fit_failed, void* ws,
                            size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_px >= 0 && n_px <= ld, "n_px must fit in ld");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_px == 0) return H3D_OK;
    std::vector<unsigned char> design(n_reps, 1);
    CondReps cr;
    int rows = 0, max_rc = 0;
    { int rc = make_cond_reps(design.data(), n_reps, 1, &cr, &rows, &max_rc); if (rc) return rc; }
    const long long seg[2] = {0, n_px};
    const int seg0 = 0;
    Workspace w(ws, ws_bytes);
    ChunkTables ct;
    { int rc = make_chunk_tables(&seg0, seg, seg + 1, 1, 1, w, st, &ct); if (rc) return rc; }
    Problem* prob = w.take<Problem>(1);
    Counters* cnt = w.take<Counters>(1);
    if (!prob || !cnt) { set_error("equalize workspace too small (%zu bytes given)", ws_bytes); return H3D_ERR_WORKSPACE; }
    single_bin_problems_kernel<<<1, 32, 0, st>>>(prob, 1, n_px, alpha, 0.5, ST_NEED_EQ, cnt);
    H3D_LAUNCHED("single_bin_problems_kernel");
    const size_t eq_smem = (size_t)kEqTasks * (sizeof(EqTask) + 2 * sizeof(int));
    { int rc = set_equalize_smem(eq_smem); if (rc) return rc; }
    const dim3 cgrid(ct.n_chunks, 1);
#define CALL(M) equalize_kernel<M><<<cgrid, 256, eq_smem, st>>>(x, f, ld, ct.chunk_seg, ct.chunk_lo, ct.chunk_hi, cr, \
        H3D_EST_QCML, prob, pseudo_out, cnt)
    DISPATCH_RC(CALL)
#undef CALL
    H3D_LAUNCHED("equalize_kernel");
    if (n_fit_failed)
        H3D_CHECK(cudaMemcpyAsync(n_fit_failed, &cnt->n_fit_failed, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return H3D_OK;
}

// Conditional NB negative log-likelihood of ONE bin of (pseudo-)data at
// delta = disp / (1 + disp): the objective of hic3defdr/util/dispersion.py:
// 72-75, evaluated by the kernel the qCML driver launches (exact 128-bit sum).
// data: SoA (n_reps, ld); nll_out: one double on the device.
extern "C" size_t h3d_cml_nll_ws_bytes(long long n_px) {
    return h3d_equalize_ws_bytes(n_px) + ws_pad(sizeof(Fix128));
}

extern "C" int h3d_cml_nll(const double* data, long long ld, long long n_px, int n_reps, double delta,
                           double* nll_out, void* ws, size_t ws_bytes, h3d_stream_t stream) {
    H3D_REQUIRE(n_reps >= 1 && n_reps <= H3D_MAX_REPS, "n_reps out of range");
    H3D_REQUIRE(n_px >= 1 && n_px <= ld, "n_px must be in [1, ld]");
    H3D_REQUIRE(delta > 0.0 && delta < 1.0, "delta must be in (0, 1)");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<unsigned char> design(n_reps, 1);
    CondReps cr;
    int rows = 0, max_rc = 0;
    { int rc = make_cond_reps(design.data(), n_reps, 1, &cr, &rows, &max_rc); if (rc) return rc; }
    const long long seg[2] = {0, n_px};
    const int seg0 = 0;
    Workspace w(ws, ws_bytes);
    ChunkTables ct;
    { int rc = make_chunk_tables(&seg0, seg, seg + 1, 1, 1, w, st, &ct); if (rc) return rc; }
    Problem* prob = w.take<Problem>(1);
    Counters* cnt = w.take<Counters>(1);
    Fix128* acc = w.take<Fix128>(1);
    if (!prob || !cnt || !acc) { set_error("cml_nll workspace too small (%zu bytes given)", ws_bytes); return H3D_ERR_WORKSPACE; }
    H3D_CHECK(cudaMemsetAsync(acc, 0, sizeof(Fix128), st));
    single_bin_problems_kernel<<<1, 32, 0, st>>>(prob, 1, n_px, 0.0, delta, ST_IN_BRENT, cnt);
    H3D_LAUNCHED("single_bin_problems_kernel");
    const dim3 cgrid(ct.n_chunks, 1);
#define CALL(M) nll_kernel<M><<<cgrid, 256, 0, st>>>(data, ld, ct.chunk_seg, ct.chunk_lo, ct.chunk_hi, cr, prob, acc)
    DISPATCH_RC(CALL)
#undef CALL
    H3D_LAUNCHED("nll_kernel");
    fix_value_kernel<<<1, 1, 0, st>>>(acc, nll_out);
    H3D_LAUNCHED("fix_value_kernel");
    return H3D_OK;
}
