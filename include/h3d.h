/*
 * libh3d -- C ABI of the B200-native implementation of the per-pixel
 * statistical pipeline behind HiC3DeFDR.run_to_qvalues()
 * (thomasgilgenast/hic3defdr v0.2.1).
 *
 * The reference is pure Python and has no FFI layer; the drop-in boundary is
 * its Python class (hic3defdr/analysis/constructor.py:12-86) and its .npy
 * files.  Each entry point below replaces the arithmetic of the reference
 * function cited next to it; hic3defdr_b200/analysis.py (the host-side mirror
 * of the reference class) is the only caller.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every array pointer is a DEVICE pointer
 *     unless its name ends in _host; the caller owns every buffer;
 *   - every call enqueues its kernels on ``stream`` (a cudaStream_t) and
 *     returns without synchronising unless stated otherwise;
 *   - return value 0 on success, negative H3D_ERR_* otherwise;
 *     h3d_last_error() returns a thread-local message;
 *   - scratch memory: ``ws``/``ws_bytes`` is a caller-owned device buffer of at
 *     least h3d_*_ws_bytes(...) bytes;
 *   - matrices are C-order (row-major), pixels x replicates unless stated.
 */
#ifndef H3D_H
#define H3D_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* h3d_stream_t; /* cudaStream_t */

#define H3D_OK 0
#define H3D_ERR_ARG (-1)
#define H3D_ERR_CUDA (-2)
#define H3D_ERR_WORKSPACE (-3)
#define H3D_ERR_NUMERIC (-4)

#define H3D_MAX_REPS 16
#define H3D_MAX_CONDS 8

/* element types of CSR ``data`` arrays */
#define H3D_DTYPE_I64 0
#define H3D_DTYPE_F64 1
#define H3D_DTYPE_I32 2
#define H3D_DTYPE_F32 3

/* size-factor modes: hic3defdr/util/scaling.py:27-149 */
#define H3D_NORM_CONDITIONAL_MOR 0
#define H3D_NORM_CONDITIONAL_SCALING 1
#define H3D_NORM_MEDIAN_OF_RATIOS 2
#define H3D_NORM_SIMPLE_SCALING 3

/* ---- library ---------------------------------------------------------- */
int h3d_version(void);
const char* h3d_last_error(void);
/* number of kernels this library has launched since load / last reset */
unsigned long long h3d_launch_count(void);
void h3d_reset_launch_count(void);
/* measured FP64 FMA rate of the current device in TFLOP/s (FMA = 2 flops):
 * the roofline denominator of the FP64-bound kernels.  scratch: >= 8 bytes on
 * the device; tflops_out: HOST.  Synchronises. */
int h3d_fp64_peak(double* scratch, double* tflops_out_host, h3d_stream_t stream);

/* ---- prepare_data ------------------------------------------------------ */

/* CoreHiC3DeFDR.load_bias filter, hic3defdr/analysis/core.py:58-59.
 * bias: (n_bins, n_reps), filtered in place. */
int h3d_bias_filter(double* bias, int n_bins, int n_reps, double bias_thresh,
                    h3d_stream_t stream);

/* sparse_union + deconvolute + wipe_distances, hic3defdr/util/matrices.py:92-129,
 * 8-38, 41-62 (mean_thresh = 0), pass 1: number of union pixels per row and
 * their exclusive prefix.  indptr_host/indices_host/data_host are HOST arrays
 * of n_reps DEVICE pointers (one CSR matrix per replicate, canonical format).
 * row_offset: (n_bins + 1) int32; row_offset[n_bins] = number of union pixels. */
int h3d_union_count(int n_reps, const void* const* indptr_host, int indptr_is64,
                    const int* const* indices_host, const void* const* data_host,
                    int data_dtype, const double* bias, int n_bins, int dist_max,
                    int* row_offset, void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_union_ws_bytes(int n_bins);

/* pass 2: emits row, col (int32, (row, col) order), dist = col - row (may be
 * NULL), raw (int64, truncating; hic3defdr/analysis/analysis.py:92-95) and
 * balanced (float64; :98-101). */
int h3d_union_emit(int n_reps, const void* const* indptr_host, int indptr_is64,
                   const int* const* indices_host, const void* const* data_host,
                   int data_dtype, const double* bias, int n_bins, int dist_max,
                   const int* row_offset, int* row, int* col, int* dist,
                   long long* raw, double* balanced, h3d_stream_t stream);

/* Size factors, hic3defdr/util/scaling.py:27-149 with equal_bin
 * (util/binning.py:4-25, canonical stable tie-break).  dist: (n_px) int32,
 * balanced: (n_px, n_reps).  Output sf_table:
 *   conditional modes: (dist_max + 1, n_reps), the size factor of every
 *                      distance (n_bins > 0: interpolated between bins;
 *                      n_bins == 0: exact-distance groups, NaN where empty);
 *   other modes:       (n_reps).
 * Synchronises the stream once (group bookkeeping is read back). */
int h3d_size_factors(const int* dist, const double* balanced, long long n_px,
                     int n_reps, int dist_max, int n_bins, int norm,
                     double* sf_table, void* ws, size_t ws_bytes,
                     h3d_stream_t stream);
size_t h3d_size_factors_ws_bytes(long long n_px, int n_reps, int dist_max);

/* The stages of h3d_size_factors as separate entry points, for a chromosome
 * whose union pixels are sharded over several GPUs by row range (SURVEY.md
 * section 8(e)); hic3defdr_b200/dist.py places the collectives between them:
 *   every rank ranks its pixels by distance (h3d_stable_rank), the per-
 *   distance counts are all-gathered, which fixes every pixel's position in
 *   the chromosome-wide (distance, row, col) order and so its equal-count bin
 *   (util/binning.py:4-25); bins are owned by ranks; one all-to-all moves the
 *   per-pixel ratios (util/scaling.py:41-47) to the owner of their bin, which
 *   takes the exact per-replicate medians; the (bins, n_reps) medians are
 *   all-gathered and every rank builds the (dist_max + 1, n_reps) table.
 * n_groups = h3d_sf_num_groups(...): n_bins (conditional norms, n_bins > 0),
 * dist_max + 1 (conditional, n_bins == 0) or 1.
 * h3d_sf_group_bounds: gstart[g] = first position (in distance order, over
 *   n_total pixels) of group g; key_start = (dist_max + 2) boundaries of the
 *   distances over the same n_total pixels (conditional norms only).
 * h3d_sf_values: values[r * n_px + rank[p]] = ratio (or balanced value, for
 *   the scaling norms) of pixel p, replicate r; pixels without a ratio (a
 *   replicate <= 0) get the all-ones bit pattern, which the reducer skips.
 * h3d_sf_group_reduce: red[g, r] = median (sum for the scaling norms) of
 *   values[r * ld + gstart[g] .. gstart[g + 1]); valid[g] = ratios in g.
 * h3d_sf_table: the table of h3d_size_factors from red, gstart, key_start. */
int h3d_sf_num_groups(int dist_max, int n_bins, int norm);
int h3d_sf_group_bounds(long long n_total, int dist_max, int n_bins, int norm,
                        const long long* key_start, long long* gstart,
                        h3d_stream_t stream);
int h3d_sf_values(const double* balanced, const int* rank, long long n_px,
                  int n_reps, int norm, double* values, h3d_stream_t stream);
int h3d_sf_group_reduce(const double* values, long long ld, const long long* gstart,
                        int n_groups, int n_reps, int norm, double* red,
                        long long* valid, h3d_stream_t stream);
int h3d_sf_table(const double* red, const long long* gstart,
                 const long long* key_start, int n_groups, int n_reps, int dist_max,
                 int n_bins, int norm, double* sf_table, void* ws, size_t ws_bytes,
                 h3d_stream_t stream);
size_t h3d_sf_table_ws_bytes(int n_groups, int n_reps);

/* scaled = balanced / size_factors, per-condition means and disp_idx,
 * hic3defdr/analysis/analysis.py:109-115.  ``data`` holds balanced on entry
 * and scaled on exit.  size_factors_out ((n_px, n_reps)) may be NULL.
 * design: (n_reps, n_conds) bytes. */
int h3d_scale_filter(const int* row, const int* col, double* data,
                     const double* sf_table, int sf_per_dist,
                     const unsigned char* design_host, long long n_px, int n_reps,
                     int n_conds, int dist_max, double mean_thresh, int dist_min,
                     double* size_factors_out, unsigned char* disp_idx,
                     h3d_stream_t stream);

/* indices of the set entries of a boolean mask, in order (the device form of
 * ``array[disp_idx]``).  n_set_out: one int64 on the device. */
int h3d_mask_to_index(const unsigned char* mask, long long n, int* index_out,
                      long long* n_set_out, void* ws, size_t ws_bytes,
                      h3d_stream_t stream);
size_t h3d_mask_to_index_ws_bytes(long long n);

/* loop_idx, hic3defdr/analysis/analysis.py:117-125: membership of
 * (row, col)[index] in a sorted set of row * 2^32 + col keys. */
int h3d_loop_membership(const int* row, const int* col, const int* index,
                        long long n, const long long* sorted_keys, long long n_keys,
                        unsigned char* out, h3d_stream_t stream);

/* ---- estimate_disp ------------------------------------------------------ */

/* Combined factor f = bias[row] * bias[col] * size_factors
 * (hic3defdr/analysis/analysis.py:181-183, 272-275) and the raw counts of the
 * selected pixels, written replicate-major (SoA: x_out[r * ld + pos]) at
 * positions dest[i] (or i when dest is NULL); dist_out[pos] = col - row.
 * x_out = f_out = NULL computes the distances only. */
int h3d_gather_counts_factors(const int* row, const int* col, const int* index,
                              long long n_sel, const long long* raw,
                              const double* size_factors, int sf_per_pixel,
                              const double* bias, int n_reps, const int* dest,
                              long long ld, double* x_out, double* f_out,
                              int* dist_out, h3d_stream_t stream);

/* ---- multi-GPU pooling over peer memory ---------------------------------- */

/* Receive buffers of the distance exchange (one process per GPU on one node):
 * a cudaMalloc allocation plus the CUDA IPC handle the other ranks open.  The
 * handle is H3D_PEER_HANDLE_BYTES opaque bytes, to be sent to the peers by any
 * means (the host side uses torch.distributed). */
#define H3D_PEER_HANDLE_BYTES 64
int h3d_peer_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out);
int h3d_peer_free(void* ptr);
int h3d_peer_open(const unsigned char* handle, void** ptr_out);
int h3d_peer_close(void* ptr);

/* n_slices (<= 16) contiguous slices of local device memory into (peer)
 * buffers: slice k = bytes_host[k] bytes from src_base + src_off_host[k] to
 * dst_base_host[k] + dst_off_host[k] (all multiples of 8).  The all-to-all of
 * the distributed BH correction (analysis/analysis.py:296-303 is global):
 * every rank stores its partition of the p-values directly into the owners'
 * receive buffers, and the q-values come back the same way. */
int h3d_peer_copy(const void* src_base, const long long* src_off_host,
                  void* const* dst_base_host, const long long* dst_off_host,
                  const long long* bytes_host, int n_slices, h3d_stream_t stream);

/* Pooling by distance (hic3defdr/analysis/analysis.py:169-206: raw[dist == d],
 * f[dist == d] genome-wide), two passes.
 * Pass 1, per chromosome: tested pixel i (union index u = index[i]) drops the
 * record {u, chrom_id << 24 | row[u]} (two int32) at rec[dest[i]], dest = its
 * pooled position (stable rank by pooling key over all chromosomes).
 * Pass 2, all chromosomes: pooled position p, of key k (key_start: device,
 * n_keys + 1 boundaries; dist_of_key[k] = the distance), is gathered from the
 * chromosome table (device, n_chroms rows of four 8-byte words: raw pointer,
 * bias pointer, size-factor pointer, size-factor form 0: (R,), 1: (N, R),
 * 2: (D + 1, R)) and written to column p + shift_of_key[k] of the (2 n_reps, ld)
 * float64 matrix peer_base_host[owner_of_key[k]] (rows [0, R): counts, rows
 * [R, 2R): factors f = bias[row] bias[col] size_factor).  One process:
 * owner 0, shift 0, the matrix is the local pooled buffer.  Several processes:
 * peer_base_host holds this rank's receive buffer and the opened peer buffers
 * (HOST array of n_ranks device pointers), and the stores go over NVLink; the
 * caller orders the ranks around the call (nobody reads a buffer before all
 * writers are done, nobody writes before the previous contents are used up). */
int h3d_pool_index(const int* row, const int* index, long long n_sel, int chrom_id,
                   const int* dest, int* rec, h3d_stream_t stream);
int h3d_pool_pull(const int* rec, long long n_local, const long long* key_start,
                  int n_keys, const int* dist_of_key, const int* owner_of_key,
                  const long long* shift_of_key, const void* chrom_table,
                  int n_chroms, int n_reps, void* const* peer_base_host,
                  int n_ranks, long long ld, h3d_stream_t stream);

/* Stable rank of every element when sorting by an integer key in
 * [0, n_keys): rank_out[i] = position of element i; key_start: (n_keys + 1)
 * int64 group boundaries.  Used for equal_bin (util/binning.py:25) and for
 * pooling pixels by distance (analysis/analysis.py:196-197). */
int h3d_stable_rank(const int* keys, long long n, int n_keys, int* rank_out,
                    long long* key_start, void* ws, size_t ws_bytes,
                    h3d_stream_t stream);
size_t h3d_stable_rank_ws_bytes(long long n, int n_keys);

/* qCML / CML / MME dispersion per (distance, condition) segment,
 * hic3defdr/util/dispersion.py:10-131 with equalize/q2qnbinom
 * (util/scaled_nb.py:186-275).  x, f: SoA (n_reps, ld) pooled by distance
 * (segment d = [seg_start_host[d], seg_start_host[d + 1])).
 * disp_per_dist_host: (n_seg, n_conds) HOST output, NaN for empty segments.
 * stats_host (may be NULL): 9 int64 = {outer iterations, NLL evaluations,
 * pixel-equalisations, kernel launches, equalize launches, equalize time (us,
 * CUDA events on ``stream``), NLL launches, NLL time (us), segments stopped at
 * the outer-iteration cap}.  Synchronises.
 * The qCML fixed point ``while |disp - new| > 1e-4`` has no iteration cap in
 * the reference (util/dispersion.py:36-42: ``it`` is never incremented) and
 * cycles forever where the dispersion is large and the bin small (the bounded
 * Brent search resolves delta to 1e-5, i.e. disp = delta / (1 - delta) only
 * to ~0.05 at disp ~ 66).  Here a segment stops after H3D_QCML_MAX_OUTER
 * outer iterations with its last iterate, and is counted in stats[8].
 * The likelihood of a bin is summed in 128-bit fixed point, i.e. exactly: the
 * result does not depend on the order of the pixels inside a segment (nor,
 * therefore, on how many GPUs pooled them).  The host queues
 * H3D_QCML_AHEAD (environment, default 6) rounds ahead of the device. */
#define H3D_QCML_MAX_OUTER 100
#define H3D_EST_QCML 0
#define H3D_EST_CML 1
#define H3D_EST_MME 2
int h3d_estimate_dispersion(const double* x, const double* f, long long ld,
                            const long long* seg_start_host, int n_seg,
                            const unsigned char* design_host, int n_reps,
                            int n_conds, int estimator,
                            double* disp_per_dist_host, long long* stats_host,
                            void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_estimate_dispersion_ws_bytes(long long n_px, int n_seg, int n_reps,
                                        int n_conds);

/* The same with every segment given as a LIST OF RUNS of the pooled arrays:
 * run r = pixels [run_lo_host[r], run_hi_host[r]) of segment run_seg_host[r].
 * After the multi-GPU exchange of hic3defdr/analysis/analysis.py:169-206's
 * pooling, a distance's pixels arrive as one run per source rank; the
 * likelihood sums are exact (order independent), so the runs are consumed where
 * they land instead of being regrouped into contiguous segments. */
int h3d_estimate_dispersion_runs(const double* x, const double* f, long long ld,
                                 const int* run_seg_host,
                                 const long long* run_lo_host,
                                 const long long* run_hi_host, int n_runs,
                                 int n_seg, const unsigned char* design_host,
                                 int n_reps, int n_conds, int estimator,
                                 double* disp_per_dist_host, long long* stats_host,
                                 void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_estimate_dispersion_runs_ws_bytes(long long n_px, int n_runs, int n_seg,
                                             int n_reps, int n_conds);

/* equalize, hic3defdr/util/scaled_nb.py:186-214 (with q2qnbinom, :217-275):
 * pseudo-data of ONE bin at dispersion ``alpha``.  x, f: SoA (n_reps, ld),
 * all replicates of one condition; pseudo_out: SoA (n_reps, ld).  The very
 * kernel the qCML driver launches, exposed so that the device pseudo-data can
 * be checked element by element.  n_fit_failed (device int32, may be NULL):
 * number of pixels with all-zero counts (the reference raises there). */
int h3d_equalize(const double* x, const double* f, long long ld, long long n_px,
                 int n_reps, double alpha, double* pseudo_out, int* n_fit_failed,
                 void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_equalize_ws_bytes(long long n_px);

/* The objective of cml, hic3defdr/util/dispersion.py:72-75: conditional NB
 * negative log-likelihood of one bin of (pseudo-)data, SoA (n_reps, ld), at
 * delta = disp / (1 + disp), by the kernel the qCML driver launches (summed
 * exactly in 128-bit fixed point).  nll_out: one double on the device. */
int h3d_cml_nll(const double* data, long long ld, long long n_px, int n_reps,
                double delta, double* nll_out, void* ws, size_t ws_bytes,
                h3d_stream_t stream);
size_t h3d_cml_nll_ws_bytes(long long n_px);

/* lowess (lib5c.util.lowess.lowess as called at hic3defdr/util/lowess.py:72):
 * x sorted ascending, n points, returns fitted values y_fit (device). */
int h3d_lowess(const double* x, const double* y, int n, double frac, int it,
               double delta, double* y_fit, void* ws, size_t ws_bytes,
               h3d_stream_t stream);
size_t h3d_lowess_ws_bytes(int n);
/* n_jobs (<= 64) independent smoothing problems in one launch, one thread
 * block each (the trends of all conditions, hic3defdr/analysis/analysis.py:
 * 208-218): problem j is the slice [off_host[j], off_host[j] + n_host[j]) of
 * the packed device arrays x / y / y_fit; off, n, frac, delta: HOST arrays. */
int h3d_lowess_batch(const double* x, const double* y, const int* off_host,
                     const int* n_host, const double* frac_host, int it,
                     const double* delta_host, int n_jobs, double* y_fit,
                     void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_lowess_batch_ws_bytes(int n_total, int n_jobs);

/* disp[:, c] = disp_fn_c(dist) for integer distances: gather from a
 * (dist_max + 1, n_conds) table (hic3defdr/analysis/analysis.py:218). */
int h3d_gather_table(const int* dist, long long n, const double* table,
                     int n_cols, int n_rows, double* out, h3d_stream_t stream);

/* ---- lrt ---------------------------------------------------------------- */

/* fit_mu_hat, hic3defdr/util/scaled_nb.py:71-183.  x, b: (n, n_reps);
 * alpha: (n, n_reps) when alpha_stride_px = n_reps, broadcast forms with
 * alpha_stride_px / alpha_stride_rep in {0, 1, n_reps}.  status: number of
 * pixels without a positive root (one int32 on the device, accumulated). */
int h3d_fit_mu_hat(const double* x, const double* b, const double* alpha,
                   long long alpha_stride_px, long long alpha_stride_rep,
                   long long n, int n_reps, double* mu_out, int* n_failed,
                   h3d_stream_t stream);

/* lrt, hic3defdr/util/lrt.py:7-50: raw (n, n_reps) float64 counts, f
 * (n, n_reps), disp (n, n_conds) per-condition dispersions (the reference's
 * ``disp @ design.T`` widening happens inside). */
int h3d_lrt(const double* raw, const double* f, const double* disp,
            const unsigned char* design_host, long long n, int n_reps, int n_conds,
            int refit_mu, double* pvalues, double* llr, double* mu_hat_null,
            double* mu_hat_alt, int* n_failed, h3d_stream_t stream);

/* The same test fused with its input gathers for the pipeline
 * (hic3defdr/analysis/analysis.py:261-278): reads the union-aligned arrays
 * through ``index`` (the disp_idx positions). */
int h3d_lrt_fused(const int* row, const int* col, const int* index, long long n_sel,
                  const long long* raw, const double* size_factors, int sf_per_pixel,
                  const double* bias, const double* disp,
                  const unsigned char* design_host, int n_reps, int n_conds,
                  int refit_mu, double* pvalues, double* llr, double* mu_hat_null,
                  double* mu_hat_alt, int* n_failed, h3d_stream_t stream);

/* ---- control read-back --------------------------------------------------- */

/* Copies nbytes (multiple of 8, <= 64 KiB) from device memory to pinned,
 * device-mapped host memory with a kernel rather than the copy engine, so that
 * small control read-backs (pixel counts that size the next allocation,
 * convergence counters) never queue behind bulk output copies on the DMA
 * engine.  Valid on the host once the stream is synchronised. */
int h3d_publish(const void* dev_src, void* host_mapped_dst, size_t nbytes,
                h3d_stream_t stream);

/* ---- bh ----------------------------------------------------------------- */

/* Benjamini-Hochberg q-values over the finite entries of p
 * (lib5c adjust_pvalues, call site hic3defdr/analysis/analysis.py:300). */
int h3d_bh(const double* p, long long n, double* q, void* ws, size_t ws_bytes,
           h3d_stream_t stream);
size_t h3d_bh_ws_bytes(long long n);

/* One bucket of a multi-GPU correction (the distributed sort/rank of p-values
 * that replaces the single argsort of statsmodels' fdr_bh): ``p`` holds every
 * p-value of the genome that falls between two splitters, ``rank_offset`` of
 * the genome's finite p-values are smaller, ``n_total`` are finite in total
 * (0: the local finite count).  Writes q before the contribution of the
 * buckets above (clipped at 1) and, to ``min_out`` (device, may be NULL), the
 * smallest p / (rank / n_total) of the bucket; the caller exchanges the minima
 * and finishes with h3d_bh_apply_carry(q, n, min over the higher buckets).
 * Workspace as for h3d_bh. */
int h3d_bh_ranked(const double* p, long long n, long long rank_offset, long long n_total,
                  double* q, double* min_out, void* ws, size_t ws_bytes,
                  h3d_stream_t stream);
int h3d_bh_apply_carry(double* q, long long n, double carry, h3d_stream_t stream);
/* the same with the carry read from device memory (no host round trip) */
int h3d_bh_apply_carry_dev(double* q, long long n, const double* carry_dev,
                           h3d_stream_t stream);

/* ---- threshold / classify (the step after bh) ------------------------------ */

/* find_clusters with connectivity 1, hic3defdr/util/clusters.py:69-96 (called
 * from util/thresholding.py:7-44 and util/classification.py:7-49): the
 * 4-connected components of a set of n pixels given sorted by (row, col),
 * unique.  label[i] = position of the first pixel (in that order) of the
 * component of pixel i; size[i] = number of pixels of the component whose first
 * pixel is i, 0 for every other i.  Synchronises (input order is checked). */
int h3d_connected_components(const int* row, const int* col, long long n, int* label,
                             int* size, void* ws, size_t ws_bytes, h3d_stream_t stream);
size_t h3d_connected_components_ws_bytes(long long n);

/* ---- evaluate (ROC / FDR curves) ----------------------------------------- */

/* sklearn.metrics.roc_curve as used by hic3defdr/util/evaluation.py:44-79
 * (evaluate; called from analysis/simulation.py:146-239), in two calls.
 * h3d_roc_sort: scores 1 - q sorted descending; keys_sorted (n uint64,
 * order-preserving image of the scores), boundary[i] = 1 at the last element
 * of every run of equal scores, y_sorted = the labels in that order.
 * h3d_roc_points: for the thresholds at sorted positions thr_idx (the set
 * entries of boundary, m of them) and the sorted positions pos_idx of the
 * n_pos positives: tps, fps (int64), thresholds (the scores), and keep[t] = 1
 * for the points sklearn's drop_intermediate retains. */
int h3d_roc_sort(const double* qvalues, const unsigned char* y_true, long long n,
                 unsigned long long* keys_sorted, unsigned char* boundary,
                 unsigned char* y_sorted, void* ws, size_t ws_bytes,
                 h3d_stream_t stream);
size_t h3d_roc_sort_ws_bytes(long long n);
int h3d_roc_points(const unsigned long long* keys_sorted, const int* thr_idx,
                   long long m, const int* pos_idx, long long n_pos, long long* tps,
                   long long* fps, double* thresholds, unsigned char* keep,
                   h3d_stream_t stream);

/* ---- simulate / balance -------------------------------------------------- */

/* One simulated replicate, hic3defdr/util/simulation.py:177-202: for pixel i,
 * f = bias[row, rep] bias[col, rep] size_factor (size_factors: (n_sim,), or
 * (n_dist, n_sim) by distance when sf_by_distance), bm = mean[i] f, count ~
 * NB(mean bm, variance bm + disp bm^2) with disp = disp[col - row] (a
 * (n_dist,) table; trend 'dist') or disp[i] (disp_per_pixel; trend 'mean').
 * bias: (n_bins, n_sim) row-major.  counts_out: int64 (n,); biased_mean_out:
 * optional (n,).  The stream of pixel i, replicate rep is Philox4x32-10 keyed
 * by ``seed``: results do not depend on the launch geometry. */
int h3d_nb_simulate(const int* row, const int* col, const double* mean, long long n,
                    const double* bias, int n_sim, const double* size_factors,
                    int sf_by_distance, int n_dist, const double* disp,
                    int disp_per_pixel, int rep, unsigned long long seed,
                    long long* counts_out, double* biased_mean_out,
                    h3d_stream_t stream);

/* perturb_cluster, hic3defdr/util/simulation.py:12-67, once the footprint of
 * every cluster has been turned into (pixel key = row << 32 | col, factor)
 * pairs: mean[pixel] *= factor for the pixels present in pixel_keys (sorted). */
int h3d_perturb(const long long* pixel_keys, long long n_px, const long long* keys,
                const double* factor, long long n, double* mean, h3d_stream_t stream);

/* The iteration of kr_balance, hic3defdr/util/balancing.py:86-174, on a
 * symmetric CSR matrix without empty rows (int32 indptr / indices, float64
 * data, all on the device): returns the balancing vector x (device, n) with
 * x_i x_j A_ij summing to 1 over every row, the residual after every outer
 * iteration (host, at most res_cap values) and the number of SpMVs. */
int h3d_kr_balance(const int* indptr, const int* indices, const double* data, int n,
                   double tol, const double* x0, double delta, double ddelta,
                   int max_iter, double* x_out, double* res_host, int res_cap,
                   int* n_res_host, int* n_matvec_host, void* ws, size_t ws_bytes,
                   h3d_stream_t stream);
size_t h3d_kr_balance_ws_bytes(int n);

/* filter_sparse_rows_count, hic3defdr/util/filtering.py:50-53: per bin of an
 * upper-triangular CSR matrix (int64 indptr), the number of entries > 0 among
 * its k nearest upstream (column i, rows i-k .. i-1) and downstream (row i,
 * columns i+1 .. i+k) contacts. */
int h3d_band_nnz(const long long* indptr, const int* indices, const double* data, int n,
                 int k, int* upstream, int* downstream, h3d_stream_t stream);

/* ---- compact transfer of redundant outputs ------------------------------- */

/* size_factors (N, R) and disp (N_d, C) are functions of the pixel distance
 * (hic3defdr/util/scaling.py:92-104, analysis/analysis.py:218): the end-to-end
 * path copies the per-distance table to the host and rebuilds the array there.
 * HOST pointers: out[k, :] = table[col[i] - row[i], :] over the pixels with
 * mask[i] != 0 (all pixels when mask is NULL), with n_threads host threads. */
int h3d_host_expand_by_distance(const double* table, int n_rows, int n_cols,
                                const int* row, const int* col,
                                const unsigned char* mask, long long n,
                                double* out, int n_threads);
/* raw (N, R) is int64 in the reference's files (analysis/analysis.py:92-96)
 * but holds counts: narrowed on the device (overflow: device int, set when a
 * value does not fit 32 bits), widened again into the host buffer. */
int h3d_narrow_i64(const long long* in, long long n, int* out, int* overflow,
                   h3d_stream_t stream);
int h3d_host_widen_i32(const int* in, long long* out, long long n, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* H3D_H */
