"""
TEST INFRASTRUCTURE -- CPU restatement (numpy/scipy) of the reference's
``run_to_qvalues`` arithmetic, stage by stage, operating on in-memory arrays.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  The restatement is validated against the unmodified
reference (oracle/refrun.py) by tests/golden/make_golden.py, which also
records the fixtures under tests/golden/.

Canonical choices (documented in DESIGN.md):
  * equal-count binning uses a STABLE rank (ties broken by (row, col) order);
    the shipped reference uses an unstable argsort (util/binning.py:25).
  * estimator='mme' follows the intended semantics raw.astype(float)/f; the
    shipped reference raises on integer input (util/dispersion.py:101-102).
"""
import numpy as np
import scipy.sparse as sparse
import scipy.stats as stats
from scipy.optimize import brentq, minimize_scalar, newton
from scipy.special import gammaln

from oracle.thirdparty import adjust_pvalues, gmean, lowess


# --------------------------------------------------------------------------
# prepare_data
# --------------------------------------------------------------------------

def filter_bias(bias, bias_thresh):
    """analysis/core.py:56-60: rows where any replicate is below the threshold
    or above its reciprocal are zeroed for ALL replicates."""
    bias = np.array(bias, dtype=float, copy=True)
    bad = np.any(bias < bias_thresh, axis=1) | \
        np.any(bias > 1. / bias_thresh, axis=1)
    bias[bad, :] = 0
    return bias


def union_pixels(mats, dist_max, bias=None):
    """util/matrices.py:92-129 (sparse_union with mean_thresh=0) incl.
    deconvolute(invert=True) :29-38 and wipe_distances :58-62.

    A pixel survives iff 0 <= col-row <= dist_max, some replicate stores a
    value there whose bias-normalised value is non-zero, and the sum over
    replicates of the normalised values is finite and >= 0.  Returned in
    (row, col) order as int32.
    """
    total = None
    for r, m in enumerate(mats):
        coo = sparse.coo_matrix(m)
        v = coo.data.astype(float)
        if bias is not None:
            b = bias[:, r]
            inv = np.where(b == 0, 0.0, 1.0 / np.where(b == 0, 1.0, b))
            v = inv[coo.row] * v * inv[coo.col]
        d = coo.col - coo.row
        keep = (d >= 0) & (d <= dist_max) & (v != 0)
        part = sparse.csr_matrix((v[keep], (coo.row[keep], coo.col[keep])),
                                 shape=coo.shape)
        total = part if total is None else total + part
    total = total.tocoo()
    ok = (total.data >= 0) & np.isfinite(total.data)
    row, col = total.row[ok], total.col[ok]
    order = np.lexsort((col, row))
    return row[order].astype(np.int32), col[order].astype(np.int32)


def gather_raw_balanced(mats, row, col, bias):
    """analysis/analysis.py:92-101: raw is int64 (truncating), balanced uses
    the stored value divided by the two bias factors."""
    n_px, n_rep = len(row), len(mats)
    raw = np.zeros((n_px, n_rep), dtype=np.int64)
    balanced = np.zeros((n_px, n_rep), dtype=float)
    for r, m in enumerate(mats):
        vals = np.asarray(sparse.csr_matrix(m)[row, col]).ravel()
        raw[:, r] = vals
        with np.errstate(divide='ignore', invalid='ignore'):
            balanced[:, r] = vals / (bias[row, r] * bias[col, r])
    return raw, balanced


def stable_rank(keys):
    """rank of each element in a stable ascending sort of ``keys``."""
    order = np.argsort(keys, kind='stable')
    rank = np.empty(len(keys), dtype=np.int64)
    rank[order] = np.arange(len(keys))
    return rank


def equal_bin(dist, n_bins):
    """util/binning.py:24-25 with the canonical stable tie-break."""
    idx = np.linspace(0, n_bins, dist.size, endpoint=0, dtype=int)
    return idx[stable_rank(dist)]


def median_of_ratios(data):
    """util/scaling.py:44-47 (filter_zeros=True; gmean pseudocount 1)."""
    ok = np.all(data > 0, axis=1)
    sub = data[ok, :]
    return np.median(sub / gmean(sub, axis=1)[:, None], axis=0)


def simple_scaling(data):
    """util/scaling.py:64-65."""
    s = np.sum(data, axis=0)
    return s / gmean(s)


def interp_extrap(xp, yp, x):
    """scipy.interpolate.interp1d(kind='linear', fill_value='extrapolate',
    assume_sorted=True) as evaluated by the container's scipy 1.18
    (interpolate/_interpolate.py:491-517), used at util/scaling.py:96-100."""
    xp = np.asarray(xp, dtype=float)
    x = np.asarray(x, dtype=float)
    hi = np.clip(np.searchsorted(xp, x), 1, len(xp) - 1).astype(int)
    lo = hi - 1
    return ((x - xp[lo]) / (xp[hi] - xp[lo])) * yp[hi] + \
        ((xp[hi] - x) / (xp[hi] - xp[lo])) * yp[lo]


def conditional_size_factors(data, dist, n_bins, reducer=median_of_ratios):
    """util/scaling.py:87-105 (conditional / conditional_mor /
    conditional_scaling).  Returns (size_factors (N, R), table dict)."""
    out = np.zeros_like(data, dtype=float)
    if n_bins:
        bins = equal_bin(dist, n_bins)
        d_b, s_b = [], []
        for b in np.unique(bins):
            sel = bins == b
            d_b.append(np.mean(dist[sel]))
            s_b.append(reducer(data[sel, :]))
        d_b, s_b = np.array(d_b), np.array(s_b)
        for r in range(data.shape[1]):
            out[:, r] = interp_extrap(d_b, s_b[:, r], dist)
    else:
        for d in np.unique(dist):
            sel = dist == d
            out[sel, :] = reducer(data[sel, :])
    return out


def scale_and_filter(balanced, size_factors, dist, design, mean_thresh,
                     dist_min):
    """analysis/analysis.py:109-115."""
    scaled = balanced / size_factors
    mean = np.dot(scaled, design.astype(float)) / np.sum(design, axis=0)
    disp_idx = np.all(mean >= mean_thresh, axis=1) & (dist >= dist_min)
    return scaled, disp_idx


def prepare_chrom(mats, bias_raw, design, dist_min=4, dist_max=200,
                  bias_thresh=0.1, mean_thresh=1.0, norm='conditional_mor',
                  n_bins=-1):
    """analysis/analysis.py:63-133 for one chromosome, in memory."""
    if n_bins == -1:
        n_bins = int(dist_max / 5)
    bias = filter_bias(bias_raw, bias_thresh)
    row, col = union_pixels(mats, dist_max, bias=bias)
    raw, balanced = gather_raw_balanced(mats, row, col, bias)
    dist = col - row
    if norm == 'conditional_mor':
        sf = conditional_size_factors(balanced, dist, n_bins)
    elif norm == 'conditional_scaling':
        sf = conditional_size_factors(balanced, dist, n_bins, simple_scaling)
    elif norm == 'median_of_ratios':
        sf = median_of_ratios(balanced)
    elif norm == 'simple_scaling':
        sf = simple_scaling(balanced)
    else:
        raise KeyError(norm)
    scaled, disp_idx = scale_and_filter(balanced, sf, dist, design,
                                        mean_thresh, dist_min)
    return dict(bias=bias, row=row, col=col, raw=raw, balanced=balanced,
                size_factors=sf, scaled=scaled, disp_idx=disp_idx)


def loop_membership(row, col, clusters):
    """analysis/analysis.py:117-125."""
    px = set()
    for c in clusters:
        px |= set(tuple(e) for e in c)
    return np.array([(int(i), int(j)) in px for i, j in zip(row, col)],
                    dtype=bool)


# --------------------------------------------------------------------------
# scaled NB primitives
# --------------------------------------------------------------------------

def fit_mu_hat(x, b, alpha):
    """util/scaled_nb.py:139-183: array secant (scipy.optimize.newton without
    fprime) from mean(x/b), brentq fallback per failed pixel."""
    x = np.asarray(x)
    b = np.asarray(b, dtype=float)
    alpha = np.asarray(alpha, dtype=float)

    def g(mu):
        if hasattr(mu, 'ndim') and 0 < mu.ndim < b.ndim:
            mu = mu[:, None]
        return np.sum((x - mu * b) / (mu + alpha * mu ** 2 * b), axis=-1)

    if x.ndim != 2:
        root = np.array([-1.0])
        failed = np.array([True])
    else:
        root, conv, zero_der = newton(g, np.mean(x / b, axis=1), maxiter=100,
                                      full_output=True)
        failed = ~conv | zero_der
        failed[root <= 0] = True
        failed[root >= np.sqrt(np.finfo(float).max) / 1e10] = True
        failed[~np.isclose(g(root), 0, atol=1e-5)] = True
    for i in np.where(failed)[0]:
        lo = 10 * np.finfo(float).eps
        hi = np.mean(x[i] / b[i]) if x.ndim == 2 else np.mean(x / b)
        gi = (lambda y: g(y)) if np.isscalar(g(lo)) else (lambda y: g(y)[i])
        for _ in range(102):
            try:
                root[i] = brentq(gi, lo, hi)
                break
            except ValueError:
                hi *= 2
        else:
            raise ValueError('bracketing interval not found within 100 '
                             'doublings')
    return root


def q2q(x, mu_in, mu_out, alpha):
    """util/scaled_nb.py:239-275.  NOTE: clamps mu_in and the caller's mu_out
    IN PLACE (order-dependent across replicates, :240-242)."""
    high = (mu_in >= 0.25) & (mu_out >= 0.25)
    mu_in[~high] = 0.25
    mu_out[~high] = 0.25
    r_in, r_out = 1 + alpha * mu_in, 1 + alpha * mu_out
    v_in, v_out = mu_in * r_in, mu_out * r_out
    right = x >= mu_in
    n_in, n_out = stats.norm(mu_in, np.sqrt(v_in)), \
        stats.norm(mu_out, np.sqrt(v_out))
    g_in, g_out = stats.gamma(mu_in / r_in, scale=r_in), \
        stats.gamma(mu_out / r_out, scale=r_out)
    qn = np.where(right, n_out.isf(n_in.sf(x)), n_out.ppf(n_in.cdf(x)))
    qg = np.where(right, g_out.isf(g_in.sf(x)), g_out.ppf(g_in.cdf(x)))
    out = (qn + qg) / 2
    out[~(out >= 0)] = 0
    return out


def equalize(data, f, alpha):
    """util/scaled_nb.py:207-214."""
    f_mean = gmean(f, pseudocount=0, axis=1)
    mu_hat = fit_mu_hat(data, f, alpha)
    mu_in = mu_hat[:, None] * f
    mu_out = mu_hat * f_mean
    pseudo = np.zeros_like(data, dtype=float)
    for i in range(data.shape[1]):
        pseudo[:, i] = q2q(data[:, i], mu_in[:, i], mu_out, alpha)
    return pseudo


def cml_nll(data, delta):
    """util/dispersion.py:72-75."""
    n = data.shape[1]
    z = np.sum(data, axis=1)
    r = 1. / delta - 1
    return -np.sum(np.sum(gammaln(data + r), axis=1) + gammaln(n * r) -
                   gammaln(z + n * r) - n * gammaln(r))


def cml(data, return_nfev=False):
    """util/dispersion.py:69-80 (bounded Brent on delta in (1e-4, 100/101))."""
    res = minimize_scalar(lambda t: cml_nll(data, t),
                          bounds=(1e-4, 100. / 101), method='bounded')
    assert res.success
    out = res.x / (1 - res.x)
    return (out, res.nfev) if return_nfev else out


def qcml(data, f=None, tol=1e-4, trace=None):
    """util/dispersion.py:31-43 (``it`` is never incremented there, so the
    loop runs until |delta| <= tol)."""
    if f is None:
        f = np.ones_like(data, dtype=float)
    disp, delta = 0.01, np.inf
    while delta > tol:
        new = cml(equalize(data, f, disp))
        delta = abs(disp - new)
        disp = new
        if trace is not None:
            trace.append(new)
        if delta < tol:
            break
    return disp


def mme(data, f=None):
    """util/dispersion.py:101-105,129-131 with the intended float division."""
    d = np.asarray(data, dtype=float)
    if f is not None:
        d = d / f
    m = np.mean(d, axis=1)
    v = np.var(d, axis=1, ddof=1)
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.nanmean((v - m) / m ** 2)


# --------------------------------------------------------------------------
# trend fit
# --------------------------------------------------------------------------

def rolling_var(y, w=20):
    """pandas Series.rolling(window=w, center=True).var() exactly as the
    reference calls it (util/lowess.py:173): window [i - w//2, i + w - w//2 - 1],
    ddof=1, NaN unless the window is full.  pandas' own online algorithm is
    used on purpose: the multiplicity of the minimum-weight point downstream is
    floor(w * (1 / w)), which flips between 0 and 1 with the last bit of the
    variance, so a mathematically equal two-pass variance is NOT equivalent
    (see DESIGN.md "trend fit sensitivity")."""
    import pandas as pd
    return pd.Series(np.asarray(y, dtype=float)).rolling(
        window=w, center=True).var().values


def lowess_curve(x, y, frac, delta=0.01):
    """util/lowess.py:72-74: returns (sorted_x, sorted_y_hat)."""
    res = lowess(y, x, frac=frac, delta=(np.nanmax(x) - np.nanmin(x)) * delta)
    return res[:, 0], res[:, 1]


def weighted_trend(x, y, frac=None, auto_frac_factor=15., w=20, power=0.25,
                   left_boundary=None, weighted=True):
    """util/lowess.py:166-244 (weighted_lowess_fit) / :10-92 (lowess_fit).

    Returns a dict describing the fitted curve: raw (x, y) points, index of
    the first increase, the unique-x lowess curve, the fraction used.  Use
    ``eval_trend`` to evaluate it."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    if not weighted:
        sx, sy = lowess_curve(x, y, 0.3 if frac is None else frac)
        _, ui = np.unique(sx, return_index=True)
        return dict(x=x, y=y, inc=0, cx=sx[ui], cy=sy[ui], frac=frac,
                    left_boundary=left_boundary, first=sy[0], weighted=False)
    n = len(y)
    order = np.argsort(x)
    x, y = x[order].copy(), y[order].copy()
    var = rolling_var(y, w)
    with np.errstate(divide='ignore'):
        prec = 1 / var
    weight = np.full(n, np.nan)
    fin = np.isfinite(prec)
    weight[fin] = np.power(prec[fin], power)
    scaled = weight * (1 / np.nanmin(weight))
    max_w = np.nanmax(scaled)
    scaled[np.isinf(scaled)] = max_w
    idx = np.arange(n)
    left_w = scaled[np.argmax(np.isfinite(scaled))]
    scaled[np.isnan(scaled) & (idx < n / 2)] = left_w
    scaled[np.isnan(scaled) & (idx > n / 2)] = 1
    assert np.all(np.isfinite(scaled))
    mult = np.floor(scaled).astype(int)
    inc = int(np.argmax(np.diff(y) > 0) + 1)
    ex = np.repeat(x[inc:], mult[inc:])
    ey = np.repeat(y[inc:], mult[inc:])
    if frac is None:
        frac = max(min(auto_frac_factor / (max_w * np.nanmean(weight)),
                       2. / 3), 0.05)
    sx, sy = lowess_curve(ex, ey, frac)
    _, ui = np.unique(sx, return_index=True)
    return dict(x=x, y=y, inc=inc, cx=sx[ui], cy=sy[ui], frac=frac,
                left_boundary=left_boundary, first=sy[0], weighted=True)


def eval_trend(fit, x_star):
    """The callable returned at util/lowess.py:76-90 and :229-242."""
    x_star = np.asarray(x_star, dtype=float)
    y_hat = interp_extrap(fit['cx'], fit['cy'], x_star)
    if fit['left_boundary'] is not None:
        y_hat = np.where(x_star <= fit['left_boundary'], fit['first'], y_hat)
    if not fit['weighted']:
        return y_hat
    x, y, inc = fit['x'], fit['y'], fit['inc']
    lin = interp_extrap(x, y, x_star)
    lin = np.where(x_star < x[0], y[0], lin)
    return np.where(x_star < x[inc], lin, y_hat)


# --------------------------------------------------------------------------
# estimate_disp / lrt / bh
# --------------------------------------------------------------------------

def combined_factor(bias, row, col, size_factors):
    """analysis/analysis.py:181-183, :272-275."""
    return bias[row, :] * bias[col, :] * size_factors


def estimate_disp(raw, f, dist, design, dist_max, estimator='qcml',
                  frac=None, auto_frac_factor=15., weighted_lowess=True,
                  stats_out=None):
    """analysis/analysis.py:185-218 on pooled (already disp_idx-filtered)
    pixels.  Returns (disp (N_d, C), disp_per_dist (D+1, C), fits)."""
    est = {'qcml': qcml, 'mme': mme,
           'cml': lambda d, f=None: cml(np.asarray(d, float) / f)}[estimator] \
        if isinstance(estimator, str) else estimator
    n_cond = design.shape[1]
    per_dist = np.zeros((dist_max + 1, n_cond))
    disp = np.zeros((len(dist), n_cond))
    fits = []
    for c in range(n_cond):
        reps = design[:, c].astype(bool)
        for d in range(dist_max + 1):
            sel = dist == d
            if not sel.any():
                per_dist[d, c] = np.nan
                continue
            per_dist[d, c] = est(raw[sel][:, reps], f=f[sel][:, reps])
        ok = np.isfinite(per_dist[:, c])
        xs = np.arange(dist_max + 1)[ok]
        ys = per_dist[:, c][ok]
        fit = weighted_trend(xs, ys, frac=frac,
                             auto_frac_factor=auto_frac_factor,
                             left_boundary=ys[0], weighted=weighted_lowess)
        fits.append(fit)
        disp[:, c] = eval_trend(fit, dist)
    return disp, per_dist, fits


def nb_logpmf(k, m, phi):
    """util/scaled_nb.py:31-33."""
    r = 1. / phi
    return gammaln(r + k) - gammaln(k + 1) - gammaln(r) + \
        r * np.log(r) - r * np.log(r + m) + k * np.log(m) - k * np.log(r + m)


def lrt(raw, f, disp_wide, design, refit_mu=True):
    """util/lrt.py:33-50."""
    design = design.astype(bool)
    n_cond = design.shape[1]
    if refit_mu:
        mu0 = fit_mu_hat(raw, f, disp_wide)
        mu1 = np.array([fit_mu_hat(raw[:, design[:, c]], f[:, design[:, c]],
                                   disp_wide[:, design[:, c]])
                        for c in range(n_cond)]).T
    else:
        mu0 = np.mean(raw / f, axis=1)
        mu1 = np.array([np.mean(raw[:, design[:, c]] / f[:, design[:, c]],
                                axis=1) for c in range(n_cond)]).T
    mu1_wide = np.dot(mu1, design.T)
    ll0 = np.sum(nb_logpmf(raw, mu0[:, None] * f, disp_wide), axis=1)
    ll1 = np.sum(nb_logpmf(raw, mu1_wide * f, disp_wide), axis=1)
    llr = ll0 - ll1
    p = stats.chi2(n_cond - 1).sf(-2 * llr)
    return p, llr, mu0, mu1


def bh(pvalues):
    """analysis/analysis.py:300 -> lib5c adjust_pvalues."""
    return adjust_pvalues(pvalues)


def run_to_qvalues(chrom_inputs, design, dist_min=4, dist_max=200,
                   bias_thresh=0.1, mean_thresh=1.0, norm='conditional_mor',
                   n_bins=-1, estimator='qcml', frac=None,
                   auto_frac_factor=15., weighted_lowess=True, refit_mu=True,
                   loops=None):
    """analysis/analysis.py:359-364 in memory.

    chrom_inputs : list of (mats, bias_raw) per chromosome.
    Returns a dict of per-chromosome stage dicts plus the genome-wide pieces.
    """
    design = np.asarray(design).astype(bool)
    per = [prepare_chrom(m, b, design, dist_min, dist_max, bias_thresh,
                         mean_thresh, norm, n_bins) for m, b in chrom_inputs]
    raws, fs, dists = [], [], []
    for st in per:
        di = st['disp_idx']
        r, c = st['row'][di], st['col'][di]
        st['f'] = combined_factor(st['bias'], r, c, st['size_factors'][di])
        raws.append(st['raw'][di])
        fs.append(st['f'])
        dists.append(c - r)
    raw, f, dist = np.concatenate(raws), np.concatenate(fs), \
        np.concatenate(dists)
    disp, per_dist, fits = estimate_disp(
        raw, f, dist, design, dist_max, estimator, frac, auto_frac_factor,
        weighted_lowess)
    offs = np.cumsum([0] + [len(d) for d in dists])
    ps = []
    for i, st in enumerate(per):
        st['disp'] = disp[offs[i]:offs[i + 1]]
        di = st['disp_idx']
        p, llr, mu0, mu1 = lrt(st['raw'][di], st['f'],
                               np.dot(st['disp'], design.T.astype(float)),
                               design, refit_mu)
        st.update(pvalues=p, llr=llr, mu_hat_null=mu0, mu_hat_alt=mu1)
        if loops is not None:
            st['loop_idx'] = loop_membership(st['row'][di], st['col'][di],
                                             loops[i])
            ps.append(p[st['loop_idx']])
        else:
            ps.append(p)
    q = bh(np.concatenate(ps))
    qoffs = np.cumsum([0] + [len(p) for p in ps])
    for i, st in enumerate(per):
        st['qvalues'] = q[qoffs[i]:qoffs[i + 1]]
    return dict(chroms=per, disp_per_dist=per_dist, fits=fits)
