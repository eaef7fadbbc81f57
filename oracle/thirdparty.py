"""
TEST INFRASTRUCTURE -- restatements of the third-party arithmetic that sits on
the reference's hot path but is NOT under /root/reference.

Dependency: lib5c==0.6.0 (reference requirements.txt:15, setup.py:46), which in
turn wraps statsmodels==0.10.2 (requirements.txt:36).  Neither is installed in
the image; the algorithms below are restated from their published definitions.
Call sites in the reference that anchor each one are cited per function.
"""
import os

import numpy as np


def gmean(x, pseudocount=1, axis=None):
    """lib5c.util.mathematics.gmean.

    Call sites: hic3defdr/util/scaling.py:47,65 (default pseudocount),
    hic3defdr/util/scaled_nb.py:207 (pseudocount=0).
    Pinned: docs/median_of_ratios.rst:27-32 only reproduces with the default
    pseudocount of 1, i.e. exp(mean(log(x + 1))) - 1.
    """
    x = np.asarray(x, dtype=float)
    return np.exp(np.mean(np.log(x + pseudocount), axis=axis)) - pseudocount


def adjust_pvalues(pvalues):
    """lib5c.util.statistics.adjust_pvalues -> statsmodels multipletests
    (method='fdr_bh') applied to the finite entries; non-finite entries stay
    NaN.  Call site: hic3defdr/analysis/analysis.py:300.  PARITY UNPINNED.

    Benjamini-Hochberg: sort ascending, q_(i) = p_(i) / (i / n), enforce
    monotonicity with a running minimum from the largest p downwards, clip at
    1, undo the sort.
    """
    p = np.asarray(pvalues, dtype=float)
    q = np.full(p.shape, np.nan)
    finite = np.isfinite(p)
    pf = p[finite]
    n = pf.size
    if n == 0:
        return q
    order = np.argsort(pf, kind='stable')
    ranked = pf[order] / (np.arange(1, n + 1) / float(n))   # p / ecdf
    ranked = np.minimum.accumulate(ranked[::-1])[::-1]
    ranked[ranked > 1] = 1
    out = np.empty(n)
    out[order] = ranked
    q[finite] = out
    return q


def check_outdir(path):
    """lib5c.util.system.check_outdir: mkdir -p of dirname(path), announcing a
    creation on stdout (pinned by README.md:116).  Call site:
    hic3defdr/analysis/constructor.py:84."""
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        print('creating directory %s' % d)
        os.makedirs(d)


def _tricube(t):
    return (1.0 - t * t * t) ** 3


def lowess(endog, exog, frac=2. / 3, it=3, delta=0.0):
    """lib5c.util.lowess.lowess (a port of statsmodels 0.10.2
    nonparametric/_smoothers_lowess.pyx).  Call site:
    hic3defdr/util/lowess.py:72.  PARITY UNPINNED.

    Cleveland's robust locally weighted regression: k = int(frac*n + 1e-10)
    nearest neighbours clipped to [2, n], tricube kernel, local linear fit,
    ``it`` robustifying passes with bisquare weights of resid/(6*median|resid|),
    and the ``delta`` speed-up (regress only at anchors more than delta apart,
    linearly interpolate between anchors, copy fits across tied x).
    Returns an (n, 2) array [sorted x, fitted y].
    """
    x = np.asarray(exog, dtype=float)
    y = np.asarray(endog, dtype=float)
    ok = np.isfinite(x) & np.isfinite(y)
    x, y = x[ok], y[ok]
    order = np.argsort(x, kind='stable')
    x, y = x[order], y[order]
    n = x.size
    k = int(frac * n + 1e-10)
    k = min(max(k, 2), n)
    y_fit = np.zeros(n)
    resid_w = np.ones(n)
    for robiter in range(it + 1):
        i = 0
        last = -1
        left, right = 0, k
        y_fit[:] = 0.0
        while True:
            # slide the k-wide window right while the left gap exceeds the
            # right one
            while right < n and (x[i] - x[left]) > (x[right] - x[i]):
                left += 1
                right += 1
            xj = x[left:right]
            dij = np.abs(xj - x[i])
            radius = max(dij[0], dij[-1])
            with np.errstate(divide='ignore', invalid='ignore'):
                w = _tricube(dij / radius) * resid_w[left:right]
            sw = np.sum(w)
            if not sw > 0.0:
                y_fit[i] = y[i]
            else:
                w = w / sw
                xbar = np.sum(w * xj)
                sq = np.sum(w * (xj - xbar) ** 2)
                with np.errstate(divide='ignore', invalid='ignore'):
                    p = w * (1.0 + (x[i] - xbar) * (xj - xbar) / sq)
                y_fit[i] = np.sum(p * y[left:right])
            # fill in anchors skipped because of delta
            if last < i - 1:
                a = x[i] - x[last]
                for j in range(last + 1, i):
                    al = (x[j] - x[last]) / a
                    y_fit[j] = al * y_fit[i] + (1.0 - al) * y_fit[last]
            # choose the next anchor
            last = i
            cut = x[last] + delta
            kk = last + 1
            stopped = False
            for kk in range(last + 1, n):
                if x[kk] > cut:
                    stopped = True
                    break
                if x[kk] == x[last]:
                    y_fit[kk] = y_fit[last]
                    last = kk
            if not stopped and last + 1 < n:
                kk = n - 1
            i = max(kk - 1, last + 1)
            if last >= n - 1:
                break
        if robiter < it:
            r = np.abs(y - y_fit)
            med = np.median(r)
            if med == 0:
                r = (r > 0).astype(float)
            else:
                r = r / (6.0 * med)
            r[r >= 1.0] = 1.0
            resid_w = (1.0 - r * r) ** 2
    return np.array([x, y_fit]).T
