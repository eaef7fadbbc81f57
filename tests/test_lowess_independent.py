"""A second, structurally different implementation of Cleveland's lowess
(W. S. Cleveland, "Robust locally weighted regression and smoothing
scatterplots", JASA 74, 1979; the ``delta`` speed-up as described for the
netlib ``lowess.f``) to pin SURVEY.md section 8 row a15 beyond the builder's
restatement in oracle/thirdparty.py: lib5c's lowess (the function the reference
calls at hic3defdr/util/lowess.py:72) is absent from /root/reference and from
the image, and no reference test pins a lowess number.

What is different here on purpose: neighbourhoods come from SORTED DISTANCES
(the k nearest points, bandwidth h = k-th smallest distance) instead of a
sliding window; every local line is a weighted LEAST-SQUARES SOLVE
(np.linalg.lstsq on the sqrt-weighted design matrix) instead of the closed-form
projection weights; the robustness pass is the textbook bisquare of
residual / (6 median |residual|).  Points at distance exactly h carry tricube
weight 0, so the choice among ties at the window edge cannot matter.

CPU tests compare the oracle restatement with it, the GPU test the device
kernel (csrc/lowess.cu)."""
import numpy as np
import pytest

from oracle.thirdparty import lowess as oracle_lowess


def local_line_at(x, y, w, x0):
    """value at x0 of the line minimising sum w (y - a - b x)^2"""
    keep = w > 0
    sw = np.sqrt(w[keep])
    A = np.stack([np.ones(keep.sum()), x[keep] - x0], axis=1) * sw[:, None]
    coef, _, rank, _ = np.linalg.lstsq(A, y[keep] * sw, rcond=None)
    if rank < 2:            # all the weight on one abscissa: weighted mean
        return float(np.sum(w[keep] * y[keep]) / np.sum(w[keep]))
    return float(coef[0])


def cleveland_lowess(x, y, frac, it, delta=0.0, statsmodels_tail=False):
    """``statsmodels_tail``: statsmodels' ``update_indices`` (whose port lib5c
    ships) leaves its scan variable at n - 1 when every remaining point lies
    within delta of the last anchor and then regresses at point n - 2 before
    the last one; Cleveland's Fortran goes straight to the last point."""
    order = np.argsort(x, kind='stable')
    x, y = np.asarray(x, float)[order], np.asarray(y, float)[order]
    n = len(x)
    k = min(max(int(frac * n + 1e-10), 2), n)
    robust = np.ones(n)
    fit = np.zeros(n)
    for _ in range(it + 1):
        # anchors: the points where a regression is actually computed
        anchors = [0]
        if delta > 0:
            last = 0
            while last < n - 1:
                within = np.flatnonzero((x > x[last]) & (x <= x[last] + delta))
                # ties of the current anchor are copies of it
                nxt = within[-1] if len(within) else \
                    np.flatnonzero(x > x[last])[0]
                if statsmodels_tail and len(within) and within[-1] == n - 1:
                    nxt = max(n - 2, last + 1)
                ties = np.flatnonzero(x == x[nxt])
                anchors.append(int(nxt))
                last = int(ties[-1])
        else:
            anchors = list(range(n))
        done = {}
        for i in anchors:
            if x[i] in done:
                fit[i] = done[x[i]]
                continue
            d = np.abs(x - x[i])
            h = np.sort(d, kind='stable')[k - 1]
            with np.errstate(divide='ignore', invalid='ignore'):
                t = np.where(d < h, d / h, 1.0) if h > 0 else \
                    np.where(d == 0, 0.0, 1.0)
            w = (1 - t ** 3) ** 3 * robust
            fit[i] = local_line_at(x, y, w, x[i]) if w.sum() > 0 else y[i]
            done[x[i]] = fit[i]
        if delta > 0:
            ax = np.array(sorted(done))
            ay = np.array([done[v] for v in ax])
            fit = np.interp(x, ax, ay)
        r = np.abs(y - fit)
        s = np.median(r)
        u = np.minimum(r / (6 * s), 1.0) if s > 0 else (r > 0).astype(float)
        robust = (1 - u ** 2) ** 2
    return x, fit


def _cases():
    rng = np.random.default_rng(11)
    out = []
    for n, frac in ((60, 0.3), (197, 0.225), (400, 0.1), (150, 2. / 3)):
        x = np.sort(rng.choice(np.arange(4, 1000), size=n, replace=False)) \
            .astype(float)
        y = 0.02 + 0.01 * np.exp(-x / 60.) + 1e-5 * x + rng.normal(0, 4e-4, n)
        y[rng.integers(0, n, 4)] += 0.01
        out.append((x, y, frac))
    # duplicated points, as weighted_lowess_fit produces them
    # (hic3defdr/util/lowess.py:206-212)
    x = np.repeat(np.arange(4, 120).astype(float), rng.integers(1, 5, 116))
    y = 0.03 - 1e-4 * x + rng.normal(0, 3e-4, len(x))
    out.append((x, y, 0.15))
    return out


@pytest.mark.parametrize('it', [0, 3])
def test_oracle_lowess_matches_independent_derivation(it):
    for x, y, frac in _cases():
        want_x, want = cleveland_lowess(x, y, frac, it)
        got = oracle_lowess(y, x, frac=frac, it=it, delta=0.0)
        np.testing.assert_array_equal(got[:, 0], want_x)
        np.testing.assert_allclose(got[:, 1], want, rtol=1e-9, atol=1e-14)


def test_oracle_lowess_delta_skipping_matches_independent_derivation():
    for x, y, frac in _cases():
        delta = 0.01 * (x.max() - x.min())
        want_x, want = cleveland_lowess(x, y, frac, 3, delta,
                                        statsmodels_tail=True)
        got = oracle_lowess(y, x, frac=frac, it=3, delta=delta)
        np.testing.assert_allclose(got[:, 1], want, rtol=1e-9, atol=1e-14)
        # the tail quirk is immaterial: Cleveland's own anchor rule moves the
        # curve by less than 1e-5 relative (the trend's noise is ~1e-2)
        _, pure = cleveland_lowess(x, y, frac, 3, delta)
        np.testing.assert_allclose(pure, want, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('it,dfrac', [(0, 0.0), (3, 0.0), (3, 0.01)])
def test_device_lowess_matches_independent_derivation(it, dfrac):
    from hic3defdr_b200.trend import _device_lowess
    for x, y, frac in _cases():
        want_x, want = cleveland_lowess(x, y, frac, it,
                                        dfrac * (x.max() - x.min()),
                                        statsmodels_tail=True)
        sx, sy = _device_lowess(x, y, frac, dfrac, it=it)
        np.testing.assert_array_equal(sx, want_x)
        np.testing.assert_allclose(sy, want, rtol=1e-9, atol=1e-14)
