"""
Distance -> dispersion trend: host-side mirror of
``hic3defdr.util.lowess.lowess_fit`` / ``weighted_lowess_fit``
(hic3defdr/util/lowess.py:10-92, 95-244).

Split of work: the point weighting (rolling variance -> precision**(1/4) ->
integer multiplicities; at most dist_thresh_max + 1 numbers) is host glue kept
formula-for-formula with the reference, including its use of pandas' rolling
variance, because the multiplicity of the minimum-weight point is
``floor(w * (1 / w))`` and flips between 0 and 1 with the last bit of the
variance (DESIGN.md "trend fit sensitivity").  The smoothing itself -- the
lowess over the duplicated point set -- runs on the GPU (csrc/lowess.cu).
The fitted curve is returned as a picklable callable, the contract of
``disp_fn_<cond>.pickle`` (hic3defdr/analysis/core.py:220-253).
"""
import sys

import numpy as np


def _interp_extrap(xp, yp, x):
    """Piecewise-linear interpolation with linear extrapolation, evaluated the
    way scipy's interp1d(fill_value='extrapolate') does (same formula as the
    size-factor table kernel in csrc/rank.cu)."""
    xp = np.asarray(xp, dtype=float)
    yp = np.asarray(yp, dtype=float)
    x = np.asarray(x, dtype=float)
    if len(xp) == 1:
        return np.full(x.shape, yp[0])
    hi = np.clip(np.searchsorted(xp, x), 1, len(xp) - 1).astype(int)
    lo = hi - 1
    return ((x - xp[lo]) / (xp[hi] - xp[lo])) * yp[hi] + \
        ((xp[hi] - x) / (xp[hi] - xp[lo])) * yp[lo]


class DispersionTrend(object):
    """Callable ``disp_fn``: vectorised over distances (int or float, in or
    out of the fitted range)."""

    def __init__(self, x, y, inc, curve_x, curve_y, frac, left_boundary,
                 right_boundary, weighted, logx=False, logy=False):
        self.x = np.asarray(x, dtype=float)
        self.y = np.asarray(y, dtype=float)
        self.inc = int(inc)
        self.curve_x = np.asarray(curve_x, dtype=float)
        self.curve_y = np.asarray(curve_y, dtype=float)
        self.frac = frac
        self.left_boundary = left_boundary
        self.right_boundary = right_boundary
        self.weighted = bool(weighted)
        self.logx = bool(logx)
        self.logy = bool(logy)

    def _smooth(self, x_star):
        # the closure of lowess_fit (util/lowess.py:76-90)
        new_x = np.log(x_star) if self.logx else x_star
        y_hat = _interp_extrap(self.curve_x, self.curve_y, new_x)
        if self.left_boundary is not None:
            y_hat = np.where(x_star <= self.left_boundary, self.curve_y[0],
                             y_hat)
        if self.right_boundary is not None:
            y_hat = np.where(x_star >= self.right_boundary, self.curve_y[-1],
                             y_hat)
        return np.exp(y_hat) if self.logy else y_hat

    def __call__(self, x_star):
        x_star = np.asarray(x_star)
        xs = x_star.astype(float)
        smooth = self._smooth(xs)
        if not self.weighted:
            return smooth
        # the closure of weighted_lowess_fit (util/lowess.py:229-242)
        lin = _interp_extrap(self.x, self.y, xs)
        lin = np.where(xs < self.x[0], self.y[0], lin)
        return np.where(xs < self.x[self.inc], lin, smooth)


def point_multiplicities(y_sorted, w=20, power=0.25):
    """util/lowess.py:172-204: returns (multiplicity per point, max scaled
    weight, mean unscaled weight)."""
    import pandas as pd
    n = len(y_sorted)
    var = pd.Series(y_sorted).rolling(window=w, center=True).var().values
    with np.errstate(divide='ignore', invalid='ignore'):
        prec = 1 / var
    weight = np.full(n, np.nan)
    fin = np.isfinite(prec)
    weight[fin] = np.power(prec[fin], power)
    scaled = weight * (1 / np.nanmin(weight))
    max_w = np.nanmax(scaled)
    scaled[np.isinf(scaled)] = max_w
    pos = np.arange(n)
    left_w = scaled[np.argmax(np.isfinite(scaled))]
    scaled[np.isnan(scaled) & (pos < n / 2)] = left_w
    scaled[np.isnan(scaled) & (pos > n / 2)] = 1
    assert np.all(np.isfinite(scaled))
    return np.floor(scaled).astype(int), max_w, np.nanmean(weight)


def _device_lowess(x, y, frac, delta_frac, it=3):
    """sorted-x lowess on the GPU; returns (x_sorted, y_fit)."""
    import torch
    from hic3defdr_b200._native import lib, ptr
    ok = np.isfinite(x) & np.isfinite(y)
    x, y = x[ok], y[ok]
    order = np.argsort(x, kind='stable')
    x, y = np.ascontiguousarray(x[order]), np.ascontiguousarray(y[order])
    n = len(x)
    delta = (np.nanmax(x) - np.nanmin(x)) * delta_frac
    xd = torch.from_numpy(x).cuda()
    yd = torch.from_numpy(y).cuda()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    wsb = lib().query('h3d_lowess_ws_bytes', n)
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    lib().call('h3d_lowess', ptr(xd), ptr(yd), n, float(frac), int(it),
               float(delta), ptr(out), ptr(ws), wsb,
               torch.cuda.current_stream().cuda_stream)
    from hic3defdr_b200.ops import to_host
    return x, to_host(out)


def lowess_fit(x, y, logx=False, logy=False, left_boundary=None,
               right_boundary=None, frac=0.3, delta=0.01):
    """hic3defdr/util/lowess.py:10-92."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    fx = np.log(x) if logx else x
    fy = np.log(y) if logy else y
    sx, sy = _device_lowess(fx, fy, frac, delta)
    _, ui = np.unique(sx, return_index=True)
    return DispersionTrend(x, y, 0, sx[ui], sy[ui], frac, left_boundary,
                           right_boundary, weighted=False, logx=logx,
                           logy=logy)


def weighted_lowess_fit(x, y, logx=False, logy=False, left_boundary=None,
                        right_boundary=None, frac=None, auto_frac_factor=15.,
                        delta=0.01, w=20, power=1. / 4,
                        interpolate_before_increase=True):
    """hic3defdr/util/lowess.py:95-244."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    order = np.argsort(x)
    x, y = x[order].copy(), y[order].copy()
    mult, max_w, mean_w = point_multiplicities(y, w=w, power=power)
    inc = int(np.argmax(np.diff(y) > 0) + 1) if interpolate_before_increase \
        else 0
    ex = np.repeat(x[inc:], mult[inc:])
    ey = np.repeat(y[inc:], mult[inc:])
    if frac is None:
        frac = max(min(auto_frac_factor / (max_w * mean_w), 2. / 3), 0.05)
        print('  using auto-determined lowess fraction of %.3f' % frac,
              file=sys.stderr)
    fx = np.log(ex) if logx else ex
    fy = np.log(ey) if logy else ey
    sx, sy = _device_lowess(fx, fy, frac, delta)
    _, ui = np.unique(sx, return_index=True)
    return DispersionTrend(x, y, inc, sx[ui], sy[ui], frac, left_boundary,
                           right_boundary, weighted=True, logx=logx, logy=logy)
