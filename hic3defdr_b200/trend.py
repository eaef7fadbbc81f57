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


def _sorted_finite(x, y):
    ok = np.isfinite(x) & np.isfinite(y)
    x, y = x[ok], y[ok]
    order = np.argsort(x, kind='stable')
    return np.ascontiguousarray(x[order]), np.ascontiguousarray(y[order])


def _device_lowess_batch(jobs, it=3):
    """jobs: list of (x, y, frac, delta_frac); all smoothed in ONE kernel launch
    (one thread block per job; one upload, one read-back).  Returns a list of
    (x_sorted, y_fit)."""
    import torch
    from hic3defdr_b200._native import lib, ptr
    from hic3defdr_b200.ops import to_host
    xs, ys, offs, ns, fracs, deltas = [], [], [], [], [], []
    total = 0
    for x, y, frac, delta_frac in jobs:
        x, y = _sorted_finite(np.asarray(x, float), np.asarray(y, float))
        xs.append(x)
        ys.append(y)
        offs.append(total)
        ns.append(len(x))
        fracs.append(float(frac))
        deltas.append((np.nanmax(x) - np.nanmin(x)) * delta_frac)
        total += (len(x) + 31) // 32 * 32
    packed = np.zeros((2, total))
    for x, y, o in zip(xs, ys, offs):
        packed[0, o:o + len(x)] = x
        packed[1, o:o + len(x)] = y
    dev = torch.from_numpy(packed).cuda()
    out = torch.empty(total, dtype=torch.float64, device='cuda')
    wsb = lib().query('h3d_lowess_batch_ws_bytes', total, len(jobs))
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    f64 = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    o_h, n_h, f_h, d_h = i32(offs), i32(ns), f64(fracs), f64(deltas)
    lib().call('h3d_lowess_batch', ptr(dev[0]), ptr(dev[1]), ptr(o_h),
               ptr(n_h), ptr(f_h), int(it), ptr(d_h), len(jobs), ptr(out),
               ptr(ws), wsb, torch.cuda.current_stream().cuda_stream)
    fit = to_host(out) if total * 8 <= 65536 else out.cpu().numpy()
    return [(x, fit[o:o + len(x)].copy()) for x, o in zip(xs, offs)]


def _device_lowess(x, y, frac, delta_frac, it=3):
    """sorted-x lowess on the GPU; returns (x_sorted, y_fit)."""
    return _device_lowess_batch([(x, y, frac, delta_frac)], it=it)[0]


def _plain_job(x, y, logx, logy, frac, delta):
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    job = (np.log(x) if logx else x, np.log(y) if logy else y, frac, delta)
    return job, dict(x=x, y=y, inc=0, frac=frac, weighted=False)


def _weighted_job(x, y, logx, logy, frac, auto_frac_factor, delta, w, power,
                  interpolate_before_increase, messages):
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
        messages.append('  using auto-determined lowess fraction of %.3f'
                        % frac)
    job = (np.log(ex) if logx else ex, np.log(ey) if logy else ey, frac, delta)
    return job, dict(x=x, y=y, inc=inc, frac=frac, weighted=True)


def _finish(info, sx, sy, left_boundary, right_boundary, logx, logy):
    _, ui = np.unique(sx, return_index=True)
    return DispersionTrend(info['x'], info['y'], info['inc'], sx[ui], sy[ui],
                           info['frac'], left_boundary, right_boundary,
                           weighted=info['weighted'], logx=logx, logy=logy)


def fit_many(specs, weighted=True):
    """One trend per (x, y, kwargs) in ``specs`` -- the kwargs of
    ``weighted_lowess_fit`` / ``lowess_fit`` -- with every smoothing in the
    same kernel launch.  Returns the list of callables."""
    jobs, infos, msgs = [], [], []
    for x, y, kw in specs:
        kw = dict(kw)
        logx, logy = kw.pop('logx', False), kw.pop('logy', False)
        lb, rb = kw.pop('left_boundary', None), kw.pop('right_boundary', None)
        if weighted:
            job, info = _weighted_job(
                x, y, logx, logy, kw.get('frac'),
                kw.get('auto_frac_factor', 15.), kw.get('delta', 0.01),
                kw.get('w', 20), kw.get('power', 1. / 4),
                kw.get('interpolate_before_increase', True), msgs)
        else:
            job, info = _plain_job(x, y, logx, logy, kw.get('frac', 0.3),
                                   kw.get('delta', 0.01))
        info.update(lb=lb, rb=rb, logx=logx, logy=logy)
        jobs.append(job)
        infos.append(info)
    for m in msgs:
        print(m, file=sys.stderr)
    fits = _device_lowess_batch(jobs)
    return [_finish(i, sx, sy, i['lb'], i['rb'], i['logx'], i['logy'])
            for i, (sx, sy) in zip(infos, fits)]


def lowess_fit(x, y, logx=False, logy=False, left_boundary=None,
               right_boundary=None, frac=0.3, delta=0.01):
    """hic3defdr/util/lowess.py:10-92."""
    return fit_many([(x, y, dict(logx=logx, logy=logy,
                                 left_boundary=left_boundary,
                                 right_boundary=right_boundary, frac=frac,
                                 delta=delta))], weighted=False)[0]


def weighted_lowess_fit(x, y, logx=False, logy=False, left_boundary=None,
                        right_boundary=None, frac=None, auto_frac_factor=15.,
                        delta=0.01, w=20, power=1. / 4,
                        interpolate_before_increase=True):
    """hic3defdr/util/lowess.py:95-244."""
    return fit_many([(x, y, dict(
        logx=logx, logy=logy, left_boundary=left_boundary,
        right_boundary=right_boundary, frac=frac,
        auto_frac_factor=auto_frac_factor, delta=delta, w=w, power=power,
        interpolate_before_increase=interpolate_before_increase))],
        weighted=True)[0]
