"""
Host-side mirror of ``hic3defdr.util.evaluation`` (make_y_true, evaluate,
compute_fdr; hic3defdr/util/evaluation.py:15-100), the arithmetic of which runs
in libh3d (csrc/roc.cu): the ROC curve of sklearn.metrics.roc_curve -- one
radix sort of the q-values, prefix counts of the labels, the corners of the
curve -- and the observed FDR at a subsample of its thresholds.
"""
import numpy as np
import torch

from hic3defdr_b200 import ops
from hic3defdr_b200._native import lib, ptr


def make_y_true(row, col, clusters, labels):
    """hic3defdr/util/evaluation.py:15-41: True for the pixels (row[i], col[i])
    that lie in a cluster whose label is not 'constit'.  Returns a numpy bool
    vector; the membership test is the device kernel of ``loop_idx``."""
    labels = np.asarray(labels)
    sig = [c for c, lab in zip(clusters, np.atleast_1d(labels))
           if lab != 'constit']
    pixels = set().union(*sig) if sig else set()
    r, c = ops.dev(row, torch.int32), ops.dev(col, torch.int32)
    if not pixels or r.numel() == 0:
        return np.zeros(r.numel(), dtype=bool)
    return ops.loop_membership(r, c, None, pixels).cpu().numpy().astype(bool)


def roc_curve(y_true, qvalues):
    """sklearn.metrics.roc_curve(y_true, 1 - qvalues) (drop_intermediate=True)
    -> (fps, tps, thresholds) with the leading (0, 0, inf) point, as numpy
    arrays (int64, int64, float64)."""
    y = ops.dev(np.asarray(y_true).astype(np.uint8)
                if not isinstance(y_true, torch.Tensor) else y_true,
                torch.uint8)
    q = ops.dev(qvalues, torch.float64)
    n = q.numel()
    if y.numel() != n:
        raise ValueError('y_true and qvalues differ in length')
    if n == 0 or bool(torch.isnan(q).any()):
        raise ValueError('qvalues must be non-empty and free of NaN '
                         '(sklearn.metrics.roc_curve raises here too)')
    keys = torch.empty(n, dtype=torch.int64, device='cuda')
    boundary = torch.empty(n, dtype=torch.uint8, device='cuda')
    ys = torch.empty(n, dtype=torch.uint8, device='cuda')
    wsb = lib().query('h3d_roc_sort_ws_bytes', n)
    ws = ops.workspace(wsb)
    lib().call('h3d_roc_sort', ptr(q), ptr(y), n, ptr(keys), ptr(boundary),
               ptr(ys), ptr(ws), wsb, ops._stream())
    thr_idx = ops.mask_to_index(boundary)
    pos_idx = ops.mask_to_index(ys)
    m = thr_idx.numel()
    tps = torch.empty(m, dtype=torch.int64, device='cuda')
    fps = torch.empty_like(tps)
    thr = torch.empty(m, dtype=torch.float64, device='cuda')
    keep = torch.empty(m, dtype=torch.uint8, device='cuda')
    lib().call('h3d_roc_points', ptr(keys), ptr(thr_idx), m, ptr(pos_idx),
               pos_idx.numel(), ptr(tps), ptr(fps), ptr(thr), ptr(keep),
               ops._stream())
    sel = ops.mask_to_index(keep).long()
    fps_h = np.concatenate([[0], fps[sel].cpu().numpy()])
    tps_h = np.concatenate([[0], tps[sel].cpu().numpy()])
    thr_h = np.concatenate([[np.inf], thr[sel].cpu().numpy()])
    return fps_h, tps_h, thr_h


def evaluate(y_true, qvalues, n_fdr_points=100):
    """hic3defdr/util/evaluation.py:44-79 -> (fdr, fpr, tpr, thresh).  The
    observed FDR at a threshold is fp / (fp + tp) of the pixels whose score
    reaches it, i.e. of the curve's own counts (the reference recounts them
    with a confusion matrix per threshold)."""
    fps, tps, thresh = roc_curve(y_true, qvalues)
    with np.errstate(divide='ignore', invalid='ignore'):
        fpr = fps / float(fps[-1]) if fps[-1] > 0 else np.full(len(fps), np.nan)
        tpr = tps / float(tps[-1]) if tps[-1] > 0 else np.full(len(tps), np.nan)
    fdr = np.ones_like(fpr) * np.nan
    rate = max(int(len(thresh) / n_fdr_points), 1)
    idx = np.arange(int(np.argmax(tpr > 0)), len(thresh), rate)
    with np.errstate(divide='ignore', invalid='ignore'):
        fdr[idx] = fps[idx] / (fps[idx] + tps[idx]).astype(float)
    return fdr, fpr, tpr, thresh


def compute_fdr(y_true, y_pred):
    """hic3defdr/util/evaluation.py:82-100."""
    y_true = np.asarray(y_true).astype(bool)
    y_pred = np.asarray(y_pred).astype(bool)
    fp = int(np.count_nonzero(~y_true & y_pred))
    tp = int(np.count_nonzero(y_true & y_pred))
    return fp / float(fp + tp)
