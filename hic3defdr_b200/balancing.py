"""
Host-side mirror of ``hic3defdr.util.balancing.kr_balance``
(hic3defdr/util/balancing.py:5-208) and
``hic3defdr.util.filtering.filter_sparse_rows_count``
(hic3defdr/util/filtering.py:7-63): the structural bookkeeping (symmetrise,
drop empty rows, expand, "sum factor") is scipy on the host, the iteration --
conjugate-gradient steps around sparse matrix-vector products -- and the band
counts run in libh3d (csrc/balance.cu).
"""
import ctypes

import numpy as np
import scipy.sparse as sparse
import torch

from hic3defdr_b200 import ops
from hic3defdr_b200._native import lib, ptr


def kr_balance(array, tol=1e-6, x0=None, delta=0.1, ddelta=3, fl=1,
               max_iter=3000):
    """hic3defdr/util/balancing.py:5-208 -> (balanced csr_matrix, bias vector,
    residuals).  As in the reference the residual list is only filled (and the
    convergence table printed to stdout) when ``fl == 1``."""
    triu = sparse.tril(array, k=-1).nnz == 0
    up = sparse.triu(sparse.csr_matrix(array)).astype(np.float64)
    full = (up + up.transpose() - sparse.diags([up.diagonal()], [0])).tocsr()
    nonzero = full.getnnz(1) > 0
    a = full[nonzero, :][:, nonzero].tocsr()
    a.sort_indices()
    n = a.shape[0]
    if a.shape[0] != a.shape[1]:
        raise Exception
    if fl == 1:
        print('it in. it res')
    res = np.array([])
    x = np.ones(0)
    if n:
        indptr, indices, data = ops.dev(a.indptr.astype(np.int32)), \
            ops.dev(a.indices.astype(np.int32)), ops.dev(a.data)
        xd = torch.empty(n, dtype=torch.float64, device='cuda')
        x0d = None if x0 is None else ops.dev(
            np.ascontiguousarray(np.asarray(x0, dtype=float).reshape(-1)))
        cap = (max_iter + 2) if max_iter is not None else 100000
        res_h = np.zeros(cap)
        n_res, n_mv = ctypes.c_int(0), ctypes.c_int(0)
        wsb = lib().query('h3d_kr_balance_ws_bytes', n)
        ws = ops.workspace(wsb)
        lib().call('h3d_kr_balance', ptr(indptr), ptr(indices), ptr(data), n,
                   float(tol), ptr(x0d), float(delta), float(ddelta),
                   -1 if max_iter is None else int(max_iter), ptr(xd),
                   ptr(res_h), cap, ctypes.byref(n_res), ctypes.byref(n_mv),
                   ptr(ws), wsb, ops._stream())
        x = xd.cpu().numpy()
        if fl == 1:
            res = res_h[:min(n_res.value, cap)].copy()
            for i, r in enumerate(res):
                print('{} {} {:.3e}'.format(i + 1, '-', r))
    # expand, scale to the magnitude of the input ("sum factor"), invert
    # (balancing.py:176-200)
    bias = np.zeros(len(nonzero), dtype=float)
    bias[nonzero] = x
    scale = sparse.diags([bias], [0])
    balanced = scale.dot(full).dot(scale)
    sum_factor = np.sqrt(full.sum() / balanced.sum())
    bias *= sum_factor
    scale = sparse.diags([bias], [0])
    balanced = scale.dot(full).dot(scale)
    bias[bias != 0] = 1 / bias[bias != 0]
    if triu:
        balanced = sparse.triu(balanced).tocsr()
    return balanced, bias, res


def band_nnz(matrix, k=300):
    """Per bin: number of positive entries among the k nearest upstream and
    downstream contacts of the (symmetrised upper triangle of the) matrix."""
    m = sparse.triu(sparse.csr_matrix(matrix)).tocsr()
    m.sort_indices()
    n = m.shape[0]
    up = torch.zeros(n, dtype=torch.int32, device='cuda')
    down = torch.zeros(n, dtype=torch.int32, device='cuda')
    if m.nnz:
        # named: a temporary's device block would be reused by the next one
        ip = ops.dev(m.indptr.astype(np.int64))
        ix = ops.dev(m.indices.astype(np.int32))
        dt = ops.dev(m.data.astype(np.float64))
        lib().call('h3d_band_nnz', ptr(ip), ptr(ix), ptr(dt), n, int(k),
                   ptr(up), ptr(down), ops._stream())
    return up.cpu().numpy(), down.cpu().numpy()


def filter_sparse_rows_count(matrix, min_nnz=25, k=300):
    """hic3defdr/util/filtering.py:7-63 for scipy sparse or dense input: wipes
    the bins (rows and columns) that make fewer than ``min_nnz`` positive
    contacts with both their ``k`` nearest upstream and downstream bins."""
    if min_nnz == 0 or k == 0:
        return matrix.copy()
    dense = isinstance(matrix, np.ndarray)
    up, down = band_nnz(sparse.csr_matrix(matrix) if dense else matrix,
                        min(k, matrix.shape[0]))
    deleted = (up < min_nnz) & (down < min_nnz)
    if dense:
        out = matrix.copy()
        out[:, deleted] = 0
        out[deleted, :] = 0
        return out
    keep = sparse.diags([(~deleted).astype(int)], [0], dtype=int).tocsr()
    keep.eliminate_zeros()
    return keep.dot(sparse.csr_matrix(matrix)).dot(keep)
