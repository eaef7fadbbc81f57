"""
Host-side mirror of ``hic3defdr.util.simulation`` (perturb_cluster, simulate;
hic3defdr/util/simulation.py:12-204).  The class labels are drawn on the host
exactly as the reference draws them (one ``np.random.choice`` from numpy's
global generator), the cluster footprints are turned into per-pixel factors,
and the heavy part -- biasing and negative-binomial sampling of every pixel of
every simulated replicate -- runs in libh3d (csrc/simulate.cu).
"""
import sys

import numpy as np
import scipy.sparse as sparse
import torch

from hic3defdr_b200 import ops
from hic3defdr_b200._native import lib, ptr


def eprint(*args, **kwargs):
    if not kwargs.pop('skip', False):
        print(*args, file=sys.stderr, **kwargs)


def cluster_footprint(cluster, shape):
    """The effect footprint of hic3defdr/util/simulation.py:36-50 as a dict
    pixel -> weight: 1 under the cluster, 1/2 on the rest of its 8-connected
    dilation (clipped to the matrix)."""
    px = {(int(r), int(c)) for r, c in cluster}
    out = {}
    for r, c in px:
        for a in (-1, 0, 1):
            for b in (-1, 0, 1):
                q = (r + a, c + b)
                if 0 <= q[0] < shape[0] and 0 <= q[1] < shape[1]:
                    out[q] = 1.0 if q in px else 0.5
    return out


def perturbation_factors(clusters, effects, shape):
    """(pixel keys int64, factor) of the compounded perturbation
    ``value += value * footprint * effect`` of every cluster with a non-zero
    effect, applied in cluster order (simulation.py:151-159)."""
    acc = {}
    for cluster, effect in zip(clusters, effects):
        if effect == 0:
            continue
        for q, wgt in cluster_footprint(cluster, shape).items():
            acc[q] = acc.get(q, 1.0) * (1.0 + wgt * effect)
    if not acc:
        return np.zeros(0, np.int64), np.zeros(0)
    keys = np.array([(r << 32) | c for r, c in acc], dtype=np.int64)
    order = np.argsort(keys)
    return keys[order], np.array(list(acc.values()))[order]


def perturb_cluster(matrix, cluster, effect, respect_zeros=True):
    """hic3defdr/util/simulation.py:12-67 (in place, scipy sparse or dense)."""
    fp = cluster_footprint(cluster, matrix.shape)
    if isinstance(matrix, sparse.spmatrix):
        m = matrix.tocsr()
        for (r, c), wgt in fp.items():
            v = m[r, c]
            if v != 0 or not respect_zeros:
                matrix[r, c] = v + v * wgt * effect
    else:
        for (r, c), wgt in fp.items():
            matrix[r, c] += matrix[r, c] * wgt * effect


def simulate(row, col, mean, disp_fn, bias, size_factors, clusters, beta=0.5,
             p_diff=0.4, trend='mean', verbose=True, seed=None):
    """hic3defdr/util/simulation.py:70-204 -> (classes, generator of
    ``scipy.sparse.csr_matrix``), same arguments.  ``seed`` keys the device
    sampler (default: drawn from numpy's global generator, so that
    ``np.random.seed`` makes a run reproducible as it does for the reference).
    Every simulated matrix stores an entry for every pixel with a positive
    mean, zeros included, like the reference's."""
    eprint('  assigning cluster classes', skip=not verbose)
    p = [1 - p_diff, p_diff / 4, p_diff / 4, p_diff / 4, p_diff / 4] \
        if type(p_diff) == float else [1 - sum(p_diff)] + list(p_diff)
    classes = np.random.choice(
        np.array(['constit', 'up A', 'down A', 'up B', 'down B'], dtype='U7'),
        size=len(clusters), p=p)
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 31 - 1))
    row, col, mean = np.asarray(row), np.asarray(col), np.asarray(mean, float)
    nonzero_idx = mean > 0
    row, col, mean = row[nonzero_idx], col[nonzero_idx], mean[nonzero_idx]
    n_bins = bias.shape[0]
    # the reference re-sorts the pixels by (row, col) through a COO -> CSR ->
    # COO round trip and asserts that nothing moved (simulation.py:166-171)
    keys_h = (row.astype(np.int64) << 32) | col.astype(np.int64)
    assert np.all(np.diff(keys_h) > 0), 'pixels must be sorted by (row, col)'

    eprint('  perturbing clusters', skip=not verbose)
    r_d, c_d = ops.dev(row, torch.int32), ops.dev(col, torch.int32)
    pixel_keys = ops.dev(keys_h)
    means = {}
    for cond, up, down in (('A', 'up A', 'down A'), ('B', 'up B', 'down B')):
        eff = [beta if cl == up else (-beta if cl == down else 0.0)
               for cl in classes]
        keys, factor = perturbation_factors(clusters, eff, (n_bins, n_bins))
        m = ops.dev(mean).clone()
        if len(keys):
            kd, fd = ops.dev(keys), ops.dev(factor)
            lib().call('h3d_perturb', ptr(pixel_keys), pixel_keys.numel(),
                       ptr(kd), ptr(fd), len(keys), ptr(m), ops._stream())
        means[cond] = m

    eprint('  renaming cluster classes', skip=not verbose)
    classes[(classes == 'up A') | (classes == 'down B')] = 'A'
    classes[(classes == 'up B') | (classes == 'down A')] = 'B'

    eprint('  preparing generator', skip=not verbose)
    size_factors = np.asarray(size_factors, dtype=float)
    n_sim = size_factors.shape[-1]
    n_sim_per_cond = int(n_sim / 2)
    bias_d = ops.dev(np.ascontiguousarray(bias, dtype=float))
    sf_d = ops.dev(np.ascontiguousarray(size_factors))
    by_dist = int(size_factors.ndim == 2)
    dist = col - row
    n_dist = max(int(size_factors.shape[0]) if by_dist else 1,
                 int(dist.max()) + 1 if len(dist) else 1)
    if by_dist and size_factors.shape[0] < n_dist:
        raise IndexError('size_factors has %d distance rows, the pixels reach '
                         'distance %d' % (size_factors.shape[0], n_dist - 1))
    counts_per_row = np.bincount(row, minlength=n_bins)
    indptr = np.concatenate([[0], np.cumsum(counts_per_row)])
    n = len(row)

    def launch(m, j, disp, per_px, counts, bm):
        lib().call('h3d_nb_simulate', ptr(r_d), ptr(c_d), ptr(m), n,
                   ptr(bias_d), n_sim, ptr(sf_d), by_dist, n_dist, ptr(disp),
                   per_px, j, seed, ptr(counts), ptr(bm), ops._stream())

    def gen():
        for j in range(n_sim):
            eprint('  biasing and simulating rep %i/%i' % (j + 1, n_sim),
                   skip=not verbose)
            m = means['A'] if j < n_sim_per_cond else means['B']
            counts = torch.empty(n, dtype=torch.int64, device='cuda')
            if trend == 'mean':
                # the dispersion is a function of the biased mean: evaluate
                # the (host) callable on it
                bm = torch.empty(n, dtype=torch.float64, device='cuda')
                launch(m, j, None, 1, None, bm)
                disp = ops.dev(np.asarray(disp_fn(bm.cpu().numpy()), float))
                launch(m, j, disp, 1, counts, None)
            else:
                disp = ops.dev(np.asarray(disp_fn(np.arange(n_dist)), float))
                launch(m, j, disp, 0, counts, None)
            yield sparse.csr_matrix(
                (counts.cpu().numpy(), col.astype(np.int32), indptr),
                shape=(n_bins, n_bins))

    return classes, gen()
