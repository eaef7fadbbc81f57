"""
Fixtures for SURVEY.md section 8(f) row 2, recorded from the UNMODIFIED
reference in the build container:
  * ``perturb_cluster`` (hic3defdr/util/simulation.py:12-67) applied cluster
    after cluster to a sparse mean matrix (overlapping clusters, clusters at the
    matrix edge, positive and negative effects);
  * ``kr_balance`` (hic3defdr/util/balancing.py:5-208) on a banded synthetic
    contact matrix with empty bins: bias vector, residual trace, checksum of
    the balanced matrix;
  * ``filter_sparse_rows_count`` (hic3defdr/util/filtering.py:7-63).
The reference's NB sampler itself needs lib5c.util.distributions (absent) and
numpy's global generator; the device sampler is checked against the
distribution, not against recorded draws (tests/test_gpu_simulate.py).

    python tests/golden/make_golden_sim.py  ->  ref_sim.npz
"""
import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sparse

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402
from hic3defdr_b200.synth import make_chrom  # noqa: E402


def main():
    refrun.install()
    # environment normalisation (not a change of the reference): balancing.py:105
    # spells float64 as np.float_, an alias numpy 2 removed
    if not hasattr(np, 'float_'):
        np.float_ = np.float64
    from hic3defdr.util.simulation import perturb_cluster
    from hic3defdr.util.balancing import kr_balance
    from hic3defdr.util.filtering import filter_sparse_rows_count
    rng = np.random.default_rng(2718)
    out = {}
    # ---- perturb_cluster ---------------------------------------------------
    n = 120
    row, col = np.triu_indices(n)
    keep = (col - row <= 25) & (rng.random(len(row)) < 0.85)
    row, col = row[keep], col[keep]
    mean = rng.gamma(2.0, 3.0, len(row))
    m = sparse.coo_matrix((mean, (row, col)), shape=(n, n)).tocsr()
    clusters, effects = [], []
    for k in range(14):
        r0, d0 = int(rng.integers(0, n - 12)), int(rng.integers(2, 20))
        size = int(rng.integers(1, 9))
        px = {(r0 + int(a), min(n - 1, r0 + d0 + int(b)))
              for a, b in rng.integers(0, 3, size=(size, 2))}
        clusters.append(sorted(px))
        effects.append(float(rng.choice([0.5, -0.5, 0.0])))
    clusters.append([(0, 0), (0, 1), (1, 1)])        # matrix corner
    effects.append(0.5)
    clusters.append(list(clusters[0]))               # overlaps cluster 0
    effects.append(-0.5)
    for cl, e in zip(clusters, effects):
        if e:
            perturb_cluster(m, cl, e)
    coo = m.tocoo()
    assert np.array_equal(coo.row, row) and np.array_equal(coo.col, col)
    out['pt_row'], out['pt_col'], out['pt_mean'] = row, col, mean
    out['pt_clusters'] = np.array([[i, r, c] for i, cl in enumerate(clusters)
                                   for r, c in cl])
    out['pt_effects'] = np.array(effects)
    out['pt_out'] = coo.data
    # ---- kr_balance / filter_sparse_rows_count ------------------------------
    mats, _, _ = make_chrom(1500, 1, 200, seed=99, amp=40.0, bad_frac=0.0)
    a = mats[0].astype(float).tolil()
    for b in (7, 8, 400, 1499):                      # empty bins
        a[b, :] = 0
        a[:, b] = 0
    a = sparse.triu(a.tocsr()).tocsr()
    a.eliminate_zeros()
    out['kr_indptr'], out['kr_indices'], out['kr_data'] = \
        a.indptr, a.indices, a.data
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        balanced, bias, res = kr_balance(a, fl=1)
    out['kr_bias'], out['kr_res'] = bias, res
    out['kr_balanced_sum'] = np.array(balanced.sum())
    out['kr_balanced_rowsum'] = np.asarray(
        (balanced + balanced.T - sparse.diags([balanced.diagonal()], [0]))
        .sum(axis=1)).ravel()
    print('kr_balance: %d outer iterations, final residual %.2e'
          % (len(res), res[-1]))
    filt = filter_sparse_rows_count(a, min_nnz=50, k=100)
    out['fs_kept_rows'] = np.flatnonzero(np.diff(filt.tocsr().indptr) > 0)
    out['fs_nnz'] = np.array(filt.nnz)
    out['fs_sum'] = np.array(filt.sum())
    print('filter: %d of %d bins keep entries' % (len(out['fs_kept_rows']), 1500))
    np.savez_compressed(os.path.join(HERE, 'ref_sim.npz'), **out)


if __name__ == '__main__':
    main()
