"""Records the UNMODIFIED reference's threshold / classify arithmetic
(hic3defdr/util/thresholding.py, util/classification.py, util/clusters.py,
imported from /root/reference through oracle/refrun.py) as the fixture
tests/golden/ref_clusters.json:

  * the recorded golden pipeline outputs (ref_pipeline.npz: row, col,
    disp_idx, loop_idx, qvalues, mu_hat_alt of both chromosomes) at FDR 0.5
    and 0.3, cluster sizes 1 and 3;
  * synthetic pixel sets with large irregular components (random band
    patterns, a spiral, single pixels, an empty set).

    python tests/golden/make_golden_clusters.py        (build container only)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402


def as_lists(clusters):
    return [sorted([int(i), int(j)] for i, j in c) for c in clusters]


def main():
    refrun.install()
    from hic3defdr.util.thresholding import threshold_and_cluster, size_filter
    from hic3defdr.util.classification import classify
    from hic3defdr.util.clusters import find_clusters
    import scipy.sparse as sparse
    out = {'pipeline': [], 'find_clusters': []}
    g = np.load(os.path.join(HERE, 'ref_pipeline.npz'))
    chroms = sorted({k.split('_')[1] for k in g.files
                     if k.startswith('qvalues_')})
    for chrom in chroms:
        di, li = g['disp_idx_%s' % chrom], g['loop_idx_%s' % chrom]
        row = g['row_%s' % chrom][di][li]
        col = g['col_%s' % chrom][di][li]
        q = g['qvalues_%s' % chrom]
        mu = g['mu_hat_alt_%s' % chrom][li]
        for fdr in (0.5, 0.3):
            sig, insig = threshold_and_cluster(q, row, col, fdr)
            for size in (1, 3):
                fs, fi = size_filter(sig, size), size_filter(insig, size)
                classes = classify(row, col, mu, fs) if fs else None
                out['pipeline'].append(dict(
                    chrom=chrom, fdr=fdr, cluster_size=size,
                    sig=as_lists(fs), insig=as_lists(fi),
                    classes=[as_lists(c) for c in classes]
                    if classes is not None else None))
    rng = np.random.default_rng(42)
    cases = []
    for n, width, density in ((60, 20, 0.45), (200, 40, 0.3), (90, 90, 0.6)):
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
        keep = (j >= i) & (j - i <= width) & (rng.random((n, n)) < density)
        cases.append((i[keep], j[keep]))
    # a spiral: one long thin component, many union-find levels
    m = np.zeros((41, 41), dtype=bool)
    x = y = 20
    m[x, y] = True
    step, d = 2, 0
    dirs = [(0, 1), (1, 0), (0, -1), (-1, 0)]
    while True:
        dx, dy = dirs[d % 4]
        ok = True
        for _ in range(step):
            x, y = x + dx, y + dy
            if not (0 <= x < 41 and 0 <= y < 41):
                ok = False
                break
            m[x, y] = True
        if not ok:
            break
        d += 1
        if d % 2 == 0:
            step += 2
    cases.append(np.nonzero(m))
    cases.append((np.array([3, 7, 7]), np.array([5, 7, 9])))      # singletons
    cases.append((np.array([], dtype=int), np.array([], dtype=int)))
    for row, col in cases:
        n = int(max(row.max(), col.max()) + 1) if len(row) else 1
        coo = sparse.coo_matrix((np.ones(len(row), dtype=bool), (row, col)),
                                shape=(n, n))
        out['find_clusters'].append(dict(
            row=[int(v) for v in row], col=[int(v) for v in col],
            clusters=as_lists(find_clusters(coo))))
    with open(os.path.join(HERE, 'ref_clusters.json'), 'w') as h:
        json.dump(out, h)
    print('wrote ref_clusters.json: %d pipeline cases, %d pixel sets (%s '
          'clusters)' % (len(out['pipeline']), len(out['find_clusters']),
                         [len(c['clusters']) for c in out['find_clusters']]))


if __name__ == '__main__':
    main()
