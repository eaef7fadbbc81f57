"""
Fixture for SURVEY.md section 8(f) row 4 (``evaluate``): outputs of the
UNMODIFIED reference's ``hic3defdr.util.evaluation.evaluate`` /
``make_y_true`` (hic3defdr/util/evaluation.py:15-79, on top of
sklearn.metrics.roc_curve) recorded in the build container.

    python tests/golden/make_golden_evaluate.py  ->  ref_evaluate.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402


def cases():
    rng = np.random.default_rng(31415)
    out = {}
    # well separated classes, continuous q
    n = 20000
    y = rng.random(n) < 0.1
    q = np.where(y, rng.beta(0.3, 4.0, n), rng.beta(2.0, 1.2, n))
    out['separated'] = (y, q)
    # heavy ties (q-values after BH come in runs) and the extremes 0 and 1
    n = 5000
    y = rng.random(n) < 0.3
    q = np.round(np.where(y, rng.beta(0.5, 2.0, n), rng.random(n)), 2)
    q[:10] = 0.0
    q[10:30] = 1.0
    out['ties'] = (y, q)
    # uninformative
    n = 3000
    out['null'] = (rng.random(n) < 0.5, rng.random(n))
    # tiny
    out['tiny'] = (np.array([True, False, True, False, False]),
                   np.array([0.01, 0.5, 0.5, 0.9, 0.02]))
    return out


def main():
    refrun.install()
    import warnings
    from hic3defdr.util.evaluation import evaluate, make_y_true
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for name, (y, q) in cases().items():
            fdr, fpr, tpr, thresh = evaluate(y, q)
            out['%s_y' % name], out['%s_q' % name] = y, q
            for k, v in zip(('fdr', 'fpr', 'tpr', 'thresh'),
                            (fdr, fpr, tpr, thresh)):
                out['%s_%s' % (name, k)] = v
            print(name, len(thresh), 'points,', int(np.isfinite(fdr).sum()),
                  'fdr values')
    rng = np.random.default_rng(5)
    row = rng.integers(0, 60, 400)
    col = row + rng.integers(0, 30, 400)
    clusters = [{(int(r), int(c)) for r, c in zip(row[i::40][:5], col[i::40][:5])}
                for i in range(12)]
    labels = np.array(['constit', 'A', 'B'] * 4, dtype='U7')
    out['yt_row'], out['yt_col'] = row, col
    out['yt_clusters'] = np.array(
        [[i, r, c] for i, cl in enumerate(clusters) for r, c in sorted(cl)])
    out['yt_labels'] = labels
    out['yt_out'] = make_y_true(row, col, clusters, labels)
    np.savez_compressed(os.path.join(HERE, 'ref_evaluate.npz'), **out)


if __name__ == '__main__':
    main()
