"""GPU parity of ``evaluate``: the ROC / FDR
curves of csrc/roc.cu against outputs recorded from the unmodified reference
(tests/golden/make_golden_evaluate.py), against sklearn at a larger size, and
through the drop-in class on the files of a simulated-loop run.  All counts
are integers, so fpr / tpr / thresh / fdr must be EXACT."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', ['separated', 'ties', 'null', 'tiny'])
def test_evaluate_vs_recorded_reference(case):
    from hic3defdr_b200 import evaluation as hev
    g = np.load(os.path.join(GOLDEN, 'ref_evaluate.npz'))
    fdr, fpr, tpr, thresh = hev.evaluate(g['%s_y' % case], g['%s_q' % case])
    np.testing.assert_array_equal(thresh, g['%s_thresh' % case])
    np.testing.assert_array_equal(fpr, g['%s_fpr' % case])
    np.testing.assert_array_equal(tpr, g['%s_tpr' % case])
    np.testing.assert_array_equal(fdr, g['%s_fdr' % case])      # NaN == NaN here


def test_make_y_true_vs_recorded_reference():
    from hic3defdr_b200 import evaluation as hev
    g = np.load(os.path.join(GOLDEN, 'ref_evaluate.npz'))
    t = g['yt_clusters']
    clusters = [{(int(r), int(c)) for i, r, c in t if i == k}
                for k in range(int(t[:, 0].max()) + 1)]
    got = hev.make_y_true(g['yt_row'], g['yt_col'], clusters, g['yt_labels'])
    np.testing.assert_array_equal(got, g['yt_out'])


def test_roc_curve_vs_sklearn_large():
    """3 M q-values with BH-like runs of ties"""
    from sklearn.metrics import roc_curve
    from hic3defdr_b200 import evaluation as hev
    rng = np.random.default_rng(9)
    n = 3_000_000
    y = rng.random(n) < 0.02
    q = np.where(y, rng.beta(0.4, 3.0, n), rng.beta(1.5, 1.0, n))
    q[rng.random(n) < 0.3] = np.round(q[rng.random(n) < 0.3], 3)[0]
    fdr, fpr, tpr, thresh = hev.evaluate(y, q)
    wfpr, wtpr, wthr = roc_curve(y, 1 - q)
    np.testing.assert_array_equal(thresh, wthr)
    np.testing.assert_array_equal(fpr, wfpr)
    np.testing.assert_array_equal(tpr, wtpr)
    fin = np.flatnonzero(np.isfinite(fdr))
    i = int(fin[len(fin) // 2])
    assert fdr[i] == pytest.approx(hev.compute_fdr(y, (1 - q) >= thresh[i]),
                                   rel=1e-15)


def test_class_evaluate_on_simulated_loops(tmp_path):
    """run_to_qvalues on a data set with injected differential loops, then
    evaluate() from the files: eval.npz equals the oracle-side computation
    (sklearn on the same files) and the curve is far above chance."""
    from sklearn.metrics import roc_curve
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.analysis import load_clusters
    from hic3defdr_b200.synth import write_dataset
    root = str(tmp_path)
    kw = write_dataset(os.path.join(root, 'in'), {'cA': 900, 'cB': 700},
                       n_reps=4, dist_max=60, config=5, amp=400.0, loops=True)
    labels = os.path.join(root, 'in', 'clusters', 'labels_<chrom>.txt')
    outdir = os.path.join(root, 'out')
    h = HiC3DeFDR(outdir=outdir, dist_thresh_max=60, **kw)
    h.run_to_qvalues(n_threads=0)
    h.evaluate('A', labels)
    got = np.load(os.path.join(outdir, 'eval.npz'))
    ys, qs = [], []
    for c in kw['chroms']:
        row, col, q = h.load_data('qvalues', c, coo=True)
        clusters = load_clusters(kw['loop_patterns']['A'].replace('<chrom>', c))
        labs = np.loadtxt(labels.replace('<chrom>', c), dtype='U7')
        sig = set().union(*[cl for cl, lab in zip(clusters, labs)
                            if lab != 'constit'])
        ys.append(np.array([(r, cc) in sig for r, cc in zip(row, col)]))
        qs.append(q)
    y, q = np.concatenate(ys), np.concatenate(qs)
    wfpr, wtpr, wthr = roc_curve(y, 1 - q)
    np.testing.assert_array_equal(got['thresh'], wthr)
    np.testing.assert_array_equal(got['fpr'], wfpr)
    np.testing.assert_array_equal(got['tpr'], wtpr)
    auc = np.trapezoid(wtpr, wfpr)
    assert auc > 0.65, auc          # chance is 0.5
    # a distance-restricted evaluation with BH re-run writes its own file
    h.evaluate('A', labels, min_dist=15, max_dist=50, rerun_bh=True)
    assert os.path.isfile(os.path.join(outdir, 'eval_15_50.npz'))
