"""Shared loaders for the golden fixtures (tests only)."""
import json
import os

import numpy as np
import scipy.sparse as sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_kats():
    with open(os.path.join(GOLDEN, 'reference_kats.json')) as h:
        return json.load(h)


def load_pipeline_golden():
    g = np.load(os.path.join(GOLDEN, 'ref_pipeline.npz'))
    dist_max, n_reps, dist_min = [int(v) for v in g['meta']]
    chroms = sorted({k.split('_')[1] for k in g.files if k.startswith('in_')
                     and k.endswith('_bias')})
    inputs, loops = [], []
    for c in chroms:
        bias = g['in_%s_bias' % c]
        n = bias.shape[0]
        mats = [sparse.csr_matrix((g['in_%s_data_%d' % (c, r)],
                                   g['in_%s_indices_%d' % (c, r)],
                                   g['in_%s_indptr_%d' % (c, r)]),
                                  shape=(n, n)) for r in range(n_reps)]
        inputs.append((mats, bias))
        lp = g['in_%s_loops' % c]
        loops.append([[tuple(p) for p in lp]])
    return dict(g=g, chroms=chroms, inputs=inputs, loops=loops,
                dist_max=dist_max, dist_min=dist_min,
                design=g['design'].astype(bool))


def load_stage_golden():
    return np.load(os.path.join(GOLDEN, 'ref_stages.npz'))


def check_pvalues(got_p, got_llr, p, llr, good, df):
    """p = chi2(df).sf(-2 llr) at the north_star tolerance of 1e-9, split so
    that it is not at the mercy of round-off in llr:
      * the survival function itself, on OUR llr, everywhere: rtol 1e-9;
      * against the oracle's p-values on ``good`` (-2 llr above the floor):
        rtol 1e-9 plus what an absolute llr difference of 2e-12 moves p by
        (|dp/dllr| = 2 pdf(-2 llr)).  llr is a difference of log-likelihoods of
        order 1e2..1e3, so both codes carry ~1e-13 of cancellation noise in it;
        near llr = 0 with one degree of freedom p = erfc(sqrt(-llr)) turns that
        into a p difference of up to ~1e-9 that no implementation can avoid."""
    from scipy import stats
    x_got = np.maximum(-2.0 * got_llr, 0.0)
    np.testing.assert_allclose(got_p, stats.chi2(df).sf(x_got), rtol=1e-9,
                               atol=1e-300)
    x = -2.0 * llr[good]
    slack = 2.0 * stats.chi2(df).pdf(np.maximum(x, 1e-300)) * 2e-12
    diff = np.abs(got_p[good] - p[good])
    bad = diff > 1e-9 * p[good] + slack
    assert not bad.any(), (int(bad.sum()), float(diff[bad].max()))
