"""CPU tests: the oracle restatement (oracle/pipeline.py) against the
reference's own doctest vectors and against outputs recorded from the real
reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import scipy.sparse as sparse

from oracle import pipeline as op
from tests.helpers import load_kats, load_pipeline_golden, load_stage_golden


def test_kat_fit_mu_hat():
    for case in load_kats()['fit_mu_hat']:
        got = op.fit_mu_hat(np.array(case['x']), np.array(case['b']),
                            np.array(case['alpha']))
        np.testing.assert_allclose(got, case['doc'], rtol=0, atol=5e-9)
        np.testing.assert_allclose(got, case['full'], rtol=1e-12)


def test_kat_conditional_mor():
    k = load_kats()['conditional_mor']
    got = op.conditional_size_factors(np.array(k['data']),
                                      np.array(k['dist']), None)
    np.testing.assert_allclose(got, k['doc'], rtol=0, atol=5e-9)
    np.testing.assert_allclose(got, k['full'], rtol=1e-14)


def test_kat_sparse_union():
    k = load_kats()['sparse_union']
    mats = [sparse.csr_matrix(np.array(k['rep1'])),
            sparse.csr_matrix(np.array(k['rep2']))]
    row, col = op.union_pixels(mats, k['dist_thresh'])
    assert row.dtype == np.int32 and col.dtype == np.int32
    assert list(zip(row.tolist(), col.tolist())) == \
        [tuple(p) for p in k['pixels']]
    assert (col - row).tolist() == k['dist']
    data = np.stack([np.asarray(m[row, col]).ravel() for m in mats], axis=1)
    np.testing.assert_array_equal(data, np.array(k['data']))


@pytest.fixture(scope='module')
def oracle_run():
    gold = load_pipeline_golden()
    res = op.run_to_qvalues(gold['inputs'], gold['design'],
                            dist_min=gold['dist_min'],
                            dist_max=gold['dist_max'], loops=gold['loops'])
    return gold, res


def test_pipeline_prepare_matches_reference(oracle_run):
    gold, res = oracle_run
    g = gold['g']
    for c, st in zip(gold['chroms'], res['chroms']):
        for name in ('row', 'col', 'raw', 'disp_idx', 'loop_idx'):
            np.testing.assert_array_equal(st[name], g['%s_%s' % (name, c)])
        assert st['row'].dtype == g['row_%s' % c].dtype
        for name in ('size_factors', 'scaled'):
            np.testing.assert_array_equal(st[name], g['%s_%s' % (name, c)])


def test_pipeline_disp_lrt_bh_match_reference(oracle_run):
    gold, res = oracle_run
    g = gold['g']
    np.testing.assert_array_equal(res['disp_per_dist'], g['disp_per_dist'])
    for c, st in zip(gold['chroms'], res['chroms']):
        for name in ('disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
                     'qvalues'):
            np.testing.assert_array_equal(st[name], g['%s_%s' % (name, c)])
    for ci, cond in enumerate('AB'):
        np.testing.assert_array_equal(
            op.eval_trend(res['fits'][ci], g['disp_fn_x']),
            g['disp_fn_%s' % cond])


def test_stage_goldens():
    s = load_stage_golden()
    x, f = s['bin_x'], s['bin_f']
    for a in (0.01, 0.2, 1e-3):
        np.testing.assert_array_equal(op.fit_mu_hat(x, f, a),
                                      s['mu_hat_%g' % a])
        np.testing.assert_array_equal(op.equalize(x, f.copy(), a),
                                      s['equalize_%g' % a])
    pseudo = s['equalize_0.01']
    assert op.cml(pseudo) == float(s['cml_pseudo'])
    np.testing.assert_array_equal(
        [op.cml_nll(pseudo, t) for t in s['nll_deltas']], s['nll_values'])
    assert op.qcml(x, f=f.copy()) == float(s['qcml'])
    assert op.qcml(s['bin4_x'], f=s['bin4_f'].copy()) == float(s['qcml4'])
    np.testing.assert_array_equal(
        op.equalize(s['bin4_x'], s['bin4_f'].copy(), 0.05),
        s['equalize4_0.05'])


def test_stage_trend():
    s = load_stage_golden()
    xs, ys, xq = s['trend_x'], s['trend_y'], s['trend_q']
    fit = op.weighted_trend(xs, ys, left_boundary=ys[0])
    np.testing.assert_array_equal(op.eval_trend(fit, xq), s['trend_weighted'])
    fit = op.weighted_trend(xs, ys, left_boundary=ys[0], frac=0.2)
    np.testing.assert_array_equal(op.eval_trend(fit, xq),
                                  s['trend_weighted_frac0.2'])
    fit = op.weighted_trend(xs, ys, left_boundary=ys[0], weighted=False)
    np.testing.assert_array_equal(op.eval_trend(fit, xq), s['trend_plain'])


def test_stage_lrt():
    s = load_stage_golden()
    design = s['lrt_design'].astype(bool)
    wide = np.dot(s['lrt_disp'], design.T.astype(float))
    for tag, refit in (('refit', True), ('norefit', False)):
        p, llr, mu0, mu1 = op.lrt(s['lrt_x'], s['lrt_f'], wide, design, refit)
        np.testing.assert_array_equal(p, s['lrt_%s_p' % tag])
        np.testing.assert_array_equal(llr, s['lrt_%s_llr' % tag])
        np.testing.assert_array_equal(mu0, s['lrt_%s_mu0' % tag])
        np.testing.assert_array_equal(mu1, s['lrt_%s_mu1' % tag])


def test_bh_properties():
    rng = np.random.default_rng(5)
    p = rng.random(1000) ** 3
    p[::97] = np.nan
    q = op.bh(p)
    assert np.isnan(q[::97]).all()
    fin = np.isfinite(p)
    assert (q[fin] >= p[fin]).all() and (q[fin] <= 1).all()
    order = np.argsort(p[fin])
    assert (np.diff(q[fin][order]) >= 0).all()
    # brute force definition
    pf = p[fin]
    n = pf.size
    brute = np.array([min(1.0, min(pj * n / (np.sum(pf <= pj))
                                   for pj in pf[pf >= pi])) for pi in pf])
    np.testing.assert_allclose(q[fin], brute, rtol=1e-14)


def test_bh_against_scipy_false_discovery_control():
    """``adjust_pvalues`` is third-party code absent from /root/reference
    (lib5c -> statsmodels fdr_bh; "parity unpinned" in DESIGN.md).  An
    independent published implementation of the same procedure,
    scipy.stats.false_discovery_control(method='bh'), agrees with the
    restatement to round-off -- including ties, p = 0 and p = 1."""
    from scipy import stats
    rng = np.random.default_rng(11)
    for n in (1, 2, 17, 1000, 50_000):
        p = rng.random(n) ** 2
        if n > 10:
            p[::7] = p[3]            # ties
            p[1], p[2] = 0.0, 1.0
        np.testing.assert_allclose(
            op.bh(p), stats.false_discovery_control(p, method='bh'),
            rtol=1e-13, atol=0)
