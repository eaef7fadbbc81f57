"""GPU parity, end to end: the drop-in ``HiC3DeFDR`` class on files against
the outputs recorded from the real reference (tests/golden/ref_pipeline.npz)."""
import json
import os

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sparse

from tests.helpers import load_pipeline_golden

pytestmark = pytest.mark.gpu


def write_golden_inputs(root, gold):
    reps = ['A1', 'A2', 'B1', 'B2']
    for ci, c in enumerate(gold['chroms']):
        mats, bias = gold['inputs'][ci]
        for r, rep in enumerate(reps):
            os.makedirs(os.path.join(root, rep), exist_ok=True)
            sparse.save_npz(os.path.join(root, rep, '%s_raw.npz' % c), mats[r])
            np.savetxt(os.path.join(root, rep, '%s_kr.bias' % c), bias[:, r])
        os.makedirs(os.path.join(root, 'clusters'), exist_ok=True)
        with open(os.path.join(root, 'clusters', 'loops_%s.json' % c), 'w') as h:
            json.dump([[list(map(int, p)) for p in cl]
                       for cl in gold['loops'][ci]], h)
    design = pd.DataFrame(gold['design'], index=reps, columns=['A', 'B'])
    return dict(
        raw_npz_patterns=[os.path.join(root, r, '<chrom>_raw.npz') for r in reps],
        bias_patterns=[os.path.join(root, r, '<chrom>_kr.bias') for r in reps],
        chroms=gold['chroms'], design=design,
        loop_patterns={'A': os.path.join(root, 'clusters', 'loops_<chrom>.json'),
                       'B': os.path.join(root, 'clusters', 'loops_<chrom>.json')})


@pytest.fixture(scope='module')
def run(tmp_path_factory):
    from hic3defdr_b200 import HiC3DeFDR
    gold = load_pipeline_golden()
    root = str(tmp_path_factory.mktemp('golden_inputs'))
    kw = write_golden_inputs(root, gold)
    outdir = os.path.join(root, 'out')
    h = HiC3DeFDR(outdir=outdir, dist_thresh_max=gold['dist_max'], **kw)
    h.run_to_qvalues(n_threads=0)
    return gold, h, outdir


def test_files_dtypes_and_indices(run):
    gold, h, outdir = run
    g = gold['g']
    for c in gold['chroms']:
        for name in ('row', 'col', 'raw', 'disp_idx', 'loop_idx'):
            got = np.load(os.path.join(outdir, '%s_%s.npy' % (name, c)))
            want = g['%s_%s' % (name, c)]
            assert got.dtype == want.dtype and got.shape == want.shape, name
            np.testing.assert_array_equal(got, want)
        for name in ('size_factors', 'scaled'):
            got = np.load(os.path.join(outdir, '%s_%s.npy' % (name, c)))
            np.testing.assert_allclose(got, g['%s_%s' % (name, c)], rtol=1e-12)
        for name in ('disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
                     'qvalues'):
            got = np.load(os.path.join(outdir, '%s_%s.npy' % (name, c)))
            want = g['%s_%s' % (name, c)]
            assert got.dtype == want.dtype and got.shape == want.shape, name


def test_dispersion_and_trend(run):
    gold, h, outdir = run
    g = gold['g']
    got = np.load(os.path.join(outdir, 'disp_per_dist.npy'))
    want = g['disp_per_dist']
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = np.isfinite(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-7)
    for cond in 'AB':
        fn = h.load_disp_fn(cond)
        np.testing.assert_allclose(fn(g['disp_fn_x']), g['disp_fn_%s' % cond],
                                   rtol=1e-6)
    for c in gold['chroms']:
        np.testing.assert_allclose(
            np.load(os.path.join(outdir, 'disp_%s.npy' % c)),
            g['disp_%s' % c], rtol=1e-6)


def test_pvalues_qvalues_end_to_end(run):
    """end to end the dispersion carries ~1e-8 of optimiser noise, so the
    bar is 1e-6 relative; p near 1 is compared
    through llr (absolute)."""
    gold, h, outdir = run
    g = gold['g']
    for c in gold['chroms']:
        ld = lambda n: np.load(os.path.join(outdir, '%s_%s.npy' % (n, c)))
        np.testing.assert_allclose(ld('mu_hat_null'), g['mu_hat_null_%s' % c],
                                   rtol=1e-6)
        np.testing.assert_allclose(ld('mu_hat_alt'), g['mu_hat_alt_%s' % c],
                                   rtol=1e-6)
        np.testing.assert_allclose(ld('llr'), g['llr_%s' % c], rtol=1e-5,
                                   atol=1e-9)
        ok = -2 * g['llr_%s' % c] >= 1e-6
        np.testing.assert_allclose(ld('pvalues')[ok], g['pvalues_%s' % c][ok],
                                   rtol=1e-5)
        q, wq = ld('qvalues'), g['qvalues_%s' % c]
        np.testing.assert_allclose(q, wq, rtol=1e-5, atol=1e-12)
        for fdr in (0.05, 0.2, 0.5):
            near = np.abs(wq - fdr) <= 1e-6 * fdr
            assert np.array_equal((q < fdr)[~near], (wq < fdr)[~near])


def test_steps_work_from_disk_alone(run):
    """checkpoint/resume contract: a fresh object re-runs lrt + bh from the
    files only (analysis/core.py:15-33)."""
    from hic3defdr_b200 import HiC3DeFDR
    gold, h, outdir = run
    before = {c: np.load(os.path.join(outdir, 'qvalues_%s.npy' % c))
              for c in gold['chroms']}
    pb = {c: np.load(os.path.join(outdir, 'pvalues_%s.npy' % c))
          for c in gold['chroms']}
    h2 = HiC3DeFDR.load(outdir)
    assert h2.dist_thresh_max == gold['dist_max']
    h2.lrt(n_threads=0)
    h2.bh()
    for c in gold['chroms']:
        np.testing.assert_array_equal(
            np.load(os.path.join(outdir, 'pvalues_%s.npy' % c)), pb[c])
        np.testing.assert_array_equal(
            np.load(os.path.join(outdir, 'qvalues_%s.npy' % c)), before[c])
    h3 = HiC3DeFDR.load(outdir)
    h3.estimate_disp(n_threads=0)
    np.testing.assert_array_equal(
        np.load(os.path.join(outdir, 'disp_per_dist.npy')),
        np.load(os.path.join(outdir, 'disp_per_dist.npy')))


def test_load_data_semantics(run):
    gold, h, outdir = run
    c = gold['chroms'][0]
    di = h.load_data('disp_idx', c)
    row = h.load_data('row', c, idx=di)
    assert len(row) == di.sum()
    allp, offs = h.load_data('pvalues', 'all')
    assert offs[0] == 0 and offs[-1] == len(allp)
    r, cc, q = h.load_data('qvalues', c, coo=True)
    assert len(r) == len(q)
    a1 = h.load_data('raw', c, rep='A1')
    assert a1.ndim == 1
