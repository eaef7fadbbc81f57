"""Runs the UNMODIFIED reference (through oracle/refrun.py) and the oracle
restatement (oracle/pipeline.py) on a freshly generated small dataset and
compares every saved stage bit for bit.  A separate process: importing the
reference installs import stubs for its plotting dependencies.

    python tests/live_reference_check.py <variant>

Used by tests/test_oracle_live_reference.py (skipped where no reference tree is
available); the committed fixtures of tests/golden/ pin the same comparison on
recorded runs."""
import json
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np
import scipy.sparse as sparse

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

# Only the conditional norms: with 'median_of_ratios' / 'simple_scaling' the
# reference's own estimate_disp() raises (analysis/analysis.py:181 indexes the
# (R,) size factors with the per-pixel disp_idx mask); only estimator='qcml':
# 'cml' and 'mme' divide the integer counts in place and raise
# (util/dispersion.py:70, :104); dist_thresh_max >= ~30: the rolling window of
# weighted_lowess_fit (20 points) needs that many distances.
VARIANTS = {
    # name: (dataset kwargs, run_to_qvalues kwargs)
    'default_loops': (dict(sizes={'c1': 210, 'c2': 160}, n_reps=4, dist_max=30,
                           config=21, amp=140.0, loops=True), {}),
    'three_vs_three': (
        dict(sizes={'c1': 190, 'c2': 150}, n_reps=6, dist_max=32, config=22,
             amp=160.0, loops=False),
        dict(n_bins_norm=5)),
    'conditional_scaling_plain_lowess_no_refit': (
        dict(sizes={'c1': 230}, n_reps=4, dist_max=28, config=23, amp=120.0,
             loops=False),
        dict(norm='conditional_scaling', weighted_lowess=False,
             refit_mu=False)),
    'three_chroms_fixed_frac': (
        dict(sizes={'c1': 170, 'c2': 170, 'c3': 120}, n_reps=4, dist_max=34,
             config=24, amp=200.0, loops=False),
        dict(norm='conditional_scaling', frac=0.5, n_bins_norm=6)),
}


def compare_loaders(ref, outdir, chroms, has_loops):
    """The drop-in class opened on the directory the REFERENCE wrote
    (``HiC3DeFDR.load``: its pickle, its files): ``load_data`` must return what
    the reference's own ``load_data`` returns (analysis/core.py:62-195) for
    every way of calling it -- per chromosome, genome-wide with offsets,
    masks, chained masks, one column by replicate / condition, COO triples,
    and the same exceptions."""
    from hic3defdr_b200 import HiC3DeFDR
    ours = HiC3DeFDR.load(outdir)
    assert ours.chroms == ref.chroms and ours.design.equals(ref.design)
    rng = np.random.default_rng(5)
    n = [0]

    def same(a, b):
        if isinstance(b, tuple):
            assert isinstance(a, tuple) and len(a) == len(b)
            for x, y in zip(a, b):
                same(x, y)
            return
        a, b = np.asarray(a), np.asarray(b)
        assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype,
                                                           a.shape, b.shape)
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == 'f')

    def check(*args, **kw):
        try:
            want = ref.load_data(*args, **kw)
        except Exception as e:            # same failure, same type
            try:
                ours.load_data(*args, **kw)
            except Exception as e2:
                assert type(e2) is type(e), (args, kw, e, e2)
                n[0] += 1
                return
            raise AssertionError('reference raised %r, we did not: %r %r'
                                 % (e, args, kw))
        same(ours.load_data(*args, **kw), want)
        n[0] += 1

    rep, cond = ref.design.index[1], ref.design.columns[1]
    union = ['row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx']
    tested = ['disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt'] + \
        (['loop_idx'] if has_loops else [])
    check('disp_per_dist')
    for name in union + tested + ['qvalues']:
        check(name, 'all')
        for c in chroms:
            check(name, c)
            # (qvalues as COO without loops runs into the reference's broken
            # loop_idx shortcut, see below)
            if name not in ('row', 'col') and (has_loops or name != 'qvalues'):
                check(name, c, coo=True)
        if name in ('row', 'col'):
            check(name, chroms[0], coo=True)      # ValueError
    for name in ('raw', 'scaled', 'size_factors'):
        check(name, chroms[0], rep=rep)
        check(name, 'all', rep=rep)
        check(name, chroms[-1], rep=rep, coo=True)
    for name in ('disp', 'mu_hat_alt'):
        check(name, chroms[0], cond=cond)
        check(name, 'all', cond=cond)
        check(name, chroms[0], cond=cond, coo=True)
    # masks: per chromosome, genome-wide, chained
    disp_all, offs = ref.load_data('disp_idx', 'all')
    for c in chroms:
        di = ref.load_data('disp_idx', c)
        check('row', c, idx=di)
        check('raw', c, idx=di, rep=rep)
        sub = rng.random(int(di.sum())) < 0.4
        check('col', c, idx=(di, sub))
        check('scaled', c, idx=(di, sub), rep=rep)
        check('pvalues', c, idx=sub)
    check('row', 'all', idx=disp_all)
    check('raw', 'all', idx=disp_all, rep=rep)
    sub = rng.random(int(disp_all.sum())) < 0.3
    check('col', 'all', idx=(disp_all, sub))
    check('llr', 'all', idx=sub)
    check('disp', 'all', idx=sub, cond=cond)
    if has_loops:
        loop_all, _ = ref.load_data('loop_idx', 'all')
        check('row', 'all', idx=(disp_all, loop_all))
        for c in chroms:
            check('row', c, idx=(ref.load_data('disp_idx', c),
                                 ref.load_data('loop_idx', c)))
    # failures
    check('raw', 'all', coo=True)
    check('raw', chroms[0], idx=ref.load_data('disp_idx', chroms[0]), coo=True)
    check('bias', chroms[0], coo=True)
    check('no_such_name', chroms[0], coo=True)
    check('no_such_name', chroms[0])
    check('raw', chroms[0], rep='no_such_rep')
    if not has_loops:
        # the reference's shortcut for runs without loops calls a function
        # that does not exist (core.py:107 ``np.load_data``); the intended
        # result is one True per tested pixel
        for c in chroms:
            got = ours.load_data('loop_idx', c)
            assert got.dtype == bool and got.all() and \
                len(got) == int(ref.load_data('disp_idx', c).sum())
            same(ours.load_data('qvalues', c, coo=True)[:2],
                 (ref.load_data('row', c, idx=ref.load_data('disp_idx', c)),
                  ref.load_data('col', c, idx=ref.load_data('disp_idx', c))))
    return n[0]


def reference_opens_our_directory(Ref, ref, kw, res, dist_max, root):
    """The other direction of the file contract: a directory made by the
    drop-in class (its pickle, its ``disp_fn_<cond>.pickle``) opens in the
    REFERENCE class (analysis/core.py:15-33, 220-237), and the product's
    dispersion callable (hic3defdr_b200/trend.py::DispersionTrend, here built
    from the oracle's fit = the reference's lowess curve bit for bit) evaluates
    to what the reference's own pickled closure gives, in and out of range."""
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.trend import DispersionTrend
    outdir = os.path.join(root, 'ours')
    ours = HiC3DeFDR(outdir=outdir, dist_thresh_max=dist_max, **kw)
    xs = np.concatenate([np.arange(dist_max + 1), [dist_max + 7]]), \
        np.array([-2.0, 0.0, 0.5, 3.99, 4.0, 4.5, 7.25, dist_max - 0.5,
                  dist_max + 3.5])
    for ci, cond in enumerate(ours.design.columns):
        fit = res['fits'][ci]
        ours.save_disp_fn(cond, DispersionTrend(
            fit['x'], fit['y'], fit['inc'], fit['cx'], fit['cy'], fit['frac'],
            fit['left_boundary'], None, weighted=fit['weighted']))
    theirs = Ref.load(outdir)
    for attr in ('raw_npz_patterns', 'bias_patterns', 'chroms',
                 'dist_thresh_min', 'dist_thresh_max', 'bias_thresh',
                 'mean_thresh', 'loop_patterns', 'res'):
        assert getattr(theirs, attr) == getattr(ref, attr), attr
    assert theirs.design.equals(ref.design)
    for cond in ours.design.columns:
        fn, want = theirs.load_disp_fn(cond), ref.load_disp_fn(cond)
        for x in xs:               # integer and float arguments
            got = fn(x.copy())
            assert got.dtype == np.float64 and \
                np.array_equal(got, want(x.copy()), equal_nan=True), cond


def main(variant):
    from oracle import pipeline as op
    from oracle import refrun
    from hic3defdr_b200.synth import write_dataset
    data_kw, run_kw = VARIANTS[variant]
    data_kw = dict(data_kw)
    sizes = data_kw.pop('sizes')
    dist_max = data_kw['dist_max']
    root = tempfile.mkdtemp(prefix='h3d_live_')
    try:
        kw = write_dataset(root, sizes, **data_kw)
        if kw['loop_patterns'] is None:
            kw.pop('loop_patterns')
        Ref = refrun.reference_class()
        outdir = os.path.join(root, 'out')
        h = Ref(outdir=outdir, dist_thresh_max=dist_max, **kw)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            h.run_to_qvalues(n_threads=0, verbose=False, **run_kw)
        inputs, loops = [], []
        for c in sizes:
            mats = [sparse.load_npz(p.replace('<chrom>', c)).tocsr()
                    for p in kw['raw_npz_patterns']]
            bias = np.array([np.loadtxt(p.replace('<chrom>', c))
                             for p in kw['bias_patterns']]).T
            inputs.append((mats, bias))
            if 'loop_patterns' in kw:
                with open(kw['loop_patterns']['A'].replace('<chrom>', c)) as f:
                    loops.append([[tuple(p) for p in cl]
                                  for cl in json.load(f)])
        okw = dict(run_kw)
        if 'n_bins_norm' in okw:
            okw['n_bins'] = okw.pop('n_bins_norm')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = op.run_to_qvalues(inputs, kw['design'].values,
                                    dist_max=dist_max,
                                    loops=loops if loops else None, **okw)
        names = ['row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx',
                 'disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
                 'qvalues'] + (['loop_idx'] if loops else [])
        n_checked = 0
        for c, st in zip(sizes, res['chroms']):
            for name in names:
                want = np.load(os.path.join(outdir, '%s_%s.npy' % (name, c)))
                got = np.asarray(st[name])
                assert got.shape == want.shape, (variant, c, name, got.shape,
                                                 want.shape)
                assert np.array_equal(got, want, equal_nan=got.dtype.kind == 'f'), \
                    (variant, c, name)
                n_checked += 1
        want = np.load(os.path.join(outdir, 'disp_per_dist.npy'))
        assert np.array_equal(res['disp_per_dist'], want, equal_nan=True)
        xs = np.concatenate([np.arange(dist_max + 1.0), [-1.0, 2.5, dist_max + 4.5]])
        for ci, cond in enumerate(kw['design'].columns):
            assert np.array_equal(op.eval_trend(res['fits'][ci], xs),
                                  h.load_disp_fn(cond)(xs.copy()), equal_nan=True)
        n_loader = compare_loaders(h, outdir, list(sizes), bool(loops))
        reference_opens_our_directory(Ref, h, kw, res, dist_max, root)
        n_px = sum(len(st['row']) for st in res['chroms'])
        print('live reference check %s: ok (%d arrays bit-identical, %d union '
              'pixels, %d load_data calls equal)'
              % (variant, n_checked + 1, n_px, n_loader))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main(sys.argv[1])
