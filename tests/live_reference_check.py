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
        n_px = sum(len(st['row']) for st in res['chroms'])
        print('live reference check %s: ok (%d arrays bit-identical, %d union '
              'pixels)' % (variant, n_checked + 1, n_px))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main(sys.argv[1])
