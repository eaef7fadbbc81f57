"""
Full-size parity fixture for BASELINE.json configs[0] (chr18 + chr19,
mouse-sized at 10 kb, 2-vs-2, dist cap 200 bins: ~2.8 M union pixels), recorded
by running the UNMODIFIED reference (/root/reference through oracle/refrun.py)
in the build container.

    python tests/golden/make_golden_config1.py

The inputs are regenerated on the GPU box by the same seeded generator
(hic3defdr_b200.synth.write_dataset, config=1); the fixture holds their
checksums so that generator drift is detected instead of being reported as a
parity failure.  Recorded (ref_config1.npz, < 10 MB):
  * sha256 of row / col / raw / disp_idx per chromosome (bit-exact stages),
  * the (D + 1, R) size-factor table per chromosome (size_factors is a pure
    function of distance) and sha256 of size_factors / scaled,
  * disp_per_dist (D + 1, C), the fitted trends on the integer distances,
  * a seeded sample of 50 000 tested pixels per chromosome with every
    per-pixel output (disp, mu_hat_null, mu_hat_alt, llr, pvalues, qvalues),
  * the number of pixels with q < 0.01 / 0.05 / 0.2 and the sum of q,
  * ``disp_selfnoise`` (D + 1, C): the reference's OWN reproducibility of
    disp_per_dist -- the largest relative change of its qcml() result over
    four random permutations of the pixel order inside the bin
    (``--selfnoise``; the summation order of the likelihood changes Brent's
    path).  Where the likelihood is flat (far distances, few counts) this
    reaches 1e-6; the parity bar of a bin is max(1e-7, 3 x its self-noise).
"""
import hashlib
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import scipy.sparse as sparse

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402
from hic3defdr_b200.synth import MM10_10KB, write_dataset  # noqa: E402

CHROMS = {c: MM10_10KB[c] for c in ('chr18', 'chr19')}
DIST_MAX = 200
N_SAMPLE = 50000


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def input_checksums(kw, chroms):
    out = {}
    for c in chroms:
        h = hashlib.sha256()
        for pat in kw['raw_npz_patterns']:
            m = sparse.load_npz(pat.replace('<chrom>', c)).tocsr()
            for a in (m.indptr, m.indices, m.data):
                h.update(np.ascontiguousarray(a).tobytes())
        for pat in kw['bias_patterns']:
            h.update(np.loadtxt(pat.replace('<chrom>', c)).tobytes())
        out[c] = h.hexdigest()
    return out


def sample_index(n_d, chrom_index):
    rng = np.random.default_rng(777 + chrom_index)
    return np.sort(rng.choice(n_d, size=min(N_SAMPLE, n_d), replace=False))


_SN = {}


def _selfnoise_task(task):
    import warnings
    from hic3defdr.util.dispersion import qcml as ref_qcml
    d, c = task
    sel = _SN['dist'] == d
    if not sel.any():
        return 0.0
    reps = _SN['design'][:, c]
    x, ff = _SN['raw'][sel][:, reps], _SN['f'][sel][:, reps]
    rng = np.random.default_rng(100000 + 2 * d + c)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        base = ref_qcml(x, f=ff.copy())
        assert base == _SN['dpd'][d, c], (d, c, base, _SN['dpd'][d, c])
        worst = 0.0
        for _ in range(4):
            perm = rng.permutation(len(x))
            v = ref_qcml(x[perm], f=ff[perm].copy())
            worst = max(worst, abs(v - base) / base)
    return worst


def selfnoise():
    """adds ``disp_selfnoise`` to the existing fixture (the reference's qcml on
    every bin, 1 + 4 times; ~15 min on 8 cores)"""
    import multiprocessing as mp
    from oracle import pipeline as op
    from hic3defdr_b200.synth import BASE_SEED, make_chrom
    refrun.install()
    path = os.path.join(HERE, 'ref_config1.npz')
    g = dict(np.load(path))
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    raws, fs, dists = [], [], []
    for ci, c in enumerate(CHROMS):
        mats, bias, _ = make_chrom(CHROMS[c], 4, DIST_MAX,
                                   BASE_SEED + 1000 * 1 + 100 * ci, amp=300.0)
        st = op.prepare_chrom(mats, bias, design, dist_max=DIST_MAX)
        di = st['disp_idx']
        row, col = st['row'][di], st['col'][di]
        b = op.filter_bias(bias, 0.1)
        fs.append(op.combined_factor(b, row, col, st['size_factors'][di]))
        raws.append(st['raw'][di])
        dists.append(col - row)
    _SN.update(raw=np.concatenate(raws), f=np.concatenate(fs),
               dist=np.concatenate(dists), design=design,
               dpd=g['disp_per_dist'])
    tasks = [(d, c) for d in range(DIST_MAX + 1) for c in range(2)]
    with mp.get_context('fork').Pool(os.cpu_count()) as pool:
        res = pool.map(_selfnoise_task, tasks, chunksize=4)
    g['disp_selfnoise'] = np.array(res).reshape(DIST_MAX + 1, 2)
    np.savez_compressed(path, **g)
    sn = g['disp_selfnoise']
    print('disp_selfnoise: median %.1e, 95%% %.1e, max %.1e'
          % (np.median(sn), np.quantile(sn, 0.95), sn.max()))


def main():
    if '--selfnoise' in sys.argv:
        return selfnoise()
    Ref = refrun.reference_class()
    root = tempfile.mkdtemp(dir=os.environ.get('TMPDIR', '/tmp'))
    kw = write_dataset(root, CHROMS, n_reps=4, dist_max=DIST_MAX, config=1,
                       amp=300.0)
    kw.pop('loop_patterns')
    out = {}
    for c, s in input_checksums(kw, CHROMS).items():
        out['input_sha_%s' % c] = np.array(s)
    outdir = os.path.join(root, 'out')
    h = Ref(outdir=outdir, dist_thresh_max=DIST_MAX, **kw)
    t0 = time.perf_counter()
    h.run_to_qvalues(n_threads=int(os.environ.get('H3D_REF_THREADS', '-1')),
                     verbose=False)
    out['reference_seconds'] = np.array(time.perf_counter() - t0)
    out['reference_cores'] = np.array(os.cpu_count())
    ld = lambda n, c: np.load(os.path.join(outdir, '%s_%s.npy' % (n, c)))
    q_all = []
    for ci, c in enumerate(CHROMS):
        row, col = ld('row', c), ld('col', c)
        di = ld('disp_idx', c)
        for name in ('row', 'col', 'raw', 'disp_idx', 'size_factors',
                     'scaled'):
            out['sha_%s_%s' % (name, c)] = np.array(sha(ld(name, c)))
        out['n_%s' % c] = np.array([len(row), int(di.sum())])
        sf = ld('size_factors', c)
        dist = col - row
        first = np.full(DIST_MAX + 1, -1, dtype=np.int64)
        first[dist[::-1]] = np.arange(len(dist))[::-1]
        assert (first >= 0).all()
        table = sf[first]
        assert np.array_equal(table[dist], sf)      # pure function of distance
        out['sf_table_%s' % c] = table
        idx = sample_index(int(di.sum()), ci)
        out['sample_%s' % c] = idx
        for name in ('disp', 'mu_hat_null', 'mu_hat_alt', 'llr', 'pvalues',
                     'qvalues'):
            out['%s_%s' % (name, c)] = ld(name, c)[idx]
        out['sample_raw_%s' % c] = ld('raw', c)[di][idx]
        out['sample_scaled_%s' % c] = ld('scaled', c)[di][idx]
        q_all.append(ld('qvalues', c))
    q = np.concatenate(q_all)
    out['n_sig'] = np.array([int((q < t).sum()) for t in (0.01, 0.05, 0.2)])
    out['q_sum'] = np.array(q.sum())
    out['disp_per_dist'] = np.load(os.path.join(outdir, 'disp_per_dist.npy'))
    xs = np.arange(DIST_MAX + 1, dtype=float)
    for cond in ('A', 'B'):
        out['disp_fn_%s' % cond] = h.load_disp_fn(cond)(xs.copy())
    np.savez_compressed(os.path.join(HERE, 'ref_config1.npz'), **out)
    shutil.rmtree(root)
    print('ref_config1.npz written; reference run_to_qvalues took %.1f s on '
          '%d cores' % (float(out['reference_seconds']), os.cpu_count()))


if __name__ == '__main__':
    main()
