"""
Generates the golden fixtures in this directory by running the UNMODIFIED
reference (/root/reference, imported through oracle/refrun.py) in the build
container.  The reference cannot travel to the GPU box, the fixtures do.

    python tests/golden/make_golden.py

Outputs
  reference_kats.json   the reference's own doctest vectors
                        (util/scaled_nb.py:100-137, docs/median_of_ratios.rst,
                        docs/sparse_union.rst), re-checked against the live
                        reference before being written.
  ref_pipeline.npz      inputs + every saved stage of the real
                        ``HiC3DeFDR.run_to_qvalues(n_threads=0)`` on a small
                        synthetic two-chromosome dataset (stable equal_bin).
  ref_stages.npz        stage-isolated calls of the real reference functions
                        (fit_mu_hat, equalize, cml, qcml, weighted_lowess_fit,
                        lrt) on seeded random inputs.
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import scipy.sparse as sparse

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402
from hic3defdr_b200.synth import write_dataset  # noqa: E402

CHROMS = {'chrA': 320, 'chrB': 220}
DIST_MAX = 40
N_REPS = 4


def kats(ref):
    fit = ref.scaled_nb.fit_mu_hat
    x = np.array([[1, 2], [3, 4], [5, 6]])
    b = np.array([[0.9, 1.1], [0.8, 1.2], [0.7, 1.3]])
    a_full = np.array([[0.1, 0.2], [0.3, 0.4], [0.5, 0.6]])
    x2 = np.array([[2, 3, 4, 2], [6, 9, 3, 1]])
    b2 = np.array([[0.45, 0.53, 0.088, 0.091], [0.70, 0.83, 0.14, 0.15]])
    a2 = np.array([[0.0071, 0.0071, 0.0073, 0.0073],
                   [0.0070, 0.0070, 0.0072, 0.0072]])
    cases = [
        dict(x=x, b=b, alpha=a_full,
             doc=[1.47251127, 3.53879843, 5.86853465]),
        dict(x=x, b=b, alpha=np.array([0.1, 0.2]),
             doc=[1.47251127, 3.53749833, 5.85554075]),
        dict(x=x, b=b, alpha=np.array([0.1, 0.2, 0.3])[:, None],
             doc=[1.49544092, 3.51679438, 5.73129492]),
        dict(x=np.array([1, 2]), b=np.array([0.9, 1.1]),
             alpha=np.array([0.1, 0.2]), doc=[1.47251127]),
        dict(x=np.array([1, 2]), b=np.array([0.9, 1.1]), alpha=0.1,
             doc=[1.49544092]),
        dict(x=x2, b=b2, alpha=a2, doc=[9.5900971, 10.45962955]),
    ]
    out = []
    for c in cases:
        got = fit(c['x'], c['b'], c['alpha'], verbose=False)
        assert np.allclose(got, c['doc'], rtol=0, atol=5e-9), (got, c['doc'])
        out.append(dict(x=np.asarray(c['x']).tolist(),
                        b=np.asarray(c['b']).tolist(),
                        alpha=np.asarray(c['alpha']).tolist(),
                        doc=c['doc'], full=got.tolist()))
    # conditional_mor table
    data = np.arange(20, dtype=float).reshape((5, 4))
    dist = np.array([1, 1, 1, 2, 2])
    mor = ref.scaling.conditional_mor(data, dist)
    doc_mor = [[0.79394639, 0.93946738, 1.08498836, 1.23050934]] * 3 + \
        [[0.90390183, 0.96968472, 1.0354676, 1.10125049]] * 2
    assert np.allclose(mor, doc_mor, rtol=0, atol=5e-9)
    # sparse_union example
    rep1 = np.array([[0., 0., 3., 1.], [0., 6., 5., 0.], [0., 0., 0., 2.],
                     [0., 0., 0., 7.]])
    rep2 = np.array([[0., 1., 3., 2.], [0., 0., 0., 0.], [0., 0., 4., 2.],
                     [0., 0., 0., 3.]])
    tmp = tempfile.mkdtemp()
    names = []
    for i, m in enumerate((rep1, rep2)):
        names.append(os.path.join(tmp, 'rep%d.npz' % (i + 1)))
        sparse.save_npz(names[-1], sparse.csr_matrix(m))
    row, col = ref.matrices.sparse_union(names, dist_thresh=2)
    shutil.rmtree(tmp)
    doc_px = [(0, 1), (0, 2), (1, 1), (1, 2), (2, 2), (2, 3), (3, 3)]
    assert list(zip(row.tolist(), col.tolist())) == doc_px
    return dict(
        fit_mu_hat=out,
        conditional_mor=dict(data=data.tolist(), dist=dist.tolist(),
                             doc=doc_mor, full=mor.tolist()),
        sparse_union=dict(rep1=rep1.tolist(), rep2=rep2.tolist(),
                          dist_thresh=2, pixels=doc_px,
                          data=[[0., 1.], [3., 3.], [6., 0.], [5., 0.],
                                [0., 4.], [2., 2.], [7., 3.]],
                          dist=[1, 2, 0, 1, 0, 1, 0]))


def pipeline(Ref):
    root = tempfile.mkdtemp()
    kw = write_dataset(root, CHROMS, n_reps=N_REPS, dist_max=DIST_MAX,
                       config=7, amp=120.0, loops=True)
    out = {}
    for chrom in CHROMS:
        for r, pat in enumerate(kw['raw_npz_patterns']):
            m = sparse.load_npz(pat.replace('<chrom>', chrom)).tocsr()
            out['in_%s_indptr_%d' % (chrom, r)] = m.indptr
            out['in_%s_indices_%d' % (chrom, r)] = m.indices
            out['in_%s_data_%d' % (chrom, r)] = m.data
        out['in_%s_bias' % chrom] = np.array(
            [np.loadtxt(p.replace('<chrom>', chrom))
             for p in kw['bias_patterns']]).T
        import json as _json
        with open(kw['loop_patterns']['A'].replace('<chrom>', chrom)) as h:
            out['in_%s_loops' % chrom] = np.array(
                [p for c in _json.load(h) for p in c], dtype=np.int32)
    outdir = os.path.join(root, 'out')
    h = Ref(outdir=outdir, dist_thresh_max=DIST_MAX, **kw)
    h.run_to_qvalues(n_threads=0, verbose=False)
    stages = ['row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx',
              'loop_idx', 'disp', 'pvalues', 'llr', 'mu_hat_null',
              'mu_hat_alt', 'qvalues']
    for chrom in CHROMS:
        for s in stages:
            out['%s_%s' % (s, chrom)] = np.load(
                os.path.join(outdir, '%s_%s.npy' % (s, chrom)))
    out['disp_per_dist'] = np.load(os.path.join(outdir, 'disp_per_dist.npy'))
    xs = np.concatenate([np.arange(DIST_MAX + 1, dtype=float),
                         [-1.0, 0.5, 3.3, 4.5, 7.25, DIST_MAX + 3.5]])
    out['disp_fn_x'] = xs
    for cond in ('A', 'B'):
        out['disp_fn_%s' % cond] = h.load_disp_fn(cond)(xs.copy())
    out['design'] = kw['design'].values
    # the reference's own reproducibility floor for disp_per_dist: its qCML
    # result under random permutations of the pixel order inside each bin
    # (summation order changes the Brent path)
    import warnings
    from hic3defdr.util.dispersion import qcml as ref_qcml
    design = kw['design'].values.astype(bool)
    raws, fs, dists = [], [], []
    for chrom in CHROMS:
        di = out['disp_idx_%s' % chrom]
        bias = h.load_bias(chrom)
        r, c = out['row_%s' % chrom][di], out['col_%s' % chrom][di]
        raws.append(out['raw_%s' % chrom][di])
        fs.append(bias[r] * bias[c] * out['size_factors_%s' % chrom][di])
        dists.append(c - r)
    raw, f, dist = np.concatenate(raws), np.concatenate(fs), \
        np.concatenate(dists)
    rng = np.random.default_rng(2026)
    noise = np.zeros_like(out['disp_per_dist'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for d in range(DIST_MAX + 1):
            sel = dist == d
            if not sel.any():
                continue
            for ci in range(design.shape[1]):
                x, ff = raw[sel][:, design[:, ci]], f[sel][:, design[:, ci]]
                base = ref_qcml(x, f=ff.copy())
                assert base == out['disp_per_dist'][d, ci]
                for _ in range(4):
                    perm = rng.permutation(len(x))
                    v = ref_qcml(x[perm], f=ff[perm].copy())
                    noise[d, ci] = max(noise[d, ci], abs(v - base) / base)
    out['disp_selfnoise'] = noise
    out['meta'] = np.array([DIST_MAX, N_REPS, 4])
    shutil.rmtree(root)
    return out


def stages(ref):
    rng = np.random.default_rng(424242)
    out = {}
    # fit_mu_hat / equalize / cml / qcml on one pooled "distance bin"
    n = 1500
    mu = rng.gamma(2.0, 8.0, size=n) + 0.3
    f = rng.lognormal(0.0, 0.25, size=(n, 2))
    lam = rng.gamma(1 / 0.03, mu[:, None] * f * 0.03)
    x = rng.poisson(lam).astype(np.int64)
    x[:5] = [[0, 1], [1, 0], [0, 3], [200, 1], [0, 1]]
    keep = x.sum(axis=1) > 0
    x, f = x[keep], f[keep]
    out['bin_x'], out['bin_f'] = x, f
    for a in (0.01, 0.2, 1e-3):
        out['mu_hat_%g' % a] = ref.scaled_nb.fit_mu_hat(
            x, f, a, verbose=False)
        out['equalize_%g' % a] = ref.scaled_nb.equalize(x, f.copy(), a)
    pseudo = out['equalize_0.01']
    out['cml_pseudo'] = np.array(ref.dispersion.cml(pseudo.copy()))
    deltas = np.array([1e-4, 0.003, 0.0099, 0.05, 0.3, 0.9, 100. / 101])
    n_r = pseudo.shape[1]
    z = pseudo.sum(axis=1)
    from scipy.special import gammaln
    out['nll_deltas'] = deltas
    out['nll_values'] = np.array(
        [-np.sum(np.sum(gammaln(pseudo + (1 / t - 1)), axis=1) +
                 gammaln(n_r * (1 / t - 1)) - gammaln(z + n_r * (1 / t - 1)) -
                 n_r * gammaln(1 / t - 1)) for t in deltas])
    out['qcml'] = np.array(ref.dispersion.qcml(x, f=f.copy()))
    # four-replicate bin (R_c = 4 path)
    f4 = rng.lognormal(0.0, 0.3, size=(400, 4))
    mu4 = rng.gamma(2.0, 3.0, size=400) + 0.5
    x4 = rng.poisson(rng.gamma(1 / 0.05, mu4[:, None] * f4 * 0.05))
    x4 = x4.astype(np.int64)
    ok = x4.sum(axis=1) > 0
    x4, f4 = x4[ok], f4[ok]
    out['bin4_x'], out['bin4_f'] = x4, f4
    out['qcml4'] = np.array(ref.dispersion.qcml(x4, f=f4.copy()))
    out['equalize4_0.05'] = ref.scaled_nb.equalize(x4, f4.copy(), 0.05)
    # trend fit
    xs = np.arange(4, 201).astype(float)
    ys = 0.02 + 0.01 * np.exp(-xs / 6.) + 1e-4 * xs + \
        rng.normal(0, 1, len(xs)) * (2e-4 + 1e-5 * xs)
    ys[0] = ys.max() + 0.01
    out['trend_x'], out['trend_y'] = xs, ys
    xq = np.concatenate([np.arange(0, 206, dtype=float), [0.5, 4.5, 9.75]])
    out['trend_q'] = xq
    fn = ref.lowess.weighted_lowess_fit(xs, ys, left_boundary=ys[0],
                                        auto_frac_factor=15.)
    out['trend_weighted'] = fn(xq.copy())
    fn = ref.lowess.weighted_lowess_fit(xs, ys, left_boundary=ys[0],
                                        frac=0.2, auto_frac_factor=15.)
    out['trend_weighted_frac0.2'] = fn(xq.copy())
    fn = ref.lowess.lowess_fit(xs, ys, left_boundary=ys[0])
    out['trend_plain'] = fn(xq.copy())
    # lrt
    n = 3000
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    f = rng.lognormal(0.0, 0.25, size=(n, 4))
    mu = rng.gamma(1.5, 10.0, size=n) + 0.5
    eff = np.where(rng.random(n) < 0.2, 1.6, 1.0)
    m = mu[:, None] * f * np.where(design[:, 1][None, :], eff[:, None], 1.0)
    x = rng.poisson(rng.gamma(1 / 0.02, m * 0.02)).astype(np.int64)
    x[:4] = [[0, 1, 1, 0], [1, 0, 0, 1], [3, 0, 0, 2], [0, 1, 40, 50]]
    disp = np.stack([0.01 + rng.random(n) * 0.05,
                     0.01 + rng.random(n) * 0.05], axis=1)
    ok = (x[:, :2].sum(axis=1) > 0) & (x[:, 2:].sum(axis=1) > 0)
    x, f, disp = x[ok], f[ok], disp[ok]
    wide = np.dot(disp, design.T.astype(float))
    for refit in (True, False):
        p, llr, mu0, mu1 = ref.lrt.lrt(x, f, wide, design, refit_mu=refit)
        tag = 'refit' if refit else 'norefit'
        out['lrt_%s_p' % tag], out['lrt_%s_llr' % tag] = p, llr
        out['lrt_%s_mu0' % tag], out['lrt_%s_mu1' % tag] = mu0, mu1
    out['lrt_x'], out['lrt_f'], out['lrt_disp'] = x, f, disp
    out['lrt_design'] = design
    return out


def main():
    Ref = refrun.reference_class()
    ref = refrun.reference_modules()
    with open(os.path.join(HERE, 'reference_kats.json'), 'w') as h:
        json.dump(kats(ref), h, indent=1)
    np.savez_compressed(os.path.join(HERE, 'ref_pipeline.npz'),
                        **pipeline(Ref))
    np.savez_compressed(os.path.join(HERE, 'ref_stages.npz'), **stages(ref))
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
