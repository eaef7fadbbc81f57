"""GPU parity of the simulate -> balance loop:
cluster perturbation, KR balancing and the sparse-bin filter against outputs
recorded from the unmodified reference (tests/golden/make_golden_sim.py); the
negative-binomial sampler (a different realisation of the same distribution:
counter-based Philox streams instead of numpy's global generator) against the
distribution itself."""
import os

import numpy as np
import pytest
import scipy.sparse as sparse
import scipy.stats as stats

from tests.helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _gold():
    return np.load(os.path.join(GOLDEN, 'ref_sim.npz'))


def _clusters(table):
    return [[(int(r), int(c)) for i, r, c in table if i == k]
            for k in range(int(table[:, 0].max()) + 1)]


def test_perturbation_vs_recorded_reference():
    """value += value * footprint * effect, cluster after cluster (overlaps
    compound, the matrix edge clips the dilation): 1e-12 (the order of the
    multiplications differs)."""
    import torch
    from hic3defdr_b200 import ops, simulation as hsim
    from hic3defdr_b200._native import lib, ptr
    g = _gold()
    row, col, mean = g['pt_row'], g['pt_col'], g['pt_mean']
    keys, factor = hsim.perturbation_factors(
        _clusters(g['pt_clusters']), g['pt_effects'], (120, 120))
    pk = ops.dev((row.astype(np.int64) << 32) | col.astype(np.int64))
    m = ops.dev(mean).clone()
    kd, fd = ops.dev(keys), ops.dev(factor)
    lib().call('h3d_perturb', ptr(pk), pk.numel(), ptr(kd), ptr(fd), len(keys),
               ptr(m), ops._stream())
    np.testing.assert_allclose(m.cpu().numpy(), g['pt_out'], rtol=1e-12)
    assert (m.cpu().numpy() != mean).sum() > 50
    # the host mirror of perturb_cluster itself
    sp = sparse.coo_matrix((mean, (row, col)), shape=(120, 120)).tolil()
    for cl, e in zip(_clusters(g['pt_clusters']), g['pt_effects']):
        if e:
            hsim.perturb_cluster(sp, cl, e)
    np.testing.assert_allclose(sp.tocsr()[row, col].A1, g['pt_out'], rtol=1e-12)


def test_kr_balance_vs_recorded_reference():
    from hic3defdr_b200.balancing import kr_balance
    g = _gold()
    a = sparse.csr_matrix((g['kr_data'], g['kr_indices'], g['kr_indptr']),
                          shape=(1500, 1500))
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        balanced, bias, res = kr_balance(a, fl=1)
    assert buf.getvalue().startswith('it in. it res')
    # same inexact-Newton path: as many outer iterations, residuals within the
    # round-off of the dot products; the iteration stops at tol = 1e-6, the
    # bias vectors agree far inside that
    assert len(res) == len(g['kr_res'])
    np.testing.assert_allclose(res, g['kr_res'], rtol=1e-6)
    assert np.array_equal(bias == 0, g['kr_bias'] == 0)
    np.testing.assert_allclose(bias, g['kr_bias'], rtol=1e-9)
    assert float(balanced.sum()) == pytest.approx(float(g['kr_balanced_sum']),
                                                  rel=1e-9)
    assert sparse.tril(balanced, k=-1).nnz == 0          # upper triangular in, upper out
    _, bias0, res0 = kr_balance(a, fl=0)
    assert len(res0) == 0 and np.array_equal(bias0, bias)


def test_filter_sparse_rows_vs_recorded_reference():
    from hic3defdr_b200.balancing import filter_sparse_rows_count
    g = _gold()
    a = sparse.csr_matrix((g['kr_data'], g['kr_indices'], g['kr_indptr']),
                          shape=(1500, 1500))
    filt = filter_sparse_rows_count(a, min_nnz=50, k=100).tocsr()
    kept = np.flatnonzero(np.diff(filt.indptr) > 0)
    np.testing.assert_array_equal(kept, g['fs_kept_rows'])
    assert filt.nnz == int(g['fs_nnz'])
    assert float(filt.sum()) == float(g['fs_sum'])
    assert 0 < len(kept) < 1500
    dense = filter_sparse_rows_count(a.toarray(), min_nnz=50, k=100)
    assert np.array_equal(dense, filt.toarray())


@pytest.mark.parametrize('mu,phi', [(0.7, 0.02), (4.0, 0.3), (35.0, 0.01),
                                    (800.0, 0.05), (12.0, 0.0), (3.0, 2.5)])
def test_nb_sampler_distribution(mu, phi):
    """4e5 draws at one (mean, dispersion): mean and variance of
    NB(mu, mu + phi mu^2) (hic3defdr/util/scaled_nb.py:36-48, mvr) within 5
    standard errors, chi-square goodness of fit of the pmf."""
    import torch
    from hic3defdr_b200 import ops
    from hic3defdr_b200._native import lib, ptr
    n = 400000
    row = ops.dev(np.zeros(n, dtype=np.int32))
    col = ops.dev(np.zeros(n, dtype=np.int32))
    mean = ops.dev(np.full(n, mu))
    bias = ops.dev(np.ones((1, 1)))
    sf = ops.dev(np.ones(1))
    disp = ops.dev(np.array([phi]))
    out = torch.empty(n, dtype=torch.int64, device='cuda')
    lib().call('h3d_nb_simulate', ptr(row), ptr(col), ptr(mean), n, ptr(bias),
               1, ptr(sf), 0, 1, ptr(disp), 0, 0, 12345, ptr(out), None,
               ops._stream())
    x = out.cpu().numpy()
    var = mu + phi * mu * mu
    assert abs(x.mean() - mu) < 5 * np.sqrt(var / n)
    # variance of the sample variance ~ (kurtosis term) var^2 / n: loose bound
    assert abs(x.var() - var) < 0.03 * var + 8 * var / np.sqrt(n)
    if phi > 0:
        dist = stats.nbinom(1.0 / phi, 1.0 / (1.0 + phi * mu))
    else:
        dist = stats.poisson(mu)
    lo, hi = int(dist.ppf(1e-4)), int(dist.ppf(1 - 1e-4))
    edges = np.unique(np.round(np.linspace(lo, hi + 1, 40)).astype(int))
    obs = np.histogram(x, bins=np.concatenate([[-0.5], edges + 0.5, [np.inf]]))[0]
    cdf = np.concatenate([[0.0], dist.cdf(edges), [1.0]])
    exp = n * np.diff(cdf)
    ok = exp > 20
    chi2 = float(((obs[ok] - exp[ok]) ** 2 / exp[ok]).sum())
    assert chi2 < stats.chi2(ok.sum() - 1).ppf(1 - 1e-5), (chi2, ok.sum())


def test_sampler_streams_are_reproducible_and_distinct():
    import torch
    from hic3defdr_b200 import ops
    from hic3defdr_b200._native import lib, ptr
    n = 10000
    row = ops.dev(np.arange(n, dtype=np.int32) % 50)
    col = ops.dev(np.arange(n, dtype=np.int32) % 50 + 3)
    mean = ops.dev(np.full(n, 20.0))
    bias = ops.dev(np.ones((60, 2)))
    sf = ops.dev(np.ones(2))
    disp = ops.dev(np.full(4, 0.05))

    def draw(rep, seed):
        out = torch.empty(n, dtype=torch.int64, device='cuda')
        lib().call('h3d_nb_simulate', ptr(row), ptr(col), ptr(mean), n,
                   ptr(bias), 2, ptr(sf), 0, 4, ptr(disp), 0, rep, seed,
                   ptr(out), None, ops._stream())
        return out.cpu().numpy()
    a = draw(0, 7)
    assert np.array_equal(a, draw(0, 7))
    assert (a != draw(1, 7)).mean() > 0.8 and (a != draw(0, 8)).mean() > 0.8
    assert abs(np.corrcoef(a, draw(1, 7))[0, 1]) < 0.05


def test_class_simulate_balance_rerun(tmp_path):
    """the README loop on a small data set: run to q-values, simulate 2 + 2
    replicates from condition A's fit with injected loops, filter + KR-balance
    them, run the pipeline on the simulation, evaluate against the labels."""
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.balancing import filter_sparse_rows_count, kr_balance
    from hic3defdr_b200.synth import write_dataset
    root = str(tmp_path)
    kw = write_dataset(os.path.join(root, 'in'), {'cA': 800}, n_reps=4,
                       dist_max=50, config=5, amp=400.0, loops=True)
    h = HiC3DeFDR(outdir=os.path.join(root, 'out'), dist_thresh_max=50, **kw)
    h.prepare_data(n_threads=0)
    h.estimate_disp(n_threads=0)
    np.random.seed(42)
    sim = os.path.join(root, 'sim')
    h.simulate('A', outdir=sim, n_threads=0)
    labels = np.loadtxt(os.path.join(sim, 'labels_cA.txt'), dtype='U7')
    assert set(labels) <= {'constit', 'A', 'B'} and (labels != 'constit').any()
    design = os.path.join(sim, 'design.csv')
    assert os.path.isfile(design)
    row = np.load(os.path.join(root, 'out', 'row_cA.npy'))
    scaled = np.load(os.path.join(root, 'out', 'scaled_cA.npy'))
    n_px = int((scaled[:, :2].mean(axis=1) > 0).sum())
    reps = ['A1', 'A2', 'B1', 'B2']
    total = 0
    for rep in reps:
        m = sparse.load_npz(os.path.join(sim, '%s_cA_raw.npz' % rep))
        assert m.shape == (800, 800) and m.nnz == n_px and m.dtype == np.int64
        total += m.sum()
        _, bias, _ = kr_balance(filter_sparse_rows_count(m, min_nnz=10, k=50),
                                fl=0)
        np.savetxt(os.path.join(sim, '%s_cA_kr.bias' % rep), bias)
    # two replicates of condition A's mean, tiled twice: the depth of the
    # input's A replicates over the simulated pixels, up to the perturbations
    # and the sampling noise
    raw = np.load(os.path.join(root, 'out', 'raw_cA.npy'))
    assert abs(total / (2.0 * raw[:, :2].sum()) - 1.0) < 0.03, \
        (total, raw[:, :2].sum())
    h2 = HiC3DeFDR(
        [os.path.join(sim, '%s_<chrom>_raw.npz' % r) for r in reps],
        [os.path.join(sim, '%s_<chrom>_kr.bias' % r) for r in reps],
        ['cA'], design, os.path.join(root, 'out_sim'), dist_thresh_max=50,
        loop_patterns={'A': kw['loop_patterns']['A']})
    h2.run_to_qvalues(n_threads=0)
    h2.evaluate('A', os.path.join(sim, 'labels_<chrom>.txt'))
    ev = np.load(os.path.join(root, 'out_sim', 'eval.npz'))
    auc = np.trapezoid(ev['tpr'], ev['fpr'])
    assert auc > 0.75, auc
