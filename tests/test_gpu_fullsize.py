"""GPU, BASELINE.json configs[1] at FULL size (mouse genome-wide, 10 kb, 2-vs-2;
48 M union pixels): the oracle cannot run this in test time, so the run is
checked through size-independent properties of every step
(hic3defdr/analysis/analysis.py:28-303):

  prepare  pixels strictly sorted by (row, col), distances within the cap,
           raw == the replicate matrices at the union pixels and every stored
           in-band entry of an unfiltered bin pair is in the union, size
           factors depend on distance only, scaled * size_factors == balanced,
           disp_idx == the reference's filter recomputed from scaled
  disp     one dispersion per (distance, condition), the trend's value
  lrt      llr <= 0 up to round-off, p == chi2(1).sf(-2 llr), the likelihood
           equations of mu_hat hold
  bh       q >= p, q monotone in p, q(max p) == max p, identical ties
  run      two runs give bit-identical outputs (deterministic reductions);
           a chromosome prepared alone equals its slice of the genome run
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIST_MAX = 200


@pytest.fixture(scope='module')
def genome():
    import torch
    import bench
    from hic3defdr_b200 import engine, staging
    from hic3defdr_b200.synth import MM10_10KB
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    inputs = []
    for i, (c, n) in enumerate(MM10_10KB.items()):
        mats, bias = bench.gen_chrom_device(n, 4, DIST_MAX,
                                            20261018 + 1000 + 100 * i, 300.0)
        inputs.append((staging.csr_to_device(mats, n), bias))
    torch.cuda.synchronize()
    states, dpd, fns, stats = engine.run_to_qvalues(inputs, design,
                                                    dist_max=DIST_MAX)
    torch.cuda.synchronize()
    return dict(inputs=inputs, states=states, dpd=dpd, fns=fns, stats=stats,
                design=design, names=list(MM10_10KB))


def test_full_size_shape(genome):
    n_px = sum(int(s['row'].numel()) for s in genome['states'])
    n_d = sum(int(s['disp_index'].numel()) for s in genome['states'])
    assert 45_000_000 < n_px < 53_000_000          # band cells = 52.5 M
    assert 0.7 * n_px < n_d < 0.9 * n_px
    assert genome['stats']['capped_segments'] == 0


def test_prepare_properties(genome):
    import torch
    for (csr, bias_raw), st in zip(genome['inputs'], genome['states']):
        row, col = st['row'].long(), st['col'].long()
        n = csr.n_bins
        key = row * n + col
        assert bool((key[1:] > key[:-1]).all())                 # strictly sorted
        d = col - row
        assert int(d.min()) >= 0 and int(d.max()) <= DIST_MAX
        bias = st['bias']
        ok_bin = (bias != 0).all(dim=1)
        assert bool(ok_bin[row].all()) and bool(ok_bin[col].all())
        # raw: the replicate matrices at the union pixels; every pixel has data
        assert bool((st['raw'] >= 0).all())
        assert bool((st['raw'].sum(dim=1) > 0).all())
        nnz_in_band = 0
        for r in range(csr.n_reps):
            ip = csr.indptr[r].long()
            rows_r = torch.repeat_interleave(
                torch.arange(n, device='cuda'), ip[1:] - ip[:-1])
            cols_r = csr.indices[r].long()
            keep = (cols_r - rows_r <= DIST_MAX) & (cols_r >= rows_r) & \
                ok_bin[rows_r] & ok_bin[cols_r] & (csr.data[r] != 0)
            k_r = rows_r[keep] * n + cols_r[keep]
            pos = torch.searchsorted(key, k_r)
            assert bool((pos < key.numel()).all())
            assert bool((key[pos] == k_r).all())                # in the union
            got = torch.zeros(key.numel(), dtype=torch.int64, device='cuda')
            got[pos] = csr.data[r][keep]
            assert bool((got == st['raw'][:, r]).all())         # and nothing else
            nnz_in_band += int(keep.sum())
        assert nnz_in_band >= key.numel()
        # size factors: a function of distance, positive, finite
        sf = st['size_factors']
        assert bool(torch.isfinite(sf).all()) and bool((sf > 0).all())
        table = torch.zeros((DIST_MAX + 1, csr.n_reps), dtype=torch.float64,
                            device='cuda')
        table[d] = sf
        assert bool((table[d] == sf).all())
        # scaled = raw / (bias_i bias_j) / size_factors
        bal = st['raw'].double() / (bias[row] * bias[col])
        torch.testing.assert_close(st['scaled'] * sf, bal, rtol=1e-13, atol=0)
        # disp_idx (analysis.py:111-115)
        design = torch.from_numpy(genome['design']).cuda()
        means = torch.stack([st['scaled'][:, design[:, c]].mean(dim=1)
                             for c in range(design.shape[1])], dim=1)
        want = (means >= 1.0).all(dim=1) & (d >= 4)
        edge = ((means - 1.0).abs() < 1e-12).any(dim=1)         # mean == thresh
        got = st['disp_idx'].bool()
        assert bool((got == want)[~edge].all())
        assert bool((torch.nonzero(got).flatten() ==
                     st['disp_index'].long()).all())


def test_dispersion_lrt_properties(genome):
    import torch
    from scipy import stats as sps
    dpd, fns = genome['dpd'], genome['fns']
    assert dpd.shape == (DIST_MAX + 1, 2)
    assert np.isnan(dpd[:4]).all() and np.isfinite(dpd[4:]).all()
    assert (dpd[4:] > 0).all() and (dpd[4:] < 0.1).all()
    # generator: dispersion 0.01 + 1e-4 d; qCML recovers its scale where the
    # counts are large (the mean filter truncates the low-count distances)
    d = np.arange(4, 101)
    ratio = dpd[4:101].mean(axis=1) / (0.01 + 1e-4 * d)
    assert (ratio > 0.5).all() and (ratio < 2.0).all(), ratio
    table = np.stack([fn(np.arange(DIST_MAX + 1)) for fn in fns], axis=1)
    tdev = torch.from_numpy(table).cuda()
    for st in genome['states']:
        idx = st['disp_index'].long()
        dist = (st['col'].long() - st['row'].long())[idx]
        assert bool((st['disp'] == tdev[dist]).all())
        llr, p = st['llr'], st['pvalues']
        assert bool(torch.isfinite(llr).all()) and bool(torch.isfinite(p).all())
        assert float(llr.max()) <= 1e-9
        assert bool((p >= 0).all()) and bool((p <= 1).all())
        # spot check of p == chi2(1).sf(-2 llr) and of the likelihood equations
        sel = torch.linspace(0, idx.numel() - 1, 20000, device='cuda').long()
        l, pp = llr[sel].cpu().numpy(), p[sel].cpu().numpy()
        np.testing.assert_allclose(pp, sps.chi2(1).sf(np.maximum(-2 * l, 0)),
                                   rtol=1e-9, atol=1e-300)
        u = idx[sel]
        bias = st['bias']
        f = bias[st['row'].long()[u]] * bias[st['col'].long()[u]] * \
            st['size_factors'][u]
        x = st['raw'][u].double()
        mu0 = st['mu_hat_null'][sel][:, None]
        a = st['disp'][sel] @ torch.from_numpy(
            genome['design'].T.astype(np.float64)).cuda()
        score = ((x - mu0 * f) / (mu0 + a * mu0 * mu0 * f)).sum(dim=1)
        scale = (x / (mu0 + a * mu0 * mu0 * f)).sum(dim=1)
        assert float((score.abs() / scale).max()) < 1e-9


def test_bh_properties(genome):
    import torch
    p = torch.cat([s['pvalues'] for s in genome['states']])
    q = torch.cat([s['qvalues'] for s in genome['states']])
    assert p.numel() > 30_000_000
    assert bool((q >= p).all()) and bool((q <= 1).all())
    order = torch.argsort(p)
    qs, ps = q[order], p[order]
    assert bool((qs[1:] >= qs[:-1]).all())                      # monotone
    assert float(qs[-1]) == float(ps[-1])                       # q(max p) = p
    same = ps[1:] == ps[:-1]
    assert bool((qs[1:][same] == qs[:-1][same]).all())          # ties
    n = p.numel()
    # definition at the top end and at a few sampled ranks
    k = torch.tensor([0, n // 7, n // 2, n - 2], device='cuda')
    raw = ps * n / torch.arange(1, n + 1, device='cuda', dtype=torch.float64)
    suffix_min = torch.flip(torch.cummin(torch.flip(raw, [0]), 0).values, [0])
    torch.testing.assert_close(qs[k], suffix_min[k].clamp(max=1.0),
                               rtol=1e-12, atol=0)


def test_deterministic_and_chromosome_independent(genome):
    import torch
    from hic3defdr_b200 import engine
    states2, dpd2, _, _ = engine.run_to_qvalues(genome['inputs'],
                                                genome['design'],
                                                dist_max=DIST_MAX)
    assert np.array_equal(genome['dpd'], dpd2, equal_nan=True)
    for a, b in zip(genome['states'], states2):
        for k in ('row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx',
                  'disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
                  'qvalues'):
            assert torch.equal(a[k], b[k]), k
    # chr19 alone: prepare_data has no cross-chromosome state
    i = genome['names'].index('chr19')
    csr, bias = genome['inputs'][i]
    alone = engine.prepare_chrom(csr, bias, genome['design'],
                                 dist_max=DIST_MAX)
    for k in ('row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx'):
        assert torch.equal(alone[k], genome['states'][i][k]), k
