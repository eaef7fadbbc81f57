"""GPU parity: prepare_data kernels (union, gathers, size factors, scaling,
disp_idx, loop_idx) through the C ABI against the reference's doctest vectors,
the recorded reference outputs and the oracle."""
import numpy as np
import pytest
import scipy.sparse as sparse

from oracle import pipeline as op
from tests.helpers import load_kats, load_pipeline_golden

pytestmark = pytest.mark.gpu


def test_sparse_union_doctest():
    from hic3defdr_b200 import ops
    k = load_kats()['sparse_union']
    mats = [sparse.csr_matrix(np.array(k['rep1'])),
            sparse.csr_matrix(np.array(k['rep2']))]
    u = ops.union_gather(ops.DeviceCSR(mats), k['dist_thresh'], None)
    row, col = u['row'].cpu().numpy(), u['col'].cpu().numpy()
    assert row.dtype == np.int32
    assert list(zip(row.tolist(), col.tolist())) == \
        [tuple(p) for p in k['pixels']]
    assert u['dist'].cpu().numpy().tolist() == k['dist']
    np.testing.assert_array_equal(u['raw'].cpu().numpy(), np.array(k['data']))


def test_conditional_mor_doctest():
    from hic3defdr_b200 import ops
    k = load_kats()['conditional_mor']
    got = ops.conditional_mor(np.array(k['data']),
                              np.array(k['dist'])).cpu().numpy()
    np.testing.assert_allclose(got, k['doc'], rtol=0, atol=5e-9)
    np.testing.assert_allclose(got, k['full'], rtol=1e-12)


def test_prepare_stage_vs_recorded_reference():
    from hic3defdr_b200 import ops
    gold = load_pipeline_golden()
    g = gold['g']
    for c, (mats, bias_raw), loops in zip(gold['chroms'], gold['inputs'],
                                          gold['loops']):
        bias = ops.filter_bias(bias_raw, 0.1)
        np.testing.assert_array_equal(bias.cpu().numpy(),
                                      op.filter_bias(bias_raw, 0.1))
        u = ops.union_gather(ops.DeviceCSR(mats), gold['dist_max'], bias)
        np.testing.assert_array_equal(u['row'].cpu().numpy(), g['row_%s' % c])
        np.testing.assert_array_equal(u['col'].cpu().numpy(), g['col_%s' % c])
        np.testing.assert_array_equal(u['raw'].cpu().numpy(), g['raw_%s' % c])
        n_bins = int(gold['dist_max'] / 5)
        table = ops.size_factor_table(u['balanced'], u['dist'],
                                      gold['dist_max'], n_bins,
                                      'conditional_mor')
        scaled, sf, disp_idx = ops.scale_filter(
            u['row'], u['col'], u['balanced'], table, gold['design'],
            gold['dist_max'], 1.0, gold['dist_min'])
        # FP64 tolerance 1e-12: same IEEE operations, libm log/exp <= 1 ulp
        np.testing.assert_allclose(sf.cpu().numpy(), g['size_factors_%s' % c],
                                   rtol=1e-12)
        np.testing.assert_allclose(scaled.cpu().numpy(), g['scaled_%s' % c],
                                   rtol=1e-12)
        np.testing.assert_array_equal(disp_idx.cpu().numpy().astype(bool),
                                      g['disp_idx_%s' % c])
        index = ops.mask_to_index(disp_idx)
        np.testing.assert_array_equal(index.cpu().numpy(),
                                      np.where(g['disp_idx_%s' % c])[0])
        li = ops.loop_membership(u['row'], u['col'], index,
                                 set().union(*[set(x) for x in loops]))
        np.testing.assert_array_equal(li.cpu().numpy().astype(bool),
                                      g['loop_idx_%s' % c])


@pytest.mark.parametrize('n,n_keys', [(1, 3), (37, 5), (5000, 41), (300000, 201),
                                      (70000, 2001)])
def test_stable_rank_matches_stable_argsort(n, n_keys):
    from hic3defdr_b200 import ops
    rng = np.random.default_rng(n)
    keys = rng.integers(0, n_keys, size=n).astype(np.int32)
    rank, start = ops.stable_rank(keys, n_keys)
    np.testing.assert_array_equal(rank.cpu().numpy(), op.stable_rank(keys))
    np.testing.assert_array_equal(
        start.cpu().numpy(),
        np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=n_keys))]))


def test_equal_bin_vs_oracle():
    from hic3defdr_b200 import ops
    rng = np.random.default_rng(11)
    dist = rng.integers(0, 61, size=12345)
    np.testing.assert_array_equal(ops.equal_bin(dist, 12),
                                  op.equal_bin(dist, 12))


@pytest.mark.parametrize('norm', ['conditional_mor', 'conditional_scaling',
                                  'median_of_ratios', 'simple_scaling'])
@pytest.mark.parametrize('n_bins', [8, None])
def test_size_factor_modes_vs_oracle(norm, n_bins):
    from hic3defdr_b200 import ops
    rng = np.random.default_rng(5)
    n, r = 20000, 4
    dist = rng.integers(0, 41, size=n)
    data = rng.gamma(2.0, 5.0, size=(n, r)) * (0.8 + 0.1 * np.arange(r))
    data[rng.random((n, r)) < 0.1] = 0.0
    if norm == 'conditional_mor':
        want = op.conditional_size_factors(data, dist, n_bins)
        got = ops.conditional_mor(data, dist, n_bins).cpu().numpy()
    elif norm == 'conditional_scaling':
        want = op.conditional_size_factors(data, dist, n_bins,
                                           op.simple_scaling)
        got = ops.conditional_scaling(data, dist, n_bins).cpu().numpy()
    elif norm == 'median_of_ratios':
        want, got = op.median_of_ratios(data), \
            ops.median_of_ratios(data).cpu().numpy()
    else:
        want, got = op.simple_scaling(data), \
            ops.simple_scaling(data).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12)


def test_union_edge_cases_vs_oracle():
    """ragged rows, lower-triangle entries, explicit zeros, float data,
    negative values, zero / out-of-range / NaN bias, entries beyond the band."""
    from hic3defdr_b200 import ops
    rng = np.random.default_rng(99)
    n, r, dmax = 150, 3, 17
    mats = []
    for k in range(r):
        dense = rng.poisson(0.6, size=(n, n)).astype(float)
        dense[rng.random((n, n)) < 0.02] = -1.0
        dense[rng.random((n, n)) < 0.02] = 0.5
        m = sparse.csr_matrix(dense)
        m.data[rng.random(m.nnz) < 0.05] = 0.0       # explicit zeros
        mats.append(m)
    mats[0][10, :] = 0
    mats[0].eliminate_zeros()
    bias_raw = rng.lognormal(0, 0.3, size=(n, r))
    bias_raw[5, 1] = 0.01
    bias_raw[20, 0] = 50.0
    bias_raw[33, 2] = np.nan
    bias_raw[40, :] = 0.0
    want_bias = op.filter_bias(bias_raw, 0.1)
    bias = ops.filter_bias(bias_raw, 0.1)
    np.testing.assert_array_equal(bias.cpu().numpy(), want_bias)
    row, col = op.union_pixels(mats, dmax, bias=want_bias.copy())
    raw, bal = op.gather_raw_balanced(mats, row, col, want_bias)
    u = ops.union_gather(ops.DeviceCSR(mats), dmax, bias)
    np.testing.assert_array_equal(u['row'].cpu().numpy(), row)
    np.testing.assert_array_equal(u['col'].cpu().numpy(), col)
    np.testing.assert_array_equal(u['raw'].cpu().numpy(), raw)
    np.testing.assert_allclose(u['balanced'].cpu().numpy(), bal, rtol=1e-15)


def test_union_empty_and_int32_data():
    from hic3defdr_b200 import ops
    n = 40
    empty = sparse.csr_matrix((n, n), dtype=np.int32)
    u = ops.union_gather(ops.DeviceCSR([empty, empty]), 10, None)
    assert u['row'].numel() == 0 and u['raw'].shape == (0, 2)
    m = sparse.random(n, n, density=0.2, random_state=3, format='csr')
    m.data = (m.data * 10).astype(np.int32) + 1
    row, col = op.union_pixels([m, empty], 10)
    u = ops.union_gather(ops.DeviceCSR([m, empty.astype(np.int32)]), 10, None)
    np.testing.assert_array_equal(u['row'].cpu().numpy(), row)
    np.testing.assert_array_equal(u['col'].cpu().numpy(), col)


@pytest.mark.parametrize('norm,n_bins', [
    ('conditional_mor', 8), ('conditional_mor', 0), ('conditional_scaling', 8),
    ('median_of_ratios', 0), ('simple_scaling', 0)])
def test_size_factor_stages_equal_the_fused_call(norm, n_bins):
    """h3d_sf_group_bounds / _values / _group_reduce / _table (the entry
    points the row-sharded path chains with collectives in between) against
    h3d_size_factors and against the oracle, on one device."""
    from hic3defdr_b200 import ops
    from hic3defdr_b200.synth import make_chrom
    dist_max = 40
    mats, bias, _ = make_chrom(600, 4, dist_max, seed=991, amp=250.0)
    u = ops.union_gather(ops.DeviceCSR(mats), dist_max,
                         ops.filter_bias(bias, 0.1))
    bal, dist = u['balanced'], u['dist']
    want = ops.size_factor_table(bal, dist, dist_max, n_bins, norm)
    conditional = 'conditional' in norm
    rank, key_start = (ops.stable_rank(dist, dist_max + 1) if conditional
                       else (None, None))
    gstart = ops.sf_group_bounds(bal.shape[0], dist_max, n_bins, norm,
                                 key_start)
    assert gstart.numel() == ops.sf_num_groups(dist_max, n_bins, norm) + 1
    values = ops.sf_values(bal, rank, norm)
    red, valid = ops.sf_group_reduce(values, gstart, norm)
    got = ops.sf_table(red, gstart, key_start, dist_max, n_bins, norm)
    np.testing.assert_array_equal(got.cpu().numpy(), want.cpu().numpy())
    b, d = bal.cpu().numpy(), dist.cpu().numpy()
    if conditional:
        reducer = op.median_of_ratios if norm.endswith('mor') else \
            op.simple_scaling
        ref = op.conditional_size_factors(b, d, n_bins, reducer)
        np.testing.assert_allclose(got.cpu().numpy()[d], ref, rtol=1e-12)
        if norm.endswith('mor') and n_bins:
            bins = op.equal_bin(d, n_bins)
            n_ok = [int(np.all(b[bins == g] > 0, axis=1).sum())
                    for g in range(n_bins)]
            assert valid.cpu().numpy().tolist() == n_ok
    else:
        ref = op.median_of_ratios(b) if norm == 'median_of_ratios' else \
            op.simple_scaling(b)
        np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-12)
