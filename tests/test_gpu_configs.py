"""GPU parity for the shapes of the other BASELINE.json configurations, at
sizes the oracle finishes in seconds:

  config 3  4-vs-4 replicates (R = 8: the <4>/<8> template instances)
  config 4  one chromosome with a wide band (dist cap 1000-2000 bins,
            n_bins_norm = 400: the generic union path, many empty distances)
  config 5  loop clusters -> loop_idx, BH restricted to loop pixels

Protocol: prepare outputs against the oracle
directly; every later stage against the oracle fed with OUR previous-stage
output (stage-isolated), tolerances as in tests/test_gpu_pipeline.py.
"""
import numpy as np
import pytest

from oracle import pipeline as op
from tests.helpers import check_pvalues

pytestmark = pytest.mark.gpu


def _design(n_reps):
    d = np.zeros((n_reps, 2), dtype=bool)
    d[:n_reps // 2, 0] = True
    d[n_reps // 2:, 1] = True
    return d


def _device_inputs(inputs):
    from hic3defdr_b200 import ops
    return [(ops.DeviceCSR(m), ops.dev(b)) for m, b in inputs]


def _check_prepare(st, want):
    for k in ('row', 'col', 'raw'):
        np.testing.assert_array_equal(st[k].cpu().numpy(), want[k], err_msg=k)
    np.testing.assert_array_equal(st['disp_idx'].cpu().numpy().astype(bool),
                                  want['disp_idx'])
    for k in ('size_factors', 'scaled'):
        np.testing.assert_allclose(st[k].cpu().numpy(), want[k], rtol=1e-12,
                                   err_msg=k)


def _check_later_stages(states, oracle_states, dpd, design, dist_max,
                        loops=None, llr_floor=1e-8):
    """disp_per_dist against the oracle's; trend / LRT / BH stage-isolated."""
    n_cond = design.shape[1]
    fits = []
    for c in range(n_cond):
        ok = np.isfinite(dpd[:, c])
        xs, ys = np.arange(dist_max + 1)[ok], dpd[:, c][ok]
        fits.append(op.weighted_trend(xs, ys, left_boundary=ys[0]))
    ps = []
    for i, (st, ost) in enumerate(zip(states, oracle_states)):
        di = ost['disp_idx']
        row, col = ost['row'][di], ost['col'][di]
        disp = st['disp'].cpu().numpy()
        want_disp = np.stack([op.eval_trend(f, col - row) for f in fits], 1)
        np.testing.assert_allclose(disp, want_disp, rtol=1e-9)
        f = op.combined_factor(ost['bias'], row, col, ost['size_factors'][di])
        p, llr, mu0, mu1 = op.lrt(ost['raw'][di], f,
                                  np.dot(disp, design.T.astype(float)), design)
        np.testing.assert_allclose(st['mu_hat_null'].cpu().numpy(), mu0,
                                   rtol=1e-9)
        np.testing.assert_allclose(st['mu_hat_alt'].cpu().numpy(), mu1,
                                   rtol=1e-9)
        np.testing.assert_allclose(st['llr'].cpu().numpy(), llr, rtol=0,
                                   atol=1e-10)
        good = -2 * llr >= llr_floor
        got_p = st['pvalues'].cpu().numpy()
        got_llr = st['llr'].cpu().numpy()
        check_pvalues(got_p, got_llr, p, llr, good, n_cond - 1)
        if loops is not None:
            li = op.loop_membership(row, col, loops[i])
            np.testing.assert_array_equal(
                st['loop_idx'].cpu().numpy().astype(bool), li)
            ps.append(got_p[li])
        else:
            ps.append(got_p)
    q = op.bh(np.concatenate(ps))
    got_q = np.concatenate([st['qvalues'].cpu().numpy() for st in states])
    np.testing.assert_allclose(got_q, q, rtol=1e-12)


def test_config3_eight_replicates():
    from hic3defdr_b200 import engine
    from hic3defdr_b200.synth import make_chrom
    design = _design(8)
    dist_max = 40
    inputs = []
    for ci, n in enumerate((520, 430)):
        mats, bias, _ = make_chrom(n, 8, dist_max, seed=31000 + 100 * ci,
                                   amp=150.0, res_scale=0.5)
        inputs.append((mats, bias))
    want = op.run_to_qvalues(inputs, design, dist_max=dist_max)
    states, dpd, fns, stats = engine.run_to_qvalues(
        _device_inputs(inputs), design, dist_max=dist_max)
    for st, ost in zip(states, want['chroms']):
        _check_prepare(st, ost)
    ok = np.isfinite(want['disp_per_dist'])
    assert np.array_equal(ok, np.isfinite(dpd))
    # ~900-pixel bins: the reference's own permutation self-noise is ~1e-6
    # (tests/golden disp_selfnoise); absolute term: Brent's xatol (see smoke())
    np.testing.assert_allclose(dpd[ok], want['disp_per_dist'][ok], rtol=5e-6,
                               atol=5e-8)
    # with 8 replicates the two log-likelihoods (~1e3) cancel to ~1e-13, so the
    # relative p check starts at -2 llr = 1e-7 (d p / d x ~ 1 / sqrt(2 pi x));
    # below that the absolute llr check above is the meaningful one
    _check_later_stages(states, want['chroms'], dpd, design, dist_max,
                        llr_floor=1e-7)


@pytest.mark.parametrize('n,n_reps,dist_max,n_bins,amp', [
    (2300, 4, 2000, 400, 30.0),      # config 4 shape: generic union path
    (1100, 8, 1000, 200, 40.0),      # R = 8 and a band too wide to stage
    (900, 4, 700, 140, 60.0),        # staged path, one warp per block
])
def test_config4_wide_band_prepare(n, n_reps, dist_max, n_bins, amp):
    """prepare_data on a single chromosome with a wide, sparse band"""
    from hic3defdr_b200 import engine, ops
    from hic3defdr_b200.synth import make_chrom
    design = _design(n_reps)
    mats, bias, _ = make_chrom(n, n_reps, dist_max, seed=41000 + n, amp=amp,
                               res_scale=0.1)
    want = op.prepare_chrom(mats, bias, design, dist_max=dist_max,
                            n_bins=n_bins)
    st = engine.prepare_chrom(ops.DeviceCSR(mats), ops.dev(bias), design,
                              dist_max=dist_max, n_bins=n_bins)
    _check_prepare(st, want)
    assert int(want['disp_idx'].sum()) > 0


def test_config4_wide_band_dispersion_has_empty_distances():
    """dispersion over a band where most distances hold no tested pixel:
    NaN rows exactly where the oracle has them, values within tolerance on a
    sample of the occupied distances (the oracle's per-distance optimisations
    are slow, so only those are run)"""
    from hic3defdr_b200 import engine
    from hic3defdr_b200.synth import make_chrom
    design = _design(4)
    dist_max, n = 600, 1500
    mats, bias, _ = make_chrom(n, 4, dist_max, seed=42000, amp=60.0,
                               res_scale=0.1)
    want = op.prepare_chrom(mats, bias, design, dist_max=dist_max, n_bins=120)
    states = [engine.prepare_chrom(*_device_inputs([(mats, bias)])[0], design,
                                   dist_max=dist_max, n_bins=120)]
    dpd, fns, stats = engine.estimate_disp(states, design, dist_max)
    # bins of a handful of pixels with dispersion ~ 60: the reference's loop
    # never ends there (include/h3d.h, H3D_QCML_MAX_OUTER); they are reported
    assert stats['capped_segments'] > 0
    di = want['disp_idx']
    row, col = want['row'][di], want['col'][di]
    dist = col - row
    f = op.combined_factor(want['bias'], row, col, want['size_factors'][di])
    raw = want['raw'][di]
    occupied = np.bincount(dist, minlength=dist_max + 1) > 0
    assert np.array_equal(np.isfinite(dpd[:, 0]), occupied)
    assert (~occupied).sum() > 100                    # the point of the test
    counts = np.bincount(dist, minlength=dist_max + 1)
    for d in np.flatnonzero(counts >= 20)[::25][:6]:
        sel = dist == d
        for c in range(2):
            ref = op.qcml(raw[sel][:, design[:, c]], f=f[sel][:, design[:, c]])
            assert dpd[d, c] == pytest.approx(ref, rel=5e-6, abs=5e-8)


def test_config5_loops_restrict_bh():
    from hic3defdr_b200 import engine, ops
    from hic3defdr_b200.synth import make_chrom
    design = _design(4)
    dist_max = 50
    inputs, loops = [], []
    for ci, n in enumerate((600, 450)):
        mats, bias, clusters = make_chrom(n, 4, dist_max,
                                          seed=51000 + 100 * ci, amp=250.0,
                                          loops=True)
        inputs.append((mats, bias))
        loops.append(clusters)
    want = op.run_to_qvalues(inputs, design, dist_max=dist_max, loops=loops)
    dev_in = _device_inputs(inputs)
    states = []
    for (csr, b), cl in zip(dev_in, loops):
        px = [tuple(p) for c in cl for p in c]
        states.append(engine.prepare_chrom(csr, b, design, dist_max=dist_max,
                                           loop_pixels=px))
    dpd, fns, stats = engine.estimate_disp(states, design, dist_max)
    for st in states:
        engine.lrt_chrom(st, design)
    engine.bh(states, use_loop_idx=True)
    for st, ost in zip(states, want['chroms']):
        _check_prepare(st, ost)
        assert st['qvalues'].numel() == int(ost['loop_idx'].sum())
    ok = np.isfinite(want['disp_per_dist'])
    np.testing.assert_allclose(dpd[ok], want['disp_per_dist'][ok], rtol=5e-6,
                               atol=5e-8)
    _check_later_stages(states, want['chroms'], dpd, design, dist_max,
                        loops=loops)
    n_loop = sum(int(st['loop_idx'].sum()) for st in states)
    assert 0 < n_loop < sum(st['loop_idx'].numel() for st in states)


@pytest.mark.parametrize('norm', ['simple_scaling', 'median_of_ratios',
                                  'conditional_scaling'])
def test_other_norms_through_prepare(norm):
    from hic3defdr_b200 import engine, ops
    from hic3defdr_b200.synth import make_chrom
    design = _design(4)
    mats, bias, _ = make_chrom(500, 4, 40, seed=61000, amp=200.0)
    want = op.prepare_chrom(mats, bias, design, dist_max=40, norm=norm)
    st = engine.prepare_chrom(ops.DeviceCSR(mats), ops.dev(bias), design,
                              dist_max=40, norm=norm)
    _check_prepare(st, want)
